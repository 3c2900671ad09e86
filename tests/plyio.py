import numpy as np


def write_ply(path, xyzn):
    a = np.ascontiguousarray(xyzn, dtype="<f4").reshape(-1, 6)
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                 "property float nx\nproperty float ny\nproperty float nz\nend_header\n" % len(a)).encode())
        f.write(a.tobytes())
