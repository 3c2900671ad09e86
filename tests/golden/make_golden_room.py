"""Real-data fixture for BASELINE config 2 (sample_data/room_source.ply -> room_target.ply, the swap path): the 2.3 M-point
source scan does not travel to the GPU box, so every 8th point is kept (289 451 points, still >= 1.2 x the 94 052-point target,
hence still swapped by the file overload).  Run in the build container (needs /root/reference and oracle/_ref):

    python tests/golden/make_golden_room.py

Stores the two clouds, the authors' ground truth (room_source_groundtruth.txt), their published result on the full pair
(file_pairs_results.txt:11-15) and the reference's own result on the decimated pair for RANSAC seeds 1-3."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref import Ref, load_ply  # noqa: E402
from tests.golden.make_golden import parse_result_file  # noqa: E402

SAMPLE = "/root/reference/sample_data"
OUT = os.path.dirname(os.path.abspath(__file__))
STEP = 8

tgt = load_ply(os.path.join(SAMPLE, "room_target.ply"))
src = np.ascontiguousarray(load_ply(os.path.join(SAMPLE, "room_source.ply"))[::STEP])
gt = np.loadtxt(os.path.join(SAMPLE, "room_source_groundtruth.txt"))
published = parse_result_file(os.path.join(SAMPLE, "file_pairs_results.txt"), 1)
ref_T = []
for seed in (1, 2, 3):
    r = Ref(quiet=True)
    r.set_seed(seed)
    # registration(T, target_file, source_file): source >= 1.2 x target -> swapped, inverse returned (PLADE/plade.cpp:689-704)
    assert len(src) >= 1.2 * len(tgt)
    ok, T = r.registration_clouds(src, tgt)
    assert ok
    ref_T.append(np.linalg.inv(T.astype(np.float64)).astype(np.float32))
np.savez_compressed(os.path.join(OUT, "room_decimated.npz"), tgt=tgt.astype(np.float32), src=src.astype(np.float32),
                    gt=gt.astype(np.float64), published=published.astype(np.float64), ref_T=np.stack(ref_T), step=np.array([STEP]))
print("room_decimated.npz:", tgt.shape, src.shape)
