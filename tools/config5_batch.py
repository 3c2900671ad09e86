"""BASELINE config 5: 64 synthetic 1 M-point pairs (make_pair seeds 1000..1063, 20 planes) written as PLY files and listed in a
file_pairs.txt, registered by the C++ CLI in batch mode (plade_b200_cli file_pairs.txt results.txt: the reference's pair-list
loop, PLADE/main.cpp:97-159, spread over all visible GPUs, PLY reading included), next to the reference's own CLI
(oracle/_ref_fast/plade_ref, PLADE/main.cpp unchanged) run as one single-threaded process per host core over the same list.
Prints and writes a JSON `batch` object: whole-box pairs/s of both sides, successes / symmetric flips / failures per side.

    python tools/config5_batch.py [--pairs 64] [--points 1000000] [--gpus 8] [--out gpurun_out/config5.json] [--skip-reference]

bench.py's cpu_baseline / reference arm rules apply: the reference binary is a prebuilt file of the oracle, only timed here."""
import argparse, json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from plade_b200.synth import make_pair, transform_error
sys.path.insert(0, os.path.join(ROOT, "tests"))
from plyio import write_ply


def _make_one(job):
    k, points, d = job
    t, s, gt = make_pair(n_points=points, n_planes=20, seed=1000 + k)
    tp, sp = os.path.join(d, "t%d.ply" % k), os.path.join(d, "s%d.ply" % k)
    write_ply(tp, t); write_ply(sp, s)
    return tp, sp, gt, float(np.linalg.norm(np.ptp(t[:, :3], axis=0)))


def parse_results(path, n):
    """[(ok, 4x4)] in list order (PLADE/main.cpp:136-147 text format)"""
    if not os.path.exists(path):
        return []
    out, lines = [], open(path).read().split("\n")
    i = 0
    while i < len(lines) and len(out) < n:
        if lines[i].startswith("target:"):
            ok = lines[i + 2].startswith("transformation")
            M = np.array([[float(x) for x in lines[i + 3 + r].split()] for r in range(4)])
            out.append((ok, M))
            i += 7
        else:
            i += 1
    return out


def judge(res, gts):
    ok = landed = flipped = failed = 0
    for (good, M), (gt, diag) in zip(res, gts):
        if not good:
            failed += 1
            continue
        ok += 1
        rot, tr = transform_error(M, gt, diag)
        if rot <= 0.5 and tr <= 5e-3:
            landed += 1
        elif rot > 20:
            flipped += 1
    return {"pairs": len(gts), "returned_true": ok, "within_0.5deg_5e-3": landed, "symmetric_flips": flipped, "failed": failed + (len(gts) - len(res))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=64)
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--gpus", type=int, default=0, help="0 = all visible")
    ap.add_argument("--workers-per-gpu", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "config5.json"))
    ap.add_argument("--skip-reference", action="store_true")
    a = ap.parse_args()
    d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    gts, names = [], []
    t0 = time.perf_counter()
    import multiprocessing as mp
    with mp.get_context("fork").Pool(min(os.cpu_count() or 1, 16)) as pool:
        made = pool.map(_make_one, [(k, a.points, d) for k in range(a.pairs)])
    for tp, sp, gt, diag in made:
        names += [tp, sp]
        gts.append((gt, diag))
    pairs_file = os.path.join(d, "file_pairs.txt")
    open(pairs_file, "w").write("\n".join(names) + "\n")
    print("wrote %d pairs (%.1f GB of PLY) in %.0f s" % (a.pairs, sum(os.path.getsize(n) for n in names) / 1e9, time.perf_counter() - t0), flush=True)
    env = dict(os.environ)
    if a.gpus:
        env["PLADE_DEVICES"] = str(a.gpus)
    if a.workers_per_gpu:
        env["PLADE_WORKERS_PER_GPU"] = str(a.workers_per_gpu)
    cli = os.path.join(ROOT, "plade_b200", "plade_b200_cli")
    doc = {"config": "BASELINE config 5: %d synthetic %d-pt pairs (seeds 1000..%d) from PLY files, whole box" % (a.pairs, a.points, 999 + a.pairs),
           "host_cores": os.cpu_count()}
    # (a) the CLI as a user runs it: ONE process, start to finish -- includes creating a CUDA context on every GPU, loading the
    #     kernels there and allocating / page-locking the scratch of every worker, which for 64 pairs outweighs the work
    out_file = os.path.join(d, "results_gpu.txt")
    for label, extra in (("gpu_cli_one_shot", {}), ("gpu_cli_one_shot_process_per_gpu", {"PLADE_CLI_FORK": "1"})):
        e2 = dict(env, PLADE_TIMING="1", **extra)
        t0 = time.perf_counter()
        r = subprocess.run([cli, pairs_file, out_file], env=e2, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        dt = time.perf_counter() - t0
        res = parse_results(out_file, a.pairs)
        workers = [l for l in r.stderr.split("\n") if l.startswith("[plade batch worker")]
        doc[label] = {"seconds": dt, "pairs_per_s": a.pairs / dt, "exit_code": r.returncode, **judge(res, gts),
                      "command": "plade_b200_cli file_pairs.txt results.txt" + ("  (PLADE_CLI_FORK=1: one child process per GPU)" if extra else "  (all GPUs from one process)"),
                      "gpus": a.gpus or "all visible", "worker_lines": workers[:4],
                      "note": "includes process start-up (CUDA contexts, scratch allocation, page-locking)"}
        print(label, json.dumps({k: v for k, v in doc[label].items() if k != "worker_lines"}), flush=True)
        for w in workers[:2]:
            print("   ", w[:300], flush=True)
    # (b) the same call the CLI makes (plade_register_batch, PLY reading included) in a process that is already warm: the
    #     steady-state whole-box throughput of a service that keeps its contexts
    import plade_b200
    n_dev = a.gpus or plade_b200.device_count()
    per_gpu = a.workers_per_gpu or max(1, min(4, (os.cpu_count() or 1) // (2 * n_dev)))
    devs = [g for _ in range(per_gpu) for g in range(n_dev)]
    pairs = [(names[2 * k], names[2 * k + 1]) for k in range(a.pairs)]
    fd, dn = os.dup(1), os.open(os.devnull, os.O_WRONLY)
    os.dup2(dn, 1)
    try:
        plade_b200.register_batch(pairs[:len(devs)], devices=devs)          # warm-up: contexts, scratch, page-locked buffers
        t0 = time.perf_counter()
        ok, T = plade_b200.register_batch(pairs, devices=devs)
        dt = time.perf_counter() - t0
    finally:
        os.dup2(fd, 1)
    res = [(bool(ok[k]), np.asarray(T[k], dtype=np.float64)) for k in range(a.pairs)]
    doc["gpu_steady_state"] = {"seconds": dt, "pairs_per_s": a.pairs / dt, **judge(res, gts), "call": "plade_register_batch (what the CLI calls), warm process",
                               "gpus": n_dev, "workers_per_gpu": per_gpu}
    print("steady", json.dumps(doc["gpu_steady_state"]), flush=True)
    ref_bin = os.path.join(ROOT, "oracle", "_ref_fast", "plade_ref")
    if not os.path.exists(ref_bin):
        ref_bin = os.path.join(ROOT, "oracle", "_ref", "plade_ref")
    if not a.skip_reference and os.path.exists(ref_bin):
        # the reference CLI is single-threaded: one process per host core, each with its share of the list
        W = min(os.cpu_count() or 1, a.pairs)
        procs = []
        t0 = time.perf_counter()
        for w in range(W):
            mine = [k for k in range(a.pairs) if k % W == w]
            lf = os.path.join(d, "pairs_ref_%d.txt" % w)
            open(lf, "w").write("\n".join(n for k in mine for n in (names[2 * k], names[2 * k + 1])) + "\n")
            procs.append((mine, os.path.join(d, "results_ref_%d.txt" % w), subprocess.Popen([ref_bin, lf, os.path.join(d, "results_ref_%d.txt" % w)],
                                                                                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)))
        res_ref = [None] * a.pairs
        for mine, of, p in procs:
            p.wait()
            for k, r in zip(mine, parse_results(of, len(mine))):
                res_ref[k] = r
        dt = time.perf_counter() - t0
        res_ref = [r if r is not None else (False, np.eye(4)) for r in res_ref]
        doc["reference"] = {"seconds": dt, "pairs_per_s": a.pairs / dt, "processes": W, "binary": os.path.relpath(ref_bin, ROOT), **judge(res_ref, gts)}
        print("reference", json.dumps(doc["reference"]), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(doc, open(a.out, "w"), indent=1)
    for n in os.listdir(d):
        os.unlink(os.path.join(d, n))
    os.rmdir(d)


if __name__ == "__main__":
    main()
