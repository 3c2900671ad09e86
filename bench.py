#!/usr/bin/env python
"""bench.py — registration pairs/s on the synthetic 2M-vs-2M indoor pair (BASELINE.json configs[2], the
configuration the metric is quoted on: "registration pairs/sec (2M-pt clouds)").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one complete registration() of one pair: RANSAC plane extraction on both clouds, spacing,
voxel grids, line/descriptor construction, descriptor matching, hypothesis generation, penetration
filter and hypothesis verification.  `value` keeps both clouds resident in HBM; `e2e` goes through the
reference-facing C-ABI call with host buffers (H2D of both clouds and D2H of the 4x4 inside the timed
region).  N > 1: pairs are independent, every rank registers its own pair (batch mode of the
reference CLI, PLADE/main.cpp:97-159): weak scaling, no data-path collective.  The hypothesis-sharded
verification with its NCCL max-allreduce (BASELINE config 4) is measured as the extra `verify_sharded`.
--impl reference times the reference's own CPU code (oracle/_ref, built from /root/reference by
oracle/Makefile) on the same pair, one single-threaded process per host core (it has no parallelism).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "registration pairs/sec (2M-pt clouds)"
SEED = 20240611


def workload(n_points):
    from plade_b200.synth import make_pair
    tgt, src, gt = make_pair(n_points=n_points, n_planes=20, seed=SEED)
    return tgt, src, gt


def workload_name(n_points, nt, ns):
    return ("synthetic indoor pair (plade_b200.synth.make_pair seed %d): %d-pt target vs %d-pt source, 20 planes, "
            "sigma 0.001, 40%% overlap" % (SEED, nt, ns))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def pinned_copy(a):
    """numpy view of page-locked memory holding a copy of `a` (so the H2D inside the C ABI is a pinned copy)."""
    import torch
    t = torch.empty(a.shape, dtype=torch.float32, pin_memory=True)
    v = t.numpy()
    v[...] = a
    return t, v


def run_ours(args):
    import torch
    import torch.distributed as dist
    import plade_b200
    from plade_b200.synth import transform_error, perturbed_hypotheses

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    tgt, src, gt = workload(args.points)
    diag = float(np.linalg.norm(np.ptp(tgt[:, :3], axis=0)))
    ctx = plade_b200.Context(local)
    quiet = open(os.devnull, "w")
    saved = os.dup(1)

    def hush(on):   # the library prints the reference's progress lines (one per stage) on stdout: drop them while timing
        sys.stdout.flush()
        os.dup2(quiet.fileno() if on else saved, 1)

    # ---- device-resident throughput ("value") ---------------------------------------------------------
    ht, hs = ctx.upload(tgt), ctx.upload(src)
    hush(True)
    for _ in range(args.warmup):
        ok, T = ctx.register_resident(ht, hs)
    hush(False)
    stage_acc, k5_ms, k5_shape = {}, [], None
    barrier()
    l0 = ctx.launch_count()
    with ClockSampler(local) as clocks:
        hush(True)
        ctx.timer_start()
        for _ in range(args.steps):
            ok, T = ctx.register_resident(ht, hs)
            st = ctx.stage_times()
            for k, v in st.items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v
            k5_ms.append(st["verify_kernel_ms"])
            k5_shape = (int(st["verify_h"]), int(st["verify_ns"]), int(st["verify_nt"]))
        ms = ctx.timer_stop_ms()
        hush(False)
    barrier()
    launches = ctx.launch_count() - l0
    t_step = max_over_ranks(ms / 1e3 / args.steps)
    value = world / t_step
    rot, tr = transform_error(T, gt, diag)
    if args.profile:     # short run under ncu: never a bench value
        if rank == 0:
            emit({"profile_run": True, "ms_per_step": t_step * 1e3, "gpu_launches": int(launches), "ok": bool(ok),
                  "rot_err_deg": rot, "stage_ms": {k: 1e3 * v / args.steps for k, v in stage_acc.items()}})
        return

    # ---- end to end through the C ABI with host buffers ("e2e") -------------------------------------------
    keep_t, ptgt = pinned_copy(tgt)
    keep_s, psrc = pinned_copy(src)
    hush(True)
    for _ in range(min(args.warmup, 2)):
        ctx.register_clouds(ptgt, psrc)
    hush(False)
    barrier()
    hush(True)
    ctx.timer_start()
    for _ in range(args.steps):
        ok_e, T_e = ctx.register_clouds(ptgt, psrc)
    ms_e = ctx.timer_stop_ms()
    hush(False)
    barrier()
    t_e2e = max_over_ranks(ms_e / 1e3 / args.steps)
    rot_e, tr_e = transform_error(T_e, gt, diag)

    # ---- hypothesis-sharded verification + NCCL max-allreduce (BASELINE config 4 shape) -----------------------
    spacing = ctx.average_spacing(src)
    leaf = 4 * spacing
    ds_t, ds_s = ctx.voxel_downsample(tgt[:, :3], leaf), ctx.voxel_downsample(src[:, :3], leaf)
    H = args.hypotheses
    Rh, Th, true_idx = perturbed_hypotheses(gt, H, seed=7)
    rc, c_src, whd, _ = ctx.bounding_box(ds_s)
    cen = (np.einsum("hij,j->hi", Rh, c_src) + Th).astype(np.float32)
    ball = float(max(whd) / 2)
    mine = np.arange(rank, H, world)
    ctx.verify_upload(ds_s, ds_t, leaf)
    key_t = torch.zeros(1, dtype=torch.int64, device="cuda")

    def sharded_once():
        counts, kms = ctx.verify_resident(Rh[mine], Th[mine], cen[mine], ball, leaf)
        b = int(np.argmax(counts)) if len(counts) else 0
        # packed key {count, ~index}: max picks the best count, ties -> lowest hypothesis index
        key = (int(counts[b]) << 32) | (0xFFFFFFFF - int(mine[b])) if len(counts) else 0
        key_t.fill_(key)
        if world > 1:
            dist.all_reduce(key_t, op=dist.ReduceOp.MAX)
        k = int(key_t.item())
        return 0xFFFFFFFF - (k & 0xFFFFFFFF), k >> 32, kms

    for _ in range(2):
        sharded_once()
    barrier()
    t0 = time.perf_counter()
    vk = []
    for _ in range(3):
        best_idx, best_cnt, kms = sharded_once()
        vk.append(kms)
    torch.cuda.synchronize()
    t_sh = max_over_ranks((time.perf_counter() - t0) / 3)
    k_sh = max_over_ranks(float(np.mean(vk)))

    peak, peak_src = measured_peak()
    Hk, nsk, ntk = k5_shape
    k5_mean_ms = float(np.mean(k5_ms)) if k5_ms else 0.0
    alg_bytes = Hk * 16.0 * nsk + 16.0 * ntk
    achieved = alg_bytes / (k5_mean_ms / 1e3) / 1e9 if k5_mean_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k5_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    out = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.points, len(tgt), len(src)), "pairs_per_step_per_gpu": 1,
                   "parallelism": "pairs sharded over GPUs, no data-path collective" if world > 1 else "single GPU",
                   "l2": "per-step working set (2 clouds x 2 float4 streams = %d MB + sort scratch) exceeds the 126 MB L2" % ((len(tgt) + len(src)) * 32 // 2**20)},
        "e2e": {"value": world / t_e2e, "unit": "pairs/s", "h2d_bytes_per_step": int((len(tgt) + len(src)) * 24), "d2h_bytes_per_step": 64,
                "ms_per_step": t_e2e * 1e3, "rot_err_deg": rot_e, "trans_err_rel": tr_e},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "result": {"ok": bool(ok), "rot_err_deg": rot, "trans_err_rel_diag": tr},
        "stage_ms": {k: 1e3 * v / args.steps for k, v in stage_acc.items() if not k.startswith("verify_")},
        "roofline": {"kernel": "verify_kernel (K5 hypothesis verification)", "bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": k5_mean_ms,
                     "shape": {"hypotheses": Hk, "src_ds_points": nsk, "tgt_ds_points": ntk}},
        "verify_sharded": {"hypotheses": H, "src_ds_points": int(len(ds_s)), "tgt_ds_points": int(len(ds_t)), "ms": t_sh * 1e3,
                           "kernel_ms_max_rank": k_sh, "hyps_per_s": H / t_sh, "scaling": "strong", "collective": "ncclAllReduce(max, 1 x i64)" if world > 1 else "none",
                           "best_index": int(best_idx), "best_count": int(best_cnt), "best_is_true_transform": bool(best_idx == true_idx),
                           "achieved_gbs": (H * 16.0 * len(ds_s) / world + 16.0 * len(ds_t)) / (k_sh / 1e3) / 1e9 if k_sh > 0 else None},
    }
    # ---- CPU baseline: the reference's own code on the same pair, rank 0 at N = 1 only --------------------------
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_once(tgt, src, gt, diag)
    ctx.free_cloud(ht)
    ctx.free_cloud(hs)
    ctx.close()
    if rank == 0:
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_once(tgt, src, gt, diag):
    from oracle import ref as oref
    from plade_b200.synth import transform_error
    if not oref.have_ref():
        return {"value": None, "unit": "pairs/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/libplade_ref.so not present on this box"}
    r = oref.Ref(quiet=True)
    r.set_seed(1)
    t0 = time.perf_counter()
    ok, T = r.registration_clouds(tgt, src)
    dt = time.perf_counter() - t0
    rot, tr = transform_error(T, gt, diag)
    return {"value": 1.0 / dt, "unit": "pairs/s", "cores": 1, "kind": "reference",
            "sample": "1 full registration(T, target, source) of the same pair by the reference's own sources (oracle/_ref), "
                      "single thread (the reference has no parallelism); %.1f s" % dt,
            "seconds_per_pair": dt, "ok": bool(ok), "rot_err_deg": rot, "trans_err_rel_diag": tr,
            "host": {"nproc": os.cpu_count()}}


def _ref_worker(args):
    path, seed = args
    from oracle import ref as oref
    d = np.load(path)
    r = oref.Ref(quiet=True)
    r.set_seed(seed)
    t0 = time.perf_counter()
    ok, T = r.registration_clouds(d["tgt"], d["src"])
    return bool(ok), T, time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import ref as oref
    from plade_b200.synth import transform_error
    import multiprocessing as mp
    import tempfile
    if not oref.have_ref():
        emit({"impl": "reference", "unavailable": "oracle/_ref/libplade_ref.so missing (built only where /root/reference is mounted)"})
        return
    tgt, src, gt = workload(args.points)
    diag = float(np.linalg.norm(np.ptp(tgt[:, :3], axis=0)))
    workers = max(1, min(os.cpu_count() or 1, args.ref_workers))
    tmp = tempfile.NamedTemporaryFile(suffix=".npz", delete=False)
    np.savez(tmp.name, tgt=tgt, src=src)
    pool = mp.get_context("spawn").Pool(workers)
    try:
        times = []
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_ref_worker, [(tmp.name, 1 + step * workers + w) for w in range(workers)])
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    finally:
        pool.terminate()
        os.unlink(tmp.name)
    t_step = float(np.mean(times))
    value = workers / t_step
    oks = [r[0] for r in res]
    errs = [transform_error(r[1], gt, diag) for r in res]
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.points, len(tgt), len(src)), "pairs_per_step": workers},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": workers, "kind": "reference",
                         "sample": "each step = %d concurrent single-threaded registration() calls of the reference's own sources "
                                   "(oracle/_ref) on the same pair, one per host core used" % workers,
                         "seconds_per_pair_single_thread": float(np.mean([r[2] for r in res])), "host": {"nproc": os.cpu_count()}},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "result": {"ok": bool(all(oks)), "rot_err_deg": float(np.median([e[0] for e in errs])), "trans_err_rel_diag": float(np.median([e[1] for e in errs]))},
    }
    emit(out)


_REAL_STDOUT = None


def emit(obj):
    """The one JSON line goes to the process's original stdout; everything else (library progress lines,
    NCCL's version banner, ...) was routed to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=2_000_000)
    ap.add_argument("--hypotheses", type=int, default=10000)
    ap.add_argument("--ref-workers", type=int, default=8)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="short run for ncu: 1 warm-up, no e2e / sharded / cpu arms")
    args = ap.parse_args()
    if args.impl == "ours" and not args.profile:
        args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
