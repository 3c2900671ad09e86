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
REF_BUILD = "g++ -O3 -march=x86-64-v3 (oracle/_ref_fast: the timing build of the reference's own sources)"
REF_BUILD_PARITY = "g++ -O2 -ffp-contract=off (oracle/_ref: the parity build; the timing build is missing)"
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

    def __init__(self, index, enabled=True):
        self.index, self.rows, self.proc, self.enabled = index, [], None, enabled

    def __enter__(self):
        if not self.enabled:      # one sampler per box is enough: NVML queries from 8 ranks at once stall each other's launches
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "500"],
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


def shared_config(n_points, nt, ns):
    """the `config` object: identical in both arms (the driver compares them)"""
    return {"workload": workload_name(n_points, nt, ns), "points_per_cloud": int(n_points), "planes": 20, "scene_seed": SEED}


PAIRS_PER_GPU = 4      # pairs registered concurrently per GPU and step: the same at every N


def run_ours(args):
    import torch
    import torch.distributed as dist
    import plade_b200
    from plade_b200.synth import make_pair, transform_error, perturbed_hypotheses

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.pin_cores and world > 1 and hasattr(os, "sched_setaffinity"):
        # every rank keeps to its own share of the host cores (no migration between the ranks' threads)
        allc = sorted(os.sched_getaffinity(0))
        per = max(1, len(allc) // world)
        os.sched_setaffinity(0, set(allc[local * per:(local + 1) * per]))

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

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    tgt, src, gt = workload(args.points)
    diag = float(np.linalg.norm(np.ptp(tgt[:, :3], axis=0)))
    # Pairs are independent, and one registration is a chain of dependent launches (the RANSAC rounds), so ONE pair cannot
    # fill 148 SMs: a step registers B pairs per GPU concurrently, one context (own streams + scratch) and one host thread
    # per pair -- the batch mode of the reference CLI (PLADE/main.cpp:97-159, plade_register_batch).  B is the same at
    # every N; when the box has fewer cores than waiting host threads the waits block instead of spinning.
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)      # cores this process may use
    B = args.pairs_per_gpu if args.pairs_per_gpu > 0 else PAIRS_PER_GPU
    host_threads = 3 * B * world          # per pair: the calling thread + the two plane-extraction lanes
    blocking = (2 if host_threads > cores else 0) if args.blocking_sync < 0 else int(args.blocking_sync)      # 0 spin, 1 blocking event, 2 poll + yield
    ctxs = [plade_b200.Context(local) for _ in range(B)]
    for kv in args.param:
        for c in ctxs:
            c.set_param(kv.split("=")[0], float(kv.split("=")[1]))
    ctx = ctxs[0]
    if blocking:
        ctx.set_param("blocking_sync", blocking)          # process-wide
    quiet = open(os.devnull, "w")
    saved = os.dup(1)

    def hush(on):   # the library prints the reference's progress lines (one per stage) on stdout: drop them while timing
        sys.stdout.flush()
        os.dup2(quiet.fileno() if on else saved, 1)

    step_no = [0]

    def run_concurrent(fn, reps, timed):
        """every context runs fn(k) `reps` times on its own host thread, each registration with its own RANSAC seed;
        returns (max device ms over contexts, wall ms, every result)"""
        dev_ms, res = [0.0] * B, [[] for _ in range(B)]
        base = step_no[0]
        step_no[0] += reps

        def work(k):
            if timed:
                ctxs[k].timer_start()
            for r in range(reps):
                ctxs[k].set_param("seed", SEED + 1000 * rank + 100 * k + base + r)
                res[k].append(fn(k))
            if timed:
                dev_ms[k] = ctxs[k].timer_stop_ms()
        th = [threading.Thread(target=work, args=(k,)) for k in range(B)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return max(dev_ms), (time.perf_counter() - t0) * 1e3, [r for rr in res for r in rr]

    def judge(results):
        errs = [transform_error(T, gt, diag) for ok, T in results if ok]
        landed = sum(1 for (r, t) in errs if r <= 0.5 and t <= 5e-3)
        return {"registrations": len(results), "ok": sum(1 for ok, _ in results if ok), "within_0.5deg_5e-3": landed,
                "max_rot_err_deg": max([e[0] for e in errs], default=None), "max_trans_err_rel_diag": max([e[1] for e in errs], default=None)}

    # ---- device-resident throughput ("value") ---------------------------------------------------------
    resident = [(c.upload(tgt), c.upload(src)) for c in ctxs]
    hush(True)
    run_concurrent(lambda k: ctxs[k].register_resident(*resident[k]), args.warmup, False)
    hush(False)
    barrier()
    l0 = sum(c.launch_count() for c in ctxs)
    ctx.set_param("kernel_clock", 1)      # context 0 also times its own kernels with CUDA events on their streams (the live roofline figures)
    with ClockSampler(local, enabled=(rank == 0)) as clocks:
        hush(True)
        ms, wall_ms, results = run_concurrent(lambda k: ctxs[k].register_resident(*resident[k]), args.steps, True)
        hush(False)
    barrier()
    ctx.set_param("kernel_clock", 0)
    launches = sum(c.launch_count() for c in ctxs) - l0
    stage = ctx.stage_times()        # context 0, last step
    k1 = ctx.kernel_times("score_candidates")
    k1r, k1b = ctx.kernel_times("refine_cluster"), ctx.kernel_times("band_compact")
    k5_in = ctx.kernel_times("verify")
    t_step = max_over_ranks(ms / 1e3 / args.steps)
    value = world * B / t_step
    verdict = judge(results)
    if args.profile:     # short run under ncu: never a bench value
        if rank == 0:
            emit({"profile_run": True, "ms_per_step": t_step * 1e3, "pairs_per_step": B, "gpu_launches": int(launches), "result": verdict,
                  "stage_ms": {k: 1e3 * v for k, v in stage.items() if not k.startswith("verify_")}})
        return

    # ---- end to end through the C ABI with host buffers ("e2e") -------------------------------------------
    keep_t, ptgt = pinned_copy(tgt)
    keep_s, psrc = pinned_copy(src)
    hush(True)
    run_concurrent(lambda k: ctxs[k].register_clouds(ptgt, psrc), min(args.warmup, 2), False)
    hush(False)
    barrier()
    hush(True)
    ms_e, wall_e, results_e = run_concurrent(lambda k: ctxs[k].register_clouds(ptgt, psrc), args.steps, True)
    hush(False)
    barrier()
    t_e2e = max_over_ranks(ms_e / 1e3 / args.steps)
    verdict_e = judge(results_e)
    for k in range(1, B):        # the remaining legs use context 0 only
        ctxs[k].free_cloud(resident[k][0]); ctxs[k].free_cloud(resident[k][1])
        ctxs[k].close()
    ht, hs = resident[0]
    ctx.free_cloud(ht)
    ctx.free_cloud(hs)
    ctx.set_param("seed", SEED)

    # ---- BASELINE config 4: 10 K hypotheses on the 5 M-point scene, sharded over the ranks; the collective is the library's own
    # ncclAllReduce(ncclUint64, ncclMax) (plade_shard_init_nccl + plade_verify_sharded): torch only carries the unique id ------------
    c4 = None
    if not args.skip_config4:
        t4, s4, gt4 = make_pair(n_points=args.config4_points, n_planes=24, seed=SEED)
        spacing = ctx.average_spacing(s4)
        leaf = 4 * spacing
        ds_t, ds_s = ctx.voxel_downsample(t4[:, :3], leaf), ctx.voxel_downsample(s4[:, :3], leaf)
        H = args.hypotheses
        Rh, Th, true_idx = perturbed_hypotheses(gt4, H, seed=7)
        rc, c_src, whd, _ = ctx.bounding_box(ds_s)
        cen = (np.einsum("hij,j->hi", Rh, c_src) + Th).astype(np.float32)
        ball = float(max(whd) / 2)
        ctx.verify_upload(ds_s, ds_t, leaf)
        if world > 1:
            uid = [plade_b200.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            ctx.shard_init_nccl(uid[0], rank, world)
        for _ in range(2):
            ctx.verify_sharded(Rh, Th, cen, ball, leaf)
        barrier()
        dms = []
        t0 = time.perf_counter()
        for _ in range(3):
            best_idx, best_cnt, d_ms = ctx.verify_sharded(Rh, Th, cen, ball, leaf)
            dms.append(d_ms)
        t_sh = max_over_ranks((time.perf_counter() - t0) / 3)
        k_sh = max_over_ranks(float(np.mean(dms)))
        if world > 1:
            ctx.shard_finalize()
        nsk, ntk = int(len(ds_s)), int(len(ds_t))
        c4 = {"scene": "synthetic %d-pt scene, 24 planes (make_pair seed %d): %d-pt target vs %d-pt source" % (args.config4_points, SEED, len(t4), len(s4)),
              "hypotheses": H, "src_ds_points": nsk, "tgt_ds_points": ntk, "ms": t_sh * 1e3, "device_ms_max_rank": k_sh, "hyps_per_s": H / t_sh,
              "scaling": "strong", "collective": "ncclAllReduce(1 x ncclUint64, ncclMax) inside libplade_b200 (plade_verify_sharded)" if world > 1 else "none (1 rank)",
              "best_index": int(best_idx), "best_count": int(best_cnt), "best_is_true_transform": bool(best_idx == true_idx),
              "algorithmic_bytes_per_rank": (H // world) * 16.0 * nsk + 16.0 * ntk,
              "achieved_gbs_per_gpu": ((H / world) * 16.0 * nsk + 16.0 * ntk) / (k_sh / 1e3) / 1e9 if k_sh > 0 else None}

    peak, peak_src = measured_peak()

    def kernel_line(name, kt, share_of_ms):
        per = kt["ms"] / kt["launches"] if kt["launches"] else 0.0
        gbs = kt["algorithmic_bytes"] / (kt["ms"] / 1e3) / 1e9 if kt["ms"] > 0 else 0.0
        return {"kernel": name, "launches": kt["launches"], "ms": kt["ms"], "avg_launch_ms": per,
                "algorithmic_bytes": kt["algorithmic_bytes"], "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak if peak else None,
                "share_of_registration": kt["ms"] / share_of_ms if share_of_ms > 0 else None}
    # kernels of context 0 in the LAST timed step (CUDA events around every launch, on the launching stream, while the other
    # contexts share the GPU); the two plane-extraction lanes of a registration run concurrently, so shares can add up to > 1
    reg_ms_total = ms / args.steps
    lines = [kernel_line("accept_loop_kernel (K1b-d: per pool entry a band pass of the 16-CTA cluster over the cloud + the acceptance chain "
                         "(flags, raster, components, select, covariance) + accept; one launch per scoring round)", k1r, reg_ms_total),
             kernel_line("score_candidates_kernel x 2 (K1a: the live slots of 16384 candidate draws x 4096 points, then the best 256 x 65536, per round)", k1, reg_ms_total),
             kernel_line("band_compact_kernel (host-driven fallback path only)", k1b, reg_ms_total),
             kernel_line("verify_kernel (K5) inside the registration (H = %d surviving hypotheses)" % int(stage["verify_h"]), k5_in, reg_ms_total)]
    dom = max(lines[:2], key=lambda l: l["ms"])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("kernel", "") in dom["kernel"]:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"kernel": dom["kernel"], "bound": "hbm",
                "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dom["achieved_gbs"] / peak if peak else None, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": dom["algorithmic_bytes"] / dom["launches"] if dom["launches"] else None,
                "launch_ms": dom["avg_launch_ms"], "share_of_step": dom["share_of_registration"],
                "note": "the dominant kernel of the timed step by device time (context 0, last timed step, live CUDA events); algorithmic bytes per "
                        "SURVEY.md 8(d): 20 B per cloud point for every band pass + 28 B per band point per evaluation (accept loop), 28 B per "
                        "subsample point per pass (K1a).  Both are latency / issue bound on a working set that lives in L2 -- see DESIGN.md section 4"}
    out = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": shared_config(args.points, len(tgt), len(src)),
        "run": {"pairs_per_step_per_gpu": B,
                "concurrency": "%d contexts (one host thread + two plane-extraction lanes each) per GPU register %d pairs concurrently per step; "
                               "every registration has its own RANSAC seed" % (B, B),
                "parallelism": "pairs sharded over GPUs, no data-path collective" if world > 1 else "single GPU",
                "host_cores": cores, "host_threads": host_threads, "wait_mode": ["spin", "blocking event", "poll + yield"][blocking], "blocking_sync": bool(blocking),
                "l2": "per-step working set (%d pairs x 2 clouds x 2 float4 streams = %d MB + sort scratch) exceeds the 126 MB L2" % (B, B * (len(tgt) + len(src)) * 32 // 2**20)},
        "e2e": {"value": world * B / t_e2e, "unit": "pairs/s", "h2d_bytes_per_step": int(B * (len(tgt) + len(src)) * 24), "d2h_bytes_per_step": 64 * B,
                "ms_per_step": t_e2e * 1e3, "wall_ms_per_step": wall_e / args.steps, "result": verdict_e},
        "wall_ms_per_step": wall_ms / args.steps,
        "latency_ms_per_pair": ms / args.steps,
        "gpu_launches": int(sum_over_ranks(float(launches))),
        "clocks": clocks.summary(),
        "result": verdict,
        "stage_ms": {k: 1e3 * v for k, v in stage.items() if not k.startswith("verify_")},
        "roofline": roofline,
        "step_kernels": lines,
    }
    if c4 is not None:
        out["config4_verify_sharded"] = c4
        out["roofline_k5_config4"] = {"kernel": "verify_kernel (K5) at BASELINE config 4: %d hypotheses per rank on the 5 M-point scene" % (c4["hypotheses"] // world),
                                      "bound": "hbm", "achieved": c4["achieved_gbs_per_gpu"], "peak": peak, "unit": "GB/s",
                                      "frac": c4["achieved_gbs_per_gpu"] / peak if (peak and c4["achieved_gbs_per_gpu"]) else None,
                                      "note": "kernel + device-side key + collective, CUDA events on the context stream; L2-resident working set: issue-bound, "
                                              "the HBM figure is the SURVEY.md 8(d) algorithmic-bytes convention (16 B per hypothesis and source point)"}
    # ---- CPU baseline: the reference's own code on the same pair, rank 0 at N = 1 only --------------------------
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_once(tgt, src, gt, diag)
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
    r = oref.Ref(quiet=True, fast=True)
    r.set_seed(1)
    t0 = time.perf_counter()
    ok, T = r.registration_clouds(tgt, src)
    dt = time.perf_counter() - t0
    rot, tr = transform_error(T, gt, diag)
    return {"value": 1.0 / dt, "unit": "pairs/s", "cores": 1, "kind": "reference", "build": REF_BUILD if r.fast else REF_BUILD_PARITY,
            "sample": "1 full registration(T, target, source) of the same pair by the reference's own sources (oracle/_ref_fast), "
                      "single thread (the reference has no parallelism); %.1f s" % dt,
            "seconds_per_pair": dt, "ok": bool(ok), "rot_err_deg": rot, "trans_err_rel_diag": tr,
            "host": {"nproc": os.cpu_count()}}


def _ref_worker(args):
    path, seed = args
    from oracle import ref as oref
    d = np.load(path)
    r = oref.Ref(quiet=True, fast=True)
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
    workers = max(1, min(os.cpu_count() or 1, args.ref_workers if args.ref_workers > 0 else 64))
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
        "config": shared_config(args.points, len(tgt), len(src)),
        "run": {"pairs_per_step": workers, "host_cores": os.cpu_count()},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": workers, "kind": "reference", "build": REF_BUILD if oref.have_ref_fast() else REF_BUILD_PARITY,
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
    ap.add_argument("--ref-workers", type=int, default=0, help="concurrent single-threaded reference processes (0 = every host core, at most 64)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--pairs-per-gpu", type=int, default=0, help="pairs registered concurrently per GPU and step (0 = 4, the same at every N)")
    ap.add_argument("--skip-config4", action="store_true", help="skip the sharded verification of BASELINE config 4")
    ap.add_argument("--pin-cores", type=int, default=0, help="1: with N > 1 ranks, every rank keeps to its own 1/N of the host cores")
    ap.add_argument("--blocking-sync", type=int, default=-1, help="-1: poll + yield instead of spinning when host threads > cores; 0 spin / 1 blocking event / 2 poll + yield: force")
    ap.add_argument("--config4-points", type=int, default=5_000_000)
    ap.add_argument("--profile", action="store_true", help="short run for ncu: 1 warm-up, no e2e / sharded / cpu arms")
    ap.add_argument("--param", action="append", default=[], help="name=value passed to plade_set_param on every context (diagnostic runs)")
    args = ap.parse_args()
    if args.impl == "ours" and not args.profile:
        args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
