"""Turn the ncu reports of one evidence run (tools/gpu_evidence.sh TAG -> gpurun_out/TAG_*) into the tracked
summaries under profiles/: launch list, one markdown table + hottest SASS lines per --set full capture, the bench
lines, and profiles/k5_traffic.json (DRAM bytes of one K5 launch, read by bench.py for roofline.traffic)."""
import csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def run(*a):
    return subprocess.run(list(a), capture_output=True, text=True).stdout


def raw_metrics(rep):
    rows = list(csv.reader(run("ncu", "-i", rep, "--page", "raw", "--csv").splitlines()))
    return {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}


def num(m, k):
    return float(m[k][0].replace(",", ""))


def to_bytes(m, k):
    v, u = num(m, k), m[k][1].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


open(os.path.join(P, tag + "_launches_one_registration.txt"), "w").write(
    "# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --profile --steps 1 --warmup 1 --pairs-per-gpu 1\n"
    "# (two registrations: the warm-up and the step)\n" + run(sys.executable, os.path.join(ROOT, "tools", "summarize_launches.py"), os.path.join(G, tag + "_launches_raw.csv")))
for f in ("bench.json", "bench_reference.json", "lscpu.txt", "smi.txt"):
    if os.path.exists(os.path.join(G, tag + "_" + f)):
        shutil.copy(os.path.join(G, tag + "_" + f), os.path.join(P, tag + "_" + f))

caps = {"verify_kernel_h10000": "ncu --set full --clock-control none --import-source on -k regex:verify_kernel -s 1 -c 1 python tools/run_verify.py 2000000 10000 2",
        "verify_kernel_config4": "ncu --set full --clock-control none --import-source on -k regex:verify_kernel -s 1 -c 1 python tools/run_verify.py 5000000 10000 2 24",
        "refine_cluster_kernel": "ncu --set full --clock-control none --import-source on -k regex:refine_cluster_kernel -s 70 -c 1 python bench.py --profile --steps 1 --warmup 1 --pairs-per-gpu 1",
        "accept_loop_kernel": "ncu --set full --clock-control none --import-source on -k regex:accept_loop_kernel -s 5 -c 1 python bench.py --profile --steps 1 --warmup 1 --pairs-per-gpu 1",
        "score_candidates_kernel": "ncu --set full --clock-control none --import-source on -k regex:score_candidates_kernel -s 6 -c 1 python bench.py --profile --steps 1 --warmup 1 --pairs-per-gpu 1",
        "score_points_kernel": "ncu --set full --clock-control none --import-source on -k regex:score_points_kernel -s 6 -c 1 python bench.py --profile --steps 1 --warmup 1 --pairs-per-gpu 1",
        "match_kernel": "ncu --set full --clock-control none --import-source on -k regex:match_kernel -s 1 -c 1 python tools/run_match.py 118000 118000"}
traffic = {}
for name, cmd in caps.items():
    rep = os.path.join(G, "%s_%s.ncu-rep" % (tag, name))
    if not os.path.exists(rep):
        continue
    body = "# %s — %s (ncu --set full)\n\ncommand: `%s`\n\n" % (tag, name, cmd)
    body += run(sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep)
    body += "\nHottest SASS lines by stall samples:\n\n```\n" + run(sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), rep, "16") + "```\n"
    m = raw_metrics(rep)
    try:
        dram = to_bytes(m, "dram__bytes_read.sum") + to_bytes(m, "dram__bytes_write.sum")
        traffic[name] = dram
        body += "\nDRAM traffic of this launch: %.2f MB (dram__bytes_read.sum + dram__bytes_write.sum)\n" % (dram / 1e6)
    except Exception:
        dram = None
    if name == "verify_kernel_h10000" and dram is not None:
        shape = {"hypotheses": 10000, "src_ds_points": 193001, "tgt_ds_points": 481539}
        alg = shape["hypotheses"] * 16.0 * shape["src_ds_points"] + 16.0 * shape["tgt_ds_points"]
        json.dump({"dram_bytes_per_launch": dram, "shape": shape, "source": "profiles/%s_%s.md" % (tag, name)}, open(os.path.join(P, "k5_traffic.json"), "w"))
        body += ("\nalgorithmic bytes per launch = H*16*N_s + 16*N_t = %.0f B; DRAM traffic per launch = %.1f MB (%.3f %% of the algorithmic bytes): "
                 "the ds clouds and the grid are L2-resident, the kernel is bound by instruction issue (see issue_active / lanes per instruction above).\n"
                 % (alg, dram / 1e6, 100 * dram / alg))
    open(os.path.join(P, "%s_%s.md" % (tag, name)), "w").write(body)
if "accept_loop_kernel" in traffic:      # read by bench.py for roofline.traffic of the step's dominant kernel
    json.dump({"kernel": "accept_loop_kernel", "dram_bytes_per_launch": traffic["accept_loop_kernel"], "source": "profiles/%s_accept_loop_kernel.md" % tag},
              open(os.path.join(P, "dominant_kernel_traffic.json"), "w"))
print("profiles written for", tag)
