"""Hottest SASS lines (by stall samples) of the kernel in an ncu report captured with --import-source on."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
h = rows[hdr]
si, ie, te, ws = h.index('Source'), h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('Warp Stall Sampling (All Samples)')
data = []
for r in rows[hdr + 1:]:
    try: data.append((int(r[ie]), int(r[te]), int(r[ws]), r[si]))
    except Exception: pass
tot = sum(d[0] for d in data); tots = sum(d[2] for d in data)
print('total warp instructions', tot, 'stall samples', tots, 'avg lanes %.1f' % (sum(d[1] for d in data) / tot))
for d in sorted(data, key=lambda d: -d[2])[:top]:
    print('%5.1f%% inst %5.1f%% stall lanes %4.1f | %s' % (100 * d[0] / tot, 100 * d[2] / max(tots, 1), d[1] / max(d[0], 1), d[3][:100]))
