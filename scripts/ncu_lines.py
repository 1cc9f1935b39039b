"""Dev tool: per-source-line instruction / stall-sample shares from an ncu report (cuda,sass view)."""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] +
                     (["--kernel-name", "regex:" + pat] if pat else []), capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fn = fp = None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fp = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        fn = r[1]
        continue
    if r[0] == "Line No":
        continue
    if r[0] != "" and r[0].isdigit() and len(r) > 8:
        try:
            s, i = int(r[6]), int(r[7])
        except ValueError:
            continue
        key = (fn, fp, int(r[0]), r[1].strip()[:100])
        a = agg.get(key, (0, 0))
        agg[key] = (a[0] + s, a[1] + i)
for f in sorted(set(k[0] for k in agg)):
    items = [(k, v) for k, v in agg.items() if k[0] == f]
    tot = sum(v[0] for k, v in items) or 1
    toti = sum(v[1] for k, v in items) or 1
    print("=====", f, "samples", tot, "inst", toti)
    for k, v in sorted(items, key=lambda kv: -kv[1][0])[:top]:
        print("%6d %5.1f%% inst %5.1f%%  %s:%d  %s" % (v[0], 100 * v[0] / tot, 100 * v[1] / toti, k[1], k[2], k[3]))
