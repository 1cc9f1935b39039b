"""Dev tool: shared-memory wavefronts / global sectors per source line from an ncu report (cuda,sass view)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
agg = {}
fp = None
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fp = r[1].split("/")[-1]
        continue
    if r[0] in ("Line No", "Address") or (len(r) > 3 and "Source" in r[:3] and hdr is None):
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    d = dict(zip(hdr, r))
    if r[0].isdigit():
        cur = (fp, int(r[0]), r[1].strip()[:90])
    try:
        ws = int(d.get("L1 Wavefronts Shared", "0") or 0)
        wi = int(d.get("L1 Wavefronts Shared Ideal", "0") or 0)
        gs = int(d.get("L2 Theoretical Sectors Global", "0") or 0)
        tg = int(d.get("L1 Tag Requests Global", "0") or 0)
    except ValueError:
        continue
    if r[0].isdigit():
        a = agg.get(cur, [0, 0, 0, 0])
        agg[cur] = [a[0] + ws, a[1] + wi, a[2] + gs, a[3] + tg]
tot = [sum(v[i] for v in agg.values()) for i in range(4)]
print("total shared wavefronts %d (ideal %d), global sectors %d, global tag requests %d" % tuple(tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%10d %5.1f%% ideal %10d | gsect %9d tag %9d  %s:%d  %s" % (v[0], 100.0 * v[0] / max(tot[0], 1), v[1], v[2], v[3], k[0], k[1], k[2]))
print("---- by global tag requests")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][3])[:12]:
    print("%10d %5.1f%% sectors %9d  %s:%d  %s" % (v[3], 100.0 * v[3] / max(tot[3], 1), v[2], k[0], k[1], k[2]))
