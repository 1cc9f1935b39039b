"""Dev tool: table of kernel launches from an ncu --csv --log-file (long format)."""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
i = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[i]
tab = OrderedDict()
for r in rows[i + 1:]:
    d = dict(zip(hdr, r))
    tab.setdefault(d["ID"], {"name": d["Kernel Name"][:64], "grid": d["Grid Size"], "block": d["Block Size"]})[d["Metric Name"]] = d["Metric Value"]
for k, v in tab.items():
    print(k, v["name"], v["grid"], v["block"], " ".join("%s=%s" % (m.split(".")[0].split("__")[-1], x) for m, x in v.items() if m not in ("name", "grid", "block")))
