"""Summarise an ncu --csv launch list (gpurun_out/r2_launches_c5_tc.csv by default): one line per MPD kernel launch."""
import csv
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2_launches_c5_tc.csv"
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
h = rows[0]
ki, vi, mi, ii = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("ID")
seen = {}
for r in rows[1:]:
    n = r[ki]
    if "mpd" in n or "dmol" in n:
        seen.setdefault((int(r[ii]), n.split("(")[0][-28:]), {})[r[mi].split(".")[0][-24:]] = r[vi]
for k, v in sorted(seen.items())[-6:]:
    print(k[1], " ".join("%s=%s" % (a, b) for a, b in v.items()))
