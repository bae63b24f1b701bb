"""Aggregate an ncu source page (CSV) by CUDA source line: instructions executed and stall samples.
usage: ncu -i rep.ncu-rep --page source --csv --print-source sass,cuda | python scripts/ncu_lines.py [topN]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
hdr = None
agg = {}
tot_inst = tot_samp = 0
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip():
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    src = r[1].strip()
    try:
        samp = int(r[hdr.index("# Samples")] or 0)
        inst = int(r[hdr.index("Instructions Executed")] or 0)
    except ValueError:
        continue
    key = (line, src)
    a = agg.setdefault(key, [0, 0])
    a[0] += inst; a[1] += samp
    tot_inst += inst; tot_samp += samp
print("total inst %.3e  samples %d" % (tot_inst, tot_samp))
for (line, src), (inst, samp) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5d  inst %5.1f%%  samples %5.1f%%  %s" % (line, 100.0 * inst / max(tot_inst, 1), 100.0 * samp / max(tot_samp, 1), src[:110]))
