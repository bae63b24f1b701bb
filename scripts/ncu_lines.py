"""Top source lines of an ncu report by warp-stall samples: python scripts/ncu_lines.py report.ncu-rep [N] [kernel-regex]
(needs -lineinfo at compile time and --import-source on at capture time)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, lines = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 4 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        d = dict(zip(hdr[4:], r[4:]))
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
        lines.append((int(d["Warp Stall Sampling (All Samples)"]), fname, int(r[0]), r[1].strip()[:100], int(d["Instructions Executed"]), stalls))
tot = sum(l[0] for l in lines)
print("total samples %d" % tot)
for s, f, ln, src, ne, st in sorted(lines, key=lambda l: -l[0])[:N]:
    top = " ".join("%s:%d" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print("%5.1f%% %7d inst %9d  %s:%d  %s   [%s]" % (100.0 * s / tot, s, ne, f, ln, src, top))
