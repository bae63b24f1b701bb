"""DRAM traffic per sample of a profiled rollout kernel -> profiles/r02_dram_traffic.json (read by bench.py's roofline).
usage: python scripts/ncu_traffic.py key=report.ncu-rep:samples [...]     e.g. swarm50:tensor=gpurun_out/prof.ncu-rep:9472"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
tab = json.load(open(path)) if os.path.exists(path) else {}
for arg in sys.argv[1:]:
    key, rest = arg.split("=")
    rep, n = rest.rsplit(":", 1)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = float(m["dram__bytes_read.sum"][1]) * scale[m["dram__bytes_read.sum"][0]]
    wr = float(m["dram__bytes_write.sum"][1]) * scale[m["dram__bytes_write.sum"][0]]
    tab[key] = {"bytes_per_sample": (rd + wr) / int(n), "dram_read_bytes": rd, "dram_write_bytes": wr, "samples": int(n),
                "kernel": m["Kernel Name"][1] if "Kernel Name" in m else "", "source": os.path.basename(rep)}
    print(key, tab[key])
json.dump(tab, open(path, "w"), indent=1, sort_keys=True)
