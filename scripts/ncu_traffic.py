"""dram__bytes_read.sum + dram__bytes_write.sum per launch from .ncu-rep files of prof_class.py runs:
  python scripts/ncu_traffic.py out.json CLASS=file.ncu-rep[:launch_quartets] ...   (merges into out.json)"""
import csv, json, os, subprocess, sys
out = sys.argv[1]
res = json.load(open(out)) if os.path.exists(out) else {}
unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
for a in sys.argv[2:]:
    cl, f = a.split("=")
    nq = 1 << 20
    if ":" in f:
        f, nq = f.split(":")
    txt = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(name)
        tot += float(r[i].replace(",", "")) * unit[units[i]]
    res[cl] = {"dram_bytes": tot, "launch_quartets": int(nq), "kernel": r[hdr.index("Kernel Name")],
               "gpu_time_us": r[hdr.index("gpu__time_duration.sum")], "source": os.path.basename(f)}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
