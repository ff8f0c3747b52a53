"""Top SASS instructions by stall samples: python scripts/ncu_hot.py file.ncu-rep [N]"""
import csv, subprocess, sys
f = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", f, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = rows[2:]
iS = hdr.index("# Samples"); iSrc = hdr.index("Source"); iEx = hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    op = r[iSrc].split()[0] if not r[iSrc].strip().startswith("@") else r[iSrc].split()[1]
    a = agg.setdefault(op, [0, 0]); a[0] += int(r[iS]); a[1] += int(r[iEx])
print("by opcode (samples, executed):")
for op, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    print("  %-14s %7d %5.1f%%  exec %d" % (op, s, 100.0 * s / tot, e))
print("top instructions:")
for idx, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][iS]))[:N]:
    st = {hdr[i][6:]: int(r[i]) for i in stalls if int(r[i]) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("  #%4d %6s %5.1f%% %-60s %s" % (idx, r[iS], 100.0 * int(r[iS]) / tot, r[iSrc].strip()[:60], top))
