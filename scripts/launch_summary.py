"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv):
   python scripts/launch_summary.py launches.csv
Per kernel: launches, total device time, share of all captured launches, and -- for the
store-mode class kernels (template MODE = 0, the sweep step of bench.py) -- the share within
the sweep, which is the figure to compare with bench.py's roofline.share_of_step."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
agg = collections.OrderedDict()
for r in rows[1:]:
    a = agg.setdefault(r[ik], [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", "")) * scale[r[iu]]
tot = sum(a[1] for a in agg.values())
store = {n: a for n, a in agg.items() if re.search(r"eri_rowreg_kernel<\d, \d, \d, \d, 0>|eri_rowreg_prim_kernel<", n)}
stot = sum(a[1] for a in store.values())
print("%d launches captured, %.3f ms device time; store-mode class kernels %.3f ms" % (len(rows) - 1, tot, stot))
print("%7s %11s %7s %9s  kernel" % ("count", "total ms", "share", "of sweep"))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    short = re.sub(r"\(lb200::EriParams, const lb200::RowInfo \*\)", "", n).replace("void lb200::", "")
    print("%7d %11.3f %6.1f%% %9s  %s" % (c, t, 100 * t / tot, ("%.1f%%" % (100 * t / stot)) if n in store else "", short[:90]))
