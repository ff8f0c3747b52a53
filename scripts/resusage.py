"""Registers / stack of every class kernel in the built library: python scripts/resusage.py [filter] [lib]"""
import re, subprocess, sys
flt = sys.argv[1] if len(sys.argv) > 1 else ""
lib = sys.argv[2] if len(sys.argv) > 2 else "libint_b200/_lib/liblibint_b200.so"
out = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
name = None
rows = []
for ln in out.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m:
        name = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+)", ln)
    if m and name:
        k = re.search(r"(eri_rowreg\w*kernel)ILi(\d)ELi(\d)ELi(\d)ELi(\d)E(?:Li(\d)E)?", name)
        if k:
            tag = "%s %s%s%s%s m%s" % ("prim" if "prim" in k.group(1) else "gen ", k.group(2), k.group(3), k.group(4), k.group(5), k.group(6) or "-")
            rows.append((tag, int(m.group(1)), int(m.group(2))))
        name = None
for tag, r, s in sorted(rows):
    if flt in tag:
        print("%-16s reg %3d stack %4d" % (tag, r, s))
