#!/bin/bash
# Fock build per library variant: VARIANTS="default _x _y" bash scripts/gpu_fock_variants.sh tag
TAG=${1:-fv}
O=gpurun_out/$TAG
mkdir -p $O
for suf in ${VARIANTS:-default}; do
  s=$suf; [ "$suf" = "default" ] && s=""
  LB200_LIB_SUFFIX=$s LB200_FOCK_PROFILE=1 timeout 600 python scripts/fock_profile.py def2-tzvp 4,4,4 > $O/fock_$suf.log 2>&1
  echo "$suf: $(grep '^build' $O/fock_$suf.log | cut -c1-60)"
  LB200_LIB_SUFFIX=$s timeout 600 python scripts/fock_once.py def2-tzvp 4,4,4 | tail -1 | cut -c1-120
done
python - $O ${VARIANTS:-default} <<'P'
import re, sys, collections
O, names = sys.argv[1], sys.argv[2:]
tab = collections.OrderedDict()
for n in names:
    for ln in open("%s/fock_%s.log" % (O, n)):
        m = re.match(r"\s+\((\d\d\|\d\d)\) buckets (\d),(\d)\s+([\d.]+) ms", ln)
        if m:
            tab.setdefault(m.group(1), {}).setdefault(n, 0.0)
            tab[m.group(1)][n] += float(m.group(4))
for k, v in sorted(tab.items(), key=lambda kv: -max(kv[1].values()))[:30]:
    print(k, "  ".join("%s %.1f" % (n, v.get(n, 0)) for n in names))
P
