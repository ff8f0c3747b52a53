#!/bin/bash
# A/B of full-library variants: sweep per class, one (H2O)_64 Fock build, and the GPU tests on the variant
O=gpurun_out/${1:-ab3}; shift
mkdir -p $O
for suf in "$@"; do
  s=$suf; [ "$suf" = "default" ] && s=""
  LB200_LIB_SUFFIX=$s timeout 400 python bench.py --steps 3 --warmup 3 --no-fock --no-cpu-baseline --no-df3c --no-grad --e2e-quartets 65536 > $O/bench_$suf.json 2> $O/bench_$suf.err
  LB200_LIB_SUFFIX=$s timeout 400 python scripts/fock_once.py def2-tzvp 4,4,4 > $O/fock_$suf.log 2>&1
  tail -1 $O/fock_$suf.log
done
python - $O "$@" <<'P'
import json, sys
O, names = sys.argv[1], sys.argv[2:]
d = {}
for n in names:
    try:
        d[n] = json.loads(open("%s/bench_%s.json" % (O, n)).read().strip().splitlines()[-1])
    except Exception as e:
        print("variant", n, "failed:", e)
names = [n for n in names if n in d]
print("ms/step   " + "  ".join("%s %.2f" % (n, d[n]["ms_per_step"]) for n in names))
for k in d[names[0]]["per_class"]:
    print(k, "  ".join("%s %.3f" % (n, d[n]["per_class"][k]["ms"]) for n in names))
P
last=${@: -1}; s=$last; [ "$last" = "default" ] && s=""
LB200_LIB_SUFFIX=$s timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest_$last.log 2>&1; tail -3 $O/pytest_$last.log
