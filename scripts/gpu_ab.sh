#!/bin/bash
# A/B of the sweep: pipelined uncontracted kernel vs the general kernel, plus GPU parity tests.
TAG=${1:-ab}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-fock --no-cpu-baseline --e2e-quartets 65536 > $O/bench_prim.json 2> $O/bench_prim.err
LB200_NO_PRIM_KERNEL=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-fock --no-cpu-baseline --e2e-quartets 65536 > $O/bench_gen.json 2> $O/bench_gen.err
TAG=$TAG python - <<'P' | tee $O/ab.txt
import json, os
T = os.environ["TAG"]
a = json.loads(open("gpurun_out/%s/bench_prim.json" % T).read().strip().splitlines()[-1])
b = json.loads(open("gpurun_out/%s/bench_gen.json" % T).read().strip().splitlines()[-1])
print("sweep ms/step prim %.2f general %.2f" % (a["ms_per_step"], b["ms_per_step"]))
for k in a["per_class"]:
    print(k, "prim %.3f ms  general %.3f ms  x%.2f  fp64 %.3f hbm %.3f" % (a["per_class"][k]["ms"], b["per_class"][k]["ms"], b["per_class"][k]["ms"] / a["per_class"][k]["ms"], a["per_class"][k]["fp64_frac"], a["per_class"][k]["hbm_frac"]))
P
