#!/bin/bash
# Sweep timing per library variant (LB200_LIB_SUFFIX list in $VARIANTS), ncu --set full of chosen
# classes (text summaries only), Fock-mode kernel ncu summary.
TAG=${1:-v1}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
for suf in ${VARIANTS:-"default"}; do
  s=$suf; [ "$suf" = "default" ] && s=""
  LB200_LIB_SUFFIX=$s timeout 600 python bench.py --steps 2 --warmup 1 --no-fock --no-cpu-baseline --e2e-quartets 65536 > $O/bench_$suf.json 2> $O/bench_$suf.err
done
python - $O ${VARIANTS:-"default"} <<'P' | tee $O/variants.txt
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
print("e2e       " + "  ".join("%s %.3e" % (n, d[n]["e2e"]["value"]) for n in names))
for k in d[names[0]]["per_class"]:
    print(k, "  ".join("%s %.3f" % (n, d[n]["per_class"][k]["ms"]) for n in names),
          " fp64 %.3f hbm %.3f" % (d[names[0]]["per_class"][k]["fp64_frac"], d[names[0]]["per_class"][k]["hbm_frac"]))
P
for c in ${NCU_CLASSES:-"2222"}; do
  a=$(echo $c | sed 's/./& /g')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:eri_rowreg -s 1 -c 1 -f -o /tmp/prof_$c \
    python scripts/prof_class.py $a 1048576 2 > $O/prof_$c.log 2>&1
  { python scripts/ncu_summary.py /tmp/prof_$c.ncu-rep; echo; echo "## hot instructions"; python scripts/ncu_hot.py /tmp/prof_$c.ncu-rep 45; } > $O/ncu_full_$c.txt 2>&1
  python scripts/ncu_traffic.py $O/traffic.json $c=/tmp/prof_$c.ncu-rep > /dev/null 2>&1
done
if [ -n "$FOCK_NCU" ]; then
  timeout 900 ncu --section LaunchStats --section Occupancy --section SpeedOfLight --section WarpStateStats \
    --section SchedulerStats --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis \
    --clock-control none --kernel-name-base demangled -k 'regex:eri_rowreg_kernel<.*, 2>' -c 60 -f -o /tmp/fock_kernels \
    python scripts/fock_profile.py def2-tzvp $FOCK_NCU > $O/fock_ncu.log 2>&1
  python scripts/ncu_summary.py /tmp/fock_kernels.ncu-rep > $O/fock_kernels_summary.txt 2>&1
fi
ls -la $O
