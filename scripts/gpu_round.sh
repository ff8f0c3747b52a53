#!/bin/bash
# One GPU-box visit: parity tests, bench lines (ours + reference arm), ncu launch list, full ncu
# captures of the dominant class kernels (text summaries only: .ncu-rep files stay on the box),
# Fock per-class profile.   Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
CLS=${NCU_CLASSES:-2222 2121 1111 0000}
for c in $CLS; do
  a=$(echo $c | sed 's/./& /g')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:eri_rowreg -s 1 -c 1 -f -o /tmp/prof_$c \
    python scripts/prof_class.py $a 1048576 2 > $O/prof_$c.log 2>&1
  { echo "# ncu --set full --clock-control none --import-source on -k regex:eri_rowreg -s 1 -c 1 python scripts/prof_class.py $a 1048576 2";
    python scripts/ncu_summary.py /tmp/prof_$c.ncu-rep; echo; echo "## hot instructions (ncu --page source)"; python scripts/ncu_hot.py /tmp/prof_$c.ncu-rep 30; } > $O/ncu_full_$c.txt 2>&1
  python scripts/ncu_traffic.py $O/traffic.json $c=/tmp/prof_$c.ncu-rep > /dev/null 2>&1
done
mkdir -p profiles; cp $O/traffic.json profiles/traffic.json 2>/dev/null
timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv \
  python bench.py --steps 1 --warmup 1 --quartets 1048576 --e2e-quartets 65536 --no-cpu-baseline --fock-waters 2,2,2 --df3c-carbons 8 > $O/bench_under_ncu.log 2>&1
python scripts/launch_summary.py $O/launches.csv > $O/launches_summary.txt 2>&1; gzip -f $O/launches.csv
LB200_FOCK_PROFILE=1 timeout 900 python scripts/fock_profile.py def2-tzvp 4,4,4 > $O/fock_profile.log 2>&1
ls -la $O
tail -3 $O/pytest_gpu.log; head -c 1500 $O/bench.json
