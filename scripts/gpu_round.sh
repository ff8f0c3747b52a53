#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list, full ncu captures, Fock profile.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 1200 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 1 --warmup 1 --quartets 1048576 --e2e-quartets 65536 --no-cpu-baseline --fock-waters 2,2,2 > $O/bench_under_ncu.log 2>&1
for c in "2 2 2 2" "1 1 1 1" "0 0 0 0" "2 1 2 1"; do
  n=$(echo $c | tr -d ' ')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:eri_rowreg -s 1 -c 1 -f -o $O/prof_$n \
    python scripts/prof_class.py $c 1048576 2 > $O/prof_$n.log 2>&1
done
LB200_FOCK_PROFILE=1 timeout 900 python scripts/fock_profile.py def2-tzvp 4,4,4 > $O/fock_profile.log 2>&1
ls -la $O
tail -3 $O/pytest_gpu.log; head -c 1500 $O/bench.json
