#!/bin/bash
# Fock build A/B over contraction-bucket settings + Fock-mode kernel ncu summary + 3-centre sweep
TAG=${1:-f1}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
for b in ${BUCKETS:-"0 1,6"}; do
  LB200_FOCK_BUCKETS=$b LB200_FOCK_PROFILE=1 timeout 600 python scripts/fock_profile.py def2-tzvp 4,4,4 > $O/fock_b$b.log 2>&1
  echo "buckets $b: $(grep '^build' $O/fock_b$b.log | cut -c1-160)"
done
timeout 600 python scripts/df3c_bench.py 40 > $O/df3c.log 2>&1; tail -16 $O/df3c.log
timeout 900 ncu --section LaunchStats --section Occupancy --section SpeedOfLight --section WarpStateStats \
  --section SchedulerStats --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis \
  --clock-control none --kernel-name-base demangled -k 'regex:eri_rowreg_kernel<.*\(int\)2>' -c 80 -f -o /tmp/fock_kernels \
  python scripts/fock_profile.py def2-tzvp 2,2,2 > $O/fock_ncu.log 2>&1
python scripts/ncu_summary.py /tmp/fock_kernels.ncu-rep > $O/fock_kernels_summary.txt 2>&1
ls -la $O
