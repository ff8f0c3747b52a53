#!/bin/bash
# ncu over the Fock-mode class kernels of one small direct Fock build; only text summaries are
# kept (the .ncu-rep is too large to travel back).
# Usage (under gpurun): bash scripts/gpu_fock_ncu.sh [tag] [nx,ny,nz] [count]
TAG=${1:-r01f}
W=${2:-2,2,2}
N=${3:-60}
O=gpurun_out/$TAG
mkdir -p $O
LB200_FOCK_PROFILE=1 timeout 600 python scripts/fock_profile.py def2-tzvp $W > $O/fock_profile_small.log 2>&1
timeout 900 ncu --section LaunchStats --section Occupancy --section SpeedOfLight --section WarpStateStats \
  --section SchedulerStats --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis \
  --clock-control none -k regex:eri_rowreg -c $N -f -o /tmp/fock_kernels \
  python scripts/fock_profile.py def2-tzvp $W > $O/fock_ncu.log 2>&1
python scripts/ncu_summary.py /tmp/fock_kernels.ncu-rep > $O/fock_kernels_summary.txt 2>&1
ls -la $O
