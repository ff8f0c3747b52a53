#!/bin/bash
# ncu --set full over the Fock-mode class kernels of one small direct Fock build.
# Usage (under gpurun): bash scripts/gpu_fock_ncu.sh [tag] [nx,ny,nz] [count]
TAG=${1:-r01f}
W=${2:-2,2,2}
N=${3:-70}
O=gpurun_out/$TAG
mkdir -p $O
LB200_FOCK_PROFILE=1 timeout 600 python scripts/fock_profile.py def2-tzvp $W > $O/fock_profile_small.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:eri_rowreg -c $N -f -o $O/fock_kernels \
  python scripts/fock_profile.py def2-tzvp $W > $O/fock_ncu.log 2>&1
ls -la $O
tail -5 $O/fock_profile_small.log
