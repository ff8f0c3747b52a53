#!/bin/bash
# A/B of library variants on the (H2O)_64 / def2-TZVP Fock build: bash scripts/gpu_ab.sh tag "" _variant ...
TAG=$1; shift
O=gpurun_out/$TAG
mkdir -p $O
for v in "$@"; do
  n=${v:-default}
  LB200_LIB_SUFFIX=$v timeout 600 python scripts/fock_once.py def2-tzvp 4,4,4 > $O/fock_once_$n.log 2>&1
  LB200_LIB_SUFFIX=$v LB200_FOCK_PROFILE=1 timeout 600 python scripts/fock_profile.py def2-tzvp 4,4,4 > $O/fock_profile_$n.log 2>&1
  tail -1 $O/fock_once_$n.log
done
