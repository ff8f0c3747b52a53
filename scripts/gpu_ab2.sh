#!/bin/bash
# A/B of library variants on two Fock workloads (warm build, 4 issue streams): bash scripts/gpu_ab2.sh "" _variant ...
for v in "$@"; do
  for w in "def2-tzvp 4,4,4" "cc-pvtz 4,4,4"; do
    echo "variant [$v] $w: $(LB200_LIB_SUFFIX=$v python scripts/fock_profile.py $w 1 2>/dev/null | tail -1)"
  done
done
