#!/bin/bash
# Fock build A/B over primitive-loop split settings (LB200_SPLIT_DIV) after the GPU parity tests
TAG=${1:-f2}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
for d in ${DIVS:-"0 4"}; do
  LB200_SPLIT_DIV=$d LB200_FOCK_PROFILE=1 timeout 600 python scripts/fock_profile.py def2-tzvp 4,4,4 > $O/fock_d$d.log 2>&1
  echo "split div $d: $(grep '^build' $O/fock_d$d.log | cut -c1-160)"
done
ls -la $O
