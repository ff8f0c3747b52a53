#!/bin/bash
# Fock builds (def2-TZVP (H2O)_64 and cc-pVTZ (H2O)_27) per library variant
TAG=${1:-fv2}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
for suf in ${VARIANTS:-default}; do
  s=$suf; [ "$suf" = "default" ] && s=""
  echo "== $suf"
  LB200_LIB_SUFFIX=$s timeout 600 python scripts/fock_once.py def2-tzvp 4,4,4 | tail -1 | cut -c1-110
  LB200_LIB_SUFFIX=$s timeout 600 python scripts/fock_once.py cc-pvtz 3,3,3 | tail -1 | cut -c1-110
done
