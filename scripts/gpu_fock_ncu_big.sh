#!/bin/bash
# ncu --set full of the Fock-mode kernels of chosen classes inside a full-size build
# (text summaries only).  Usage: CLASSES="1010 0010" bash scripts/gpu_fock_ncu_big.sh [tag] [nx,ny,nz] [count]
TAG=${1:-fb}
W=${2:-4,4,4}
N=${3:-9}
O=gpurun_out/$TAG
mkdir -p $O
for c in ${CLASSES:-1010}; do
  re=$(echo $c | sed -E 's/(.)(.)(.)(.)/\\(int\\)\1, \\(int\\)\2, \\(int\\)\3, \\(int\\)\4, \\(int\\)2>/')
  timeout 1200 ncu --set full --import-source on --clock-control none --kernel-name-base demangled \
    -k "regex:eri_rowreg_kernel<$re" -c $N -f -o /tmp/fk_$c python scripts/fock_profile.py def2-tzvp $W > $O/fk_$c.log 2>&1
  python scripts/ncu_summary.py /tmp/fk_$c.ncu-rep > $O/fk_${c}_summary.txt 2>&1
  python scripts/ncu_hot.py /tmp/fk_$c.ncu-rep 40 > $O/fk_${c}_hot.txt 2>&1
done
ls -la $O
