#!/bin/bash
# Round-2 GPU visit B: correctness of the 64-byte records / staged digestion / Boys recursion, then
# Fock (H2O)_64 timing of the default build vs the LB200_BOYS_RECUR=0 variant, profile with the
# warp-aggregated counters, and Fock-mode ncu of the kernels that dominate.
TAG=${1:-r02b}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
for v in "" _norecur; do
  LB200_LIB_SUFFIX=$v timeout 600 python scripts/fock_once.py def2-tzvp 4,4,4 > $O/fock_once$v.log 2>&1
done
LB200_FOCK_PROFILE=1 timeout 900 python scripts/fock_profile.py def2-tzvp 4,4,4 > $O/fock_profile.log 2>&1
timeout 600 python scripts/parity_truth.py --quartets 20000 --out $O/parity_truth.json > $O/parity_truth.log 2>&1
timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
for c in ${FOCK_NCU:-0010 1020 0020}; do
  re=$(echo $c | sed -E 's/(.)(.)(.)(.)/(\\(int\\))?\1, (\\(int\\))?\2, (\\(int\\))?\3, (\\(int\\))?\4,/')
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:eri_rowreg.*kernel<$re.*(2|true)>" -c 1 -f -o /tmp/fk_$c python scripts/fock_once.py def2-tzvp 4,4,4 > $O/fk_$c.log 2>&1
  { echo "# ncu --set full, first Fock-mode launch of kernel<$c> inside scripts/fock_once.py def2-tzvp 4,4,4";
    python scripts/ncu_summary.py /tmp/fk_$c.ncu-rep; echo; echo "## hot instructions (ncu --page source)"; python scripts/ncu_hot.py /tmp/fk_$c.ncu-rep 40; } > $O/ncu_fock_$c.txt 2>&1
done
ls -la $O
tail -4 $O/pytest_gpu.log; tail -2 $O/fock_once*.log; tail -2 $O/parity_truth.log
