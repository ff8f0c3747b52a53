#!/bin/bash
# Round-2 first GPU visit: GPU tests (incl. purity, Libint_t boundary), parity vs the arbiter on the
# bench inputs (FMA and -fmad=false builds), bench line, Fock-mode ncu captures of representative classes.
TAG=${1:-r02a}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
nproc >> $O/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> $O/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 900 python scripts/parity_truth.py --quartets 100000 --out $O/parity_truth.json > $O/parity_truth.log 2>&1
timeout 600 python scripts/parity_truth.py --quartets 20000 --lib-suffix _nofma --out $O/parity_truth_nofma.json > $O/parity_truth_nofma.log 2>&1
timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
# Fock-mode captures inside a full-size (H2O)_64 / def2-TZVP build: first launch of each kernel
for c in ${FOCK_NCU:-1010 1020 0010 3210 1021 3300}; do
  re=$(echo $c | sed -E 's/(.)(.)(.)(.)/(\\(int\\))?\1, (\\(int\\))?\2, (\\(int\\))?\3, (\\(int\\))?\4,/')
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:eri_rowreg.*kernel<$re" -c 1 -f -o /tmp/fk_$c python scripts/fock_once.py def2-tzvp 4,4,4 > $O/fk_$c.log 2>&1
  { echo "# ncu --set full --clock-control none --import-source on -k regex:eri_rowreg.*kernel<$re -c 1 python scripts/fock_once.py def2-tzvp 4,4,4";
    python scripts/ncu_summary.py /tmp/fk_$c.ncu-rep; echo; echo "## hot instructions (ncu --page source)"; python scripts/ncu_hot.py /tmp/fk_$c.ncu-rep 40; } > $O/ncu_fock_$c.txt 2>&1
done
ls -la $O
tail -5 $O/pytest_gpu.log; tail -3 $O/parity_truth.log; head -c 1200 $O/bench.json
