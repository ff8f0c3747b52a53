#!/bin/bash
# Round-2 evidence visit: smoke, GPU tests, bench lines (ours + reference arm), ncu launch list, full ncu captures
# of the dominant store kernel and of three Fock-mode kernels, compute-sanitizer runs.  Text summaries only.
TAG=${1:-r02z}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
nproc >> $O/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> $O/gpu.txt
timeout 600 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err
# launch list of a reduced bench step (timing pass: cold-cache, serialised launches)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches.csv \
  python bench.py --steps 1 --warmup 1 --quartets 1048576 --e2e-quartets 65536 --no-cpu-baseline --fock-waters 2,2,2 --df3c-carbons 8 > $O/bench_under_ncu.log 2>&1
python scripts/launch_summary.py $O/launches.csv > $O/launches_summary.txt 2>&1; gzip -f $O/launches.csv
# dominant store-mode kernel
for c in 2222 2122; do
  a=$(echo $c | sed 's/./& /g')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:eri_rowreg -s 1 -c 1 -f -o /tmp/prof_$c \
    python scripts/prof_class.py $a 1048576 2 > $O/prof_$c.log 2>&1
  { echo "# ncu --set full --clock-control none --import-source on -k regex:eri_rowreg -s 1 -c 1 python scripts/prof_class.py $a 1048576 2";
    python scripts/ncu_summary.py /tmp/prof_$c.ncu-rep; echo; echo "## hot instructions (ncu --page source)"; python scripts/ncu_hot.py /tmp/prof_$c.ncu-rep 30; } > $O/ncu_full_$c.txt 2>&1
  python scripts/ncu_traffic.py $O/traffic.json $c=/tmp/prof_$c.ncu-rep > /dev/null 2>&1
done
# Fock-mode kernels inside a full-size build (first Fock-mode launch of each kernel)
for c in 0010 1010 1020 2021; do
  re=$(echo $c | sed -E 's/(.)(.)(.)(.)/(\\(int\\))?\1, (\\(int\\))?\2, (\\(int\\))?\3, (\\(int\\))?\4,/')
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:eri_rowreg.*kernel<$re.*(2|true)>" -c 1 -f -o /tmp/fk_$c python scripts/fock_once.py def2-tzvp 4,4,4 > $O/fk_$c.log 2>&1
  { echo "# ncu --set full, first Fock-mode launch of kernel<$c> inside scripts/fock_once.py def2-tzvp 4,4,4";
    python scripts/ncu_summary.py /tmp/fk_$c.ncu-rep; echo; echo "## hot instructions (ncu --page source)"; python scripts/ncu_hot.py /tmp/fk_$c.ncu-rep 30; } > $O/ncu_fock_$c.txt 2>&1
done
# sanitizers on small cases
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest -q -x \
  "tests/test_gpu_fock.py::test_fock_h2o_vs_golden" "tests/test_gpu_fock.py::test_fock_cartesian_d_basis" \
  "tests/test_gpu_eri.py::test_every_class_vs_committed_goldens" "tests/test_gpu_iface.py::test_c_api_port_vs_engine" \
  "tests/test_gpu_df3c.py::test_df_slab_metric_and_fock_vs_reference_formulas" "tests/test_scf.py::test_onebody_device_matches_host" \
  > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest -q -x \
  "tests/test_gpu_fock.py::test_fock_h2o_vs_golden" "tests/test_gpu_eri.py::test_uncontracted_pipeline_ragged" \
  > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 3 python -m pytest -q -x \
  "tests/test_gpu_fock.py::test_fock_h2o_vs_golden" "tests/test_gpu_eri.py::test_uncontracted_pipeline_ragged" \
  > $O/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" >> $O/sanitizer_synccheck.log
ls -la $O
tail -2 $O/smoke.log; tail -3 $O/pytest_gpu.log; tail -4 $O/sanitizer_*.log
