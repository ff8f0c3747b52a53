#!/bin/bash
# Round-2 (second session) evidence visit: GPU tests, smoke, bench lines (ours + reference arm), ncu launch list of a
# reduced bench step, one full ncu capture of the gradient contraction kernel.  Text summaries only.
TAG=${1:-r03}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
nproc >> $O/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
if [ -z "$SKIP_BENCH" ]; then
timeout 1200 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv \
  python bench.py --steps 1 --warmup 1 --quartets 1048576 --e2e-quartets 65536 --no-cpu-baseline --fock-waters 2,2,2 --df3c-carbons 8 --grad-waters 2,1,1 > $O/bench_under_ncu.log 2>&1
python scripts/launch_summary.py $O/launches.csv > $O/launches_summary.txt 2>&1; gzip -f $O/launches.csv
# dominant store-mode kernels, full captures (the kernel changed in this session: cross-term prefetch)
for c in 2222 2122; do
  a=$(echo $c | sed 's/./& /g')
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:eri_rowreg -s 1 -c 1 -f -o /tmp/prof_$c \
    python scripts/prof_class.py $a 1048576 2 > $O/prof_$c.log 2>&1
  { echo "# ncu --set full --clock-control none --import-source on -k regex:eri_rowreg -s 1 -c 1 python scripts/prof_class.py $a 1048576 2";
    python scripts/ncu_summary.py /tmp/prof_$c.ncu-rep; echo; echo "## hot instructions (ncu --page source)"; python scripts/ncu_hot.py /tmp/prof_$c.ncu-rep 30; } > $O/ncu_full_$c.txt 2>&1
  python scripts/ncu_traffic.py $O/traffic.json $c=/tmp/prof_$c.ncu-rep > /dev/null 2>&1
done
# compute-sanitizer memcheck on the code added in this session (derivatives, forces, device pair records, eri1 boundary)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest -q -x \
  "tests/test_gpu_deriv.py::test_deriv1_batch_vs_closed_form" "tests/test_gpu_deriv.py::test_forces_2body_vs_golden" \
  "tests/test_gpu_deriv.py::test_pair_records_built_on_device_match_host" \
  "tests/test_gpu_iface.py::test_reference_engine_first_derivatives_on_gpu_library" \
  "tests/test_gpu_eri.py::test_primitive_screening_matches_engine" "tests/test_gpu_eri.py::test_uncontracted_pipeline_ragged" \
  > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/sanitizer_memcheck.log
timeout 300 python scripts/forces_once.py cc-pvdz 3,3,3 > $O/forces_once.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:deriv_grad_kernel -s 40 -c 1 -f -o /tmp/prof_grad \
  python scripts/forces_once.py cc-pvdz 2,2,2 > $O/prof_grad.log 2>&1
{ echo "# ncu --set full --clock-control none --import-source on -k regex:deriv_grad_kernel -s 40 -c 1 python scripts/forces_once.py cc-pvdz 2,2,2";
  python scripts/ncu_summary.py /tmp/prof_grad.ncu-rep; } > $O/ncu_full_deriv_grad.txt 2>&1
fi
ls -la $O
tail -2 $O/smoke.log; tail -5 $O/pytest_gpu.log; tail -3 $O/forces_once.log; tail -4 $O/sanitizer_memcheck.log
