#!/bin/bash
# 2-GPU check of the bench under torchrun (NCCL all-reduce of the partial Fock matrices)
O=gpurun_out/${1:-n2}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 2 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?"
tail -3 $O/bench_n2.err; head -c 700 $O/bench_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-seconds 5 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref rc=$?"; head -c 300 $O/bench_ref_n2.json
