#!/bin/bash
# Multi-GPU checks that need more than one visible device: run as  gpurun --gpus 2 -- bash tools/multi_gpu_check.sh
# (1) the single-process multi-GPU C entry (b200reg_batch_*, NCCL all-gather inside) against the single-handle batch,
# (2) the torchrun loop-batch bench line at N ranks (what the driver's scaling run launches).
set -u
N=${1:-2}
python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus "$N" --steps 3 --warmup 3 --workload loop \
  > "gpurun_out/bench_loop_n$N.json" 2> "gpurun_out/bench_loop_n$N.err"
echo "bench loop N=$N rc=$?"
tail -c 900 "gpurun_out/bench_loop_n$N.json"
python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus "$N" --steps 2 --warmup 3 \
  > "gpurun_out/bench_n$N.json" 2> "gpurun_out/bench_n$N.err"
echo "bench default N=$N rc=$?"
tail -c 1500 "gpurun_out/bench_n$N.json"
