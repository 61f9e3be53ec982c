#!/bin/bash
# The default bench line at N ranks, as the driver's scaling run launches it:  gpurun --gpus N -- bash tools/scale_run.sh N
set -u
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus "$N" --steps 3 --warmup 3 \
  > "gpurun_out/bench_n$N.json" 2> "gpurun_out/bench_n$N.err"
echo "bench default N=$N rc=$?"
tail -c 1300 "gpurun_out/bench_n$N.json"
