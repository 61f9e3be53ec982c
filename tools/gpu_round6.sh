#!/bin/bash
# One short gpurun call at HEAD: GPU parity tests, the outlier-filter probe + its ncu launch list, the N=1 bench line.
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 60 python tools/sor_probe.py > gpurun_out/sor_probe.json 2> gpurun_out/sor_probe.err && \
SOR_PROBE_REPS=2 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_sor.csv python tools/sor_probe.py > /dev/null 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
echo done
