#!/bin/bash
# One gpurun call: GPU parity tests, bench (plain), then ncu launch list + full captures of the top kernels.
set -x
mkdir -p gpurun_out
lscpu > gpurun_out/lscpu.txt 2>&1
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json
python tools/profile_driver.py 8 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tools/profile_driver.py 8 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ndt_align -s 3 -c 1 -f -o gpurun_out/prof_align python tools/profile_driver.py 8 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_voxel_sort_coop -s 4 -c 1 -f -o gpurun_out/prof_sort python tools/profile_driver.py 8 > gpurun_out/ncu_full2.log 2>&1
echo done
