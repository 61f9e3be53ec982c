#!/bin/bash
# One gpurun call: GPU parity tests, bench (both arms), ncu launch lists, full captures of the top kernels.
set -x
mkdir -p gpurun_out
lscpu > gpurun_out/lscpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python tools/profile_driver.py 8 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_odometry.csv python tools/profile_driver.py 8 > gpurun_out/ncu_launches.log 2>&1
python tools/profile_driver.py 8 gicp > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_gicp.csv python tools/profile_driver.py 8 gicp > /dev/null 2>&1
python tools/profile_batch.py > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_nn\|k_fitness\|k_ndt -c 300 --csv --log-file gpurun_out/launches_batch.csv python tools/profile_batch.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ndt_align -s 3 -c 1 -f -o gpurun_out/prof_align python tools/profile_driver.py 8 > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_voxel_sort_coop -s 4 -c 1 -f -o gpurun_out/prof_sort python tools/profile_driver.py 8 > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ndt_align -s 1 -c 1 -f -o gpurun_out/prof_align_batch python tools/profile_batch.py > gpurun_out/ncu_full3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gicp_align -s 2 -c 1 -f -o gpurun_out/prof_gicp_align python tools/profile_driver.py 6 gicp > gpurun_out/ncu_full4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gicp_knn$ -s 2 -c 1 -f -o gpurun_out/prof_gicp_knn python tools/profile_driver.py 6 gicp > gpurun_out/ncu_full5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_nn_search_batch -s 1 -c 1 -f -o gpurun_out/prof_nn_batch python tools/profile_batch.py > gpurun_out/ncu_full6.log 2>&1
echo done
