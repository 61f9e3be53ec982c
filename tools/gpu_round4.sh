#!/bin/bash
# Final refresh at HEAD: GPU tests, both bench arms, launch lists, captures of the two headline kernels.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python tools/profile_driver.py 8 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_odometry.csv python tools/profile_driver.py 8 > /dev/null 2>&1
python tools/profile_batch.py > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_nn\|k_fitness\|k_ndt -c 300 --csv --log-file gpurun_out/launches_batch.csv python tools/profile_batch.py > /dev/null 2>&1
cap() {
  local name=$1 k=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s $skip -c 1 -f -o gpurun_out/prof_$name "$@" > gpurun_out/ncu_full_$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_$name.ncu-rep > gpurun_out/ncu_full_$name.txt 2>/dev/null
  ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_$name.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/src_$name.csv 40 > gpurun_out/ncu_lines_$name.txt 2>/dev/null
}
cap k_ndt_align k_ndt_align 3 python tools/profile_driver.py 8
cap k_ndt_align_batch k_ndt_align 1 python tools/profile_batch.py
cap k_gicp_align k_gicp_align 2 python tools/profile_driver.py 6 gicp
python tools/ncu_traffic.py gpurun_out/ncu_traffic.json k_ndt_align_single=gpurun_out/prof_k_ndt_align.ncu-rep k_ndt_align_batch_per_registration=gpurun_out/prof_k_ndt_align_batch.ncu-rep/160 k_gicp_align=gpurun_out/prof_k_gicp_align.ncu-rep > /dev/null 2>&1
rm -f gpurun_out/prof_k_gicp_align.ncu-rep gpurun_out/prof_k_ndt_align_batch.ncu-rep
python tools/dev_timing.py 2>&1 | grep "DIRECT7 align\|DIRECT7 profile" > gpurun_out/dev_timing.txt
python tools/gicp_probe.py > gpurun_out/gicp_probe.txt 2>&1
echo done
