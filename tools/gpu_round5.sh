#!/bin/bash
# One short gpurun call: GPU parity tests, the outlier-filter probe (plain), then its ncu launch list and one
# --set full capture of the statistical filter's k-NN kernel (summarised on the box).
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 60 python tools/sor_probe.py > gpurun_out/sor_probe.json 2> gpurun_out/sor_probe.err && \
SOR_PROBE_REPS=2 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_sor.csv python tools/sor_probe.py > /dev/null 2>&1
SOR_PROBE_REPS=2 timeout 150 ncu --set full --clock-control none --import-source on -k "regex:k_gicp_knn" -s 1 -c 1 -f -o gpurun_out/prof_k_sor_knn python tools/sor_probe.py > gpurun_out/ncu_full_k_sor_knn.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_k_sor_knn.ncu-rep > gpurun_out/ncu_full_k_sor_knn.txt 2>/dev/null
ncu -i gpurun_out/prof_k_sor_knn.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_sor.csv 2>/dev/null
python tools/ncu_lines.py /tmp/src_sor.csv 40 > gpurun_out/ncu_lines_k_sor_knn.txt 2>/dev/null
rm -f gpurun_out/prof_k_sor_knn.ncu-rep
echo done
