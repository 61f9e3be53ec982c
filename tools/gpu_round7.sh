#!/bin/bash
# Last seconds of the round's GPU budget: the flat-cloud filter plain, then ONE ncu --set full pass over the launches of one call.
mkdir -p gpurun_out
timeout 30 python tools/flat_probe.py > gpurun_out/flat_probe2.json 2> gpurun_out/flat_probe2.err || exit 0
FLAT_PROBE_REPS=4 timeout 40 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_flat python tools/flat_probe.py > gpurun_out/ncu_flat.log 2>&1
timeout 15 python tools/ncu_summary.py gpurun_out/prof_flat.ncu-rep > gpurun_out/ncu_full_flat_filter.txt 2>/dev/null
rm -f gpurun_out/prof_flat.ncu-rep
echo done
