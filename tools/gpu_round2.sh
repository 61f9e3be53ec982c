#!/bin/bash
set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -25
python tools/frame_breakdown.py 300
python bench.py --steps 2 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"; tail -5 gpurun_out/bench2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench2.json"))
def show(n,o):
    print(n, "value %.1f %s"%(o["value"],o["unit"]), "e2e %.1f"%o["e2e"]["value"], "ms/step %.1f"%o["ms_per_step"], "roofline frac %.3f avg_launch_ms %.4f share %.2f"%(o["roofline"]["frac"],o["roofline"]["avg_launch_ms"],o["roofline"]["share_of_step"]), "cpu", o["cpu_baseline"] and round(o["cpu_baseline"]["value"],2), o["checks"])
show("odometry",d); show("loop",d["loop_batch"]); show("gicp",d["gicp_odometry"])
PY
