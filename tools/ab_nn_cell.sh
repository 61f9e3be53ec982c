for lib in libb200reg.so libb200reg_m8.so; do
  B200REG_LIB=$PWD/delta_graph_slam_b200/$lib python bench.py --workload loop --loop-targets 64 --steps 2 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib loop', 'pairs/s %.0f'%d['value'], 'align %.1f'%d['roofline']['avg_launch_ms'], 'fitness %.1f'%d['roofline']['fitness_ms_per_step'])"
done
