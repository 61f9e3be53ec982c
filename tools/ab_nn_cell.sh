for lib in libb200reg_old.so libb200reg.so libb200reg_c35.so libb200reg_c25.so; do
  B200REG_LIB=$PWD/delta_graph_slam_b200/$lib python bench.py --workload loop --loop-targets 64 --steps 2 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib loop', 'pairs/s %.0f'%d['value'], 'align %.1f'%d['roofline']['avg_launch_ms'], 'fitness %.1f'%d['roofline']['fitness_ms_per_step'])"
  B200REG_LIB=$PWD/delta_graph_slam_b200/$lib python bench.py --no-loop --no-dense --frames 200 --steps 2 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); g=d['gicp_odometry']
print('$lib gicp %.0f e2e %.0f align_ms %.3f' % (g['value'], g['e2e']['value'], g['roofline']['avg_launch_ms']))"
done
