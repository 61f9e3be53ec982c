python bench.py --workload loop --loop-targets 64 --steps 2 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('loop: pairs/s %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'ms/step %.1f'%d['ms_per_step'], 'align %.1f'%d['roofline']['avg_launch_ms'], 'fitness %.1f'%d['roofline']['fitness_ms_per_step'], d['checks'])"
