for t in 32 64 256; do
python bench.py --workload loop --loop-targets $t --steps 2 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('targets $t: pairs/s %.0f'%d['value'], 'ms/step %.1f'%d['ms_per_step'], 'align %.1f'%d['roofline']['avg_launch_ms'], 'fitness %.1f'%d['roofline']['fitness_ms_per_step'], d['roofline']['kernel'], d['checks']['device_and_host_legs_bit_identical'], d['checks']['pairs_within_5cm_of_ground_truth'])"
done
