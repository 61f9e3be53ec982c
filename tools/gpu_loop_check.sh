python -m pytest tests/test_gpu_loop_batch.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -4
for t in 64 128 256; do
python bench.py --workload loop --loop-targets $t --steps 2 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('targets $t: pairs/s %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'ms/step %.1f'%d['ms_per_step'], 'align %.1f'%d['roofline']['avg_launch_ms'], 'fitness %.1f'%d['roofline']['fitness_ms_per_step'], d['checks']['device_and_host_legs_bit_identical'])"
done
