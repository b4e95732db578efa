timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --no-cpu-baseline 2>gpurun_out/exp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['sub_metrics']['kernel_ms']
print('RESULT value %.1fM ms %.2f e2e %.1fM fwd %.2f back %.2f loc %.2f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, k['fwd_kernel'], k['back_kernel'], k['locate_kernel']))
print('RESULT', d['roofline'])"
