for i in 1 2 3 4; do python bench.py --no-cpu-baseline --no-batched 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), d['ms_per_step'], d['clocks'], d['e2e']['value'], d['roofline']['avg_launch_ms'], d.get('wall_ms_per_step'))
"; done
nvidia-smi --query-gpu=temperature.gpu,power.draw,clocks.sm,clocks_throttle_reasons.active --format=csv
