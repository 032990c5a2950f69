timeout 600 python -m pytest tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -2
for v in "A=0" "A=1"; do
  env $v timeout 200 python bench.py --workload train --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'], d['loss'], d['gpu_launches'])"
done
