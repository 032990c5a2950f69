for v in "YSP_UP2_PER_SM=2" "YSP_UP2_PER_SM=4" "YSP_UP2_PER_SM=8"; do
  env $v timeout 200 python bench.py --workload train --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'])"
done
