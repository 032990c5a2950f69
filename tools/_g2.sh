for v in "YSP_WGRAD_TC_MINW=129" "YSP_WGRAD_TC_MINW=161" "YSP_WGRAD_TC_MINW=100" "YSP_TRAIN_NO_WGRAD_TC=1"; do
  env $v timeout 200 python bench.py --workload train --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'])"
done
