for v in "YSP_TRAIN_NO_SPLIT=0" "YSP_TRAIN_NO_SPLIT=1" "YSP_TRAIN_NO_SPLIT=2" "YSP_TRAIN_NO_SPLIT=3" "YSP_TRAIN_NO_SPLIT=0"; do
  env $v timeout 300 python bench.py --workload train --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'], d['loss'], d['gpu_launches'])"
done
