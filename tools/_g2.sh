for v in "YSP_NO_LANES=1" "YSP_NO_OVERLAP=1"; do
  env $v timeout 100 python bench.py --steps 60 --single-mode --no-library --no-cpu --parity-slices 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
