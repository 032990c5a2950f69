mkdir -p gpurun_out/r8
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_ddp.py -x -q -m gpu > gpurun_out/r8/pytest_train.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r8/pytest_train.log
for v in "A=0" "YSP_TRAIN_NO_EPISTAT=1" "A=1"; do
  env $v timeout 300 python bench.py --workload train --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'], d['loss'], d['gpu_launches'])"
done
