mkdir -p gpurun_out/r8
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -m gpu > gpurun_out/r8/pytest_train.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r8/pytest_train.log
for v in "" ; do
  env $v timeout 300 python bench.py --workload train --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'], d['loss'], d['gpu_launches'], d['roofline']['achieved'])"
done
timeout 600 ncu -k regex:dw_bwd_fused --launch-skip 12 --launch-count 6 --set full --import-source on --clock-control none -o gpurun_out/r8/dwfused2 python bench.py --workload train --steps 1 --warmup 2 --no-cpu > gpurun_out/r8/ncu_dwf.log 2>&1; echo "ncu rc=$?"
