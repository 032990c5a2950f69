mkdir -p gpurun_out/r8
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_ddp.py tests/test_gpu_examples.py -x -q -m gpu > gpurun_out/r8/pytest_train.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r8/pytest_train.log
timeout 600 python bench.py --workload train --steps 20 --warmup 3 > gpurun_out/r8/train_n1.json 2> gpurun_out/r8/train_n1.err; echo "train bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r8/train_n1.json'));print(d['value'],d['ms_per_step'],d['gpu_launches'],d['roofline']['achieved'],d['cpu_baseline']['value'])"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
