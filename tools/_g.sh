mkdir -p gpurun_out/r8
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_ddp.py -x -q -m gpu 2>&1 | tail -2
timeout 400 python bench.py --workload train --steps 20 --warmup 3 > gpurun_out/r8/train_n1.json 2> gpurun_out/r8/train_n1.err; echo "train rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r8/train_n1.json'));print(d['value'],d['ms_per_step'],d['gpu_launches'],d['roofline']['achieved'])"
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
