mkdir -p gpurun_out/r8
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r8/pytest_all.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r8/pytest_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r8/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r8/smoke.log
timeout 400 python bench.py --workload train --steps 20 --warmup 3 > gpurun_out/r8/train_n1.json 2> gpurun_out/r8/train_n1.err; echo "train rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r8/train_n1.json'));print(d['value'],d['ms_per_step'],d['gpu_launches'],d['roofline']['achieved'])"
