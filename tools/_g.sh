mkdir -p gpurun_out/r8
timeout 600 python bench.py --workload train --steps 20 --warmup 3 > gpurun_out/r8/train_n1.json 2> gpurun_out/r8/train_n1.err; echo "train bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r8/train_n1.json'));print(d['value'],d['ms_per_step'],d['gpu_launches'],d['roofline'],d['cpu_baseline'])"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r8/launches_train.csv python bench.py --workload train --steps 1 --warmup 2 --no-cpu > gpurun_out/r8/ncu_train.log 2>&1; echo "ncu rc=$?"
python tools/train_step_table.py gpurun_out/r8/launches_train.csv > gpurun_out/r8/train_table.md; head -45 gpurun_out/r8/train_table.md
