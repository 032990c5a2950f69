mkdir -p gpurun_out/r8
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r8/pytest_all.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r8/pytest_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r8/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r8/smoke.log
timeout 900 python bench.py > gpurun_out/r8/bench_n1.json 2> gpurun_out/r8/bench_n1.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r8/bench_n1.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['logits_max_abs'], d['parity']['nms_keep_equal'], d['throughput_mode']['value'])"
