mkdir -p gpurun_out/r8
timeout 900 python -m pytest tests/test_gpu_eval.py -x -q -m gpu > gpurun_out/r8/pytest_eval.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r8/pytest_eval.log
timeout 900 python bench.py --steps 40 --no-library --no-cpu > gpurun_out/r8/bench_e2.json 2> gpurun_out/r8/bench_e2.err; echo "bench rc=$?"; tail -3 gpurun_out/r8/bench_e2.err
python -c "
import json;d=json.load(open('gpurun_out/r8/bench_e2.json'))
for k in ('value','ms_per_step','e2e','e2e_with_mask','parity','step_ms','gpu_launches'): print(k, d.get(k))
t=d['throughput_mode']; print('bf16', t['value'], t['e2e']['value'], t['parity']['logits_max_abs'])
print(d['config']['engines'])
"
timeout 900 python bench.py --steps 40 --no-library --no-cpu --engines 1 --single-mode 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read()); print('engines=1', d['value'], d['e2e']['value'])"
timeout 900 python bench.py --workload eval 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read()); print('eval', d['value'], d['ms_per_step'], d['e2e']['value'], d['metrics'], d['metrics_e2e_equal'])"
timeout 900 python bench.py --workload eval --engines 1 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read()); print('eval engines=1', d['value'], d['ms_per_step'], d['e2e']['value'], d['metrics'], d['metrics_e2e_equal'])"
