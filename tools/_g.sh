set -x
mkdir -p gpurun_out/r7
timeout 300 python tools/gpu_debug.py tc32 > gpurun_out/r7/debug_tc32b.log 2>&1; echo "debug rc=$?"
grep -E "raw|y box|seg out|model.20 " gpurun_out/r7/debug_tc32b.log
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -s > gpurun_out/r7/pytest_model.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r7/pytest_model.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r7/bench_a.json 2> gpurun_out/r7/bench_a.err; echo "bench rc=$?"
tail -3 gpurun_out/r7/bench_a.err
python -c "
import json;d=json.load(open('gpurun_out/r7/bench_a.json'))
for k in ('value','ms_per_step','dtype','e2e','e2e_with_mask','parity','step_ms','library_baseline','cpu_baseline'): print(k, d.get(k))
t=d['throughput_mode']; print('bf16', t['value'], t['e2e'], t['parity'])
print(d['roofline'])
"
