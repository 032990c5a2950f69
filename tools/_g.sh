mkdir -p gpurun_out/r8
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r8/pytest_all.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r8/pytest_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r8/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r8/smoke.log
timeout 900 python bench.py > gpurun_out/r8/bench_n1.json 2> gpurun_out/r8/bench_n1.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r8/bench_n1.json'))
for k in ('value','ms_per_step','step_ms','gpu_launches'): print(k, d.get(k))
print('e2e', d['e2e']['value'], d['e2e_with_mask']['value'], d['parity']['logits_max_abs'], d['parity']['dice_max_abs'], d['parity']['nms_keep_equal'])
t=d['throughput_mode']; print('bf16', t['value'], t['e2e']['value'])
print(d['roofline']['frac'], d['roofline']['us_per_launch'], d['roofline']['kernel']); print([ (k['kernel'],round(k['ms_per_step'],3)) for k in d['top_kernels']])
"
timeout 300 python tools/profile_layers.py tc32 256 > gpurun_out/r8/layers_tc32.md 2>&1; head -12 gpurun_out/r8/layers_tc32.md
