mkdir -p gpurun_out/r8
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_golden.py tests/test_gpu_eval.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 60 --no-library --no-cpu > gpurun_out/r8/bench_f.json 2> gpurun_out/r8/bench_f.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r8/bench_f.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['logits_max_abs'], d['parity']['nms_keep_equal'])
t=d['throughput_mode']; print('bf16', t['value'], t['e2e']['value'])"
