mkdir -p gpurun_out/r8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 100 --warmup 3 > gpurun_out/r8/bench_n8.json 2> gpurun_out/r8/bench_n8.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r8/bench_n8.json'));print(d['value'],d['ms_per_step'],d['n_gpus'],d['e2e']['value'],d['e2e_with_mask']['value'],d['throughput_mode']['value'],d['throughput_mode']['e2e']['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --workload eval > gpurun_out/r8/eval_n8.json 2> gpurun_out/r8/eval_n8.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r8/eval_n8.json'));print(d['value'],d['ms_per_step'],d['n_gpus'],d['e2e']['value'],d['metrics'],d['metrics_e2e_equal'])"
