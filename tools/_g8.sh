mkdir -p gpurun_out/r8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 100 --warmup 3 > gpurun_out/r8/bench_n2.json 2> gpurun_out/r8/bench_n2.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r8/bench_n2.json'));print(d['value'],d['ms_per_step'],d['n_gpus'],d['e2e']['value'],d['throughput_mode']['value'],d['parity']['logits_max_abs'])"
tail -2 gpurun_out/r8/bench_n2.err
