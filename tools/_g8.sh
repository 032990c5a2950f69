mkdir -p gpurun_out/r8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --workload train --gpus 8 --steps 20 --warmup 3 > gpurun_out/r8/train_n8.json 2> gpurun_out/r8/train_n8.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r8/train_n8.json'));print(d['value'],d['ms_per_step'],d['n_gpus'],d['loss'],d['clocks'])"
tail -3 gpurun_out/r8/train_n8.err
