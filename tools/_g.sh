mkdir -p gpurun_out/r8
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r8/launches_tc32.csv python bench.py --steps 2 --warmup 1 --single-mode --no-library --no-cpu --parity-slices 0 > gpurun_out/r8/ncu_bench.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r8/ncu_bench.log | cut -c1-300
grep -c mask_dice gpurun_out/r8/launches_tc32.csv
