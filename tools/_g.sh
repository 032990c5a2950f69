for c in 0 4 5; do
  echo "== YSP_DLC32_CFG=$c"
  YSP_DLC32_CFG=$c timeout 300 python tools/profile_layers.py tc32 256 2>&1 | grep -E "dlc32|sum of" 
done
YSP_DLC32_CFG=4 timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -2
