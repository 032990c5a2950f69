mkdir -p gpurun_out/r8
for v in "A=1" "YSP_TC32_ONE_GROUP=1"; do
  echo "== $v"
  env $v timeout 300 python tools/profile_layers.py tc32 256 > gpurun_out/r8/layers_$v.md 2>&1
  grep -E "sum of" gpurun_out/r8/layers_$v.md
  python - <<PY
import re
k={}
for l in open("gpurun_out/r8/layers_$v.md"):
    m=l.split('|')
    if len(m)>5:
        try: k.setdefault(m[2].strip(),[0,0.]); k[m[2].strip()][0]+=1; k[m[2].strip()][1]+=float(m[3])
        except: pass
for a,b in sorted(k.items(),key=lambda x:-x[1][1])[:6]: print(a,b)
PY
done
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -2
