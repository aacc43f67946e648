mkdir -p gpurun_out
for rep in 1 2; do
for wl in humanoid_8192 humanoid_512k; do
  python bench.py --workload $wl --steps 10 --no-extra --no-cpu-baseline 2>> gpurun_out/v4c.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('main(v1 const)', d['config']['workload'], round(d['value']))"
  BXG_FORCE_VARIANT=4 BXG_LIB=brax_b200/libbxg_v4c.so python bench.py --workload $wl --steps 10 --no-extra --no-cpu-baseline 2>> gpurun_out/v4c.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('v4 const', d['config']['workload'], round(d['value']))"
done; done
BXG_FORCE_VARIANT=4 BXG_LIB=brax_b200/libbxg_v4c.so python - <<'PY'
import torch
from brax_b200 import native, workloads
s,q,qd=workloads.reset('humanoid',0,64,0,torch.device('cuda',0))
nm=native.model_for(s,0); print('kernel id', nm.kernel_id, nm.launch_shape(8192))
PY
tail -2 gpurun_out/v4c.err
