python - <<'PY'
import torch
from brax_b200 import native, workloads
for m in ('ant','humanoid','humanoid_falls'):
    s,q,qd=workloads.reset(m,0,64,0,torch.device('cuda',0))
    nm=native.model_for(s,0)
    print(m, 'plan', native.plan(s)['kernel_id'], 'runtime', nm.kernel_id)
PY
for wl in humanoid_8192; do for e in "" "BXG_NO_SPECIALISE=1"; do env $e python bench.py --workload $wl --steps 10 --no-extra --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$e', d['config']['workload'], round(d['value']))"; done; done
