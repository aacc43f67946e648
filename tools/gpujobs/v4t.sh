mkdir -p gpurun_out
for fv in 1 4; do BXG_FORCE_VARIANT=$fv BXG_LIB=brax_b200/libbxg_timers.so python tools/phase_timers.py humanoid 23680 2>> gpurun_out/v4.err | python -c "
import json,sys; d=json.load(sys.stdin); print('variant $fv', d['launch'], 'total', d['cycles_per_warp_substep_total']); print({k: v for k, v in d['cycles_per_warp_substep'].items()})"; done
