run() { echo "$@"; env "$@" python bench.py --scaling weak --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-order22 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['config']['launch']['traj_per_cta'], d['config']['launch']['ctas'])"; }
run A=1
run HY_CUDA_TRAJ_PER_CTA=8 HY_CUDA_CTAS_PER_SM=2
run HY_CUDA_TRAJ_PER_CTA=4 HY_CUDA_CTAS_PER_SM=4
run HY_CUDA_TRAJ_PER_CTA=2 HY_CUDA_CTAS_PER_SM=8
run HY_CUDA_TRAJ_PER_CTA=14
run HY_CUDA_TRAJ_PER_CTA=12
