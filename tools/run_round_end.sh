#!/bin/bash
# What the driver runs at round end, plus the per-config bench lines and the profiles (one GPU).
export HY_CUDA_JIT_CACHE=$PWD/gpurun_out/jit_cache_final
mkdir -p $HY_CUDA_JIT_CACHE; cp heyoka.py_b200/csrc/jit_cache/*.hyjit $HY_CUDA_JIT_CACHE/ 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/r02_smoke.log
python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref rc $?"
for c in 2 3 4 5; do
  python bench.py --config $c > gpurun_out/r02_bench_cfg$c.json 2> gpurun_out/r02_bench_cfg$c.err; echo "cfg$c rc $?"; tail -c 300 gpurun_out/r02_bench_cfg$c.err
done
# launch list of the default bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-order22 > gpurun_out/r02_ncu_launches_bench.log 2>&1
echo "launch list rc $?"
