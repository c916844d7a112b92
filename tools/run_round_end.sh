#!/bin/bash
# What the driver runs at round end, plus the per-config bench lines and the profiles (one GPU).
# The kernel cache starts from what build() warmed; everything compiled on the box is kept
# (gpurun_out/jit_cache_final) so that it can be copied back into csrc/jit_cache.
export HY_CUDA_JIT_CACHE=$PWD/gpurun_out/jit_cache_final
rm -rf $HY_CUDA_JIT_CACHE; mkdir -p $HY_CUDA_JIT_CACHE; cp heyoka.py_b200/csrc/jit_cache/*.hyjit $HY_CUDA_JIT_CACHE/ 2>/dev/null
python -m pytest tests -m gpu -q > gpurun_out/r02_gputest.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r02_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc $?"
python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref rc $?"
for c in 2 3 4 5; do
  python bench.py --config $c > gpurun_out/r02_bench_cfg$c.json 2> gpurun_out/r02_bench_cfg$c.err; echo "cfg$c rc $?"; tail -c 200 gpurun_out/r02_bench_cfg$c.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-order22 > gpurun_out/r02_ncu_launches_bench.log 2>&1
echo "launch list rc $?"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_all.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?"; tail -1 gpurun_out/r02_sanitizer_$tool.log
done
