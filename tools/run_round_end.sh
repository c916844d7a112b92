#!/bin/bash
# Round-end measurements on one B200 (run through gpurun): GPU tests, bench (both arms), ncu launch list,
# ncu full captures of the N-body and CR3BP register kernels.  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py > gpurun_out/bench_r1_g.json 2> gpurun_out/bench_r1_g.err; tail -c 600 gpurun_out/bench_r1_g.json
python bench.py --impl reference > gpurun_out/bench_r1_g_ref.json 2>/dev/null; tail -c 300 gpurun_out/bench_r1_g_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_v11.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01_launches_v11_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -c 1 -o gpurun_out/r01_prof_bench_v11 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --traj-per-gpu 125000 --horizon 100 > gpurun_out/r01_prof_bench_v11.log 2>&1
QSKIP_INTERP=1 QB=1000000 ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -c 3 -o gpurun_out/r01_prof_cr3bp_reg -f python tools/gpu_cr3bp_perf.py > gpurun_out/r01_prof_cr3bp_reg.log 2>&1
ls -la gpurun_out | tail -8
