ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_v12.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01_launches_v12_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -c 1 -o gpurun_out/r01_prof_bench_v12 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --traj-per-gpu 125000 --horizon 100 > gpurun_out/r01_prof_bench_v12.log 2>&1
tail -2 gpurun_out/r01_launches_v12.csv | cut -c1-200
