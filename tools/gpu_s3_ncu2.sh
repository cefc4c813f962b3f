ncu --set full --clock-control none --import-source on -k regex:zz_run_kernel -s 3 -c 1 -o gpurun_out/r01c_prof python bench.py --steps 2 --warmup 1 --no-cpu --no-tight > gpurun_out/r01c_prof_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-tight > gpurun_out/r01c_launches_bench.log 2>&1
ls -la gpurun_out | tail -5
