python bench.py --steps 10 --warmup 3 > gpurun_out/r01e_bench.json 2> gpurun_out/r01e_bench.err; tail -c 3500 gpurun_out/r01e_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-tight > gpurun_out/r01e_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:zz_run_kernel -s 3 -c 1 -o gpurun_out/r01e_prof python bench.py --steps 2 --warmup 1 --no-cpu --no-tight > gpurun_out/r01e_prof_bench.log 2>&1
ls -la gpurun_out | grep r01e || true
