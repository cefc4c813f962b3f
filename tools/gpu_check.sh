set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; tail -c 3000 gpurun_out/s3_bench.json
python bench.py --impl reference --steps 2 --warmup 1 --cpu-T 0.5 > gpurun_out/s3_ref.json 2>&1; cat gpurun_out/s3_ref.json
nproc
