python bench.py --steps 10 --warmup 3 > gpurun_out/s3_bench2.json 2> gpurun_out/s3_bench2.err; tail -c 4000 gpurun_out/s3_bench2.json; tail -3 gpurun_out/s3_bench2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s3_ref2.json 2>&1; cat gpurun_out/s3_ref2.json
