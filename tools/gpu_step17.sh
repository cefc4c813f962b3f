timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('N=1 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'], 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'])"
