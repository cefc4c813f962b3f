timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
