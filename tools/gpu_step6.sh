timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 5 --warmup 3 2>&1 | tail -3
