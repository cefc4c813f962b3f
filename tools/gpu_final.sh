# final regression run of the whole GPU suite (log kept under profiles/)
set -x
timeout 70 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
