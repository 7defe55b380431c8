set -x
timeout 240 python -m pytest tests -m gpu -x -q --timeout 60 2>&1 | tail -3
timeout 150 python bench.py --steps 30 --warmup 5 --profile-mode 2>/dev/null | tail -1
