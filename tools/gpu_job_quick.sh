set -x
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python bench.py --steps 30 --warmup 5 --profile-mode 2>/dev/null
