set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --profile-mode 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:bm25_score -s 4 -c 1 -f -o gpurun_out/prof_cur python bench.py --steps 3 --warmup 3 --profile-mode > gpurun_out/ncu_cur.log 2>&1
