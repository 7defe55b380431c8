set -x
ncu --set full --clock-control none --import-source on -k regex:bm25_ -s 4 -c 1 -f -o gpurun_out/prof_cur timeout 300 python bench.py --steps 3 --warmup 3 --profile-mode > gpurun_out/ncu_cur.log 2>&1
