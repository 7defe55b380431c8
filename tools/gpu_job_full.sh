set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_r1_h.json 2> gpurun_out/bench_r1_h.err
cat gpurun_out/bench_r1_h.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_h_ref.json 2> gpurun_out/bench_r1_h_ref.err
cat gpurun_out/bench_r1_h_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_h.csv python bench.py --steps 5 --warmup 3 --profile-mode > gpurun_out/ncu_h1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bm25_score -s 4 -c 1 -f -o gpurun_out/prof_r1_h python bench.py --steps 3 --warmup 3 --profile-mode > gpurun_out/ncu_h2.log 2>&1
