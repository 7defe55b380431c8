set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1_g.json 2> gpurun_out/bench_r1_g.err
cat gpurun_out/bench_r1_g.json
NSB200_TRACE=1 python tools/e2e_probe.py > gpurun_out/e2e_trace.log 2>&1
tail -20 gpurun_out/e2e_trace.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_g.csv python bench.py --steps 5 --warmup 3 --profile-mode > gpurun_out/ncu_g1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bm25_score -s 4 -c 1 -f -o gpurun_out/prof_r1_g python bench.py --steps 3 --warmup 3 --profile-mode > gpurun_out/ncu_g2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:impact -s 4 -c 1 -f -o gpurun_out/prof_r1_g_impact python bench.py --steps 3 --warmup 3 --profile-mode > gpurun_out/ncu_g3.log 2>&1
ls -la gpurun_out
