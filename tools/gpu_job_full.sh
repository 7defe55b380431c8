# tests, full bench (ours + reference arm), ncu launch list, one full ncu capture of the score kernel
set -x
TAG=${1:-r1_v9}
timeout 300 python -m pytest tests -m gpu -x -q --timeout 90 2>&1 | tail -3
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
cat gpurun_out/bench_${TAG}.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
cat gpurun_out/bench_${TAG}_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 5 --warmup 3 --profile-mode > gpurun_out/ncu_${TAG}_1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bm25_score -s 4 -c 1 -f -o gpurun_out/prof_${TAG} python bench.py --steps 3 --warmup 3 --profile-mode > gpurun_out/ncu_${TAG}_2.log 2>&1
ls -la gpurun_out/prof_${TAG}.ncu-rep
