# usage: bash tools/gpu_ncu.sh <tag>   (env NSB200_* is passed through)
set -x
TAG=${1:-cur}
timeout 150 python bench.py --steps 30 --warmup 5 --profile-mode 2>/dev/null | tail -1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bm25_ -s 4 -c 1 -f -o gpurun_out/prof_${TAG} python bench.py --steps 3 --warmup 3 --profile-mode > gpurun_out/ncu_${TAG}.log 2>&1
ls -la gpurun_out/prof_${TAG}.ncu-rep
