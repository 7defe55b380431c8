# last-minutes GPU check of the round: changed / new tests first, then the rest of the GPU suite, then smoke()
mkdir -p gpurun_out
t0=$(date +%s)
timeout 90 python -m pytest tests/test_odd_lexicon.py tests/test_reload_cases.py tests/test_gpu_parity.py tests/test_gpu_engine_multi.py \
    -m gpu -q --timeout 60 -p no:cacheprovider > gpurun_out/final_tests_1.log 2>&1
echo "rc=$? after $(( $(date +%s) - t0 ))s" >> gpurun_out/final_tests_1.log; tail -25 gpurun_out/final_tests_1.log
timeout 80 python -m pytest tests/test_gpu_exchange.py tests/test_gpu_semantic.py tests/test_gpu_dist_ipc.py \
    -m gpu -q --timeout 60 -p no:cacheprovider > gpurun_out/final_tests_2.log 2>&1
echo "rc=$? after $(( $(date +%s) - t0 ))s" >> gpurun_out/final_tests_2.log; tail -8 gpurun_out/final_tests_2.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "rc=$? after $(( $(date +%s) - t0 ))s" >> gpurun_out/final_smoke.log; tail -3 gpurun_out/final_smoke.log
