# re-check after the last host-side changes: debug-library test (rebuilt with -split-compile), engine tests, smoke()
mkdir -p gpurun_out
t0=$(date +%s)
timeout 100 python -m pytest tests/test_gpu_debug_checks.py tests/test_odd_lexicon.py tests/test_reload_cases.py tests/test_gpu_engine_multi.py \
    -m gpu -q --timeout 90 -p no:cacheprovider > gpurun_out/final_tests_5.log 2>&1
echo "rc=$? after $(( $(date +%s) - t0 ))s" >> gpurun_out/final_tests_5.log; tail -6 gpurun_out/final_tests_5.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke2.log 2>&1
echo "rc=$? after $(( $(date +%s) - t0 ))s" >> gpurun_out/final_smoke2.log; tail -2 gpurun_out/final_smoke2.log
