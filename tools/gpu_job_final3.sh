# after the long-query split in the engine: parity file, engine tests, debug-library test, smoke()
mkdir -p gpurun_out
t0=$(date +%s)
timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine_multi.py tests/test_gpu_debug_checks.py tests/test_odd_lexicon.py \
    -m gpu -q --timeout 90 -p no:cacheprovider > gpurun_out/final_tests_6.log 2>&1
echo "rc=$? after $(( $(date +%s) - t0 ))s" >> gpurun_out/final_tests_6.log; tail -12 gpurun_out/final_tests_6.log
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke3.log 2>&1
echo "rc=$? after $(( $(date +%s) - t0 ))s" >> gpurun_out/final_smoke3.log; tail -2 gpurun_out/final_smoke3.log
