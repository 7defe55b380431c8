# the path against the debug library (index checks compiled into the score kernel): probe output, then the pytest form
mkdir -p gpurun_out
t0=$(date +%s)
NSB200_LIB=$PWD/nextsearch-api_b200/libnsb200_dbg.so timeout 100 python tools/debug_checks_probe.py > gpurun_out/debug_checks_probe.json 2> gpurun_out/debug_checks_probe.err
echo "rc=$? after $(( $(date +%s) - t0 ))s"; cat gpurun_out/debug_checks_probe.json; tail -5 gpurun_out/debug_checks_probe.err
timeout 100 python -m pytest tests/test_gpu_debug_checks.py -m gpu -q --timeout 90 -p no:cacheprovider > gpurun_out/final_tests_4.log 2>&1
echo "rc=$? after $(( $(date +%s) - t0 ))s" >> gpurun_out/final_tests_4.log; tail -12 gpurun_out/final_tests_4.log
