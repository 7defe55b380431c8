# the odd-lexicon GPU tests again, then compute-sanitizer (memcheck, racecheck, synccheck) over tools/sanitizer_probe.py
mkdir -p gpurun_out
t0=$(date +%s)
timeout 40 python -m pytest tests/test_odd_lexicon.py -m gpu -q --timeout 30 -p no:cacheprovider > gpurun_out/final_tests_3.log 2>&1
echo "rc=$? after $(( $(date +%s) - t0 ))s" >> gpurun_out/final_tests_3.log; tail -4 gpurun_out/final_tests_3.log
timeout 20 python tools/sanitizer_probe.py > gpurun_out/sanitizer_plain.log 2>&1
echo "rc=$? after $(( $(date +%s) - t0 ))s" >> gpurun_out/sanitizer_plain.log; tail -3 gpurun_out/sanitizer_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 45 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_probe.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "rc=$? after $(( $(date +%s) - t0 ))s" >> gpurun_out/sanitizer_$tool.log
  tail -6 gpurun_out/sanitizer_$tool.log
done
