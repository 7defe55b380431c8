# A/B of an experimental library against the product library: alternating runs of tools/ab_probe.py, then the parity tests on the experiment
mkdir -p gpurun_out
EXP=$PWD/nextsearch-api_b200/libnsb200_exp_${1:-hdr}.so
BASE=$PWD/nextsearch-api_b200/libnsb200.so
: > gpurun_out/ab_${1:-hdr}.jsonl
for r in 1 2; do
  for L in $BASE $EXP; do
    NSB200_LIB=$L AB_PARITY=$([ $r = 1 ] && echo 1 || echo 0) timeout 60 python tools/ab_probe.py 6 1 >> gpurun_out/ab_${1:-hdr}.jsonl 2>> gpurun_out/ab_${1:-hdr}.err
  done
done
for L in $BASE $EXP; do NSB200_LIB=$L AB_PARITY=0 timeout 60 python tools/ab_probe.py 6 8 >> gpurun_out/ab_${1:-hdr}.jsonl 2>> gpurun_out/ab_${1:-hdr}.err; done
cat gpurun_out/ab_${1:-hdr}.jsonl; tail -3 gpurun_out/ab_${1:-hdr}.err
NSB200_LIB=$EXP timeout 60 python -m pytest tests/test_gpu_parity.py tests/test_odd_lexicon.py -m gpu -q -x --timeout 50 -p no:cacheprovider 2>&1 | tail -3
