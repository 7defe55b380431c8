python -m pytest tests -m gpu -x -q 2>&1 | tail -3
NSB200_TRACE=1 python tools/e2e_probe.py 2>&1 | tail -16
for t in 1 4 8; do echo "host threads $t"; NSB200_HOST_THREADS=$t NSB200_TRACE=1 python tools/e2e_probe.py 2>&1 | grep resolve | tail -2; done
