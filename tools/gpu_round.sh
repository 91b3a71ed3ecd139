#!/bin/bash
# one GPU-box visit: every -m gpu test file in its own process (a faulting kernel must not poison the others), then the bench
mkdir -p gpurun_out
TAG=${1:-r2}
for f in tests/test_*.py; do
  if grep -q "pytest.mark.gpu" "$f"; then
    echo "=== $f" >> gpurun_out/tests_$TAG.log
    timeout 1500 python -m pytest "$f" -m gpu -q -x --timeout 1400 2>&1 | tail -25 >> gpurun_out/tests_$TAG.log
  fi
done
if [ -z "$NO_BENCH" ]; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
  timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
fi
tail -5 gpurun_out/tests_$TAG.log
