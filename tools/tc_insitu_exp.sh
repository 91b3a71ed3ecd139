#!/bin/bash
# in-situ bottleneck experiments on the tensor-core path: rebuild the library with an experiment switch (wrong results on purpose),
# time one 65 536-pose projection step, restore the product build.  Run under gpurun.
for v in "-DPNDF_TC_EXP_NO_EPI" "-DPNDF_TC_EXP_NO_DRAIN" "-DPNDF_TC_EXP_HALF_FEED"; do
  echo "=== $v"
  PNDF_NVCC_EXTRA="$v" python -c "import __graft_entry__ as g; g.build(force=True)" > /dev/null 2>&1
  PNDF_TILE=128 python tools/quick_bench.py lrelu 65536 2>&1 | tail -2
  python tools/tc_profile.py lrelu 65536 2>&1 | grep "us " | awk '{printf "%s ", $1} END {print ""}'
done
python -c "import __graft_entry__ as g; g.build(force=True)" > /dev/null 2>&1
