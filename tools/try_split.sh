#!/bin/bash
# First hardware run of the split chain kernel (DESIGN.md §7 "Next" item 1).  Builds the library with the barrier
# watchdog, runs the bf16x3 parity tests and a short bench with DSACT_CHAIN_SPLIT=1, everything under `timeout`.
#   gpurun --timeout 600 -- 'bash tools/try_split.sh'
set -x
rm -f dsac-v2_b200/libdsact.so
DSACT_NVCC_FLAGS="-DDSACT_MBAR_GUARD" python -c "import __graft_entry__ as g; g.build()" || exit 1
DSACT_CHAIN_SPLIT=1 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bf16x3 or tensor_core" 2>&1 | tail -5
DSACT_CHAIN_SPLIT=1 timeout 60 python tools/chain_timeline.py 2>&1 | sed -n "/step 2/,\$p" | cut -c1-330 | head -8
rm -f dsac-v2_b200/libdsact.so
python -c "import __graft_entry__ as g; g.build()"       # production build (no watchdog) for the timing A/B
for v in 0 1; do
  DSACT_CHAIN_SPLIT=$v timeout 120 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null > /tmp/split_$v.json
  python - <<PY
import json
d = json.load(open("/tmp/split_$v.json"))
print("split=$v", round(d["value"], 1), round(d["ms_per_step"], 5))
PY
done
