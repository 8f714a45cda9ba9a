#!/bin/bash
# wgrad tile / slab knobs
for cfg in "512 128" "1024 128" "512 256" "1024 256"; do
  set -- $cfg
  DSACT_SLAB_ROWS=$1 DSACT_WG_BN=$2 python bench.py --steps 200 --warmup 20 2>/dev/null > /tmp/ab.json
  python - <<PY
import json
d = json.load(open("/tmp/ab.json"))
print("slab_rows=$1 wg_bn=$2", round(d["value"], 1), round(d["ms_per_step"], 5), round(d["e2e"]["value"], 1))
PY
done
