#!/bin/bash
# A/B of programmatic dependent launch on the bench step
for p in 0 1; do
  DSACT_PDL=$p python bench.py --steps 200 --warmup 20 2>/dev/null > /tmp/pdl_$p.json
  python - <<PY
import json
d = json.load(open("/tmp/pdl_$p.json"))
print("pdl=$p", d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
done
