#!/bin/bash
# usage: tools/ab.sh VAR v1 v2 ...   -- bench.py under each value of an environment knob
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 200 --warmup 20 2>/dev/null > /tmp/ab_$v.json
  python - <<PY
import json
d = json.load(open("/tmp/ab_$v.json"))
print("$var=$v", round(d["value"], 1), round(d["ms_per_step"], 5), round(d["e2e"]["value"], 1))
PY
done
