#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -m gpu -x -q 2>&1 | tail -4
run() { # name, env...
  name=$1; shift
  env "$@" timeout 120 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_c10_$name.json
  python -c "import json;d=json.load(open('gpurun_out/bench_c10_$name.json'));print('$name',round(d['value'],1),round(d['ms_per_step'],5),d['launches_per_step'],round(d['e2e']['value'],1),d['e2e']['h2d_gbs'],round(d['roofline']['other_ms_per_step'],4),{k:round(x['ms_per_step'],4) for k,x in d['roofline']['by_kind'].items()})"
}
run jobs1 DSACT_JOBS=1
run jobs0 DSACT_JOBS=0
run c7 DSACT_LIB=$PWD/dsac-v2_b200/libdsact_c7.so
run jobs1b DSACT_JOBS=1
run wg256s8 DSACT_JOBS=1 DSACT_WG_BN=256 DSACT_WG_SLABS=8
run wg128s8 DSACT_JOBS=1 DSACT_WG_SLABS=8
run wg128s3 DSACT_JOBS=1 DSACT_WG_SLABS=3
run wg128s2 DSACT_JOBS=1 DSACT_WG_SLABS=2
DSACT_PDL=0 timeout 120 python tools/trace_step.py > gpurun_out/trace_step_c10.txt 2>/dev/null
timeout 60 python tools/chain_timeline.py humanoid 4096 bf16x3 gelu 2>&1 | sed -n "/step 2/,\$p" | cut -c1-260 | head -50 > gpurun_out/chain_timeline_c10.txt
