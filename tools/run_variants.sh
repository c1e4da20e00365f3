#!/bin/bash
# On the GPU box: run "python bench.py <args>" once per variant library, restoring the default library afterwards.
# usage: tools/run_variants.sh "<bench args>" name1 name2 ...   (results: gpurun_out/var_<name>.json)
cd "$(dirname "$0")/.."
L=hopfield_network_go_b200/lib
cp $L/libhopfield_b200.so /tmp/default_lib.so
args=$1; shift
for v in "$@"; do
  cp $L/variants/$v.so $L/libhopfield_b200.so
  python bench.py $args > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  if [ -n "$VERIFY" ]; then python -m pytest $VERIFY -q 2>&1 | tail -1; fi
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/var_{v}.json").read().strip().splitlines()[-1])
    print(v, "value %.0f" % d["value"], "relax_ms %.2f" % d["stage_ms_per_step"]["relax"], d["config"]["column_stream"], d["clocks"]["sm_mhz"])
except Exception as e:
    print(v, "FAILED", e)
PY
done
cp /tmp/default_lib.so $L/libhopfield_b200.so
