#!/bin/bash
# A/B of library builds on one box: for each variants/libsfk_<name>.so, run the 64-halo bench
# and print the stage times.  Usage: gpurun -- 'bash scripts/experiments/ab_variants.sh tag v0 old ...'
TAG=$1; shift
cp shellfish_b200/libsfk.so /tmp/libsfk_keep.so
for v in "$@"; do
  cp variants/libsfk_$v.so shellfish_b200/libsfk.so
  timeout 200 python bench.py --halos 64 --n-total 16000000 --box 125 --cells 4 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$v.json")); print("$v", round(d["value"],1), {k:round(x,2) for k,x in d["stage_ms_per_step"].items()}, d.get("halos_ok"))
except Exception as e:
    print("$v failed", e, open("gpurun_out/${TAG}_$v.err").read()[-400:])
PY
done
cp /tmp/libsfk_keep.so shellfish_b200/libsfk.so
