#!/bin/bash
# One gpurun call that gathers everything a round is judged on, cheapest first, so
# that a cut-off call still leaves the earlier artefacts in gpurun_out/:
#   1. pytest -m gpu                      -> gpurun_out/<tag>_gpu_tests.log
#   2. python bench.py (64-halo sample)   -> gpurun_out/<tag>_bench64.json
#   3. ncu launch list of the same command-> gpurun_out/<tag>_launches_bench64.csv
#   4. ncu --set full of ONE bin_kernel   -> gpurun_out/<tag>_bin.ncu-rep (+ by-line summary)
#   5. python bench.py (full configs[1])  -> gpurun_out/<tag>_bench_full.json
# Usage:  gpurun --timeout 900 -- 'bash scripts/collect_evidence.sh r02a [full]'
# Numbers printed under ncu (steps 3, 4) are never bench values.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
B64="python bench.py --halos 64 --n-total 16000000 --box 125 --cells 4 --steps 2 --warmup 3 --no-cpu"

timeout 300 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1
echo "tests rc=$? $(tail -1 $OUT/${TAG}_gpu_tests.log)"

timeout 200 $B64 > $OUT/${TAG}_bench64.json 2> $OUT/${TAG}_bench64.err
echo "bench64 rc=$? $(cut -c1-200 $OUT/${TAG}_bench64.json)"

timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file $OUT/${TAG}_launches_bench64.csv $B64 > $OUT/${TAG}_ncu_launch.log 2>&1
echo "launch list rc=$? $(wc -l < $OUT/${TAG}_launches_bench64.csv) lines"

# one launch of the dominant kernel, after the warm-up launches
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bin_kernel -s 3 -c 1 -f \
    -o $OUT/${TAG}_bin $B64 > $OUT/${TAG}_ncu_bin.log 2>&1
echo "ncu full rc=$?"
if [ -f $OUT/${TAG}_bin.ncu-rep ]; then
    python profiles/ncu_by_line.py $OUT/${TAG}_bin.ncu-rep sfk_bin bin_kernelILb0 > $OUT/${TAG}_bin_kernel_by_line.txt 2>&1 || true
fi

# A/B of the experimental sliding-moment los kernel on the same 64-halo sample (stage_ms_per_step.ms_los)
timeout 200 $B64 --no-e2e --los-moments > $OUT/${TAG}_bench64_los_moments.json 2> $OUT/${TAG}_bench64_los_moments.err
python - <<PY
import json
try:
    a = json.loads(open("$OUT/${TAG}_bench64.json").read().strip().splitlines()[-1])
    b = json.loads(open("$OUT/${TAG}_bench64_los_moments.json").read().strip().splitlines()[-1])
    print("los A/B: ms_los %.2f -> %.2f, halos/s %.1f -> %.1f" % (a["stage_ms_per_step"]["ms_los"],
          b["stage_ms_per_step"]["ms_los"], a["value"], b["value"]))
except Exception as e:
    print("los A/B: no comparison (%s)" % e)
PY

if [ "${2:-}" = "full" ]; then
    timeout 600 python bench.py > $OUT/${TAG}_bench_full.json 2> $OUT/${TAG}_bench_full.err
    echo "bench full rc=$? $(cut -c1-200 $OUT/${TAG}_bench_full.json)"
fi
