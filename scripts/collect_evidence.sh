#!/bin/bash
# One gpurun call that gathers everything a round is judged on, cheapest first, so
# that a cut-off call still leaves the earlier artefacts in gpurun_out/:
#   1. pytest -m gpu                      -> gpurun_out/<tag>_gpu_tests.log
#   2. python bench.py (64-halo sample)   -> gpurun_out/<tag>_bench64.json
#   3. ncu launch list of the same command-> gpurun_out/<tag>_launches_bench64.csv
#   4. ncu --set full of ONE bin_kernel   -> gpurun_out/<tag>_bin.ncu-rep (+ by-line summary)
#   5. ncu --set full of the other kernels-> gpurun_out/<tag>_others.ncu-rep (+ raw csv)
#   6. python bench.py (full configs[1])  -> gpurun_out/<tag>_bench_full.json
# Usage:  gpurun --timeout 1200 -- 'bash scripts/collect_evidence.sh r02a [full]'
# Numbers printed under ncu (steps 3-5) are never bench values.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
B64="python bench.py --halos 64 --n-total 16000000 --box 125 --cells 4 --steps 2 --warmup 3 --no-cpu"

timeout 400 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1
echo "tests rc=$? $(tail -1 $OUT/${TAG}_gpu_tests.log)"

timeout 200 $B64 > $OUT/${TAG}_bench64.json 2> $OUT/${TAG}_bench64.err
echo "bench64 rc=$? $(cut -c1-200 $OUT/${TAG}_bench64.json)"

timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file $OUT/${TAG}_launches_bench64.csv $B64 --no-e2e > $OUT/${TAG}_ncu_launch.log 2>&1
echo "launch list rc=$? $(wc -l < $OUT/${TAG}_launches_bench64.csv) lines"

# one launch of the dominant kernel, after the warm-up launches
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bin_kernel -s 3 -c 1 -f \
    -o $OUT/${TAG}_bin $B64 --no-e2e > $OUT/${TAG}_ncu_bin.log 2>&1
echo "ncu full (bin) rc=$?"
if [ -f $OUT/${TAG}_bin.ncu-rep ]; then
    python profiles/ncu_by_line.py $OUT/${TAG}_bin.ncu-rep sfk_bin bin_kernelILb0 > $OUT/${TAG}_bin_kernel_by_line.txt 2>&1 || true
    cp profiles/ncu_bin_constants.json $OUT/${TAG}_ncu_bin_constants_before.json
    cp $OUT/${TAG}_bin_kernel_by_line.txt profiles/${TAG}_bin_kernel_by_line.txt
    python profiles/make_ncu_constants.py profiles/${TAG}_bin_kernel_by_line.txt $OUT/${TAG}_bench64.json > $OUT/${TAG}_ncu_constants.log 2>&1 \
        && cp profiles/ncu_bin_constants.json $OUT/${TAG}_ncu_bin_constants.json || true
fi

# the other kernels of one batch: hits, los, filter, the four fit kernels
timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:'hits_kernel|hit_count_kernel|los_kernel|filter_kernel|fit_rows|fit_solve|fit_apply|fit_finish' -c 9 -f \
    -o $OUT/${TAG}_others $B64 --no-e2e > $OUT/${TAG}_ncu_others.log 2>&1
echo "ncu full (others) rc=$?"
if [ -f $OUT/${TAG}_others.ncu-rep ]; then
    ncu -i $OUT/${TAG}_others.ncu-rep --page raw --csv > $OUT/${TAG}_others_raw.csv 2>/dev/null || true
fi

if [ "${2:-}" = "full" ]; then
    timeout 600 python bench.py > $OUT/${TAG}_bench_full.json 2> $OUT/${TAG}_bench_full.err
    echo "bench full rc=$? $(cut -c1-200 $OUT/${TAG}_bench_full.json)"
fi
if [ "${3:-}" = "taploop" ]; then
    timeout 400 python bench.py --no-cpu --no-e2e --los-taploop > $OUT/${TAG}_bench_full_los_taploop.json 2> $OUT/${TAG}_bench_full_los_taploop.err
    echo "bench full (los tap loop) rc=$? $(cut -c1-200 $OUT/${TAG}_bench_full_los_taploop.json)"
fi
