#!/bin/bash
# BASELINE config 5: Order x SubsampleFactor on config 2's box (64-halo sample).
# Usage (on a GPU box): bash scripts/sweep_config5.sh > gpurun_out/sweep_config5.jsonl
for order in 2 3 4; do
  for sub in 1 2; do
    timeout 200 python bench.py --halos 64 --n-total 16000000 --box 125 --cells 4 --steps 3 --warmup 3 \
        --no-cpu --no-e2e --order $order --subsample $sub 2>/dev/null | tail -1
  done
done
