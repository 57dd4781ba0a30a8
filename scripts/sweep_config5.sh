#!/bin/bash
# BASELINE configs[4] (SURVEY "config 5"): Order x SubsampleFactor on config 2's box (64-halo sample),
# every combination checked against the oracle on its first 2 halos (bench.py's `parity` key).
# Usage (on a GPU box): bash scripts/sweep_config5.sh > gpurun_out/sweep_config5.jsonl
for order in 2 3 4; do
  for sub in 1 2; do
    timeout 300 python bench.py --halos 64 --n-total 16000000 --box 125 --cells 4 --steps 3 --warmup 3 \
        --cpu-halos 2 --no-e2e --order $order --subsample $sub 2>/dev/null | tail -1
  done
done
