"""Work counters of one small `shell` run (candidates, ring hits, intersections, stage
times): PYTHONPATH=. python scripts/work_stats.py.  With a library built from an
instrumented sfk_bin.cu that exports sfk_debug_counters, also the evaluated (record,
LoS) pairs and where they are rejected (that is how DESIGN.md's pruning numbers were
measured); the product build does not carry those counters."""
import ctypes as C

from shellfish_b200 import api, synth

halos, blocks, mass = synth.config2(seed=2, n_halos=16, n_total=4000000, box=80.0, cells=2)
params = api.default_params(seed=1337, flags=api.SFK_FLAG_TAPS)
ctx, coeffs, status = api.run_shell(params, mass, halos, blocks)
s = ctx.stats()
print(s)
print("hits/cand", s["n_ring_hits"] / s["n_candidates"], "inter/hit", s["n_intersections"] / s["n_ring_hits"])
if hasattr(api.lib(), "sfk_debug_counters"):
    out = (C.c_ulonglong * 8)()
    api.lib().sfk_debug_counters(out)
    d = list(out)
    print("passes", d[0], "pairs", d[1], "nan-skip", d[2], "early-out", d[3], "groups", d[4], "recs touching", d[5])
    print("pairs/pass", d[1] / d[0], "pairs/group", d[1] / d[4], "passes/group", d[0] / d[4],
          "inter/pair", s["n_intersections"] / d[1])
