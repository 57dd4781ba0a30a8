"""Quick A/B of the two los kernels on one default-config halo (100 rings x 256 LoS x 256 bins):
the library's own CUDA-event stage timer (sfk_get_stats), no torch.
    PYTHONPATH=. python scripts/time_los_variants.py"""
import json
import sys

sys.path.insert(0, ".")
from shellfish_b200 import api, synth  # noqa: E402

halos, blocks, mass = synth.single_halo(seed=3, n=20000, n_background=5000)
out = {}
for name, flags in (("taps", api.SFK_FLAG_LOS_TAPLOOP), ("moments", 0)):
    best = None
    for _ in range(3):
        ctx, coeffs, status = api.run_shell(api.default_params(seed=1337, flags=flags), mass, halos, blocks)
        ms = ctx.stats()["ms_los"]
        best = ms if best is None else min(best, ms)
    out[name] = {"ms_los": best, "c0": float(coeffs[0, 0])}
print(json.dumps(out))
