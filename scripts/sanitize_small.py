#!/usr/bin/env python
"""A tiny end-to-end pass over every kernel path; also the driver for compute-sanitizer where the
tool is available (it is closed on the build pool):
   compute-sanitizer --tool memcheck python scripts/sanitize_small.py
Covers the default path (TMA cull, two-ended hit queue, plain-store bin flush, moments los,
filter, fit), the tap instantiation, a ragged configuration (generic los kernel, RED flush) and
the stats passes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shellfish_b200 import api, synth  # noqa: E402

halos, blocks, mass = synth.make_box(5, 20.0, 6, 1500, 3000, 2, r200m_range=(0.6, 0.9))
for kw in (dict(rings=4), dict(rings=3, flags=api.SFK_FLAG_TAPS), dict(rings=3, spokes=100, radial_bins=200, smoothing_window=61, levels=2),
           dict(rings=3, flags=api.SFK_FLAG_LOS_TAPLOOP | api.SFK_FLAG_NO_PRUNE | api.SFK_FLAG_KEEP_GRID)):
    ctx, c, s = api.run_shell(api.default_params(seed=1337, **kw), mass, halos, blocks)
    print(kw, "status", s.tolist(), "c000", np.round(c[:, 0], 4).tolist())
ctx, c, s = api.run_shell(api.default_params(seed=1337, rings=4), mass, halos, blocks)
props = ctx.shell_props(c)
m, pairs = ctx.shell_particles(halos[:, :3], c, props[:, 9], props[:, 10], halos[:, 3].astype(np.float32), 0.05, blocks)
print("masses", np.round(m / mass).tolist(), "band pairs", sum(len(p) for p in pairs))
ctx2 = api.ShellContext(api.default_params(seed=1337, rings=4))
ctx2.comm_init(api.comm_unique_id(), 0, 1)
ctx2.begin_snapshot(blocks[0], mass)
ctx2.add_halos(halos)
for b in blocks:
    ctx2.push_block(b)
c2, s2 = ctx2.split_finish_comm()
print("comm path equal:", bool(np.array_equal(c2, c)))
