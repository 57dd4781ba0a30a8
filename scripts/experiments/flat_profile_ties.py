import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from shellfish_b200 import api, synth
from oracle import pyoracle as po
halos, blocks, mass = synth.make_box(5, 20.0, 6, 1500, 3000, 2, r200m_range=(0.6, 0.9))
res = {}
for name, fl in (("default", 0), ("noprune", api.SFK_FLAG_NO_PRUNE), ("taploop", api.SFK_FLAG_LOS_TAPLOOP), ("taps", api.SFK_FLAG_TAPS)):
    ctx, c, s = api.run_shell(api.default_params(seed=1337, rings=3, flags=fl), mass, halos, blocks)
    res[name] = (c, [ctx.rsp(h) for h in range(len(halos))])
ref = po.shell_run(po.ShellConfig.default(seed=1337, threads=1, rings=3), mass, halos, blocks, taps=("rsp",))
for name in res:
    c, rsp = res[name]
    nd = sum(int(np.sum(~((rsp[h] == res["default"][1][h]) | (np.isnan(rsp[h]) & np.isnan(res["default"][1][h]))))) for h in range(len(halos)))
    no = sum(int(np.sum(~((rsp[h] == ref["rsp"][h]) | (np.isnan(rsp[h]) & np.isnan(ref["rsp"][h]))))) for h in range(len(halos)))
    close = sum(int(np.sum(~(np.isclose(rsp[h], ref["rsp"][h], rtol=1e-5, atol=0) | (np.isnan(rsp[h]) & np.isnan(ref["rsp"][h]))))) for h in range(len(halos)))
    print(name, "LoS differing from default:", nd, " exact-differing from oracle:", no, " beyond 1e-5 from oracle:", close, "of", 6 * 3 * 256)
