import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pyoracle as po
from shellfish_b200 import api, synth

def compare(name, rings, n, nbg, cells=1, nh=1, box=62.5, seed=1, spokes=256, bins=256):
    if nh == 1:
        halos, blocks, mass = synth.make_box(seed, box, 1, n, nbg, cells, centers=[[box/2]*3], radii=[1.0])
    else:
        halos, blocks, mass = synth.make_box(seed, box, nh, n, nbg, cells, r200m_range=(0.5,1.0))
    cfg = po.ShellConfig.default(rings=rings, threads=1, spokes=spokes, radialBins=bins)
    t=time.time()
    ref = po.shell_run(cfg, mass, halos, blocks, taps=("rsp","profiles","nInsert","startEdges","endEdges","nFiltered","nCandidates","nIntersections"))
    t_or = time.time()-t
    p = api.default_params(rings=rings, seed=1337, flags=api.SFK_FLAG_TAPS, spokes=spokes, radial_bins=bins)
    t=time.time()
    ctx, coeffs, status = api.run_shell(p, mass, halos, blocks)
    t_gpu = time.time()-t
    st = ctx.stats()
    print(f"== {name}: oracle {t_or:.2f}s gpu {t_gpu:.3f}s stats {st}")
    print(" status gpu", status, "oracle", ref['status'], ref['ok'])
    for h in range(len(halos)):
        ins, s, e = ctx.counts(h)
        print(f" halo {h}: cand gpu {ctx.candidate_count(h)} ref {ref['nCandidates'][h]}; inserts gpu {ins.sum()} ref {ref['nInsert'][h].sum()} ; mismatch ins {np.count_nonzero(ins!=ref['nInsert'][h])} start {np.count_nonzero(s!=ref['startEdges'][h])} end {np.count_nonzero(e!=ref['endEdges'][h])}")
        prof = ctx.profiles(h); rp = ref['profiles'][h]
        den = np.maximum(np.abs(rp), 1e-300)
        print("  profile max rel err", np.max(np.abs(prof-rp)/den))
        rsp = ctx.rsp(h); rr = ref['rsp'][h]
        nanm = np.isnan(rsp) != np.isnan(rr)
        both = ~np.isnan(rsp) & ~np.isnan(rr)
        print("  rsp nan mismatch", nanm.sum(), "max rel", np.max(np.abs(rsp[both]-rr[both])/rr[both]) if both.any() else None, "n differing >1e-5", np.count_nonzero(np.abs(rsp[both]-rr[both])/rr[both] > 1e-5))
        c, rc = coeffs[h], ref['coeffs'][h]
        print("  coeff max abs diff", np.max(np.abs(c-rc)), "rel to max", np.max(np.abs(c-rc))/max(np.max(np.abs(rc)),1e-300), "npts", ref['nFiltered'][h])
    return ctx

compare("small r10", 10, 20000, 5000)
compare("rings3", 3, 20000, 5000)
compare("multi", 6, 8000, 20000, cells=2, nh=5, box=8.0)
compare("default100", 100, 50000, 10000)
