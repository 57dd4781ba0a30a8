"""Parity of the CUDA path (through the C ABI, shellfish_b200/libsfk.so) with
the CPU oracle on the same seeded inputs.  Runs on the B200 box only.

Bars (BASELINE.json north_star):
  * integer taps bit-exact: candidates per halo, inserts per LoS, start/end
    edges per (LoS, bin);
  * profiles (Halo.GetRhos), R_sp per LoS and Penna coefficients within 1e-5
    relative.  Profiles are compared relative to max(|ref|, background rho),
    coefficients relative to the halo's largest |coefficient| (the catalog
    prints %.6g of each, cmd/catalog/catalog.go:125-140).
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from shellfish_b200 import api, synth

pytestmark = pytest.mark.gpu

RTOL = 1e-5
ORC_STATUS = {0: 0, -2: 2, -3: 3, -4: 4, -1: 5}   # ORC_E_* -> SFK_HALO_*


def run_both(halos, blocks, mass, taps=True, **kw):
    names = {"rings": "rings", "spokes": "spokes", "radial_bins": "radialBins", "order": "order",
             "levels": "levels", "smoothing_window": "smoothingWindow",
             "subsample_factor": "subsampleFactor", "eta": "eta", "r_kernel_mult": "rKernelMult",
             "background_rho_mult": "backgroundRhoMult", "los_slope_cutoff": "losSlopeCutoff"}
    params = api.default_params(seed=1337, flags=api.SFK_FLAG_TAPS if taps else 0, **kw)
    cfg = po.ShellConfig.default(seed=1337, threads=1, **{names[k]: v for k, v in kw.items()})
    ctx, coeffs, status = api.run_shell(params, mass, halos, blocks)
    want = ("rsp", "profiles", "nInsert", "startEdges", "endEdges", "nFiltered", "nCandidates",
            "nIntersections") if taps else ("rsp", "nFiltered", "nCandidates")
    ref = po.shell_run(cfg, mass, halos, blocks, taps=want)
    return ctx, coeffs, status, ref


def assert_parity(ctx, coeffs, status, ref, n_halos, taps=True):
    for h in range(n_halos):
        assert status[h] == ORC_STATUS[int(ref["status"][h])], (h, status[h], ref["status"][h])
        assert ctx.candidate_count(h) == ref["nCandidates"][h]
        if taps:
            ins, st, en = ctx.counts(h)
            assert np.array_equal(ins, ref["nInsert"][h]), "inserts per LoS differ (halo %d)" % h
            assert np.array_equal(st, ref["startEdges"][h]), "start-edge counts differ (halo %d)" % h
            assert np.array_equal(en, ref["endEdges"][h]), "end-edge counts differ (halo %d)" % h
            prof, rp = ctx.profiles(h), ref["profiles"][h]
            fin = np.isfinite(rp)
            assert np.array_equal(np.isfinite(prof), fin)
            bg = rp[fin].min() if fin.any() else 0.0
            den = np.maximum(np.abs(rp[fin]), max(bg, 1e-300))
            assert np.max(np.abs(prof[fin] - rp[fin]) / den, initial=0.0) <= RTOL
        rsp, rr = ctx.rsp(h), ref["rsp"][h]
        assert np.array_equal(np.isnan(rsp), np.isnan(rr)), "R_sp ok-mask differs (halo %d)" % h
        m = ~np.isnan(rr)
        assert np.all(np.abs(rsp[m] - rr[m]) <= RTOL * np.abs(rr[m]))
        if status[h] == 0:
            scale = np.max(np.abs(ref["coeffs"][h]))
            assert np.all(np.abs(coeffs[h] - ref["coeffs"][h]) <= RTOL * scale), (coeffs[h], ref["coeffs"][h])
        else:
            assert np.all(coeffs[h] == 0)
    if taps:
        assert ctx.stats()["n_intersections"] == int(ref["nIntersections"].sum())


def test_single_halo_rng_rings():
    """BASELINE configs[0] shape at test size: one NFW halo, one block."""
    halos, blocks, mass = synth.single_halo(seed=1, n=20000, n_background=5000)
    ctx, c, s, ref = run_both(halos, blocks, mass, rings=10)
    assert s[0] == 0
    assert_parity(ctx, c, s, ref, 1)


def test_three_rings_bypass_rng():
    """Rings == 3 uses the axis triplet (cmd/shell.go:570-571)."""
    halos, blocks, mass = synth.single_halo(seed=3, n=20000, n_background=5000)
    ctx, c, s, ref = run_both(halos, blocks, mass, rings=3)
    assert_parity(ctx, c, s, ref, 1)


def test_default_config_full_halo():
    """Default shell.config (100 x 256 x 256) on a 5e4-particle halo."""
    halos, blocks, mass = synth.single_halo(seed=4, n=50000, n_background=10000)
    ctx, c, s, ref = run_both(halos, blocks, mass)
    assert s[0] == 0
    assert_parity(ctx, c, s, ref, 1)


def test_periodic_wrap_many_halos_many_blocks():
    """Halos near the faces of a small periodic box cut into 64 blocks: the
    cumulative in-place float32 wrap of Halo.Transform (los/halo.go:168-197)
    and SheetIntersect's periodic branch (halo.go:442-467) are exercised
    (blocks narrower than half the box, see test_sheet_intersect_periodic)."""
    box = 8.0
    centers = [[0.3, 4.0, 4.0], [7.8, 0.2, 4.1], [4.0, 4.0, 7.9], [0.1, 0.1, 0.1], [3.0, 5.0, 2.0]]
    radii = [0.8, 0.7, 0.9, 0.6, 0.75]
    halos, blocks, mass = synth.make_box(7, box, 5, 9000, 30000, 4, centers=centers, radii=radii)
    ctx, c, s, ref = run_both(halos, blocks, mass, rings=6)
    assert_parity(ctx, c, s, ref, 5)


@pytest.mark.parametrize("order", [2, 4])
def test_penna_order(order):
    halos, blocks, mass = synth.single_halo(seed=5, n=25000, n_background=5000)
    ctx, c, s, ref = run_both(halos, blocks, mass, rings=12, order=order)
    assert c.shape[1] == 2 * order * order
    assert_parity(ctx, c, s, ref, 1)


def test_subsample_factor():
    """SubsampleFactor = 2 keeps every 8th particle with 8x weight (cmd/shell.go:490-495)."""
    halos, blocks, mass = synth.single_halo(seed=6, n=80000, n_background=20000)
    ctx, c, s, ref = run_both(halos, blocks, mass, rings=8, subsample_factor=2)
    assert_parity(ctx, c, s, ref, 1)


def test_ragged_shapes():
    """Spokes not a multiple of the 32-LoS tile, bins/window of los/halo_bench_test.go."""
    halos, blocks, mass = synth.single_halo(seed=8, n=20000, n_background=4000)
    ctx, c, s, ref = run_both(halos, blocks, mass, rings=5, spokes=100, radial_bins=200,
                              smoothing_window=61, levels=2)
    assert_parity(ctx, c, s, ref, 1)


def test_window_not_smaller_than_bins():
    """Smooth returns ok=false for every LoS when bins <= window (smooth.go:51-53):
    no valid point, NewKDETree panics in the reference -> status no_points."""
    halos, blocks, mass = synth.single_halo(seed=9, n=5000, n_background=1000)
    ctx, c, s, ref = run_both(halos, blocks, mass, rings=3, radial_bins=64, smoothing_window=121)
    assert s[0] == 2
    assert_parity(ctx, c, s, ref, 1)


def test_zero_min_mass_like_gadget2():
    """Gadget-2's reader never sets MinMass (io/gadget2.go:327,402): the
    background density is 0 and nothing clamps the profile.  Bins no kernel
    covers then hold the reference's float64 cancellation residue (~1e-16 of
    the plateau, either sign, Threads-dependent) and ln() of that residue runs
    through the FIR, so a LoS that touches such a bin has no well-defined
    answer in the reference itself.  Integer taps must still be bit-exact;
    profiles are compared with a floor at 1e-6 of the halo's peak density and
    R_sp on the lines of sight whose profile stays above that floor."""
    halos, blocks, mass = synth.single_halo(seed=10, n=20000, n_background=0)
    ctx, c, s, ref = run_both(halos, blocks, 0.0, rings=6)
    ins, st, en = ctx.counts(0)
    assert np.array_equal(ins, ref["nInsert"][0])
    assert np.array_equal(st, ref["startEdges"][0]) and np.array_equal(en, ref["endEdges"][0])
    prof, rp = ctx.profiles(0), ref["profiles"][0]
    floor = 1e-6 * rp.max()   # cancellation residue reaches ~1e-16 * peak * sqrt(adds)
    assert np.all(np.abs(prof - rp) <= RTOL * np.maximum(np.abs(rp), floor))
    clean = (rp > floor).all(axis=2)
    assert clean.mean() > 0.05
    rsp, rr = ctx.rsp(0), ref["rsp"][0]
    assert np.array_equal(np.isnan(rsp[clean]), np.isnan(rr[clean]))
    m = clean & ~np.isnan(rr)
    assert np.all(np.abs(rsp[m] - rr[m]) <= RTOL * rr[m])


def test_skipped_and_empty_halos():
    """r200m <= 0 is skipped (cmd/shell.go:535-538); a halo with no particle
    in range has no valid LoS."""
    halos, blocks, mass = synth.make_box(11, 40.0, 1, 20000, 0, 1, centers=[[20, 20, 20]], radii=[1.0])
    halos = np.concatenate([halos, [[5.0, 5.0, 5.0, 0.5], [10.0, 10.0, 10.0, -1.0], [1.0, 2.0, 3.0, 0.0]]])
    params = api.default_params(seed=1337, rings=4)
    ctx, c, s = api.run_shell(params, mass, halos, blocks)
    assert list(s) == [0, 2, 1, 1]
    assert np.all(c[1:] == 0) and np.any(c[0] != 0)
    cfg = po.ShellConfig.default(seed=1337, rings=4)
    ref = po.shell_run(cfg, mass, halos, blocks)
    assert list(ref["ok"]) == [1, 0, 0, 0]
    assert np.all(np.abs(c[0] - ref["coeffs"][0]) <= RTOL * np.abs(ref["coeffs"][0]).max())


def test_bitwise_reproducible_and_batch_invariant():
    """int64 fixed-point bins make the result independent of atomic order, of
    the batch a halo lands in and of how its hit list is chunked."""
    halos, blocks, mass = synth.make_box(12, 30.0, 4, 15000, 20000, 2, r200m_range=(0.6, 0.9))
    params = api.default_params(seed=1337, rings=16)
    _, c1, s1 = api.run_shell(params, mass, halos, blocks)
    _, c2, s2 = api.run_shell(params, mass, halos, blocks)
    assert np.array_equal(c1, c2) and np.array_equal(s1, s2)
    # one halo alone: different batch composition and chunking of its hit lists
    for h in range(len(halos)):
        ctx = api.ShellContext(params)
        ctx.begin_snapshot(blocks[0], mass)
        ctx.add_halos(halos)
        owned = np.zeros(len(halos), bool)
        owned[h] = True
        ctx.set_owned(owned)
        for b in blocks:
            ctx.push_block(b)
        c, s = ctx.finish_snapshot()
        assert s[h] == s1[h] and all(s[k] == 6 for k in range(len(halos)) if k != h)
        assert np.array_equal(c[h], c1[h])


def test_device_and_host_push_agree():
    import torch
    halos, blocks, mass = synth.make_box(13, 20.0, 2, 12000, 8000, 2, r200m_range=(0.6, 0.9))
    params = api.default_params(seed=1337, rings=8)
    _, c_host, s_host = api.run_shell(params, mass, halos, blocks)
    ctx = api.ShellContext(params)
    ctx.begin_snapshot(blocks[0], mass)
    ctx.add_halos(halos)
    keep = []   # device buffers stay valid until finish_snapshot: the cull of a block is queued, not finished
    for b in blocks:
        keep.append((torch.from_numpy(b["xs"]).cuda(), torch.from_numpy(b["ms"]).cuda()))
        ctx.push_block_device(b, *keep[-1])
    c_dev, s_dev = ctx.finish_snapshot()
    assert np.array_equal(c_host, c_dev) and np.array_equal(s_host, s_dev)


def test_queued_host_push_agrees():
    """sfk_push_block_async (copies only queued, buffers held by the caller until push_sync /
    finish_snapshot) gives bitwise the result of sfk_push_block, from page-locked and from
    pageable buffers, with and without a sync in the middle of the snapshot."""
    import torch
    halos, blocks, mass = synth.make_box(13, 20.0, 2, 12000, 8000, 2, r200m_range=(0.6, 0.9))
    params = api.default_params(seed=1337, rings=8)
    _, c_host, s_host = api.run_shell(params, mass, halos, blocks)
    for pinned in (True, False):
        ctx = api.ShellContext(params)
        ctx.begin_snapshot(blocks[0], mass)
        ctx.add_halos(halos)
        keep = []
        for k, b in enumerate(blocks):
            pb = dict(b)
            if pinned:
                pb["xs"] = torch.from_numpy(b["xs"]).pin_memory()
                pb["ms"] = torch.from_numpy(b["ms"]).pin_memory()
            keep.append(pb)
            ctx.push_block(pb, wait=False)
            if k == len(blocks) // 2:
                ctx.push_sync()
        c, s = ctx.finish_snapshot()
        assert np.array_equal(c_host, c) and np.array_equal(s_host, s)


def test_ring_tables_match_oracle():
    params = api.default_params(seed=987654321, rings=100)
    ctx = api.ShellContext(params)
    norms, rot, irot = ctx.ring_tables()
    assert np.array_equal(norms, po.norm_vecs(987654321, 100))
    h = po.Halo(norms, [0, 0, 0], 0.3, 3.0, 8, 8)
    for r in range(100):
        assert np.array_equal(rot[r], h.rot(r)) and np.array_equal(irot[r], h.irot(r))


def test_validation_messages_follow_the_reference():
    """ShellConfig.validate, cmd/shell.go:143-187."""
    with pytest.raises(api.ShellfishError, match=r"The variable 'Rings' was set to 0\."):
        api.ShellContext(api.default_params(rings=0))
    with pytest.raises(api.ShellfishError, match=r"The variable 'RMinMult' was set to 3, but the variable 'RMaxMult' was set to 3\."):
        api.ShellContext(api.default_params(r_min_mult=3.0))
    with pytest.raises(api.ShellfishError, match="Order above 4"):
        api.ShellContext(api.default_params(order=5))


def test_call_order_errors():
    ctx = api.ShellContext(api.default_params(rings=3))
    with pytest.raises(api.ShellfishError):
        ctx.add_halos(np.array([[1.0, 1, 1, 1]]))
    with pytest.raises(api.ShellfishError):
        ctx.finish_snapshot()


def test_full_size_properties():
    """BASELINE-size halos (1e5 particles, default config) without the oracle:
    counters are consistent and the recovered shell tracks the planted
    truncation radius."""
    halos, blocks, mass = synth.config2(seed=21, n_halos=12, n_total=12 * 140000, box=60.0, cells=3)
    params = api.default_params(seed=1337)
    ctx, c, s = api.run_shell(params, mass, halos, blocks)
    st = ctx.stats()
    assert np.all(s == 0)
    assert st["n_candidates"] == sum(ctx.candidate_count(h) for h in range(len(halos)))
    assert st["n_intersections"] > 500 * st["n_candidates"]
    for h in range(len(halos)):
        rsp = ctx.rsp(h)
        assert np.isfinite(rsp).mean() > 0.9
        # synth truncates the NFW body at 1.25 R200m (triaxial, so a band)
        assert 0.8 < np.nanmedian(rsp) / halos[h, 3] < 1.7
        # c_000 is the mean shell radius
        assert 0.8 < c[h, 0] / halos[h, 3] < 1.7


def test_split_halo_emulated_ranks_match_single_gpu():
    """BASELINE config 3 at test size: the blocks of one halo are dealt to
    'ranks' (contexts on this GPU); the NCCL all-reduces of shard.split_halo_finish
    are emulated with tensor adds on the exchange buffers.  int64 bins make the
    result bitwise equal to the single-context run for any number of ranks."""
    import torch
    halos, blocks, mass = synth.make_box(31, 12.0, 1, 60000, 6000, 3, centers=[[6.0, 6.0, 6.0]], radii=[1.2])
    params = api.default_params(seed=1337, rings=24)
    _, c_ref, s_ref = api.run_shell(params, mass, halos, blocks)
    assert s_ref[0] == 0
    for world in (2, 3):
        ctxs = []
        for r in range(world):
            ctx = api.ShellContext(params)
            ctx.begin_snapshot(blocks[0], mass)
            ctx.add_halos(halos)
            for b in blocks[r::world]:
                ctx.push_block(b)
            ctx.split_count()
            ctxs.append(ctx)
        for which, op in ((api.SPLIT_HIT_COUNTS, torch.sum), (api.SPLIT_RHO_MAX, torch.amax)):
            bufs = [c.split_buffer(which) for c in ctxs]
            red = op(torch.stack(bufs), dim=0)
            for b in bufs:
                b.copy_(red)
        torch.cuda.synchronize()
        for c in ctxs:
            c.split_bin()
        grids = [c.split_buffer(api.SPLIT_GRID) for c in ctxs]
        total = torch.stack(grids).sum(dim=0)
        for g in grids:
            g.copy_(total)
        torch.cuda.synchronize()
        for c in ctxs:
            coeffs, status = c.split_finish()
            assert status[0] == 0
            assert np.array_equal(coeffs, c_ref)


def test_percentile_profile_mode():
    """PercentileProfile output (cmd/shell.go:629-660): bin radii + the
    percentile-th density over all rings*spokes profiles, against the oracle's
    restated msort.Percentile on its own GetRhos tap."""
    halos, blocks, mass = synth.make_box(31, 62.5, 2, 20000, 8000, 2, margin=8.0)
    cfg = po.ShellConfig.default(seed=1337, threads=1, rings=10)
    ref = po.shell_run(cfg, mass, halos, blocks, taps=("profiles",))
    for pct in (50.0, 90.0, 0.0, 100.0, 12.5):
        params = api.default_params(seed=1337, rings=10, percentile_profile=1, percentile=pct)
        ctx, rows, status = api.run_shell(params, mass, halos, blocks)
        assert rows.shape == (2, 512) and np.all(status == 0)
        for h in range(2):
            want = po.calc_percentile(halos[h, 3], ref["profiles"][h], pct)
            assert np.array_equal(rows[h, :256], want[:256]), "GetRs radii differ"
            assert np.all(np.abs(rows[h, 256:] - want[256:]) <= RTOL * np.abs(want[256:])), (pct, h)
    with pytest.raises(RuntimeError, match="percentile must be in the range"):
        api.ShellContext(api.default_params(percentile_profile=1, percentile=120.0))


def test_dense_core_centre_containing_hits():
    """A cloud inside 0.35 R200m: almost every ring hit is centre-containing
    (insertToRing's first branch, halo.go:241-250), the case whose LoS ranges are
    trimmed.  Integer taps must stay bit-exact."""
    rng = np.random.default_rng(77)
    n = 120000
    v = rng.normal(size=(n, 3))
    v *= (0.35 * rng.random(n) ** (1 / 3) / np.linalg.norm(v, axis=1))[:, None]
    box = 62.5
    c = np.array([30.0, 31.0, 32.0])
    xs = (c + v).astype(np.float32)
    halos = np.array([[c[0], c[1], c[2], 1.0]])
    mass = float(synth.uniform_mass(n, box))
    blocks = [dict(synth.COSMO, TotalWidth=box, Origin=(np.float32(0),) * 3, Width=(np.float32(box),) * 3,
                   xs=xs, ms=np.full(n, mass, np.float32))]
    params = api.default_params(seed=1337, rings=16, flags=api.SFK_FLAG_TAPS)
    ctx, coeffs, status = api.run_shell(params, mass, halos, blocks)
    ref = po.shell_run(po.ShellConfig.default(seed=1337, threads=4, rings=16), mass, halos, blocks,
                       taps=("nInsert", "startEdges", "endEdges", "profiles", "nIntersections"))
    ins, st, en = ctx.counts(0)
    assert np.array_equal(ins, ref["nInsert"][0])
    assert np.array_equal(st, ref["startEdges"][0])
    assert np.array_equal(en, ref["endEdges"][0])
    assert ctx.stats()["n_intersections"] == ref["nIntersections"][0]
    prof, rp = ctx.profiles(0), ref["profiles"][0]
    den = np.maximum(np.abs(rp), rp.min())
    assert np.max(np.abs(prof - rp) / den) <= RTOL


def test_nan_line_of_sight_is_reproduced():
    """oneValIntrDist takes sqrt(dist2 - b*b) (halo.go:337-346).  For a particle at
    (-a, a) in the plane of the z ring and the LoS at phi = pi/4, b = a (cos + sin) and
    (cos(pi/4) + sin(pi/4))^2 rounds to 2.0000000000000004 > 2: the radicand is negative,
    rHi is NaN, and ProfileRing.Insert adds rho to bin 0 without ever taking it out
    (profile_ring.go:93-110).  The pruning in make_hit must keep exactly these LoS."""
    halos, blocks, mass = synth.single_halo(seed=5, n=30000, n_background=5000)
    c = halos[0, :3]
    assert np.all(c == c.astype(np.float32))          # box centre 31.25: exact in float32
    extra = []
    for a in (2.0 ** -4, 2.0 ** -5, 3 * 2.0 ** -6, 2.0 ** -3):
        for cz in (2.0 ** -5, -2.0 ** -6):
            extra += [[-a, a, cz], [a, -a, cz], [a, a, cz], [-a, -a, cz]]
    extra = (c + np.array(extra)).astype(np.float32)
    b = blocks[0]
    b["xs"] = np.concatenate([b["xs"], extra]).astype(np.float32)
    b["ms"] = np.concatenate([b["ms"], np.full(len(extra), mass, np.float32)])
    # the crafted particles do hit the condition (ring 0 of the 3-ring set is the z ring and its
    # rotation is the identity in float32, so the ring-frame coordinates are x - origin)
    d = (extra - c.astype(np.float32)).astype(np.float64)
    events = 0
    for i in (32, 96, 160, 224):
        phi = float(i) / 256.0 * (2 * np.pi)
        bb = d[:, 1] * np.cos(phi) - d[:, 0] * np.sin(phi)
        events += int(np.count_nonzero(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] - bb * bb < 0))
    assert events >= 4, events
    ctx, coeffs, status, ref = run_both(halos, blocks, mass, rings=3)   # rings == 3: axis normals, ring 0 = z
    assert_parity(ctx, coeffs, status, ref, 1)
    # and they show: those LoS keep rho out to rMax, which a complete chord of a particle at
    # r < 0.2 R200m never does -- the outermost bin of LoS 32 sits above the background
    prof = ctx.profiles(0)
    assert prof[0, 32, -1] > 1.5 * prof[0, :, -1].min() or prof[0, 160, -1] > 1.5 * prof[0, :, -1].min()


def test_sliding_moment_smoothing_equals_the_tap_loop():
    """The default los kernel evaluates Smooth's two Savitzky-Golay filters through sliding
    polynomial moments and a table logarithm; SFK_FLAG_LOS_TAPLOOP selects ConvolveAt's literal
    tap loop (filter.go:104-161).  Same R_sp for every line of sight -- and hence as the oracle
    -- and the same coefficients to rounding; with MinMass = 0 (rho_bg = 0) both run the tap loop."""
    halos, blocks, mass = synth.single_halo(seed=21, n=30000, n_background=8000)
    base, c0, s0 = api.run_shell(api.default_params(rings=20, seed=1337, flags=api.SFK_FLAG_LOS_TAPLOOP), mass, halos, blocks)
    exp, c1, s1 = api.run_shell(api.default_params(rings=20, seed=1337), mass, halos, blocks)
    assert s0[0] == 0 and s1[0] == 0
    r0, r1 = base.rsp(0), exp.rsp(0)
    assert np.array_equal(np.isnan(r0), np.isnan(r1))
    assert np.array_equal(r0[~np.isnan(r0)], r1[~np.isnan(r1)])
    np.testing.assert_allclose(c1, c0, rtol=0, atol=1e-9 * np.abs(c0).max())
    _, cz0, _ = api.run_shell(api.default_params(rings=6, seed=1337, flags=api.SFK_FLAG_LOS_TAPLOOP), 0.0, halos, blocks)
    _, cz1, _ = api.run_shell(api.default_params(rings=6, seed=1337), 0.0, halos, blocks)
    assert np.array_equal(cz0, cz1, equal_nan=True)


def _taps_of(flags, halos, blocks, mass, **kw):
    ctx, c, s = api.run_shell(api.default_params(seed=1337, flags=flags, **kw), mass, halos, blocks)
    return ctx, c, s


def test_pruned_and_literal_hit_records_give_identical_integer_bins():
    """Differential test of the exact pruning in make_hit (sfk_kernels.cu) on > 1e7 ring hits.
    SFK_FLAG_NO_PRUNE builds every record as insertToRing / idxRange write it (los/halo.go:
    241-310: all n LoS for a centre-containing circle, float64 asin / atan2, the scan up to LoS
    n - 1 of a wedge that atan2 puts below zero); the default path trims LoS ranges and drops
    hits.  Inserts per LoS, start / end edges per (LoS, bin), the int64 bin grid and every
    downstream number must be identical: the pruning only removes work that reaches no bin."""
    rng = np.random.default_rng(5)
    halos, blocks, mass = synth.single_halo(seed=41, n=260000, n_background=30000)
    c = halos[0, :3]
    # a dense core (centre-containing circles) and a sheet of particles around phi = +-pi of
    # the z axis frame (wedges that wrap), on top of the NFW + infall halo
    n_core = 40000
    v = rng.normal(size=(n_core, 3))
    v *= (0.3 * rng.random(n_core) ** (1 / 3) / np.linalg.norm(v, axis=1))[:, None]
    ang = np.pi + rng.normal(scale=0.05, size=20000)
    rad = rng.uniform(0.25, 3.1, 20000)
    sheet = np.stack([rad * np.cos(ang), rad * np.sin(ang), rng.normal(scale=0.3, size=20000)], axis=1)
    extra = (c + np.concatenate([v, sheet])).astype(np.float32)
    b = blocks[0]
    b["xs"] = np.concatenate([b["xs"], extra]).astype(np.float32)
    b["ms"] = np.concatenate([b["ms"], np.full(len(extra), mass, np.float32)])
    T, K, N = api.SFK_FLAG_TAPS, api.SFK_FLAG_KEEP_GRID, api.SFK_FLAG_NO_PRUNE
    a, ca, sa = _taps_of(T | K, halos, blocks, mass)
    l, cl, sl = _taps_of(T | K | N, halos, blocks, mass)
    assert a.stats()["n_ring_hits"] > 10_000_000
    assert a.stats()["n_ring_hits"] == l.stats()["n_ring_hits"]   # raw sphereIntersectRing successes
    assert a.stats()["n_intersections"] == l.stats()["n_intersections"]
    for x, y in zip(a.counts(0), l.counts(0)):
        assert np.array_equal(x, y)
    ga, qa = a.grid(0)
    gl, ql = l.grid(0)
    assert qa == ql and np.array_equal(ga, gl)
    assert np.array_equal(a.rsp(0), l.rsp(0), equal_nan=True)
    assert np.array_equal(ca, cl) and np.array_equal(sa, sl)


@pytest.mark.parametrize("rings,n", [(100, 50000), (7, 30000)])
def test_product_bin_kernel_writes_the_grid_whose_taps_are_checked(rings, n):
    """The integer taps (inserts, edges) are read from bin_kernel<TAPS = true>; the product
    runs bin_kernel<false>.  Both must leave the same int64 grid, and the grid must be the
    prefix-differenced profile the oracle holds: sum over bins of grid * quantum == GetRhos
    before the clamp."""
    halos, blocks, mass = synth.single_halo(seed=43, n=n, n_background=8000)
    K = api.SFK_FLAG_KEEP_GRID
    prod, cp, sp_ = _taps_of(K, halos, blocks, mass, rings=rings)
    taps, ct, st = _taps_of(K | api.SFK_FLAG_TAPS, halos, blocks, mass, rings=rings)
    gp, qp = prod.grid(0)
    gt, qt = taps.grid(0)
    assert qp == qt and np.array_equal(gp, gt)
    assert np.array_equal(cp, ct) and np.array_equal(prod.rsp(0), taps.rsp(0), equal_nan=True)
    ref = po.shell_run(po.ShellConfig.default(seed=1337, threads=4, rings=rings), mass, halos, blocks,
                       taps=("profiles",))
    rho = np.cumsum(gp, axis=2).astype(np.float64) * qp
    rp = ref["profiles"][0]
    bg = rp.min()
    rho = np.maximum(rho, bg)
    assert np.max(np.abs(rho - rp) / np.maximum(np.abs(rp), bg)) <= RTOL


@pytest.mark.parametrize("spokes", [100, 96, 40, 300])
def test_wide_and_narrow_bin_tiles_leave_the_same_grid(spokes):
    """RadialBins = 256 without taps runs bin_kernel<false, 256, 64> (64 lines of sight per CTA), the
    tap instantiation 32 per CTA.  With line-of-sight counts that are not a multiple of either tile
    (a ragged last tile, a ring narrower than one tile) both must leave the same int64 grid and
    the same rows."""
    halos, blocks, mass = synth.single_halo(seed=47, n=20000, n_background=5000)
    K = api.SFK_FLAG_KEEP_GRID
    prod, cp, sp_ = _taps_of(K, halos, blocks, mass, rings=5, spokes=spokes)
    taps, ct, st = _taps_of(K | api.SFK_FLAG_TAPS, halos, blocks, mass, rings=5, spokes=spokes)
    gp, qp = prod.grid(0)
    gt, qt = taps.grid(0)
    assert qp == qt and np.array_equal(gp, gt)
    assert np.array_equal(cp, ct) and np.array_equal(sp_, st)
    assert np.array_equal(prod.rsp(0), taps.rsp(0), equal_nan=True)
    assert prod.stats()["n_intersections"] == taps.stats()["n_intersections"]


def test_sparse_halos_with_flat_profile_windows():
    """Halos of 1500 particles: many lines of sight are clamped to the background, or see no kernel
    edge, over more than a smoothing window.  There the reference's smoothed values at consecutive
    bins are bit-identical (identical window contents) and SplashbackRadius' strict comparisons see
    exact ties; sliding moments would break them (129 of 4608 lines of sight differed before such
    lines were handed to the literal tap loop, scripts/experiments/flat_profile_ties.py)."""
    halos, blocks, mass = synth.make_box(5, 20.0, 6, 1500, 3000, 2, r200m_range=(0.6, 0.9))
    ctx, c, s, ref = run_both(halos, blocks, mass, rings=3)
    assert_parity(ctx, c, s, ref, len(halos))
    # the product instantiation (no taps) takes the same decisions
    _, c2, s2 = api.run_shell(api.default_params(seed=1337, rings=3), mass, halos, blocks)
    assert np.array_equal(c2, c) and np.array_equal(s2, s)


@pytest.mark.parametrize("case", range(20))
def test_randomized_small_configurations(case):
    """Seeded random shapes and sizes against the oracle: sparse and dense halos, ragged ring /
    spoke / bin counts (fast and generic los kernels, stored and RED-added bin tiles), windows,
    KDE levels, orders, subsampling, several blocks with periodic wrap.

    Sparse halos expose what float64 rounding decides in the reference: a profile made of a few
    particles' plateaus repeats the same density levels and the same smoothed step patterns, and
    SplashbackRadius orders them with strict comparisons.  Two roundings take part: (1) the
    smoothing arithmetic -- the CUDA path hands every line of sight with long plateaus to a literal
    instantiation (libm log, multiply-then-add taps, exp), so this part matches; (2) the bin sums --
    the reference adds rho*(1-rem) + rho*rem, equal to rho only up to an ulp, where the CUDA path
    adds exact integers (rho - q) + q.  The oracle's "exact bins" variant (orc_los.c) carries the
    bin sums exactly and is otherwise the literal restatement: the CUDA path must agree with it on
    EVERY line of sight, and the literal oracle may differ from both only on a handful of lines."""
    rng = np.random.default_rng(1000 + case)
    n_halos = int(rng.integers(1, 5))
    n_per = int(rng.choice([400, 1500, 6000, 20000]))
    box = float(rng.choice([6.0, 12.0, 25.0]))
    cells = int(rng.integers(1, 4))
    halos, blocks, mass = synth.make_box(2000 + case, box, n_halos, n_per, int(rng.integers(0, 4000)), cells,
                                         r200m_range=(0.5, 0.9))
    bins = int(rng.choice([256, 256, 128, 200]))
    kw = dict(rings=int(rng.integers(3, 9)), spokes=int(rng.choice([256, 64, 100, 160])), radial_bins=bins,
              smoothing_window=int(rng.choice([w for w in (31, 61, 121) if w < bins - 8])),
              levels=int(rng.integers(1, 4)), order=int(rng.integers(1, 5)),
              subsample_factor=int(rng.choice([1, 1, 2])), eta=float(rng.choice([10.0, 6.0, 14.0])))
    ctx, c, s, literal = run_both(halos, blocks, mass, **kw)
    po.set_exact_bins(True)
    try:
        _, _, _, exact = run_both(halos, blocks, mass, **kw)
    finally:
        po.set_exact_bins(False)
    assert_parity(ctx, c, s, exact, n_halos)
    lines = differing = 0
    for h in range(n_halos):
        a, b = literal["rsp"][h], exact["rsp"][h]
        bad = (np.isnan(a) != np.isnan(b)) | (~np.isnan(a) & ~np.isnan(b) & (np.abs(a - b) > RTOL * np.abs(a)))
        lines += a.size
        differing += int(bad.sum())
    assert differing <= max(2, lines // 200), (differing, lines)   # rounding-decided lines of the reference
    if differing == 0:
        assert_parity(ctx, c, s, literal, n_halos)


def test_uniform_mass_push_equals_mass_array():
    """sfk_push_block[_device] with mass == NULL: every particle has the min_mass of
    sfk_begin_snapshot (what the LGadget-2 and gotetra readers fill ms with).  Bitwise the same
    result as pushing the array, from host and from device memory."""
    import torch
    halos, blocks, mass = synth.make_box(17, 20.0, 3, 9000, 6000, 2, r200m_range=(0.6, 0.9))
    params = api.default_params(seed=1337, rings=8)
    _, c_ref, s_ref = api.run_shell(params, mass, halos, blocks)
    assert all(np.all(b["ms"] == np.float32(mass)) for b in blocks)
    for device in (False, True):
        ctx = api.ShellContext(params)
        ctx.begin_snapshot(blocks[0], mass)
        ctx.add_halos(halos)
        keep = []
        for b in blocks:
            if device:
                keep.append(torch.from_numpy(b["xs"]).cuda())
                ctx.push_block_device(b, keep[-1], None)
            else:
                ctx.push_block(dict(b, ms=None))
        c, s = ctx.finish_snapshot()
        assert np.array_equal(c, c_ref) and np.array_equal(s, s_ref)
