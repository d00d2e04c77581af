"""Randomised differential cases: one seed = one random configuration (non-cubic, non-dyadic
boxes with negative corners, depth caps from 0 up, tiny leaf sizes, NEE on/off), a few
splat + refine iterations on records salted with the inputs the reference's masks exist for
(positions outside the box / NaN, directions outside [0,1]^2 / NaN, negative, NaN and infinite
radiance, zero / negative / NaN woPdf, inactive lanes), then every query, one guided bounce
(sdt_guided with a random bsdfSamplingFraction), the NEE MIS weights and the fused path-data splat on
salted inputs -- all held against the oracle bit for bit (general fp32 energies: 2e-4) through the
same C ABI the other cases use.

Energies stay multiples of 1/8 (see sdt_cases.dyadic_records) so the fp32 sums do not depend
on the order of the atomics.  Run many seeds by hand with
    python tests/fuzz_cases.py --seeds 300            (host emulation, no GPU needed; --gpu = the CUDA library)
the suites run a handful (tests/test_hostemu_parity.py, tests/test_gpu_parity.py)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sdt_cases as cases  # noqa: E402
from oracle import sdtree_oracle as so  # noqa: E402

F = np.float32
U = np.uint32
QUAD_CAPACITY = 1 << 21


def random_records(rng, n, lo, hi, nee, specials=True, negative=True):
    ext = hi - lo
    pos = (lo + rng.random((n, 3)) ** rng.choice([1.0, 1.5, 3.0]) * ext).astype(F)
    d = rng.random((n, 2)).astype(F)
    nl = int(rng.integers(0, 4))
    k = rng.integers(0, nl + 1, n)
    for j in range(nl):
        cx, cy, s = rng.random(), rng.random(), 10.0 ** rng.uniform(-4, -1)
        m = k == j
        d[m] = np.clip(np.stack([cx + s * rng.standard_normal(m.sum()), cy + s * rng.standard_normal(m.sum())], 1), 0, 1).astype(F)
    radiance = (rng.integers(0, 17, n) / 8.0).astype(F)
    wo = rng.choice(np.array([0.25, 0.5, 1.0, 2.0], F), n).astype(F)
    active = rng.random(n) < rng.choice([1.0, 0.9, 0.3])
    if specials and n >= 64:
        def pick(frac=0.01):
            return rng.random(n) < frac
        pos[pick()] += (ext * 1.5).astype(F)                       # outside the box
        pos[pick(0.005), int(rng.integers(0, 3))] = np.nan
        m = pick(0.005)
        pos[m] = np.where(rng.random((int(m.sum()), 3)) < 0.5, lo, hi).astype(F)     # box corners
        m = pick(0.01)
        pos[m, 0] = (lo[0] + ext[0] * F(0.5)).astype(F)            # on the first split plane
        d[pick(0.005)] += F(1.5)                                    # outside the canonical square
        d[pick(0.005), 0] = np.nan
        m = pick(0.01)
        d[m] = (rng.integers(0, 5, (int(m.sum()), 2)) / 4.0).astype(F)              # on quadrant borders
        if negative:
            radiance[pick(0.01)] *= F(-1.0)
        if rng.random() < 0.3:
            radiance[pick(0.002)] = np.nan
        if rng.random() < 0.3:
            radiance[pick(0.002)] = np.inf
        wo[pick(0.01)] = 0.0
        wo[pick(0.005)] = -1.0
        wo[pick(0.003)] = np.nan
    rec = so.SurfaceInteractionRecord(pos, d, radiance, wo)
    if nee:
        # NEE radiance with an exactly dyadic luminance (sdt_cases.NEE_GREEN_UNIT): non-zero NEE energy, still
        # order-independent sums; non-dyadic energies are covered by sdt_cases.case_splat_float_tolerance
        rec.radiance_nee = cases.dyadic_nee(rng, n)
        rec.direction_nee = rng.random((n, 2)).astype(F)
        if specials and n >= 64:
            rec.direction_nee[rng.random(n) < 0.01] += F(1.5)
            rec.direction_nee[rng.random(n) < 0.005, 1] = np.nan
            rec.radiance_nee[rng.random(n) < 0.005, 0] = np.nan
    return rec, active


def fuzz_one(ctx, seed, verbose=False):
    rng = np.random.default_rng(100000 + seed)
    lo = rng.uniform(-5, 5, 3).astype(F)
    hi = (lo + rng.uniform(0.1, 20, 3)).astype(F)
    kd_max_depth = int(rng.choice([0, 1, 2, 3, 5, 8, 12]))
    quad_max_depth = int(rng.choice([0, 1, 2, 4, 7, 12, 20]))
    store_nee = bool(rng.random() < 0.4)
    max_leaf = float(rng.choice([1, 3, 20, 150, 1000]))
    iters = int(rng.integers(1, 5))
    n = int(rng.choice([0, 1, 33, 700, 5000, 12000]))
    cfg = dict(bbox_min=tuple(float(x) for x in lo), bbox_max=tuple(float(x) for x in hi), kd_max_depth=kd_max_depth,
               quad_max_depth=quad_max_depth, store_nee=store_nee, kd_capacity=1 << 14, quad_capacity=QUAD_CAPACITY)
    if verbose:
        print(seed, cfg, max_leaf, iters, n)
    t = ctx.make(**cfg)
    cur, prev = cases.oracle_pair(lo, hi, kd_max_depth, quad_max_depth, store_nee)
    peak = 0
    for it in range(iters):
        # negative radiance can cancel a tree's total to exactly 0 -> threshold 0 -> every leaf with positive energy splits
        # down to the depth cap (up to 4^depth nodes per tree, in the reference too; sdt_cases.case_zero_total_energy) --
        # only affordable with shallow caps
        rec, active = random_records(rng, n, lo, hi, store_nee, negative=quad_max_depth <= 4 and kd_max_depth <= 8)
        cases.splat(t, ctx, rec, active)
        cur.addDataPropagate(_compress(rec, active))
        if it == iters - 1 and rng.random() < 0.5:
            # leave the last iteration unrefined: `current` carries statistics, `prev` the older topology
            break
        t.set_max_leaf_size(max_leaf)
        t.refine()
        cases.oracle_refine(cur, prev, max_leaf)
        peak = max(peak, prev.quadTree.quadTreeNode.getWidth())
    if t.sizes()["error"] != 0:
        # the only legitimate error: the refined forest does not fit the arena (sticky bit 2, sdtree.h)
        assert t.sizes()["error"] == 2 and peak > QUAD_CAPACITY
        return
    cases.assert_tree_equal(t.download(0), prev)
    cases.assert_tree_equal(t.download(1), cur)
    _queries(ctx, t, prev, rng, lo, hi)
    _bounce_and_path_data(ctx, t, cur, prev, rng, lo, hi, store_nee)


def _compress(rec, active):
    idx = np.nonzero(active)[0]
    out = so.SurfaceInteractionRecord(rec.position[idx], rec.direction[idx], rec.radiance[idx], rec.woPdf[idx])
    out.radiance_nee = rec.radiance_nee[idx]
    out.direction_nee = rec.direction_nee[idx]
    return out


def _queries(ctx, t, prev, rng, lo, hi, n=1500):
    ext = hi - lo
    pos = (lo - 0.01 * ext + rng.random((n, 3)) * ext * 1.02).astype(F)
    pos[:4] = np.stack([lo, hi, lo + ext * F(0.5), lo + ext * F(0.25)]).astype(F)
    pos[4, 1] = np.nan
    active = rng.random(n) < 0.9
    a8 = active.astype(np.uint8)
    leaf, root = t.locate(ctx.dev(pos), ctx.dev(a8))
    o_leaf = prev.getLeafNodeIndex(pos, active)
    assert np.array_equal(ctx.host(leaf).view(U), o_leaf)
    assert np.array_equal(ctx.host(root).view(U), so.gather(prev.kdTreeNode.quadTreeRootIndex, o_leaf, active))
    depth = prev.quadTree.maxDepth
    u = rng.random((n, 3 * (depth + 2))).astype(F)
    u[:16, 2::3] = np.array([0.0, 0.25, 0.5, 0.75] * 4, F)[:, None]
    d, p, dbg = t.sample(ctx.dev(pos), ctx.dev(a8), u=ctx.dev(u), debug=True)
    od, op, odbg = prev.sample(pos, so.ExplicitSampler(u=u), active, return_debug=True)
    dbg = ctx.host(dbg).view(U)
    for col, key in ((0, 'leaf'), (1, 'root'), (2, 'sample_node'), (3, 'pdf_node')):
        assert np.array_equal(dbg[active, col], odbg[key][active]), key
    assert cases.beq(ctx.host(d), od) and cases.beq(ctx.host(p), op)
    d2, p2 = t.sample(ctx.dev(pos), ctx.dev(a8), seed=99, lane_offset=3)
    od2, op2 = prev.sample(pos, so.ExplicitSampler(seed=99, n=n, lane_offset=3), active)
    assert cases.beq(ctx.host(d2), od2) and cases.beq(ctx.host(p2), op2)
    dirs = rng.standard_normal((n, 3)).astype(F)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dirs[:6] = np.array([[1, 0, 0], [0, 0, 1], [0, 0, -1], [0, 0, 0], [np.nan, 0, 1], [-1, 0, 0]], F)
    pp, pdbg = t.pdf(ctx.dev(pos), ctx.dev(dirs), ctx.dev(a8), debug=True)
    opp, opdbg = prev.pdf(pos, dirs, active, return_debug=True)
    assert np.array_equal(ctx.host(pdbg).view(U)[active, 2], opdbg['pdf_node'][active])
    assert cases.beq(ctx.host(pp), opp)
    # the same pdf through the path-product jump table (no debug output), and sample + pdf fused into one call
    assert cases.beq(ctx.host(t.pdf(ctx.dev(pos), ctx.dev(dirs), ctx.dev(a8))), opp)
    fd, fp, fq = t.sample_pdf(ctx.dev(pos), ctx.dev(dirs), ctx.dev(a8), seed=99, lane_offset=3)
    assert cases.beq(ctx.host(fd), od2) and cases.beq(ctx.host(fp), op2) and cases.beq(ctx.host(fq), opp)


def _salt(rng, a, frac=0.02):
    """in place: a few zeros, negatives, NaNs and infinities"""
    flat = a.reshape(-1)
    for v in (0.0, -1.0, np.nan, np.inf):
        flat[rng.random(flat.shape[0]) < frac / 4] = v
    return a


def _bounce_and_path_data(ctx, t, cur, prev, rng, lo, hi, store_nee, n=1200):
    """one guided bounce (sdt_guided: sample / pdf + mixture / idle lanes), the NEE MIS weights, and the fused
    processPathData + filter + splat call, all on salted inputs, against the oracle's separate steps"""
    ext = hi - lo
    pos = (lo - 0.01 * ext + rng.random((n, 3)) * ext * 1.02).astype(F)
    mode = rng.integers(0, 3, n).astype(np.uint8)
    wo = rng.standard_normal((n, 3)).astype(F)
    wo /= np.linalg.norm(wo, axis=1, keepdims=True)
    wo[:3] = np.array([[0, 0, 1], [np.nan, 0, 0], [0, 0, 0]], F)
    bp = _salt(rng, (rng.random(n) * 2).astype(F))
    bv = _salt(rng, rng.random((n, 3)).astype(F))
    frac = float(rng.choice([0.5, 0.25, 0.9]))
    # the emitter direction's pdf rides along on a random subset of the lanes, whatever their mode (idle lanes included)
    em = rng.standard_normal((n, 3)).astype(F)
    em /= np.linalg.norm(em, axis=1, keepdims=True)
    em[:3] = np.array([[0, 0, -1], [0, np.nan, 0], [1, 0, 0]], F)
    em_act = rng.random(n) < 0.7
    d, sp, wp, wt, ep = t.guided(ctx.dev(pos), ctx.dev(mode), wo=ctx.dev(wo), seed=5, lane_offset=11, bsdf_pdf=ctx.dev(bp),
                                 bsdf_value=ctx.dev(bv), bsdf_sampling_fraction=frac, em_dir=ctx.dev(em),
                                 em_active=ctx.dev(em_act.astype(np.uint8)))
    d, sp, wp, wt, ep = ctx.host(d), ctx.host(sp), ctx.host(wp), ctx.host(wt), ctx.host(ep)
    oep = prev.pdf(pos, em, em_act)
    assert cases.beq(ep[em_act], oep[em_act]) and np.all(ep[~em_act] == 1.0), "emitter-direction pdf fused into the bounce"
    m1, m2 = mode == 1, mode == 2
    od, op = prev.sample(pos, so.ExplicitSampler(seed=5, n=n, lane_offset=11), m1)
    assert cases.beq(d[m1], od[m1]) and cases.beq(sp[m1], op[m1])
    op2 = prev.pdf(pos, wo, m2)
    assert cases.beq(sp[m2], op2[m2])
    owo, ow = so.mixture(bp, op2, bv, m2, frac)
    assert cases.beq(wp[m2], owo[m2]) and cases.beq(wt[m2], ow[m2])
    # NEE MIS weight (src/path_guiding_integrator.py:241-253), learning and guiding iterations
    a = [_salt(rng, (rng.random(n) * 3).astype(F)) for _ in range(5)]
    delta = rng.random(n) < 0.2
    for it in (1, 2):
        s, m = t.mis_nee(ctx.dev(a[0]), ctx.dev(a[1]), ctx.dev(a[2]), ctx.dev(a[3]), ctx.dev(a[4]), ctx.dev(delta.astype(np.uint8)), frac, it)
        os_, om = so.nee_mis(a[0], a[1], a[2], a[3], a[4], delta, frac, it)
        assert cases.beq(ctx.host(s), os_) and cases.beq(ctx.host(m), om)
    # fused path-data splat on top of whatever `current` holds (general fp32 values: tolerance on the energies)
    md = int(rng.integers(1, 6))
    rays = n // md
    slots = rays * md
    Lf = _salt(rng, (rng.random((rays, 3)) * 4).astype(F), 0.01)
    tr = _salt(rng, (rng.random((slots, 3)) * 2).astype(F))
    tb = _salt(rng, rng.random((slots, 3)).astype(F))
    bsdf = _salt(rng, rng.random((slots, 3)).astype(F))
    p2 = pos[:slots]
    d2 = _salt(rng, rng.random((slots, 2)).astype(F), 0.01)
    wop = _salt(rng, (rng.random(slots) * 2).astype(F))
    nee = _salt(rng, (rng.random((slots, 3)) * (rng.random((slots, 1)) < 0.5)).astype(F), 0.01)
    dnee = rng.random((slots, 2)).astype(F)
    active = rng.random(slots) < 0.7
    before = t.download(1)
    rad = t.splat_path_data(md, ctx.dev(Lf), ctx.dev(tr), ctx.dev(tb), ctx.dev(bsdf), ctx.dev(p2), ctx.dev(d2), ctx.dev(wop),
                            ctx.dev(nee), ctx.dev(dnee), ctx.dev(active.astype(np.uint8)), want_radiance=True)
    _, orad = so.process_path_data(Lf, tr, tb, bsdf, md)
    keep, orad2, onee = so.filter_records(active, orad, nee, wop)
    assert cases.beq(ctx.host(rad), orad2)
    sub = so.SurfaceInteractionRecord(p2[keep], d2[keep], orad2[keep], wop[keep], onee[keep], dnee[keep])
    fresh = so.KDTree(maxDepth=cur.maxDepth)
    fresh.copyFrom(cur)
    fresh.resetTreeVertCount()
    fresh.resetAllQuadTreeIrradiance()
    fresh.addDataPropagate(sub, exact=True)
    got = t.download(1)
    np.testing.assert_array_equal(got['kdtree_vertCount'] - before['kdtree_vertCount'], fresh.kdTreeNode.vertCount)
    e = fresh.quadTree.quadTreeNode.irradiance.astype(np.float64) + before['quadtree_irradiance'].astype(np.float64)
    with np.errstate(invalid='ignore'):
        fin = np.isfinite(e) & np.isfinite(before['quadtree_irradiance'])
    np.testing.assert_allclose(got['quadtree_irradiance'][fin], e[fin], rtol=2e-4, atol=1e-5)


def make_case(seed):
    def case(ctx):
        fuzz_one(ctx, seed)
    case.__name__ = f"case_fuzz_seed{seed}"
    return case


SUITE_SEEDS = (0, 1, 3, 12, 13, 14, 18, 20, 33, 41, 43, 47, 59)     # empty / single-record / shallow / deep / NEE mixes
SUITE_CASES = [make_case(s) for s in SUITE_SEEDS]


if __name__ == "__main__":
    import argparse
    import traceback
    from hostemu.build_hostemu import build as build_hostemu
    from practical_path_guiding_lab_b200 import SDTree
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=100)
    ap.add_argument("--start", type=int, default=0)
    ap.add_argument("--gpu", action="store_true", help="run on libsdtree.so / cuda:0 (host-pointer calls) instead of the host emulation")
    ap.add_argument("--lib", default=None, help="with --gpu: an alternative build of libsdtree.so")
    a = ap.parse_args()
    if a.gpu:
        ctx = cases.Ctx(make=lambda **kw: SDTree(device=0, lib_path=a.lib, **kw))
    else:
        lib = build_hostemu()
        ctx = cases.Ctx(make=lambda **kw: SDTree(lib_path=lib, **kw))
    bad = []
    for s in range(a.start, a.start + a.seeds):
        try:
            print("seed", s, flush=True)
            fuzz_one(ctx, s)
        except Exception:
            bad.append(s)
            print("seed", s, "FAILED")
            traceback.print_exc(limit=3)
    print("failed seeds:", bad)
    sys.exit(1 if bad else 0)
