"""GPU suite (-m gpu): the parity cases of tests/sdt_cases.py on libsdtree.so, called
through the C ABI, against the oracle.  Two buffer modes: torch CUDA tensors (device
pointers, the integrator's path) and numpy arrays (host pointers staged by the library,
SDT_HOST_PTRS -- the path bench.py's e2e number times)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sdt_cases as cases  # noqa: E402
import fuzz_cases  # noqa: E402

pytestmark = pytest.mark.gpu


def _ctx(mode):
    import torch
    from practical_path_guiding_lab_b200 import SDTree
    from practical_path_guiding_lab_b200.build import build
    build()
    make = lambda **kw: SDTree(device=0, **kw)
    if mode == "host":
        return cases.Ctx(make=make)

    def dev(x):
        if x is None:
            return None
        a = np.ascontiguousarray(x)
        if a.dtype == np.uint32:
            a = a.view(np.int32)
        return torch.from_numpy(a).cuda()

    def host(x):
        if x is None:
            return None
        if isinstance(x, np.ndarray):
            return x
        torch.cuda.synchronize()
        return x.cpu().numpy()
    return cases.Ctx(make=make, dev=dev, host=host)


@pytest.fixture(scope="module", params=["device", "host"])
def ctx(request):
    return _ctx(request.param)


@pytest.mark.parametrize("case", cases.ALL_CASES + fuzz_cases.SUITE_CASES, ids=lambda c: c.__name__)
def test_case(ctx, case):
    case(ctx)


def test_npz_roundtrip(ctx, tmp_path):
    cases.case_npz_roundtrip(ctx, tmp_path)


def test_library_is_cuda():
    """the product path is the CUDA library: kernels were launched, an L2 probe runs"""
    c = _ctx("device")
    t = c.make(kd_capacity=64, quad_capacity=256)
    before = t.kernel_launches()
    import torch
    pos = torch.rand(1000, 3, device="cuda")
    t.sample(pos, seed=1)
    torch.cuda.synchronize()
    assert t.kernel_launches() > before
    assert t.measure_l2(16 << 20, 20) > 500.0      # GB/s; anything CPU-like would be far below


def test_large_wavefront_properties():
    """full-size wavefront (2^22 lanes): size-independent properties -- determinism,
    pdf(sample) consistency, conservation of splatted energy and counts"""
    import torch
    c = _ctx("device")
    t, cur, prev = cases.train(c, iters=3)
    n = 1 << 22
    g = torch.Generator(device="cuda").manual_seed(5)
    pos = torch.rand(n, 3, device="cuda", generator=g)
    d1, p1 = t.sample(pos, seed=42)
    d2, p2 = t.sample(pos, seed=42)
    assert torch.equal(d1, d2) and torch.equal(p1, p2)
    p3 = t.pdf(pos, d1)
    assert torch.equal(p1, p3)                      # KDTree.sample's pdf IS KDTree.pdf of the sampled direction
    assert torch.isfinite(p1).all() and (p1 >= 0).all()
    nrm = d1.norm(dim=1)
    assert (nrm - 1).abs().max() < 1e-5
    # splat: dyadic energies -> exact, order-independent sums
    dirs = torch.rand(n, 2, device="cuda", generator=g)
    rad = torch.randint(0, 9, (n,), device="cuda", generator=g).float() / 8
    wo = torch.full((n,), 0.5, device="cuda")
    t.splat_records(pos, dirs, rad, wo)
    got = t.download(1)
    R = t.sizes()['n_roots']
    total = float((rad.double() / 0.5).sum())
    assert float(got['quadtree_irradiance'][:R].astype(np.float64).sum()) == total
    assert float(got['quadtree_irradiance'][got['quadtree_isLeaf']].astype(np.float64).sum()) == total
    assert got['kdtree_vertCount'][0] == n == got['kdtree_vertCount'][got['kdtree_isLeaf']].sum()


def test_cornell_box_train_and_render_gpu():
    """config-1 shape on the GPU: the reference driver loop (iteration doubling, refine per
    iteration, guiding from iteration 2) on the analytic Cornell box, MSE against the reference's
    own ground truth (TungstenRender.exr, box-downsampled fixture).  Stated bound: the final image of the
    64-spp budget has per-pixel luminance MSE (clamped like computeMSE) below 2e-3 at 128x128 -- 2.6x the
    7.7e-4 measured on the B200 (seed 3; the noise of 36 blended samples per pixel) -- and a mean colour within
    2 % of the ground truth's (measured 0.35 %): a biased estimator (a wrong mixture pdf or MIS weight) fails both."""
    import torch
    from practical_path_guiding_lab_b200 import driver
    from practical_path_guiding_lab_b200.build import build
    from practical_path_guiding_lab_b200.cornell import CornellBox
    build()
    gt = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cornell_box_tungsten_256.npy")).astype(np.float32)
    gt = torch.from_numpy(gt.reshape(128, 2, 128, 2, 3).mean(axis=(1, 3))).cuda()
    r = CornellBox(128, 128, max_depth=30, device="cuda")
    r.setup()
    res = driver.train_and_render(r, 64, seed=3, ground_truth=gt)
    assert [s for _, s, _ in res["iterations"]] == [4, 8, 16, 32, 4]
    s = r.core.tree.sizes()
    assert s["error"] == 0 and s["refine_count"] == 3 and s["n_quad"] > 100
    img = res["image"]
    assert torch.isfinite(img).all()
    rel = float((img.mean((0, 1)) - gt.mean((0, 1))).abs().max() / gt.mean())
    assert rel < 0.02, rel                                   # unbiased up to noise
    print("cornell 128x128 budget 64: final mse_groundTruth", res["records"][-1]["mse_groundTruth"], "mean colour error", rel)
    assert res["records"][-1]["mse_groundTruth"] < 2e-3, res["records"][-1]


def test_two_handles_on_two_devices():
    """one handle per GPU in ONE process: the >48 KB shared-memory opt-in and the occupancy answer are per device, and every
    entry point runs on its handle's device whatever device the caller has current (advisor finding, round 1)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from practical_path_guiding_lab_b200 import SDTree
    rec = cases.dyadic_records(20000, 5, ((0.3, 0.7, 0.02),))
    trees = []
    for dev in (0, 1):
        t = SDTree(device=dev, kd_capacity=1 << 14, quad_capacity=1 << 18, store_nee=False)
        torch.cuda.set_device(1 - dev)                 # the OTHER device is current while this handle works
        d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(f"cuda:{dev}")
        for it in range(3):
            t.splat_records(d(rec.position), d(rec.direction), d(rec.radiance), d(rec.woPdf))     # ~80 KB of dynamic shared memory
            t.set_max_leaf_size(400)
            t.refine()
        trees.append(t)
    a, b = trees[0].download(0), trees[1].download(0)
    assert a['kdtree_depth'].shape[0] > 15
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k
    pos = np.random.default_rng(1).random((4096, 3)).astype(np.float32)
    torch.cuda.set_device(0)
    d1, p1 = trees[1].sample(torch.from_numpy(pos).to("cuda:1"), seed=9)
    d0, p0 = trees[0].sample(torch.from_numpy(pos).to("cuda:0"), seed=9)
    assert cases.beq(d0.cpu().numpy(), d1.cpu().numpy()) and cases.beq(p0.cpu().numpy(), p1.cpu().numpy())


def test_tile_sizes_agree():
    """The compacting kernel sorts 256-lane tiles per warp on large wavefronts and 64-lane tiles on wavefronts too small
    to give every resident thread two lanes (launch_wavefront): the same masked bounce / masked sample evaluated as ONE
    large launch and as small launches over slices of it (lane_offset keeps the generator keys) gives the same bits --
    the small-tile path is the one the oracle-checked cases run."""
    import torch
    c = _ctx("device")
    t, cur, prev = cases.train(c, iters=3)
    n, chunk = 1 << 20, 1 << 16
    g = torch.Generator(device="cuda").manual_seed(17)
    pos = torch.rand(n, 3, device="cuda", generator=g)
    mode = (torch.rand(n, device="cuda", generator=g) * 3).to(torch.uint8).clamp_(max=2)
    wo = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda", generator=g), dim=1)
    bp = torch.rand(n, device="cuda", generator=g) * 2
    bv = torch.rand(n, 3, device="cuda", generator=g)
    act = (torch.rand(n, device="cuda", generator=g) < 0.4).to(torch.uint8)
    big = t.guided(pos, mode, wo=wo, seed=11, bsdf_pdf=bp, bsdf_value=bv, em_dir=wo, em_active=act)
    sd, sp = t.sample(pos, active=act, seed=5)
    for a in range(0, n, chunk):
        s = slice(a, a + chunk)
        small = t.guided(pos[s], mode[s], wo=wo[s], seed=11, lane_offset=a, bsdf_pdf=bp[s], bsdf_value=bv[s], em_dir=wo[s], em_active=act[s])
        for x, y in zip(big, small):
            assert torch.equal(x[s], y), a
        d2, p2 = t.sample(pos[s], active=act[s], seed=5, lane_offset=a)
        assert torch.equal(sd[s], d2) and torch.equal(sp[s], p2), a
