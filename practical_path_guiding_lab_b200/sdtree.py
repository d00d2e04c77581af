"""Python face of libsdtree.so: one `SDTree` = one sdt_handle = the reference integrator's
(sdTree_prev, sdTree_current) pair (/root/reference/src/path_guiding_integrator.py:68-69).

Buffers are either torch CUDA tensors (device pointers handed straight to the kernels,
work enqueued on torch's current stream) or numpy arrays (host pointers; the library
stages them H2D/D2H inside the call, SDT_HOST_PTRS).  torch is only the allocator here.
"""
import ctypes as C
import math

import numpy as np

from . import _lib as L

NPZ_KEYS = ('kdtree_maxLeafSize', 'kdtree_maxDepth', 'kdtree_bbox_min', 'kdtree_bbox_max',
            'kdtree_depth', 'kdtree_vertCount', 'kdtree_isLeaf', 'kdtree_quadTreeRootIndex',
            'kdtree_child_left_index', 'kdtree_child_right_index',
            'quadtree_maxDepth', 'quadtree_isStoreNEERadiance', 'quadtree_rootNodeIndex',
            'quadtree_bbox_min', 'quadtree_bbox_max', 'quadtree_depth', 'quadtree_irradiance',
            'quadtree_isLeaf', 'quadtree_refinementThreshold', 'quadtree_child_1_index',
            'quadtree_child_2_index', 'quadtree_child_3_index', 'quadtree_child_4_index')


class SDTreeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsdtree error {code}: {msg}")
        self.code = code


def _is_torch(x):
    return type(x).__module__.startswith("torch")


_TORCH_DTYPES = None       # numpy dtype -> torch dtype of a device buffer, filled on first use (torch is imported lazily)


def _torch_dtypes():
    global _TORCH_DTYPES
    if _TORCH_DTYPES is None:
        import torch
        _TORCH_DTYPES = (torch, {np.float32: torch.float32, np.uint32: torch.int32, np.uint8: torch.uint8})
    return _TORCH_DTYPES


class _Buf:
    """Resolves numpy / torch buffers to raw pointers; remembers whether the call is a
    host-pointer call and keeps temporaries alive."""

    def __init__(self):
        self.host = None
        self.keep = []
        self.device = None

    def _kind(self, x):
        host = not _is_torch(x)
        if not host and not x.is_cuda:
            raise TypeError("torch tensors passed to SDTree must live on a CUDA device (use numpy for host data)")
        if self.host is None:
            self.host = host
            if not host:
                self.device = x.device
        elif self.host != host:
            raise TypeError("do not mix numpy (host) and torch CUDA (device) buffers in one call")
        return host

    def arr(self, x, dtype, shape=None):
        """pointer to a contiguous array of `dtype` (None passes through)"""
        if x is None:
            return None
        if self._kind(x):
            a = np.ascontiguousarray(x, dtype=dtype)
            if shape is not None:
                a = a.reshape(shape)
            self.keep.append(a)
            return a.ctypes.data
        torch, tds = _torch_dtypes()
        td = tds[dtype]
        t = x
        if t.dtype is td and t.is_contiguous():              # the common case: nothing to convert (a wavefront loop makes
            self.keep.append(t)                              # a dozen of these per call)
            return t.data_ptr()
        if dtype is np.uint8 and t.dtype == torch.bool:
            t = t.view(torch.uint8) if t.is_contiguous() else t.contiguous().view(torch.uint8)
        if dtype is np.uint32 and t.dtype == torch.uint32:
            td = torch.uint32
        if t.dtype != td:
            t = t.to(td)
        t = t.contiguous()
        self.keep.append(t)
        return t.data_ptr()

    def vec(self, x, k):
        """(n,k) interleaved array, or a k-tuple of (n,) planes -> Vec3 / Vec2"""
        V = L.Vec3 if k == 3 else L.Vec2
        v = V()
        if x is None:
            return v
        if isinstance(x, (tuple, list)) and len(x) == k and not np.isscalar(x[0]) and getattr(x[0], "ndim", 0) == 1:
            ptrs = [self.arr(c, np.float32) for c in x]
            v.x, v.y = ptrs[0], ptrs[1]
            if k == 3:
                v.z = ptrs[2]
            v.stride = 1
            return v
        p = self.arr(x, np.float32)
        v.x, v.y = p, p + 4
        if k == 3:
            v.z = p + 8
        v.stride = k
        return v

    def new(self, shape, dtype):
        """output buffer of the same kind as the inputs"""
        if self.host or self.host is None:
            self.host = True
            a = np.empty(shape, dtype=dtype)
            return a, a.ctypes.data
        import torch
        td = {np.float32: torch.float32, np.uint32: torch.int32, np.uint8: torch.uint8}[dtype]
        t = torch.empty(shape, dtype=td, device=self.device)
        return t, t.data_ptr()

    def flags(self, extra=0):
        return (L.SDT_HOST_PTRS if self.host else 0) | extra

    def stream(self):
        if self.host or self.host is None:
            return None
        import torch
        return torch.cuda.current_stream(self.device).cuda_stream


def _n_of(x):
    if isinstance(x, (tuple, list)):
        return int(x[0].shape[0])
    return int(x.shape[0])


class _PreparedGuided:
    """an sdt_guided call with its argument struct already built (SDTree.prepare_guided)"""
    __slots__ = ("_tree", "_g", "_ref", "_n", "_flags", "_buf", "_outs", "_fn", "_h")

    def __init__(self, tree, g, n, flags, buf, outs):
        self._tree, self._g, self._n, self._flags, self._buf, self._outs = tree, g, n, flags, buf, outs
        self._ref = C.byref(g)
        self._fn, self._h = tree._lib.sdt_guided, tree._h

    def __call__(self, seed=None):
        if seed is not None:
            self._g.seed = int(seed) & 0xFFFFFFFF
        rc = self._fn(self._h, self._ref, self._n, self._flags, self._buf.stream())       # on the caller's current stream
        if rc != 0:
            self._tree._ck(rc)
        return self._outs


class SDTree:
    def __init__(self, bbox_min=(0, 0, 0), bbox_max=(1, 1, 1), kd_max_depth=20, quad_max_depth=20,
                 store_nee=True, device=0, kd_capacity=0, quad_capacity=0, lib_path=None):
        self._lib = L.load_library(lib_path)
        self._lib_is_cuda = lib_path is None or not str(lib_path).endswith("libsdtree_hostemu.so")
        cfg = L.Config()
        cfg.bbox_min[:] = [float(np.float32(v)) for v in bbox_min]
        cfg.bbox_max[:] = [float(np.float32(v)) for v in bbox_max]
        cfg.kd_max_depth = int(kd_max_depth)
        cfg.quad_max_depth = int(quad_max_depth)
        cfg.store_nee = int(bool(store_nee))
        cfg.device = int(device)
        cfg.kd_capacity = int(kd_capacity)
        cfg.quad_capacity = int(quad_capacity)
        h = C.c_void_p()
        rc = self._lib.sdt_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise SDTreeError(rc, (self._lib.sdt_last_error(None) or b"").decode())
        self._h = h
        self.host_wait = True       # False: host-pointer calls pass SDT_NO_WAIT; call synchronize() before reading outputs
        self.store_nee = bool(store_nee)
        self.device = int(device)

    # ---- plumbing -----------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.sdt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _flags(self, b, extra=0):
        """flags of one call; host-pointer calls return before their outputs landed when host_wait is False"""
        return b.flags(extra) | (L.SDT_NO_WAIT if (b.host and not self.host_wait) else 0)

    def synchronize(self, stream=None):
        """completes every call enqueued so far (needed after host-pointer calls made with host_wait = False)"""
        self._ck(self._lib.sdt_synchronize(self._h, stream))

    def _ck(self, rc):
        if rc != 0:
            raise SDTreeError(rc, (self._lib.sdt_last_error(self._h) or b"").decode())

    def sizes(self):
        s = L.Sizes()
        self._ck(self._lib.sdt_get_sizes(self._h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in L.Sizes._fields_}

    def set_tuning(self, key, value):
        self._ck(self._lib.sdt_set_tuning(self._h, key.encode(), int(value)))

    def kernel_launches(self):
        return int(self._lib.sdt_kernel_launches(self._h))

    def measure_l2(self, nbytes=32 << 20, passes=50):
        g = C.c_float()
        self._ck(self._lib.sdt_measure_l2(self._h, int(nbytes), int(passes), C.byref(g), None))
        return float(g.value)

    def measure_gather(self, nbytes=16 << 20, iters=200, via_l1=True):
        """random 32-byte-sector gather bandwidth over an L2-resident set (GB/s of sectors delivered)"""
        g = C.c_float()
        self._ck(self._lib.sdt_measure_gather(self._h, int(nbytes), int(iters), int(bool(via_l1)), C.byref(g), None))
        return float(g.value)

    def _stream(self, stream):
        """stream of a call that takes no arrays: the caller's, else torch's current stream on this device (so that the
        call is ordered after the splats / queries issued through torch tensors), else the NULL stream"""
        if stream is not None:
            return stream
        import sys
        torch = sys.modules.get("torch")
        if torch is not None and self._lib_is_cuda and torch.cuda.is_available():
            return torch.cuda.current_stream(self.device).cuda_stream
        return None

    # ---- queries on prev --------------------------------------------------------------
    def locate(self, pos, active=None, sync=True):
        """KDTree.getLeafNodeIndex + quadTreeRootIndex gather -> (leaf, root)"""
        b = _Buf()
        n = _n_of(pos)
        p = b.vec(pos, 3)
        a = b.arr(active, np.uint8)
        leaf, lp = b.new((n,), np.uint32)
        root, rp = b.new((n,), np.uint32)
        self._ck(self._lib.sdt_locate(self._h, C.byref(p), a, n, lp, rp, self._flags(b, L.SDT_SYNC if sync and b.host else 0), b.stream()))
        return leaf, root

    def sample(self, pos, active=None, u=None, seed=0, lane_offset=0, debug=False, out=None):
        """KDTree.sample -> (dir (n,3), pdf (n,)[, dbg (n,4)])"""
        b = _Buf()
        n = _n_of(pos)
        p = b.vec(pos, 3)
        a = b.arr(active, np.uint8)
        up, us = None, 0
        if u is not None:
            us = int(u.shape[1])
            up = b.arr(u, np.float32)
        if out is not None:
            d, pdf = out
            dv = b.vec(d, 3)
            pp = b.arr(pdf, np.float32)
        else:
            d, dp = b.new((n, 3), np.float32)
            dv = L.Vec3(dp, dp + 4, dp + 8, 3)
            pdf, pp = b.new((n,), np.float32)
        dbg, gp = (b.new((n, 4), np.uint32) if debug else (None, None))
        self._ck(self._lib.sdt_sample(self._h, C.byref(p), a, n, up, us, int(seed) & 0xFFFFFFFF, int(lane_offset),
                                      C.byref(dv), pp, gp, self._flags(b), b.stream()))
        return (d, pdf, dbg) if debug else (d, pdf)

    def sample_pdf(self, pos, qdir, active=None, u=None, seed=0, lane_offset=0, out=None):
        """KDTree.sample and KDTree.pdf of the given directions `qdir` at the same vertices, one spatial descent
        -> (dir (n,3), pdf (n,), qpdf (n,)); identical to sample(...) + pdf(pos, qdir)"""
        b = _Buf()
        n = _n_of(pos)
        p = b.vec(pos, 3)
        q = b.vec(qdir, 3)
        a = b.arr(active, np.uint8)
        up, us = None, 0
        if u is not None:
            us = int(u.shape[1])
            up = b.arr(u, np.float32)
        if out is not None:
            d, pdf, qpdf = out
            dv = b.vec(d, 3)
            pp = b.arr(pdf, np.float32)
            qp = b.arr(qpdf, np.float32)
        else:
            d, dp = b.new((n, 3), np.float32)
            dv = L.Vec3(dp, dp + 4, dp + 8, 3)
            pdf, pp = b.new((n,), np.float32)
            qpdf, qp = b.new((n,), np.float32)
        self._ck(self._lib.sdt_sample_pdf(self._h, C.byref(p), a, n, up, us, int(seed) & 0xFFFFFFFF, int(lane_offset),
                                          C.byref(dv), pp, C.byref(q), qp, self._flags(b), b.stream()))
        return d, pdf, qpdf

    def pdf(self, pos, direction, active=None, debug=False, out=None):
        """KDTree.pdf -> pdf (n,)[, dbg (n,3)]"""
        b = _Buf()
        n = _n_of(pos)
        p = b.vec(pos, 3)
        dv = b.vec(direction, 3)
        a = b.arr(active, np.uint8)
        if out is not None:
            pdf = out
            pp = b.arr(out, np.float32)
        else:
            pdf, pp = b.new((n,), np.float32)
        dbg, gp = (b.new((n, 3), np.uint32) if debug else (None, None))
        self._ck(self._lib.sdt_pdf(self._h, C.byref(p), C.byref(dv), a, n, pp, gp, self._flags(b), b.stream()))
        return (pdf, dbg) if debug else pdf

    def guided(self, pos, mode, wo=None, u=None, seed=0, lane_offset=0, bsdf_pdf=None, bsdf_value=None,
               bsdf_sampling_fraction=0.5, dir_out=None, sdtree_pdf_out=None, wo_pdf_out=None, weight_out=None,
               em_dir=None, em_active=None, sdtree_pdf_em_out=None):
        """one bounce: mode 1 lanes are sampled, mode 2 lanes get the pdf of `wo` (+ fused mixture).
        Output arrays are updated in place on the lanes concerned; fresh ones are zero-filled.
        With em_dir the lanes of em_active (all when None) also get the tree's pdf of that direction from the same
        spatial descent; the return value then has a fifth element, sdtree_pdf_em (fresh: 1 on the other lanes, like
        KDTree.pdf leaves inactive lanes)."""
        call = self.prepare_guided(pos, mode, wo, u, seed, lane_offset, bsdf_pdf, bsdf_value, bsdf_sampling_fraction, dir_out,
                                   sdtree_pdf_out, wo_pdf_out, weight_out, em_dir, em_active, sdtree_pdf_em_out)
        return call()

    def prepare_guided(self, pos, mode, wo=None, u=None, seed=0, lane_offset=0, bsdf_pdf=None, bsdf_value=None,
                       bsdf_sampling_fraction=0.5, dir_out=None, sdtree_pdf_out=None, wo_pdf_out=None, weight_out=None,
                       em_dir=None, em_active=None, sdtree_pdf_em_out=None):
        """guided() with the marshalling done ONCE: returns a callable that issues the same sdt_guided call on the same
        buffers every time it is invoked (`call(seed=...)` changes the generator seed) and returns the same output arrays.
        For wavefront loops that keep their buffers between bounces: no argument marshalling per call (measured on the
        65 k-lane wavefronts of a 256 x 256 pass: 30 bounce calls 0.74 -> 0.68 ms; the rest is the launches' own latency,
        DESIGN.md section 7).  The arrays must stay alive and in place (the callable holds references to them)."""
        b = _Buf()
        n = _n_of(pos)
        g = L.GuidedArgs()
        g.pos = b.vec(pos, 3)
        g.wo = b.vec(wo, 3)
        g.mode = b.arr(mode, np.uint8)
        if u is not None:
            g.u = b.arr(u, np.float32)
            g.u_stride = int(u.shape[1])
        g.seed = int(seed) & 0xFFFFFFFF
        g.lane_offset = int(lane_offset)
        g.bsdf_pdf = b.arr(bsdf_pdf, np.float32)
        g.bsdf_value = b.vec(bsdf_value, 3)
        g.bsdf_sampling_fraction = float(bsdf_sampling_fraction)

        def out_or_new(x, shape):
            if x is not None:
                return x
            t, _ = b.new(shape, np.float32)
            if b.host:
                t[...] = 0
            else:
                t.zero_()
            return t
        d = out_or_new(dir_out, (n, 3))
        sp = out_or_new(sdtree_pdf_out, (n,))
        fused = bsdf_pdf is not None
        wp = out_or_new(wo_pdf_out, (n,)) if fused else None
        wt = out_or_new(weight_out, (n, 3)) if fused and bsdf_value is not None else None
        ep = None
        if em_dir is not None:
            g.em_dir = b.vec(em_dir, 3)
            g.em_active = b.arr(em_active, np.uint8)
            ep = sdtree_pdf_em_out
            if ep is None:
                ep, _ = b.new((n,), np.float32)
                if b.host:
                    ep[...] = 1
                else:
                    ep.fill_(1)
        if b.host:
            for t in (d, sp, wp, wt, ep):
                if t is not None and not (t.flags.c_contiguous and t.dtype == np.float32):
                    raise TypeError("guided(): host output arrays must be C-contiguous float32")
            g.dir = L.Vec3(d.ctypes.data, d.ctypes.data + 4, d.ctypes.data + 8, 3)
            g.sdtree_pdf = sp.ctypes.data
            if wp is not None:
                g.wo_pdf = wp.ctypes.data
            if wt is not None:
                g.weight = L.Vec3(wt.ctypes.data, wt.ctypes.data + 4, wt.ctypes.data + 8, 3)
            if ep is not None:
                g.sdtree_pdf_em = ep.ctypes.data
        else:
            g.dir = L.Vec3(d.data_ptr(), d.data_ptr() + 4, d.data_ptr() + 8, 3)
            g.sdtree_pdf = sp.data_ptr()
            if wp is not None:
                g.wo_pdf = wp.data_ptr()
            if wt is not None:
                g.weight = L.Vec3(wt.data_ptr(), wt.data_ptr() + 4, wt.data_ptr() + 8, 3)
            if ep is not None:
                g.sdtree_pdf_em = ep.data_ptr()
        outs = (d, sp, wp, wt) if em_dir is None else (d, sp, wp, wt, ep)
        return _PreparedGuided(self, g, n, self._flags(b), b, outs)

    def mis_nee(self, bsdf_pdf_em, sdtree_pdf_em, pdf_with_delta, pdf_without_delta, ds_pdf, ds_delta,
                bsdf_sampling_fraction, iteration):
        b = _Buf()
        n = _n_of(bsdf_pdf_em)
        ins = [b.arr(x, np.float32) for x in (bsdf_pdf_em, sdtree_pdf_em, pdf_with_delta, pdf_without_delta, ds_pdf)]
        dl = b.arr(ds_delta, np.uint8)
        s, sp = b.new((n,), np.float32)
        m, mp = b.new((n,), np.float32)
        self._ck(self._lib.sdt_mis_nee(self._h, n, *ins, dl, float(bsdf_sampling_fraction), int(iteration), sp, mp,
                                       self._flags(b), b.stream()))
        return s, m

    def mis_mixture(self, bsdf_pdf, sdtree_pdf, bsdf_value, do_mis, bsdf_sampling_fraction):
        b = _Buf()
        n = _n_of(bsdf_pdf)
        bp = b.arr(bsdf_pdf, np.float32)
        sp = b.arr(sdtree_pdf, np.float32)
        bv = b.vec(bsdf_value, 3)
        dm = b.arr(do_mis, np.uint8)
        wo, wop = b.new((n,), np.float32)
        w, wp = b.new((n, 3), np.float32)
        wv = L.Vec3(wp, wp + 4, wp + 8, 3)
        self._ck(self._lib.sdt_mis_mixture(self._h, n, bp, sp, C.byref(bv), dm, float(bsdf_sampling_fraction), wop,
                                           C.byref(wv), self._flags(b), b.stream()))
        return wo, w

    def dir_to_canonical(self, direction):
        """dirToCanonical (src/common.py:132-158): (n,3) -> (n,2)"""
        b = _Buf()
        n = _n_of(direction)
        dv = b.vec(direction, 3)
        o, op = b.new((n, 2), np.float32)
        self._ck(self._lib.sdt_dir_to_canonical(self._h, C.byref(dv), n, op, self._flags(b), b.stream()))
        return o

    def canonical_to_dir(self, pos2):
        """canonicalToDir (src/common.py:100-129): (n,2) -> (n,3)"""
        b = _Buf()
        n = _n_of(pos2)
        pv = b.vec(pos2, 2)
        o, op = b.new((n, 3), np.float32)
        ov = L.Vec3(op, op + 4, op + 8, 3)
        self._ck(self._lib.sdt_canonical_to_dir(self._h, C.byref(pv), n, C.byref(ov), self._flags(b), b.stream()))
        return o

    # ---- splat into current -----------------------------------------------------------
    def splat_records(self, position, direction, radiance, wo_pdf, radiance_nee=None, direction_nee=None, active=None):
        """KDTree.addDataPropagate on already-filtered records"""
        b = _Buf()
        n = _n_of(position)
        r = L.Records()
        r.position = b.vec(position, 3)
        r.direction = b.vec(direction, 2)
        r.radiance = b.arr(radiance, np.float32)
        r.wo_pdf = b.arr(wo_pdf, np.float32)
        r.radiance_nee = b.vec(radiance_nee, 3)
        r.direction_nee = b.vec(direction_nee, 2)
        r.active = b.arr(active, np.uint8)
        self._ck(self._lib.sdt_splat_records(self._h, C.byref(r), n, self._flags(b), b.stream()))

    def splat_path_data(self, max_depth, l_final, throughput_radiance, throughput_bsdf, bsdf, position, direction,
                        wo_pdf, radiance_nee=None, direction_nee=None, active=None, want_radiance=False):
        """processPathData + scatterDataIntoSDTree + addDataPropagate in one pass"""
        b = _Buf()
        n = _n_of(position)
        p = L.PathData()
        p.slots = n
        p.max_depth = int(max_depth)
        p.l_final = b.vec(l_final, 3)
        p.throughput_radiance = b.vec(throughput_radiance, 3)
        p.throughput_bsdf = b.vec(throughput_bsdf, 3)
        p.bsdf = b.vec(bsdf, 3)
        p.position = b.vec(position, 3)
        p.direction = b.vec(direction, 2)
        p.wo_pdf = b.arr(wo_pdf, np.float32)
        p.radiance_nee = b.vec(radiance_nee, 3)
        p.direction_nee = b.vec(direction_nee, 2)
        p.active = b.arr(active, np.uint8)
        rad = None
        if want_radiance:
            rad, rp = b.new((n,), np.float32)
            p.radiance_out = rp
        self._ck(self._lib.sdt_splat_path_data(self._h, C.byref(p), self._flags(b), b.stream()))
        return rad

    # ---- refine -------------------------------------------------------------------------
    def set_iteration_threshold(self, iteration):
        """KDTree.setRefinementThreshold: maxLeafSize = 12000 * sqrt(2**iteration)"""
        self._ck(self._lib.sdt_set_iteration_threshold(self._h, int(iteration)))

    def set_max_leaf_size(self, v):
        self._ck(self._lib.sdt_set_max_leaf_size(self._h, float(np.float32(v))))

    def refine(self, iteration=None, kd=True, quad=True, sync=False, stream=None, check=False):
        """refineAndPrepareSDTreeForNextIteration, on the device.  check=True waits for it and raises when an arena
        overflowed (the device-side sticky error flag): the refined tree would be truncated, i.e. not the reference's"""
        if iteration is not None:
            self.set_iteration_threshold(iteration)
        flags = (0 if kd else L.SDT_REFINE_NO_KD) | (0 if quad else L.SDT_REFINE_NO_QUAD) | (L.SDT_SYNC if sync else 0)
        self._ck(self._lib.sdt_refine(self._h, flags, self._stream(stream)))
        if check:
            self.check_error()

    def check_error(self):
        """raises SDTreeError when the device-side error flag is set (bit 0: spatial arena, bit 1: quadtree arena
        exhausted; bit 2: a refine scan stalled -- an internal error, the tree is invalid)"""
        e = self.sizes()['error']
        if e & 4:
            raise SDTreeError(-2,    # SDT_ERR_CUDA
                              f"a refine scan gave up waiting for another block (device error flag {e}): the tree is invalid")
        if e & 8:
            raise SDTreeError(-1,    # SDT_ERR_INVALID
                              f"hint_records() was given fewer records than were splatted (device error flag {e}): "
                              "spatial splits are missing from the tree")
        if e:
            what = {1: "spatial node arena", 2: "quadtree node arena"}.get(e, "arena")
            raise SDTreeError(-3,    # SDT_ERR_CAPACITY
                              f"refine ran out of the {what} (device error flag {e}): the tree is truncated; "
                              "create the SDTree with a larger kd_capacity / quad_capacity")

    def reset_stats(self, stream=None):
        self._ck(self._lib.sdt_reset_stats(self._h, self._stream(stream)))

    # ---- multi-GPU ------------------------------------------------------------------------
    def comm_unique_id(self):
        buf = (C.c_char * 128)()
        rc = self._lib.sdt_comm_unique_id(buf)
        if rc != 0:
            raise SDTreeError(rc, (self._lib.sdt_last_error(None) or b"").decode())
        return bytes(buf)

    def comm_init(self, unique_id, rank, nranks):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        self._ck(self._lib.sdt_comm_init(self._h, buf, int(rank), int(nranks)))

    def allreduce(self, stream=None, records_all_ranks=None):
        """one NCCL all-reduce over current's leaf statistics; records_all_ranks: see hint_records"""
        self._ck(self._lib.sdt_allreduce(self._h, self._stream(stream)))
        if records_all_ranks is not None:
            self.hint_records(records_all_ranks)

    def hint_records(self, records_all_ranks):
        """upper bound of the records splatted into `current` over all ranks since its statistics were zero (e.g. passes x
        rays x max_depth): lets the refine skip the split rounds no leaf can reach (sdt_hint_records)"""
        self._ck(self._lib.sdt_hint_records(self._h, int(records_all_ranks)))

    def stat_buffers(self):
        """(ptr_q_energy, n_quad, ptr_kd_count, n_kd) of current's statistics"""
        q, k = C.c_void_p(), C.c_void_p()
        nq, nk = C.c_uint32(), C.c_uint32()
        self._ck(self._lib.sdt_stat_buffers(self._h, C.byref(q), C.byref(nq), C.byref(k), C.byref(nk)))
        return q.value, int(nq.value), k.value, int(nk.value)

    # ---- tree exchange in the reference's npz schema ------------------------------------------
    def upload(self, d):
        """d: mapping with the 23 keys of KDTree.saveToFile (src/kdtree.py:575-602)"""
        f32 = lambda k, sh=None: np.ascontiguousarray(np.asarray(d[k], np.float32).reshape(sh) if sh else np.asarray(d[k], np.float32))
        u32 = lambda k: np.ascontiguousarray(np.asarray(d[k]).astype(np.uint32))
        u8 = lambda k: np.ascontiguousarray(np.asarray(d[k]).astype(np.uint8))
        keep = dict(
            kd_bbox_min=f32('kdtree_bbox_min', (-1, 3)), kd_bbox_max=f32('kdtree_bbox_max', (-1, 3)),
            kd_depth=u32('kdtree_depth'), kd_vert_count=f32('kdtree_vertCount'), kd_is_leaf=u8('kdtree_isLeaf'),
            kd_quad_root=u32('kdtree_quadTreeRootIndex'), kd_child_left=u32('kdtree_child_left_index'),
            kd_child_right=u32('kdtree_child_right_index'), q_root_node=u32('quadtree_rootNodeIndex'),
            q_bbox_min=f32('quadtree_bbox_min', (-1, 2)), q_bbox_max=f32('quadtree_bbox_max', (-1, 2)),
            q_depth=u32('quadtree_depth'), q_irradiance=f32('quadtree_irradiance'), q_is_leaf=u8('quadtree_isLeaf'),
            q_threshold=f32('quadtree_refinementThreshold'))
        qc = [u32(f'quadtree_child_{k}_index') for k in (1, 2, 3, 4)]
        a = L.Arrays()
        a.n_kd = keep['kd_depth'].shape[0]
        a.n_quad = keep['q_depth'].shape[0]
        a.n_roots = keep['q_root_node'].shape[0]
        a.kd_max_leaf_size = float(np.float32(np.asarray(d['kdtree_maxLeafSize']).item()))
        a.kd_max_depth = int(np.asarray(d['kdtree_maxDepth']).item())
        a.quad_max_depth = int(np.asarray(d['quadtree_maxDepth']).item())
        a.quad_store_nee = int(bool(np.asarray(d['quadtree_isStoreNEERadiance']).item()))
        for k, v in keep.items():
            setattr(a, k, v.ctypes.data)
        for k in range(4):
            a.q_child[k] = qc[k].ctypes.data
        self._ck(self._lib.sdt_upload(self._h, C.byref(a)))
        self.store_nee = bool(a.quad_store_nee)

    def upload_stats(self, q_irradiance=None, kd_vert_count=None):
        q = None if q_irradiance is None else np.ascontiguousarray(q_irradiance, np.float32)
        k = None if kd_vert_count is None else np.ascontiguousarray(kd_vert_count, np.float32)
        self._ck(self._lib.sdt_upload_stats(self._h, None if q is None else q.ctypes.data, None if k is None else k.ctypes.data))

    def download(self, which=L.SDT_TREE_PREV):
        """-> dict with the 23 npz keys (quadtree in the canonical clearTreeUnusedNode layout)"""
        s = self.sizes()
        nk, nq, R = s['n_kd'], s['n_quad'], s['n_roots']
        o = dict(
            kd_bbox_min=np.zeros((nk, 3), np.float32), kd_bbox_max=np.zeros((nk, 3), np.float32),
            kd_depth=np.zeros(nk, np.uint32), kd_vert_count=np.zeros(nk, np.float32), kd_is_leaf=np.zeros(nk, np.uint8),
            kd_quad_root=np.zeros(nk, np.uint32), kd_child_left=np.zeros(nk, np.uint32), kd_child_right=np.zeros(nk, np.uint32),
            q_root_node=np.zeros(R, np.uint32), q_bbox_min=np.zeros((nq, 2), np.float32), q_bbox_max=np.zeros((nq, 2), np.float32),
            q_depth=np.zeros(nq, np.uint32), q_irradiance=np.zeros(nq, np.float32), q_is_leaf=np.zeros(nq, np.uint8),
            q_threshold=np.zeros(nq, np.float32))
        qc = [np.zeros(nq, np.uint32) for _ in range(4)]
        a = L.Arrays()
        a.n_kd, a.n_quad, a.n_roots = nk, nq, R
        for k, v in o.items():
            setattr(a, k, v.ctypes.data)
        for k in range(4):
            a.q_child[k] = qc[k].ctypes.data
        self._ck(self._lib.sdt_download(self._h, int(which), C.byref(a)))
        return dict(
            kdtree_maxLeafSize=np.asarray(np.float32(a.kd_max_leaf_size)), kdtree_maxDepth=np.asarray(int(a.kd_max_depth)),
            kdtree_bbox_min=o['kd_bbox_min'], kdtree_bbox_max=o['kd_bbox_max'], kdtree_depth=o['kd_depth'],
            kdtree_vertCount=o['kd_vert_count'], kdtree_isLeaf=o['kd_is_leaf'].astype(bool),
            kdtree_quadTreeRootIndex=o['kd_quad_root'], kdtree_child_left_index=o['kd_child_left'],
            kdtree_child_right_index=o['kd_child_right'],
            quadtree_maxDepth=np.asarray(int(a.quad_max_depth)), quadtree_isStoreNEERadiance=np.asarray(bool(a.quad_store_nee)),
            quadtree_rootNodeIndex=o['q_root_node'], quadtree_bbox_min=o['q_bbox_min'], quadtree_bbox_max=o['q_bbox_max'],
            quadtree_depth=o['q_depth'], quadtree_irradiance=o['q_irradiance'], quadtree_isLeaf=o['q_is_leaf'].astype(bool),
            quadtree_refinementThreshold=o['q_threshold'],
            quadtree_child_1_index=qc[0], quadtree_child_2_index=qc[1], quadtree_child_3_index=qc[2], quadtree_child_4_index=qc[3])

    def save_npz(self, file_name, which=L.SDT_TREE_PREV):
        """KDTree.saveToFile (src/kdtree.py:539-602)"""
        np.savez_compressed(file_name, **self.download(which))

    def load_npz(self, file_name):
        """KDTree.loadFromFile + loadSDTreeFromFile (src/kdtree.py:156-170, integrator :597-608)"""
        d = dict(np.load(file_name))
        d['kdtree_maxLeafSize'] = np.asarray(int(np.asarray(d['kdtree_maxLeafSize']).item()))   # :161 truncates
        self.upload(d)
