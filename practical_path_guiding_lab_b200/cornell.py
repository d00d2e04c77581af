"""Stand-in vertex source for images without Mitsuba: a wavefront path
tracer for the analytic Cornell box of /root/reference/scenes/cornell-box/scene.xml (five
rectangles, two cubes, one area light, two-sided diffuse BSDFs, perspective sensor) written with
torch tensor ops, driving `PathGuidingCore` exactly where the reference integrator calls its
SD-tree (src/path_guiding_integrator.py:126-431).  It replaces `scene.ray_intersect`,
`bsdf.sample/eval_pdf` and `scene.sample_emitter_direction` only -- every tree operation (pdf,
sample, MIS mixture, record splat, refine) goes through libsdtree.so.  With Mitsuba present the
registered integrator (integrator.py) takes over and this file is not used.
"""
import math

import numpy as np
import torch

from .integrator import PathGuidingCore

# scene.xml:60-120 -- to_world matrices (row-major) and reflectances
_RECTS = {
    "Floor": ("-4.37114e-008 1 4.37114e-008 0 0 -8.74228e-008 2 0 1 4.37114e-008 1.91069e-015 0 0 0 0 1", (0.725, 0.71, 0.68)),
    "Ceiling": ("-1 7.64274e-015 -1.74846e-007 0 8.74228e-008 8.74228e-008 -2 2 0 -1 -4.37114e-008 0 0 0 0 1", (0.725, 0.71, 0.68)),
    "BackWall": ("1.91069e-015 1 1.31134e-007 0 1 3.82137e-015 -8.74228e-008 1 -4.37114e-008 1.31134e-007 -2 -1 0 0 0 1", (0.725, 0.71, 0.68)),
    "RightWall": ("4.37114e-008 -1.74846e-007 2 1 1 3.82137e-015 -8.74228e-008 1 3.82137e-015 1 2.18557e-007 0 0 0 0 1", (0.14, 0.45, 0.091)),
    "LeftWall": ("-4.37114e-008 8.74228e-008 -2 -1 1 3.82137e-015 -8.74228e-008 1 0 -1 -4.37114e-008 0 0 0 0 1", (0.63, 0.065, 0.05)),
}
_CUBES = {
    "ShortBox": ("0.0851643 0.289542 1.31134e-008 0.328631 3.72265e-009 1.26563e-008 -0.3 0.3 -0.284951 0.0865363 5.73206e-016 0.374592 0 0 0 1", (0.725, 0.71, 0.68)),
    "TallBox": ("0.286776 0.098229 -2.29282e-015 -0.335439 -4.36233e-009 1.23382e-008 -0.6 0.6 -0.0997984 0.282266 2.62268e-008 -0.291415 0 0 0 1", (0.725, 0.71, 0.68)),
}
_LIGHT = ("0.235 -1.66103e-008 -7.80685e-009 -0.005 -2.05444e-008 3.90343e-009 -0.0893 1.98 2.05444e-008 0.19 8.30516e-009 -0.03 0 0 0 1", (17.0, 12.0, 4.0))
_CAMERA = "-1 0 0 0 0 1 0 1 0 0 -1 6.8 0 0 0 1"
_FOV = 19.5
LUM = (0.212671, 0.715160, 0.072169)


def _mat(s):
    return np.array([float(v) for v in s.split()], np.float64).reshape(4, 4)


def _quads():
    """-> centre (Q,3), eu (Q,3), ev (Q,3), albedo (Q,3), emission (Q,3); the light is the LAST quad"""
    c, eu, ev, alb, em = [], [], [], [], []
    for m, a in _RECTS.values():
        M = _mat(m)
        c.append(M[:3, 3]); eu.append(M[:3, 0]); ev.append(M[:3, 1]); alb.append(a); em.append((0, 0, 0))
    for m, a in _CUBES.values():
        M = _mat(m)
        for ax in range(3):
            o1, o2 = (ax + 1) % 3, (ax + 2) % 3
            for sgn in (1.0, -1.0):
                c.append(M[:3, 3] + sgn * M[:3, ax]); eu.append(M[:3, o1]); ev.append(sgn * M[:3, o2]); alb.append(a); em.append((0, 0, 0))
    M = _mat(_LIGHT[0])
    c.append(M[:3, 3]); eu.append(M[:3, 0]); ev.append(M[:3, 1]); alb.append((0, 0, 0)); em.append(_LIGHT[1])
    return [np.asarray(x, np.float32) for x in (c, eu, ev, alb, em)]


class CornellBox:
    def __init__(self, width=256, height=256, max_depth=30, rr_depth=8, device="cuda", lib_path=None,
                 kd_capacity=0, quad_capacity=0):
        self.W, self.H = int(width), int(height)
        self.dev = torch.device(device)
        host = self.dev.type != "cuda"
        c, eu, ev, alb, em = _quads()
        t = lambda a: torch.from_numpy(a).to(self.dev)
        self.qc, self.qu, self.qv, self.alb, self.em = t(c), t(eu), t(ev), t(alb), t(em)
        n = torch.linalg.cross(self.qu, self.qv)
        self.qn = n / n.norm(dim=1, keepdim=True)
        self.qu2 = (self.qu * self.qu).sum(1)
        self.qv2 = (self.qv * self.qv).sum(1)
        self.light = self.qc.shape[0] - 1
        self.light_area = float(4.0 * self.qu[self.light].norm() * self.qv[self.light].norm())
        self.cam = torch.from_numpy(_mat(_CAMERA).astype(np.float32)).to(self.dev)
        self.lum = torch.tensor(LUM, device=self.dev)
        dev_index = self.dev.index if self.dev.type == "cuda" and self.dev.index is not None else 0
        self.core = PathGuidingCore(max_depth=max_depth, rr_depth=rr_depth, device=dev_index, lib_path=lib_path,
                                    kd_capacity=kd_capacity, quad_capacity=quad_capacity)
        self._host = host
        pts = torch.cat([self.qc + a * self.qu + b * self.qv for a in (-1, 1) for b in (-1, 1)])
        self.bbox_min = pts.min(0).values.cpu().numpy()
        self.bbox_max = pts.max(0).values.cpu().numpy()
        self.sumL = None
        self.sumL2 = None
        self._gL = self._gL2 = None
        self.p0, self.p1 = 0, self.W * self.H          # this rank's tile of the film: pixels [p0, p1) (driver.TorchDistRanks, tile sharding)
        self._rank = 0
        self._pin = None

    def set_tile(self, rank, world):
        """tile sharding (main.py:208-218 split by image tiles): from now on this rank renders pixels
        [p0, p1) -- a band of rows -- of the passes it is given; lanes keep globally unique RNG keys (generator seed and
        lane offset per tile).  The driver sets it per iteration (driver.shard_plan)."""
        P = self.W * self.H
        self.p0, self.p1 = rank * P // world, (rank + 1) * P // world
        self._rank = int(rank)
        self.core.numRays = self.p1 - self.p0             # the per-pass record buffers follow the tile (resetRayPathData)
        self.core.array_size = self.core.numRays * self.core.max_depth

    # ---- what main.py does before the loop (main.py:45-64) -----------------------------------
    def setup(self, sdTreeMaxDepth=20, quadTreeMaxDepth=20, isStoreNEERadiance=True, bsdfSamplingFraction=0.5):
        eps = 1e-4
        self.core.setup(self.p1 - self.p0, self.bbox_min - eps, self.bbox_max + eps, sdTreeMaxDepth, quadTreeMaxDepth,
                        isStoreNEERadiance, bsdfSamplingFraction)
        self.resetVarianceCounter()

    def resetVarianceCounter(self):
        self.sumL = torch.zeros(self.W * self.H, 3, device=self.dev)
        self.sumL2 = torch.zeros(self.W * self.H, 3, device=self.dev)
        self._gL = self._gL2 = None          # all-rank sums (multi-GPU), else the local counters are the totals

    def _sums(self):
        return (self.sumL, self.sumL2) if self._gL is None else (self._gL, self._gL2)

    # buffers handed to the library: torch CUDA tensors, or numpy views of CPU tensors (host emulation in tests)
    def _x(self, t):
        if t is None:
            return None
        return t.contiguous().numpy() if self._host else t.contiguous()

    def _t(self, a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.dev) if isinstance(a, np.ndarray) else a

    # ---- geometry ------------------------------------------------------------------------------
    def intersect(self, o, d, tmax=None):
        """closest hit of rays (n,3)+(n,3) -> (t (n,), quad (n,) long, valid (n,) bool)"""
        dn = d @ self.qn.T                                        # (n,Q)
        num = ((self.qc[None] - o[:, None]) * self.qn[None]).sum(2)
        t = num / dn
        p = o[:, None] + t[..., None] * d[:, None] - self.qc[None]
        a = (p * self.qu[None]).sum(2) / self.qu2[None]
        b = (p * self.qv[None]).sum(2) / self.qv2[None]
        ok = (t > 1e-4) & (a.abs() <= 1) & (b.abs() <= 1) & torch.isfinite(t)
        t = torch.where(ok, t, torch.full_like(t, float("inf")))
        tt, q = t.min(1)
        valid = torch.isfinite(tt)
        if tmax is not None:
            valid = valid & (tt < tmax)
        return tt, q, valid

    def camera_rays(self, spp, gen):
        n = (self.p1 - self.p0) * spp
        pix = torch.arange(self.p0, self.p1, device=self.dev).repeat_interleave(spp)   # samples of a pixel adjacent
        jit = torch.rand(n, 2, device=self.dev, generator=gen)
        sx = ((pix % self.W).float() + jit[:, 0]) / self.W
        sy = ((pix // self.W).float() + jit[:, 1]) / self.H
        tf = math.tan(math.radians(_FOV) / 2)
        dl = torch.stack([(1 - 2 * sx) * tf, (1 - 2 * sy) * tf * self.H / self.W, torch.ones_like(sx)], 1)
        d = dl @ self.cam[:3, :3].T
        d = d / d.norm(dim=1, keepdim=True)
        o = self.cam[:3, 3].expand(n, 3).contiguous()
        return o, d, pix

    @staticmethod
    def _frame(nrm):
        a = torch.where(nrm[:, 0:1].abs() > 0.9, torch.tensor([0.0, 1.0, 0.0], device=nrm.device), torch.tensor([1.0, 0.0, 0.0], device=nrm.device))
        s = torch.linalg.cross(nrm, a.expand_as(nrm))
        s = s / s.norm(dim=1, keepdim=True)
        return s, torch.linalg.cross(nrm, s)

    # ---- one mi.render(scene, spp, seed) call: PathGuidingIntegrator.sample on W*H*spp lanes -----
    def render(self, spp, seed):
        core = self.core
        dev = self.dev
        gen = torch.Generator(device=dev).manual_seed(int(seed) + self._rank * 0x9E3779B1)      # rank 0 / one GPU: the plain seed
        rnd = lambda *s: torch.rand(*s, device=dev, generator=gen)
        o, d, pix = self.camera_rays(spp, gen)
        n = o.shape[0]
        f = core.bsdfSamplingFraction
        thr = torch.ones(n, 3, device=dev)
        L = torch.zeros(n, 3, device=dev)
        depth = torch.zeros(n, dtype=torch.long, device=dev)
        active = torch.ones(n, dtype=torch.bool, device=dev)
        ray_index = torch.arange(n, device=dev)
        prev_p = o.clone()
        prev_pdf = torch.ones(n, device=dev)
        prev_delta = torch.ones(n, dtype=torch.bool, device=dev)
        training = not core.isFinalIter
        if training:
            assert spp == 1, "training passes are 1 spp (array_size = numRays * max_depth)"
            core.resetRayPathData(self._x(L))
        core._pass_seed += 1
        light_n = self.qn[self.light]
        bounce = 0
        lane0 = self.p0 * spp
        # "any lane still active?" is read back without stalling the queue: the flag of bounce b is copied to pinned memory
        # asynchronously and looked at two bounces later (at most two fully masked bounces run past the end of the last path)
        if self._pin is None and dev.type == "cuda":
            self._pin = torch.zeros(core.max_depth + 2, dtype=torch.bool).pin_memory()
        pending = []
        while bounce < core.max_depth:
            if dev.type == "cuda":
                self._pin[bounce].copy_(active.any(), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                pending.append((ev, bounce))
                stop = False
                while pending and (len(pending) > 2 or pending[0][0].query()):
                    e, b = pending.pop(0)
                    e.synchronize()
                    if not bool(self._pin[b]):
                        stop = True
                if stop:
                    break
            elif not bool(active.any()):
                break
            bounce += 1
            t, q, valid = self.intersect(o, d)
            valid = valid & active
            p = o + t.nan_to_num(posinf=0.0)[:, None] * d
            ng = self.qn[q]
            nf = torch.where(((ng * d).sum(1) < 0)[:, None], ng, -ng)           # two-sided: face the incoming ray
            alb = self.alb[q]
            # -- emission seen by the ray, MIS against emitter sampling (:189-200)
            on_light = valid & (q == self.light) & ((light_n * d).sum(1) < 0)
            dist2 = ((p - prev_p) ** 2).sum(1)
            cosl = (-(light_n * d).sum(1)).clamp_min(1e-20)
            em_pdf = torch.where(on_light & ~prev_delta, dist2 / (self.light_area * cosl), torch.zeros_like(t))
            a2 = prev_pdf * prev_pdf
            mis = torch.where(prev_pdf > 0, a2 / (em_pdf * em_pdf + a2), torch.zeros_like(a2)).nan_to_num(0.0)
            Le = thr * (mis * on_light)[:, None] * self.em[self.light][None]
            active_next = valid & (depth + 1 < core.max_depth)
            # -- next event estimation (:208-256)
            uv = rnd(n, 2) * 2 - 1
            lp = self.qc[self.light] + uv[:, 0:1] * self.qu[self.light] + uv[:, 1:2] * self.qv[self.light]
            wl = lp - p
            ld2 = (wl * wl).sum(1)
            ld = ld2.sqrt()
            wl = wl / ld[:, None]
            cos_l = -(wl * light_n).sum(1)
            ds_pdf = torch.where(cos_l > 0, ld2 / (self.light_area * cos_l.clamp_min(1e-20)), torch.zeros_like(ld2))
            active_em = active_next & (ds_pdf > 0)
            st, _, sv = self.intersect(p + nf * 1e-4, wl, tmax=ld - 2e-4)
            vis = active_em & ~sv
            em_weight = torch.where(vis[:, None], self.em[self.light][None] / ds_pdf.clamp_min(1e-20)[:, None], torch.zeros(n, 3, device=dev))
            cos_s = (wl * nf).sum(1)
            bsdf_pdf_em = torch.where(active_em & (cos_s > 0), cos_s / math.pi, torch.zeros_like(cos_s))
            bsdf_val_em = alb * bsdf_pdf_em[:, None]
            act_sd_em = active_em & core.guiding
            # -- continuation: BSDF sample (:272-281); the tree is asked once per vertex, below
            u1, u2 = rnd(n), rnd(n)
            r = u1.sqrt()
            ph = 2 * math.pi * u2
            s_, t_ = self._frame(nf)
            cz = (1 - u1).clamp_min(0).sqrt()
            wo = (r * ph.cos())[:, None] * s_ + (r * ph.sin())[:, None] * t_ + cz[:, None] * nf
            bsdf_pdf = torch.where(active_next, cz / math.pi, torch.zeros_like(cz))
            bsdf_value = alb * bsdf_pdf[:, None]
            woPdf = bsdf_pdf.clone()
            bsdf_weight = torch.where((bsdf_pdf > 0)[:, None], alb, torch.zeros_like(alb))
            do_mis = active_next & core.guiding
            choose_u = rnd(n)
            no_delta = self._x(torch.zeros(n, dtype=torch.uint8, device=dev))
            if core.guiding:
                # ONE library call / spatial descent per vertex: tree pdf of the emitter direction (:244) + guided / BSDF choice (:283-307)
                sd_em, mode, sd_dir, sd_pdf = core.bounce(self._x(p), self._x(wl), self._x(act_sd_em), self._x(wo), self._x(do_mis), self._x(choose_u),
                                                          seed=(core._pass_seed * 1315423911 + bounce) & 0xFFFFFFFF, lane_offset=lane0)
                mis_em = self._t(core.nee_mis_from_pdf(sd_em, self._x(bsdf_pdf_em), self._x(bsdf_pdf_em), self._x(bsdf_pdf_em), self._x(ds_pdf), no_delta))
            else:
                mis_em = self._t(core.nee_mis(self._x(p), self._x(wl), self._x(act_sd_em), self._x(bsdf_pdf_em), self._x(bsdf_pdf_em),
                                              self._x(bsdf_pdf_em), self._x(ds_pdf), no_delta))
            Lr_dir = thr * mis_em[:, None] * bsdf_val_em * em_weight
            L = L + Le + Lr_dir
            # -- one-sample mixture of the guided / BSDF choice (:301-311)
            if core.guiding:
                mode, sd_dir, sd_pdf = self._t(mode), self._t(sd_dir), self._t(sd_pdf)
                g = mode == 1
                wo = torch.where(g[:, None], sd_dir, wo)
                cg = (wo * nf).sum(1)
                p2 = torch.where(cg > 0, cg / math.pi, torch.zeros_like(cg))
                bsdf_pdf = torch.where(g, p2, bsdf_pdf)
                bsdf_value = torch.where(g[:, None], alb * p2[:, None], bsdf_value)
                mw, mwt = core.mixture(self._x(bsdf_pdf), self._x(sd_pdf), self._x(bsdf_value), self._x(do_mis))
                mw, mwt = self._t(mw), self._t(mwt)
                woPdf = torch.where(do_mis, mw, woPdf)
                bsdf_weight = torch.where(do_mis[:, None], mwt.nan_to_num(0.0, 0.0, 0.0), bsdf_weight)
            # -- record (:318-346)
            if training:
                core.store_vertex(self._x(ray_index), self._x(depth), self._x(valid), self._x(p), self._x(wo), self._x(bsdf_weight),
                                  self._x(thr), self._x(L), self._x(mis_em[:, None] * bsdf_val_em * em_weight), self._x(wl), self._x(woPdf))
            # -- next ray, Russian roulette (:352-381)
            side = torch.where((wo * nf).sum(1) >= 0, 1e-4, -1e-4)
            o = p + nf * side[:, None]
            d = wo
            thr = thr * bsdf_weight
            prev_p = torch.where(valid[:, None], p, prev_p)
            prev_pdf = woPdf
            prev_delta = torch.zeros(n, dtype=torch.bool, device=dev)
            tmax_ = thr.max(1).values
            active_next = active_next & (tmax_ != 0)
            rr_prob = tmax_.clamp_max(0.95)
            rr_active = depth >= core.rr_depth
            active_next = active_next & (~rr_active | (rnd(n) < rr_prob))
            active = active_next
            depth = depth + valid.long()
        if training:
            core.end_of_pass(self._x(L))
        # film + variance counters (:400-429)
        Ls = L.view(self.p1 - self.p0, spp, 3)
        self.sumL[self.p0:self.p1] += Ls.sum(1)
        self.sumL2[self.p0:self.p1] += (Ls * Ls).sum(1)
        if self.p1 - self.p0 == self.W * self.H:
            return Ls.mean(1).view(self.H, self.W, 3)
        img = torch.zeros(self.W * self.H, 3, device=dev)             # this rank's tile; the driver sums the ranks' images
        img[self.p0:self.p1] = Ls.mean(1)
        return img.view(self.H, self.W, 3)

    # ---- src/path_guiding_integrator.py:503-550 ------------------------------------------------
    def _lum(self, x):
        return (x * self.lum).sum(-1)

    def computeMSE(self, spp, groundTruth):
        sL, _ = self._sums()
        mse = self._lum((sL / spp - groundTruth.view(-1, 3)) ** 2).clamp_max(10000)
        return float(mse.mean())

    def computeVariance(self, spp, groundTruth=None):
        sL, sL2 = self._sums()
        if groundTruth is not None:
            v = self._lum(sL2 / spp - groundTruth.view(-1, 3) ** 2).clamp_max(10000)
            return float(v.mean()) / spp
        Lm = sL / spp
        v = float(self._lum(sL2 / spp - Lm * Lm).clamp_max(10000).mean())
        return v / (spp - 1) if spp > 1 else v

    # ---- multi-GPU (driver.TorchDistRanks) -------------------------------------------------------
    def zero_image(self):
        return torch.zeros(self.H, self.W, 3, device=self.dev)

    def comm_init(self, dist):
        ids = [self.core.tree.comm_unique_id() if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        self.core.tree.comm_init(ids[0], dist.get_rank(), dist.get_world_size())

    def allreduce_statistics(self, dist, passes=None):
        """one exchange per iteration: SD-tree statistics (sdt_allreduce, NCCL) + variance counters.  passes = passes of
        the iteration over all ranks: passes x pixels x max_depth record slots bound what any spatial leaf can have
        counted, which every rank knows without asking the others (sdt_hint_records)"""
        bound = None if passes is None else int(passes) * self.W * self.H * int(self.core.max_depth)
        if self._host:
            # host emulation (CPU tests, gloo): the same exchange through sdt_stat_buffers + torch.distributed
            import ctypes
            qp, nq, kp, nk = self.core.tree.stat_buffers()
            q = np.ctypeslib.as_array(ctypes.cast(qp, ctypes.POINTER(ctypes.c_float)), shape=(nq,))
            k = np.ctypeslib.as_array(ctypes.cast(kp, ctypes.POINTER(ctypes.c_float)), shape=(nk,))
            buf = torch.from_numpy(np.concatenate([q, k]))
            dist.all_reduce(buf)
            q[:] = buf[:nq].numpy()
            k[:] = buf[nq:].numpy()
            if bound is not None:
                self.core.tree.hint_records(bound)
        else:
            self.core.tree.allreduce(torch.cuda.current_stream().cuda_stream, records_all_ranks=bound)
        self._gL, self._gL2 = self.sumL.clone(), self.sumL2.clone()      # local counters stay local
        dist.all_reduce(self._gL)
        dist.all_reduce(self._gL2)

    # forwarded integrator interface used by the driver
    def setIteration(self, iteration, isFinalIter):
        self.core.setIteration(iteration, isFinalIter)

    def refineAndPrepareSDTreeForNextIteration(self):
        self.core.refineAndPrepareSDTreeForNextIteration()

    def saveSDTreeToFile(self, f):
        self.core.saveSDTreeToFile(f)

    def loadSDTreeFromFile(self, f):
        self.core.loadSDTreeFromFile(f)

    def saveSDTreeOBJ(self, f):
        self.core.saveSDTreeOBJ(f)
