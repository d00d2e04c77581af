"""Drop-in for /root/reference/src/path_guiding_integrator.py: the same integrator interface
(setup / setIteration / sample / refineAndPrepareSDTreeForNextIteration / save / load ...,
registered under the same plugin name), with the SD-tree replaced by libsdtree.so.

Two layers:
  * `PathGuidingCore` -- everything that does not need Mitsuba: the (prev, current) tree pair
    behind one sdt_handle, the per-pass record buffers (SurfaceInteractionRecord,
    src/common.py:14-40, only the fields the tree consumes), the guided bounce, the NEE MIS,
    the end-of-pass splat and the per-iteration refine.  Works on torch CUDA tensors (device
    pointers) or numpy arrays (host pointers) -- the array type of the inputs decides.
  * `PathGuidingIntegrator(mi.SamplingIntegrator)` -- defined and registered as
    'path_guiding_integrator' only when `mitsuba` imports (it does not in this image, SURVEY.md
    header item 2).  Ray intersection, BSDFs and emitters stay with Mitsuba's cuda variant; the
    recorded mi.Loop of the reference becomes a wavefront loop with one sdt_* call per tree
    operation.  Dr.Jit <-> torch exchange is zero-copy (`.torch()` / `dr.cuda` DLPack).
"""
import numpy as np

from .sdtree import SDTree
from . import _lib as L

EPSILON = 0.00001      # src/path_guiding_integrator.py:14


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class PathGuidingCore:
    FUSED_BOUNCE_MAX_LANES = 1 << 22      # bounce(): one fused library call up to this many lanes, two calls beyond

    def __init__(self, max_depth=30, rr_depth=8, device=0, lib_path=None, kd_capacity=0, quad_capacity=0):
        # props checks of src/path_guiding_integrator.py:34-41
        if max_depth < 0 and max_depth != -1:
            raise Exception("\"max_depth\" must be set to -1 (infinite) or a value >= 0")
        if rr_depth < 0:
            raise Exception("\"rr_depth\" must be set to >= 0")
        self.max_depth = max_depth
        self.rr_depth = rr_depth
        self.numRays = 0
        self.array_size = 0
        self.isStoreNEERadiance = False
        self.bsdfSamplingFraction = 0.5
        self.iteration = 0
        self.isFinalIter = False
        self.tree = None
        self.record = None
        self._device = device
        self._lib_path = lib_path
        self._caps = dict(kd_capacity=kd_capacity, quad_capacity=quad_capacity)
        self._pass_seed = 0

    # ---- src/path_guiding_integrator.py:77-123 ----------------------------------------
    def setup(self, numRays, bbox_min, bbox_max, sdTreeMaxDepth=10, quadTreeMaxDepth=30,
              isStoreNEERadiance=True, bsdfSamplingFraction=0.5):
        self.numRays = int(numRays)
        self.array_size = self.numRays * self.max_depth
        self.isStoreNEERadiance = bool(isStoreNEERadiance)
        self.bsdfSamplingFraction = float(bsdfSamplingFraction)
        self.tree = SDTree([float(v) for v in bbox_min], [float(v) for v in bbox_max], kd_max_depth=sdTreeMaxDepth,
                           quad_max_depth=quadTreeMaxDepth, store_nee=isStoreNEERadiance, device=self._device,
                           lib_path=self._lib_path, **self._caps)
        self.record = None

    def setIteration(self, iteration, isFinalIter):
        self.iteration = int(iteration)
        self.isFinalIter = bool(isFinalIter)

    # ---- record buffers (dr.zeros(SurfaceInteractionRecord, array_size), :111-118) -----
    def resetRayPathData(self, like):
        """zero-filled SoA record of `array_size` slots, same array kind as `like`"""
        n = 1 if self.isFinalIter else self.array_size + 1       # + one spare slot: the target of masked-out scatters
        if _is_torch(like):
            import torch
            z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=like.device)
            u8 = torch.uint8
        else:
            z = lambda *s, dt=np.float32: np.zeros(s, dtype=dt)
            u8 = np.uint8
        self.record = dict(position=z(n, 3), direction=z(n, 2), bsdf=z(n, 3), throughputBsdf=z(n, 3),
                           throughputRadiance=z(n, 3), radiance_nee=z(n, 3), direction_nee=z(n, 2),
                           woPdf=z(n), active=z(n, dt=u8))
        return self.record

    def store_vertex(self, ray_index, depth, store_flag, position, wo_world, bsdf_weight, throughput_weight, L,
                     radiance_nee, nee_dir_world, woPdf):
        """the scatters of :318-346 at globalIndex = ray_index*max_depth + depth"""
        if self.isFinalIter:
            return
        m = store_flag != 0                             # uint8 / bool, numpy or torch -> boolean mask
        n = m.shape[0]

        def wide(x):        # Dr.Jit literals (e.g. the initial throughput Spectrum(1)) have width 1: dr.scatter broadcasts them
            if x.shape[0] == n:
                return x
            return x.expand(n, *x.shape[1:]) if _is_torch(x) else np.broadcast_to(x, (n,) + tuple(x.shape[1:]))
        position, wo_world, bsdf_weight, throughput_weight, L, radiance_nee, nee_dir_world, woPdf = (
            wide(x) for x in (position, wo_world, bsdf_weight, throughput_weight, L, radiance_nee, nee_dir_world, woPdf))
        r = self.record
        if _is_torch(ray_index):
            # masked scatter without a data-dependent shape (no host synchronisation): lanes that store nothing write
            # into the spare slot behind the last real one
            import torch
            gi = torch.where(m, ray_index.long() * self.max_depth + depth.long(), torch.full_like(ray_index.long(), self.array_size))
            one = torch.ones((), dtype=r['active'].dtype, device=gi.device).expand(n)
            r['position'].index_copy_(0, gi, position.to(r['position'].dtype))
            r['direction'].index_copy_(0, gi, self.tree.dir_to_canonical(wo_world.contiguous()))
            r['active'].index_copy_(0, gi, one)
            r['bsdf'].index_copy_(0, gi, bsdf_weight.to(r['bsdf'].dtype))
            r['throughputBsdf'].index_copy_(0, gi, throughput_weight.to(r['bsdf'].dtype))
            r['throughputRadiance'].index_copy_(0, gi, L.to(r['bsdf'].dtype))
            if self.isStoreNEERadiance:
                r['radiance_nee'].index_copy_(0, gi, radiance_nee.to(r['bsdf'].dtype))
                r['direction_nee'].index_copy_(0, gi, self.tree.dir_to_canonical(nee_dir_world.contiguous()))
            r['woPdf'].index_copy_(0, gi, woPdf.to(r['woPdf'].dtype))
            return
        if not bool(m.any()):
            return
        ray_index, depth = np.asarray(ray_index, np.int64), np.asarray(depth, np.int64)
        gi = (ray_index * self.max_depth + depth)[m]
        r['position'][gi] = position[m]
        r['direction'][gi] = self.tree.dir_to_canonical(wo_world[m])
        r['active'][gi] = 1
        r['bsdf'][gi] = bsdf_weight[m]
        r['throughputBsdf'][gi] = throughput_weight[m]
        r['throughputRadiance'][gi] = L[m]
        if self.isStoreNEERadiance:
            r['radiance_nee'][gi] = radiance_nee[m]
            r['direction_nee'][gi] = self.tree.dir_to_canonical(nee_dir_world[m])
        r['woPdf'][gi] = woPdf[m]

    # ---- tree operations of one bounce ------------------------------------------------------
    @property
    def guiding(self):
        return self.iteration > 1                                     # :223,:250,:283

    def nee_mis(self, position, ds_d, active_em, bsdf_pdf_em, pdf_with_delta, pdf_without_delta, ds_pdf, ds_delta):
        """:241-253 -> mis_em.  sdTree_prev.pdf only runs when guiding (iteration > 1)."""
        if self.guiding:
            sd = self.tree.pdf(position, ds_d, active_em)
        else:
            sd = bsdf_pdf_em
        _, mis = self.tree.mis_nee(bsdf_pdf_em, sd, pdf_with_delta, pdf_without_delta, ds_pdf, ds_delta,
                                   self.bsdfSamplingFraction, self.iteration)
        return mis

    def nee_mis_from_pdf(self, sdtree_pdf_em, bsdf_pdf_em, pdf_with_delta, pdf_without_delta, ds_pdf, ds_delta):
        """:241-253 with the tree's pdf of the emitter direction already known (from `bounce`) -> mis_em"""
        _, mis = self.tree.mis_nee(bsdf_pdf_em, sdtree_pdf_em, pdf_with_delta, pdf_without_delta, ds_pdf, ds_delta,
                                   self.bsdfSamplingFraction, self.iteration)
        return mis

    def bounce(self, position, ds_d, active_em, wo_world_bsdf, do_mis, choose_u, seed, lane_offset=0, u=None):
        """Every tree query of one path vertex in ONE library call and one spatial descent (guiding iterations): the tree's
        pdf of the emitter direction for the NEE MIS weight (:244) on the `active_em` lanes, and `choose_and_sample`
        (:283-307) -> (sdtree_pdf_em, mode, sdtree_dir, sdtree_pdf)"""
        do_mis = do_mis != 0
        guided = (choose_u > self.bsdfSamplingFraction) & do_mis
        mode = guided.astype(np.uint8) if not _is_torch(guided) else guided.to(dtype=__import__("torch").uint8)
        bs = do_mis & ~guided
        mode = mode + 2 * (bs.astype(np.uint8) if not _is_torch(bs) else bs.to(dtype=mode.dtype))
        em = active_em != 0
        em = em.astype(np.uint8) if not _is_torch(em) else em.to(dtype=mode.dtype)
        if mode.shape[0] > self.FUSED_BOUNCE_MAX_LANES:
            # very wide wavefronts: the two launches are ~8 % faster on the device than the fused one (the pdf kernel runs at
            # 64 warps per SM, the fused kernel at the sampler's 48) and a launch more no longer matters -- same results
            ep = self.tree.pdf(position, ds_d, em)
            d, sp, _, _ = self.tree.guided(position, mode, wo=wo_world_bsdf, u=u, seed=seed, lane_offset=lane_offset,
                                           bsdf_sampling_fraction=self.bsdfSamplingFraction)
            return ep, mode, d, sp
        d, sp, _, _, ep = self.tree.guided(position, mode, wo=wo_world_bsdf, u=u, seed=seed, lane_offset=lane_offset,
                                           bsdf_sampling_fraction=self.bsdfSamplingFraction, em_dir=ds_d, em_active=em)
        return ep, mode, d, sp

    def choose_and_sample(self, position, wo_world_bsdf, do_mis, choose_u, seed, lane_offset=0, u=None):
        """:283-307: lanes with choose_u > bsdfSamplingFraction (and do_mis) are sampled from the
        tree, the other do_mis lanes get the tree pdf of the BSDF-sampled direction.
        -> (mode, sdtree_dir, sdtree_pdf); mode 1 = guided sample, 2 = BSDF sample with MIS, 0 = no MIS"""
        do_mis = do_mis != 0
        guided = (choose_u > self.bsdfSamplingFraction) & do_mis
        mode = guided.astype(np.uint8) if not _is_torch(guided) else guided.to(dtype=__import__("torch").uint8)
        bs = do_mis & ~guided
        mode = mode + 2 * (bs.astype(np.uint8) if not _is_torch(bs) else bs.to(dtype=mode.dtype))
        d, sp, _, _ = self.tree.guided(position, mode, wo=wo_world_bsdf, u=u, seed=seed, lane_offset=lane_offset,
                                       bsdf_sampling_fraction=self.bsdfSamplingFraction)
        return mode, d, sp

    def mixture(self, bsdf_pdf, sdtree_pdf, bsdf_value, do_mis):
        """:310-311 -> (woPdf, bsdf_weight) (bsdf_weight valid on do_mis lanes)"""
        return self.tree.mis_mixture(bsdf_pdf, sdtree_pdf, bsdf_value, do_mis, self.bsdfSamplingFraction)

    # ---- end of pass / iteration ---------------------------------------------------------------
    def end_of_pass(self, Lfinal):
        """processPathData + scatterDataIntoSDTree + addDataPropagate (:388-395, :434-500)"""
        if self.isFinalIter or self.record is None:
            return
        r = {k: v[:self.array_size] for k, v in self.record.items()}          # without the spare slot
        self.tree.splat_path_data(self.max_depth, Lfinal, r['throughputRadiance'], r['throughputBsdf'], r['bsdf'],
                                  r['position'], r['direction'], r['woPdf'], r['radiance_nee'], r['direction_nee'], r['active'])

    def refine(self):
        raise NotImplementedError("refine() alone (src/path_guiding_integrator.py:553-564) is folded into "
                                  "refineAndPrepareSDTreeForNextIteration(); the device refine rebuilds prev and current together")

    def refineAndPrepareSDTreeForNextIteration(self):
        """:566-586 -- one sdt_refine on the device.  Once per training iteration the sizes are read back: an
        exhausted arena (device error flag) raises instead of silently training on a truncated tree."""
        self.tree.refine(iteration=self.iteration, check=True)

    def saveSDTreeToFile(self, fileName):
        self.tree.save_npz(fileName, L.SDT_TREE_PREV)

    def loadSDTreeFromFile(self, fileName):
        self.tree.load_npz(fileName)

    def saveSDTreeOBJ(self, fileName):
        """KDTree.saveOBJ (src/kdtree.py:605-663): wireframe boxes of every spatial node"""
        d = self.tree.download(L.SDT_TREE_PREV)
        bmin, bmax = d['kdtree_bbox_min'], d['kdtree_bbox_max']
        name = fileName.split('/')[-1].split('.')[0]
        v = 1
        with open(fileName, 'w') as f:
            f.write('# OBJ file of KDTree Bounding Boxes\n')
            f.write(f'o {name}\n')
            for a, b in zip(bmin, bmax):
                for y in (a[1], b[1]):
                    f.write(f'v {a[0]} {y} {a[2]}\nv {b[0]} {y} {a[2]}\nv {b[0]} {y} {b[2]}\nv {a[0]} {y} {b[2]}\n')
                f.write(f'l {v} {v + 1} {v + 2} {v + 3} {v}\n')
                f.write(f'l {v + 4} {v + 5} {v + 6} {v + 7} {v + 4}\n')
                for k in range(4):
                    f.write(f'l {v + k} {v + 4 + k}\n')
                v += 8


# =============================================================================== Mitsuba plugin
try:                                                   # pragma: no cover - Mitsuba is not in this image
    import drjit as dr
    import mitsuba as mi
    _HAVE_MITSUBA = True
except Exception:                                      # ModuleNotFoundError here
    _HAVE_MITSUBA = False


if _HAVE_MITSUBA:                                      # pragma: no cover

    def mis_weight(pdf_a, pdf_b):
        a2 = dr.sqr(pdf_a)
        result = dr.select(pdf_a > 0, a2 / dr.fma(pdf_b, pdf_b, a2), 0)
        result[dr.isnan(result)] = 0
        return result

    def _t(x):
        """Dr.Jit array -> torch CUDA tensor (zero-copy); Vector/Color -> (n, k)"""
        dr.eval(x)
        dr.sync_thread()
        t = x.torch()
        return t

    def _sync():
        """the library ran on torch's current stream, Dr.Jit reads on its own: drain before handing results over"""
        import torch
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()

    class PathGuidingIntegrator(mi.SamplingIntegrator):
        """Same constructor props, methods and plugin name as the reference class
        (src/path_guiding_integrator.py:27-628)."""

        def __init__(self, props):
            super().__init__(props)
            self.core = PathGuidingCore(props.get('max_depth', 30), props.get('rr_depth', 8))
            self.max_depth = self.core.max_depth
            self.rr_depth = self.core.rr_depth
            self.sumL = mi.Spectrum(0)
            self.sumL2 = mi.Spectrum(0)

        # -- forwarded interface
        def setup(self, numRays, bbox_min, bbox_max, sdTreeMaxDepth=10, quadTreeMaxDepth=30,
                  isStoreNEERadiance=True, bsdfSamplingFraction=0.5):
            self.core.setup(numRays, [bbox_min[0], bbox_min[1], bbox_min[2]], [bbox_max[0], bbox_max[1], bbox_max[2]],
                            sdTreeMaxDepth, quadTreeMaxDepth, isStoreNEERadiance, bsdfSamplingFraction)

        def setIteration(self, iteration, isFinalIter):
            self.core.setIteration(iteration, isFinalIter)

        def resetVarianceCounter(self):
            self.sumL = mi.Spectrum(0)
            self.sumL2 = mi.Spectrum(0)

        def refine(self):
            self.core.refine()

        def refineAndPrepareSDTreeForNextIteration(self):
            self.core.refineAndPrepareSDTreeForNextIteration()

        def saveSDTreeToFile(self, fileName):
            self.core.saveSDTreeToFile(fileName)

        def loadSDTreeFromFile(self, fileName):
            self.core.loadSDTreeFromFile(fileName)

        def saveSDTreeOBJ(self, fileName):
            self.core.saveSDTreeOBJ(fileName)

        def aov_names(self):
            return ["depth.Y"]

        def to_string(self):
            return "path_guiding_integrator"

        def computeMSE(self, spp, groundTruth):
            mse = mi.luminance((self.sumL / spp - groundTruth) ** 2)
            return dr.mean(dr.minimum(mse, 10000))[0]

        def computeVariance(self, spp, groundTruth=None):
            if groundTruth is not None:
                v = mi.luminance((self.sumL2 / spp) - (groundTruth * groundTruth))
                return dr.mean(dr.minimum(v, 10000))[0] / spp
            Lm = self.sumL / spp
            v = mi.luminance(self.sumL2 / spp - Lm * Lm)
            v = dr.mean(dr.minimum(v, 10000))[0]
            return v / (spp - 1) if spp > 1 else v

        # -- the path loop (src/path_guiding_integrator.py:126-431) as a wavefront loop
        def sample(self, scene, sampler, ray, medium=None, active=True, aovs=None):
            import torch
            core = self.core
            f = core.bsdfSamplingFraction
            bsdf_ctx = mi.BSDFContext()
            ray = mi.Ray3f(ray)
            throughput = mi.Spectrum(1)
            depth = mi.UInt32(0)
            L = mi.Spectrum(0)
            ior = mi.Float(1)
            active = mi.Bool(active)
            n = dr.width(ray)
            ray_index = dr.arange(mi.UInt32, n)
            prev_si = dr.zeros(mi.SurfaceInteraction3f)
            prev_bsdf_pdf = mi.Float(1.0)
            prev_bsdf_delta = mi.Bool(True)
            core._pass_seed += 1
            rec_ready = False
            it = 0
            while dr.any(active) and (self.max_depth < 0 or it < self.max_depth):
                it += 1
                # the reference's recorded mi.Loop masks every state update by `active`; this unrolled wavefront loop
                # masks what a finished lane could still change: its intersection and its radiance
                si = scene.ray_intersect(ray, ray_flags=mi.RayFlags.All, coherent=dr.eq(depth, 0), active=active)
                bsdf = si.bsdf()
                ds_direct = mi.DirectionSample3f(scene, si=si, ref=prev_si)
                emitter_pdf = scene.pdf_emitter_direction(prev_si, ds_direct, ~prev_bsdf_delta)
                Le = throughput * mis_weight(prev_bsdf_pdf, emitter_pdf) * ds_direct.emitter.eval(si)
                active_next = (depth + 1 < self.max_depth) & si.is_valid()
                active_em = active_next & mi.has_flag(bsdf.flags(), mi.BSDFFlags.Smooth)
                ds, em_weight = scene.sample_emitter_direction(si, sampler.next_2d(), True, active_em)
                active_em &= dr.neq(ds.pdf, 0.0)
                wo = si.to_local(ds.d)
                bsdf_value_em, bsdf_pdf_em = bsdf.eval_pdf(bsdf_ctx, si, wo, active_em)
                act_sd_em = active_em & (core.iteration > 1)
                prev_component = bsdf_ctx.component
                bsdf_ctx.component &= ~mi.BSDFFlags.Delta
                bs_nd, _ = bsdf.sample(bsdf_ctx, si, mi.Float(0.5), mi.Vector2f(0.5, 0.5), act_sd_em)
                pdf_without_delta = bsdf.pdf(bsdf_ctx, si, bs_nd.wo, act_sd_em)
                bsdf_ctx.component = prev_component
                pdf_with_delta = bsdf.pdf(bsdf_ctx, si, bs_nd.wo, act_sd_em)
                # continuation: BSDF sample (the sampler is advanced in the reference's order; the tree is asked once, below)
                bsdf_sample, bsdf_weight = bsdf.sample(bsdf_ctx, si, sampler.next_1d(active_next), sampler.next_2d(active_next), active_next)
                bsdf_pdf = mi.Float(bsdf_sample.pdf)
                bsdf_value = bsdf_weight * bsdf_pdf
                woPdf = mi.Float(bsdf_pdf)
                wo_local = mi.Vector3f(bsdf_sample.wo)
                wo_world = si.to_world(wo_local)
                delta = mi.has_flag(bsdf_sample.sampled_type, mi.BSDFFlags.Delta)
                do_mis = active_next & ~delta & (core.iteration > 1)
                choose_u = sampler.next_1d(active_next)
                # A4 + A5 on the device library: every tree query of the vertex in ONE call and one spatial descent
                p_t = _t(si.p)
                if core.iteration > 1:
                    sd_em_t, mode_t, sd_dir_t, sd_pdf_t = core.bounce(p_t, _t(ds.d), _t(act_sd_em), _t(wo_world), _t(do_mis), _t(choose_u),
                                                                      seed=core._pass_seed * 1315423911 + it)
                    mis_em_t = core.nee_mis_from_pdf(sd_em_t, _t(bsdf_pdf_em), _t(pdf_with_delta), _t(pdf_without_delta), _t(ds.pdf), _t(ds.delta))
                else:
                    mis_em_t = core.nee_mis(p_t, _t(ds.d), _t(act_sd_em), _t(bsdf_pdf_em), _t(pdf_with_delta),
                                            _t(pdf_without_delta), _t(ds.pdf), _t(ds.delta))
                _sync()
                mis_em = mi.Float(mis_em_t)
                Lr_dir = throughput * mis_em * bsdf_value_em * em_weight
                L = dr.select(active, L + Le + Lr_dir, L)
                if core.iteration > 1:
                    guided = mi.Bool(mode_t == 1)
                    sd_dir = dr.unravel(mi.Vector3f, mi.Float(sd_dir_t.reshape(-1)))
                    wo_world[guided] = sd_dir
                    wo_local[guided] = si.to_local(sd_dir)
                    v2, p2 = bsdf.eval_pdf(bsdf_ctx, si, wo_local, guided)
                    bsdf_value[guided] = v2
                    bsdf_pdf[guided] = p2
                    woPdf_t, w_t = core.mixture(_t(bsdf_pdf), sd_pdf_t, _t(bsdf_value), _t(do_mis))
                    _sync()
                    woPdf[do_mis] = mi.Float(woPdf_t)
                    bsdf_weight[do_mis] = dr.unravel(mi.Spectrum, mi.Float(w_t.reshape(-1)))
                # record (:318-346)
                store = active & si.is_valid()
                if not core.isFinalIter:
                    if not rec_ready:
                        core.resetRayPathData(p_t)
                        rec_ready = True
                    core.store_vertex(_t(mi.Int32(ray_index)), _t(mi.Int32(depth)), _t(store), p_t, _t(wo_world), _t(bsdf_weight),
                                      _t(throughput), _t(L), _t(Lr_dir / throughput), _t(ds.d), _t(woPdf))
                ray = si.spawn_ray(wo_world)
                ior *= bsdf_sample.eta
                throughput *= bsdf_weight
                prev_si = si
                prev_bsdf_pdf = woPdf
                prev_bsdf_delta = delta
                tmax = dr.max(throughput)
                active_next &= dr.neq(tmax, 0)
                rr_prob = dr.minimum(tmax * ior ** 2, 0.95)
                rr_active = depth >= self.rr_depth
                rr_continue = sampler.next_1d() < rr_prob
                active_next &= ~rr_active | rr_continue
                active = active_next
                depth[si.is_valid()] += 1
            if not core.isFinalIter and rec_ready:
                core.end_of_pass(_t(L))
            spp = sampler.sample_count()
            if spp == 1:
                self.sumL += L
                self.sumL2 += L * L
            else:
                one = int(dr.width(L) / spp)
                base = dr.arange(mi.UInt32, one) * spp
                for i in range(spp):
                    s = dr.gather(mi.Spectrum, L, base + i)
                    self.sumL += s
                    self.sumL2 += s * s
            return (L, dr.neq(depth, 0), [1])

    mi.register_integrator('path_guiding_integrator', lambda props: PathGuidingIntegrator(props))
