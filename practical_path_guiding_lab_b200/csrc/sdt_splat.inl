// Radiance-record splatting into `current` (KDTree.addDataPropagate src/kdtree.py:180-225,
// QuadTree.addDataPropagate src/quadtree.py:389-464) and the bottom-up sweeps that turn
// leaf statistics into the per-node statistics the reference accumulates with one
// atomic per visited node.
//
// The reference adds the count / energy at EVERY node of the root->leaf path.  Here a
// record issues ONE atomic per descent -- on the leaf -- and the interior sums are
// rebuilt level by level afterwards (interior = sum of its children, in fp32, inside the
// 1e-4 relative tolerance the reference's own atomic ordering already has).  Lanes of a
// warp that hit the same leaf are combined first (match_any + shuffle reduction), so a
// hot leaf costs one atomic per warp instead of 32 serialised ones.

struct SplatTarget {
    float* kd_count;
    float* q_ecur;
    uint32_t store_nee;
};

// one fp32 add per record; with `aggregate` the lanes of a warp that hit the same address are summed by the lowest
// lane first (match_any + shuffles: pays when neighbouring lanes share leaves, i.e. on pixel-coherent wavefronts;
// on incoherent ones it finds no peers and only costs instructions -- "splat_aggregate" tuning, off by default)
SDT_HD void sdt_splat_add(float* base, uint32_t idx, float v, bool valid, uint32_t aggregate) {
#if defined(__CUDA_ARCH__)
    if (!aggregate) { if (valid) atomicAdd(base + idx, v); return; }
    const uint32_t act = __ballot_sync(__activemask(), valid);
    if (!valid) return;
    const uint32_t peers = __match_any_sync(act, idx);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t leader = __ffs(peers) - 1u;
    if (peers == (1u << lane)) { atomicAdd(base + idx, v); return; }
    // sum the peer group in ascending lane order (deterministic within the warp)
    float acc = 0.0f;
    uint32_t rest = peers;
    while (rest) {
        const uint32_t src = __ffs(rest) - 1u;
        acc += __shfl_sync(peers, v, src);
        rest &= rest - 1u;
    }
    if (lane == leader) atomicAdd(base + idx, acc);
#else
    (void)aggregate;
    if (valid) base[idx] += v;
#endif
}

// one ACTIVE record through both trees
template <int KD>
SDT_HD void sdt_splat_one(const TreeView& t, const SplatTarget& tg, const KdCtx& k,
                          float px, float py, float pz, float dx, float dy, float radiance, float wo_pdf,
                          float nr, float ng, float nb, float ndx, float ndy) {
    const KdResult r = sdt_kd_descend<KD>(k, px, py, pz);
    // src/kdtree.py:199: +1.0f (fp32 counter; exact below 2^24, clamped there by the sweep like the reference's sticks)
    if (k.cnt_s) {
#if defined(__CUDA_ARCH__)
        if (r.inbox) atomicAdd(k.cnt_s + r.leaf, 1u);                      // shared-memory counter of this CTA
#endif
    } else {
        sdt_splat_add(tg.kd_count, r.leaf, 1.0f, r.inbox, k.aggregate);
    }
    // src/kdtree.py:224: the root id is gathered UNMASKED -- out-of-box records go to the tree of node 0
    const uint32_t ri = r.rootrec;
    // a single-leaf tree has no record: its only node is the root, whose id is in kd_root
    uint32_t root = 0;
    if (ri == SDT_NONE) root = SDT_LDG(t.kd_root + r.leaf);
    const float irr = (wo_pdf > 0.0f) ? radiance / wo_pdf : 0.0f;                      // src/quadtree.py:451
    uint32_t leaf = SDT_NONE;
    if (irr != 0.0f) leaf = sdt_quad_leaf(t, ri, root, dx, dy);
    sdt_splat_add(tg.q_ecur, leaf, irr, leaf != SDT_NONE, k.aggregate);
    if (tg.store_nee) {                                                               // :455-464
        const float lum = sdt_luminance(nr, ng, nb);
        const float irr2 = (wo_pdf > 0.0f) ? lum / wo_pdf : 0.0f;
        uint32_t leaf2 = SDT_NONE;
        if (irr2 != 0.0f) leaf2 = sdt_quad_leaf(t, ri, root, ndx, ndy);
        sdt_splat_add(tg.q_ecur, leaf2, irr2, leaf2 != SDT_NONE, k.aggregate);
    }
}

struct SplatRecordsLane {
    static constexpr bool kSmemCounts = true, kGrid = true;
    static constexpr int kModes = 1, kMaxThreads = SDT_SPLAT_THREADS;
    SDT_HD void flush_count(uint32_t node, float c) const { sdt_atomic_add_f32(tg.kd_count + node, c); }
    TreeView t; SplatTarget tg; sdt_records r;
    SDT_HD uint32_t mode_of(uint32_t i) const { return r.active ? (SDT_LDG(r.active + i) != 0 ? 1u : 0u) : 1u; }
    SDT_HD void idle(uint32_t) const {}
    template <int KD>
    SDT_HD void run_mode(const KdCtx& k, uint32_t i, uint32_t) const {
        float nr = 0.0f, ng = 0.0f, nb = 0.0f, ndx = 0.0f, ndy = 0.0f;
        if (tg.store_nee && r.radiance_nee.x && r.direction_nee.x) {
            nr = sdt_ld(r.radiance_nee.x, r.radiance_nee.stride, i);
            ng = sdt_ld(r.radiance_nee.y, r.radiance_nee.stride, i);
            nb = sdt_ld(r.radiance_nee.z, r.radiance_nee.stride, i);
            ndx = sdt_ld(r.direction_nee.x, r.direction_nee.stride, i);
            ndy = sdt_ld(r.direction_nee.y, r.direction_nee.stride, i);
        }
        sdt_splat_one<KD>(t, tg, k,
                          sdt_ld(r.position.x, r.position.stride, i), sdt_ld(r.position.y, r.position.stride, i), sdt_ld(r.position.z, r.position.stride, i),
                          sdt_ld(r.direction.x, r.direction.stride, i), sdt_ld(r.direction.y, r.direction.stride, i),
                          SDT_LDG(r.radiance + i), SDT_LDG(r.wo_pdf + i), nr, ng, nb, ndx, ndy);
    }
};

SDT_HD float sdt_nan0(float v) { return (v != v) ? 0.0f : v; }

// processPathData + scatterDataIntoSDTree (src/path_guiding_integrator.py:434-500) fused in front
// of the splat.  The reference compacts the surviving records (dr.compress + 9 gathers + a host
// sync); here the slots of a tile are sorted by their `active` flag and the filter masks the rest.
struct SplatPathLane {
    static constexpr bool kSmemCounts = true, kGrid = true;
    static constexpr int kModes = 1, kMaxThreads = SDT_SPLAT_THREADS;
    SDT_HD void flush_count(uint32_t node, float c) const { sdt_atomic_add_f32(tg.kd_count + node, c); }
    TreeView t; SplatTarget tg; sdt_path_data p;
    struct Vals { float radiance, wo_pdf, nr, ng, nb; };
    SDT_HD bool prepare(uint32_t i, Vals& v) const {
        const uint32_t ray = i / p.max_depth;                                          // :440
        float inc[3];
        const float* lf[3] = {p.l_final.x, p.l_final.y, p.l_final.z};
        const float* tr[3] = {p.throughput_radiance.x, p.throughput_radiance.y, p.throughput_radiance.z};
        const float* tb[3] = {p.throughput_bsdf.x, p.throughput_bsdf.y, p.throughput_bsdf.z};
        const float* bs[3] = {p.bsdf.x, p.bsdf.y, p.bsdf.z};
        for (int c = 0; c < 3; ++c) {
            float out = (sdt_ld(lf[c], p.l_final.stride, ray) - sdt_ld(tr[c], p.throughput_radiance.stride, i))
                        / sdt_ld(tb[c], p.throughput_bsdf.stride, i);                 // :443
            out = sdt_nan0(out);                                                       // :444
            inc[c] = sdt_nan0(out / sdt_ld(bs[c], p.bsdf.stride, i));                  // :448-449
        }
        v.radiance = sdt_nan0(sdt_luminance(inc[0], inc[1], inc[2]));                 // :452, :466
        v.nr = v.ng = v.nb = 0.0f;
        if (p.radiance_nee.x) {
            v.nr = sdt_nan0(sdt_ld(p.radiance_nee.x, p.radiance_nee.stride, i));       // :467
            v.ng = sdt_nan0(sdt_ld(p.radiance_nee.y, p.radiance_nee.stride, i));
            v.nb = sdt_nan0(sdt_ld(p.radiance_nee.z, p.radiance_nee.stride, i));
        }
        v.wo_pdf = SDT_LDG(p.wo_pdf + i);
        const bool both_zero = (v.radiance == 0.0f) && (sdt_luminance(v.nr, v.ng, v.nb) == 0.0f);  // :470-472
        const bool act = p.active ? SDT_LDG(p.active + i) != 0 : true;
        return act && !both_zero && !(v.wo_pdf == 0.0f) && !(v.wo_pdf != v.wo_pdf);    // :475-478
    }
    // Lanes are classified by the `active` flag alone (one byte per slot): a pass has numRays*max_depth
    // slots and most of them lie beyond the end of their path, so idle slots cost that byte and nothing
    // else (unless the caller wants SurfaceInteractionRecord.radiance for every slot); the radiance
    // back-propagation and the reference's filter run once, on the active slots.
    SDT_HD uint32_t mode_of(uint32_t i) const { return p.active ? (SDT_LDG(p.active + i) != 0 ? 1u : 0u) : 1u; }
    SDT_HD void idle(uint32_t i) const {
        if (!p.radiance_out) return;
        Vals v;
        prepare(i, v);
        p.radiance_out[i] = v.radiance;
    }
    template <int KD>
    SDT_HD void run_mode(const KdCtx& k, uint32_t i, uint32_t) const {
        Vals v;
        const bool keep = prepare(i, v);
        if (p.radiance_out) p.radiance_out[i] = v.radiance;       // SurfaceInteractionRecord.radiance, every slot
        if (!keep) return;
        float ndx = 0.0f, ndy = 0.0f;
        if (p.direction_nee.x) {
            ndx = sdt_ld(p.direction_nee.x, p.direction_nee.stride, i);
            ndy = sdt_ld(p.direction_nee.y, p.direction_nee.stride, i);
        }
        sdt_splat_one<KD>(t, tg, k,
                          sdt_ld(p.position.x, p.position.stride, i), sdt_ld(p.position.y, p.position.stride, i), sdt_ld(p.position.z, p.position.stride, i),
                          sdt_ld(p.direction.x, p.direction.stride, i), sdt_ld(p.direction.y, p.direction.stride, i),
                          v.radiance, v.wo_pdf, v.nr, v.ng, v.nb, ndx, ndy);
    }
};

// ---- bottom-up sweeps -------------------------------------------------------------
// quadtree level l (deepest first): interior energy = ((c1+c2)+c3)+c4
struct QuadSweepItem {
    const DevHeader* hdr; const uint32_t* child; float* e; uint32_t level;
    SDT_HD void operator()(uint32_t i) const {
        const uint32_t id = hdr->level_off[level] + i;
        const uint32_t cb = child[id];
        if (cb) e[id] = ((e[cb] + e[cb + 1u]) + e[cb + 2u]) + e[cb + 3u];
    }
};
// Two levels per launch (the sweep is a chain of dependent launches that pays latency, not bandwidth): a thread owns a node
// of level `level`, first completes those of its children that are interior from THEIR children (level + 2, final by
// then), then sums its own.  Same additions in the same order as two single-level passes.
struct QuadSweep2Item {
    const DevHeader* hdr; const uint32_t* child; float* e; uint32_t level;
    SDT_HD void operator()(uint32_t i) const {
        const uint32_t id = hdr->level_off[level] + i;
        const uint32_t cb = child[id];
        if (!cb) return;
        uint32_t gcb[4];
        for (uint32_t k = 0; k < 4u; ++k) gcb[k] = child[cb + k];
        float s[4];
        for (uint32_t k = 0; k < 4u; ++k) {
            const uint32_t g = gcb[k];
            if (g) { s[k] = ((e[g] + e[g + 1u]) + e[g + 2u]) + e[g + 3u]; e[cb + k] = s[k]; }
            else s[k] = e[cb + k];
        }
        e[id] = ((s[0] + s[1]) + s[2]) + s[3];
    }
};
// spatial nodes of depth d: interior count = left + right, sticking at 2^24 like a
// chain of fp32 "+1.0f" atomics does
struct KdSweepItem {
    const uint32_t* kd_word; const uint32_t* kd_depth; float* cnt; uint32_t depth;
    SDT_HD void operator()(uint32_t i) const {
        const uint32_t w = kd_word[i];
        if (kd_depth[i] != depth) return;
        // a leaf counter fed by per-CTA partial sums can pass 2^24, where the reference's chain of
        // "+1.0f" sticks: clamp leaves and interiors alike
        const float s = (w & SDT_KD_LEAF_BIT) ? cnt[i] : cnt[w] + cnt[w + 1u];
        cnt[i] = s > 16777216.0f ? 16777216.0f : s;
    }
};

// quadtree levels the energy sweep has to visit: `current` has the topology of `prev`, whose level count the host knows
// exactly after the last refine's (non-blocking) header read-back, and bounds by levels_hint before that
static inline uint32_t sdt_sweep_levels(sdt_handle h) {
    (void)tree_view(h);                                  // picks up a header read-back that has landed
    return (h->levels_known && h->levels_known < h->levels_hint) ? h->levels_known : h->levels_hint;
}

// interior statistics of `current` from its leaf statistics.  with_kd = false (the refine): only the quadtree energies --
// the refine reads spatial counts of LEAVES only, so the 21 per-depth passes over the spatial tree are left to whoever
// wants to SEE interior counts (a download of either tree; sdt_sweep_prev_counts for the counts the refine rolled into prev)
static int sdt_complete_stats(sdt_handle h, cudaStream_t st, bool with_kd = true) {
    if (h->stats_complete && (h->kd_complete || !with_kd)) return SDT_OK;
    const ExecCtx x = exec_ctx(h, st);
    const QuadSet& s = h->set[h->cur];
    // level sizes live on the device; the number of levels in use is exact once the last refine's header has been read back
    if (!h->stats_complete) {
        // the deepest level in use holds leaves only; from there upwards in pairs of levels, a single one if one is left
        int l = (int)sdt_sweep_levels(h) - 2;
        for (; l >= 1; l -= 2)
            launch_items(x, &s.hdr->level_cnt[l - 1], 0, QuadSweep2Item{s.hdr, s.child, h->q_ecur, (uint32_t)(l - 1)});
        if (l == 0) launch_items(x, &s.hdr->level_cnt[0], 0, QuadSweepItem{s.hdr, s.child, h->q_ecur, 0u});
    }
    if (with_kd && !h->kd_complete) {
        for (int d = h->cfg.kd_max_depth; d >= 0; --d)
            launch_items(x, &s.hdr->n_kd, 0, KdSweepItem{h->kd_word, h->kd_depth, h->kd_count, (uint32_t)d});
        h->kd_complete = true;
    }
    SDT_TRY(sdt_post_launch(h, "sdt_complete_stats"));
    h->stats_complete = true;
    return SDT_OK;
}
// prev.vertCount as the refine left it holds leaf counts only (see above): interiors on demand
static int sdt_sweep_prev_counts(sdt_handle h, cudaStream_t st) {
    if (!h->prev_kd_dirty) return SDT_OK;
    const ExecCtx x = exec_ctx(h, st);
    const QuadSet& s = h->set[h->cur];
    for (int d = h->cfg.kd_max_depth; d >= 0; --d)
        launch_items(x, &s.hdr->n_kd, 0, KdSweepItem{h->kd_word, h->kd_depth, h->kd_prev_count, (uint32_t)d});
    SDT_TRY(sdt_post_launch(h, "sdt_sweep_prev_counts"));
    h->prev_kd_dirty = false;
    return SDT_OK;
}

// ---------------------------------------------------------------------------- entry points
extern "C" int sdt_splat_records(sdt_handle h, const sdt_records* rec, uint32_t n, uint32_t flags, sdt_stream stream) {
    SDT_ENTER(h);
    if (n == 0) return SDT_OK;                          // empty wavefront: nothing to do (pointers may be NULL)
    SDT_CHECK(h, rec && rec->position.x && rec->direction.x && rec->radiance && rec->wo_pdf, SDT_ERR_INVALID, "sdt_splat_records: missing field");
    cudaStream_t st = (cudaStream_t)stream;
    const bool nee = h->cfg.store_nee != 0;
    h->stats_complete = false; h->kd_complete = false;
    h->splat_bound += n;
    const size_t per_lane = 12 + 8 + 4 + 4 + (nee ? 20 : 0) + 1 + 8;
    return sdt_run_chunked(h, st, flags, n, per_lane, false, [&](Stager& sg, uint32_t off, uint32_t cnt) -> int {
        sdt_records d = *rec;
        d.position = sg.in3(sdt_off3(rec->position, off), cnt);
        d.direction = sg.in2(sdt_off2(rec->direction, off), cnt);
        d.radiance = sg.in_t(sdt_offp(rec->radiance, off), cnt);
        d.wo_pdf = sg.in_t(sdt_offp(rec->wo_pdf, off), cnt);
        if (nee) { d.radiance_nee = sg.in3(sdt_off3(rec->radiance_nee, off), cnt); d.direction_nee = sg.in2(sdt_off2(rec->direction_nee, off), cnt); }
        d.active = sg.in_t(sdt_offp(rec->active, off), cnt);
        if (sg.status != SDT_OK) return sg.status;
        SplatRecordsLane f{tree_view(h), SplatTarget{h->kd_count, h->q_ecur, (uint32_t)nee}, d};
        sg.before_launch();
        return launch_wavefront(h, st, cnt, f, h->splat_block, h->splat_ctas_per_sm, rec->active != nullptr);
    });
}

extern "C" int sdt_splat_path_data(sdt_handle h, const sdt_path_data* pd, uint32_t flags, sdt_stream stream) {
    SDT_ENTER(h);
    if (pd && pd->slots == 0) return SDT_OK;            // empty wavefront: nothing to do (pointers may be NULL)
    SDT_CHECK(h, pd && pd->max_depth > 0 && pd->l_final.x && pd->throughput_radiance.x && pd->throughput_bsdf.x && pd->bsdf.x &&
                     pd->position.x && pd->direction.x && pd->wo_pdf, SDT_ERR_INVALID, "sdt_splat_path_data: missing field");
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t n = pd->slots;
    const uint32_t rays = (n + pd->max_depth - 1u) / pd->max_depth;
    Stager sg(h, st, flags);
    SDT_TRY(sg.reserve((size_t)n * (36 + 12 + 8 + 4 + 12 + 8 + 1 + 4) + (size_t)rays * 12 + 32768));
    sdt_path_data d = *pd;
    d.l_final = sg.in3(pd->l_final, rays);
    d.throughput_radiance = sg.in3(pd->throughput_radiance, n);
    d.throughput_bsdf = sg.in3(pd->throughput_bsdf, n);
    d.bsdf = sg.in3(pd->bsdf, n);
    d.position = sg.in3(pd->position, n);
    d.direction = sg.in2(pd->direction, n);
    d.wo_pdf = sg.in_t(pd->wo_pdf, n);
    d.radiance_nee = sg.in3(pd->radiance_nee, n);
    d.direction_nee = sg.in2(pd->direction_nee, n);
    d.active = sg.in_t(pd->active, n);
    d.radiance_out = sg.out_t(pd->radiance_out, n);
    if (sg.status != SDT_OK) return sg.status;
    SplatPathLane f{tree_view(h), SplatTarget{h->kd_count, h->q_ecur, (uint32_t)(h->cfg.store_nee != 0)}, d};
    h->stats_complete = false; h->kd_complete = false;
    h->splat_bound += n;
    SDT_TRY(launch_wavefront(h, st, n, f, h->splat_block, h->splat_ctas_per_sm, pd->active != nullptr));
    return sg.finish(flags);
}
