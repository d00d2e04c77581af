// Per-iteration refine, entirely on the device: refineAndPrepareSDTreeForNextIteration
// (src/path_guiding_integrator.py:566-586).  The reference runs host loops with one
// round trip and one reallocation per BFS level; here every pass is a grid-wide
// "items" or "exclusive scan" launch whose sizes are read from the device header, so the
// whole refine is one stream of launches with no host synchronisation.
//
//  spatial split  (KDTree.refine/split, src/kdtree.py:229-358): a leaf with count C at
//     depth d ends up as a perfect subtree of s = min{s : C/2^s <= T} (capped by
//     maxDepth-d) levels.  Round r of the reference's loop appends, for every original
//     leaf with s >= r in ascending id, the 2^r children of its level-(r-1) descendants in
//     path order -- so one scan over the original leaves per round yields every new node
//     id and every new quadtree root id in the reference's numbering.
//  quadtrees  (setRefinementThreshold :512-560, refine :563-637, clearTreeUnusedNode
//     :844-851, copyTree :695-828): the new forest is built level by level straight into
//     the canonical layout; a level's nodes decide (merge / keep / split) and one scan of
//     the non-leaf flags places their children.  A leaf with energy E at depth d becomes a
//     perfect 4-ary subtree of s = min{s : E/4^s <= thr} (capped by maxDepth-d) levels.

#define SDT_KIND_VIRTUAL 1u   // node created by a split in this refine (no source node)
#define SDT_KIND_REACHED 2u   // the reference's merge BFS reaches this node (:574-611)

struct RefineCtx {
    DevHeader* H0;            // header of the tree being refined
    DevHeader* H1;            // header of the tree being built (also holds the scratch counters)
    // spatial
    uint32_t* kd_word; float* kd_count; uint32_t* kd_depth; uint32_t* kd_root; float* kd_bmin; float* kd_bmax;
    float* kd_prev_count; uint8_t* kd_s; uint32_t* kd_sel; uint32_t* kd_rank_cur; uint32_t* kd_rank_prev; uint32_t* root_src;
    // quadtrees: old set + current statistics, new set + scratch
    const uint32_t* child0; const float* thr0; const float* e_cur;
    uint32_t* child1; float* energy1; float* thr1; uint32_t* iidx1; QRec* rec1; uint32_t* root_iidx1;
    float* pp1;               // per-node pdf products of the tree being built (sdt_core.h), written top-down with the levels
    uint32_t* s_src; uint8_t* s_kind; uint8_t* s_srem;
    uint32_t no_quad;
    uint32_t kd_rounds;   // split rounds the host launches (sdt_kd_rounds_bound)
    uint32_t thr_recip;   // threshold = E * fp32(1/100) instead of E / 100 (how Dr.Jit may lower the literal division)
};

// header of the tree being built <- header of the tree being refined, with the refine's counters reset: one thread per
// 32-bit word (a single thread copying the ~700-byte struct was an 8-11 us link of the launch chain)
struct RefineInit {
    RefineCtx c;
    SDT_HD void operator()(uint32_t w) const {
        const DevHeader* H0 = c.H0;
        uint32_t v = reinterpret_cast<const uint32_t*>(H0)[w];
#define SDT_HDR_WORD(field) (offsetof(DevHeader, field) / 4u)
        if (w == SDT_HDR_WORD(kd_n_old) || w == SDT_HDR_WORD(kd_round_base) || w == SDT_HDR_WORD(kd_prev_round_base)) v = H0->n_kd;
        else if (w == SDT_HDR_WORD(n_roots_old) || w == SDT_HDR_WORD(kd_round_root_base)) v = H0->n_roots;
        else if (w == SDT_HDR_WORD(n_quad_old)) v = H0->n_quad;
        else if (w == SDT_HDR_WORD(kd_sel) || w == SDT_HDR_WORD(kd_round_new) || w == SDT_HDR_WORD(kd_stop) || w == SDT_HDR_WORD(lvl_trunc)) v = 0u;
        else if (w == SDT_HDR_WORD(refine_count)) v = H0->refine_count + 1u;
#undef SDT_HDR_WORD
        reinterpret_cast<uint32_t*>(c.H1)[w] = v;
    }
};

// ---- spatial --------------------------------------------------------------------
struct KdLevelsItem {       // s per original leaf (KDTree.refine's condition, :346-347)
    RefineCtx c;
    SDT_HD void operator()(uint32_t i) const {
        if (i < c.H1->n_roots_old) c.root_src[i] = i;           // every tree starts as a copy of itself (roots <= nodes ...
        if (i == 0u) for (uint32_t r = c.H1->kd_n_old; r < c.H1->n_roots_old; ++r) c.root_src[r] = r;   // ... in any well-formed tree)
        uint32_t s = 0;
        if (c.kd_word[i] & SDT_KD_LEAF_BIT) {
            const float T = c.H1->max_leaf_size;
            const uint32_t d = c.kd_depth[i], maxd = c.H1->kd_max_depth;
            // a leaf counter fed by per-CTA partial sums can pass 2^24, where the reference's chain of "+1.0f" sticks
            float v = c.kd_count[i];
            if (v > 16777216.0f) { v = 16777216.0f; c.kd_count[i] = v; }
            while (v > T && d + s < maxd) { if (v > 0.0f) v = v / 2.0f; ++s; }    // :261-264
            if (s > c.kd_rounds) {            // the host's bound of the rounds was wrong (sdt_hint_records): say so
#if defined(__CUDA_ARCH__)
                atomicOr(&c.H1->error, (uint32_t)DEV_ERR_KD_ROUNDS);
#else
                c.H1->error |= DEV_ERR_KD_ROUNDS;
#endif
            }
        }
        c.kd_s[i] = (uint8_t)s;
    }
};
struct RootIdentityItem { uint32_t* root_src; SDT_HD void operator()(uint32_t r) const { root_src[r] = r; } };

struct KdRoundFlag {
    RefineCtx c; uint32_t r;
    // kd_stop = the round whose nodes no longer fitted the arena (set by that round's own fin, which may run while
    // blocks of the same scan still emit: only LATER rounds look at it)
    SDT_HD uint32_t operator()(uint32_t i) const { return (!(c.H1->kd_stop && r > c.H1->kd_stop) && c.kd_s[i] >= r) ? 1u : 0u; }
};
struct KdRoundEmit {
    RefineCtx c;
    SDT_HD void operator()(uint32_t i, uint32_t rank, uint32_t v) const {
        if (v) { c.kd_sel[rank] = i; c.kd_rank_cur[i] = rank; }
    }
};
struct KdRoundFin {
    RefineCtx c; uint32_t r;
    SDT_HD void operator()(uint32_t total) const {
        DevHeader* H = c.H1;
        uint64_t splits = (uint64_t)total << (r - 1u);
        if ((uint64_t)H->n_kd + 2u * splits > H->kd_cap) {     // arena exhausted: stop splitting, flag it
            if (total) { H->error |= DEV_ERR_KD_CAPACITY; if (!H->kd_stop) H->kd_stop = r; }
            splits = 0; total = 0;
        }
        H->kd_sel = total;
        H->kd_round_new = (uint32_t)(2u * splits);
        H->kd_prev_round_base = H->kd_round_base;
        H->kd_round_base = H->n_kd;
        H->kd_round_root_base = H->n_roots;
        H->n_kd += (uint32_t)(2u * splits);
        H->n_roots += (uint32_t)splits;
    }
};
// one thread per node appended in round r (KDTree.split, :229-323)
struct KdMakeNodeItem {
    RefineCtx c; uint32_t r;
    SDT_HD void operator()(uint32_t j) const {
        const DevHeader* H = c.H1;
        const uint32_t i = j >> 1, b = j & 1u;                 // i-th split node of the round, left/right
        const uint32_t sh = r - 1u;
        const uint32_t lr = i >> sh, k = i & ((1u << sh) - 1u);
        const uint32_t leaf0 = c.kd_sel[lr];                   // original leaf this subtree hangs from
        uint32_t parent = leaf0;
        if (r > 1u) parent = H->kd_prev_round_base + 2u * ((c.kd_rank_prev[leaf0] << (sh - 1u)) + (k >> 1)) + (k & 1u);
        const uint32_t node = H->kd_round_base + j;
        if (b == 0u) c.kd_word[parent] = H->kd_round_base + 2u * i;       // :243-249, isLeaf <- False
        const uint32_t d = c.kd_depth[parent];
        c.kd_depth[node] = d + 1u;                                        // :255-258
        const float v = c.kd_count[parent];
        c.kd_count[node] = v > 0.0f ? v / 2.0f : v;                       // :261-264
        const uint32_t axis = d % 3u;                                     // :277
        for (uint32_t a = 0; a < 3u; ++a) {
            const float lo = c.kd_bmin[3u * parent + a], hi = c.kd_bmax[3u * parent + a];
            const float mid = (lo + hi) / 2.0f;                           // :270
            c.kd_bmin[3u * node + a] = (a == axis && b == 1u) ? mid : lo;
            c.kd_bmax[3u * node + a] = (a == axis && b == 0u) ? mid : hi;
        }
        // left inherits the parent's quadtree, right owns a copy with a new root id (:316-323)
        uint32_t root = c.kd_root[parent];
        if (b == 1u) {
            root = H->kd_round_root_base + i;
            c.root_src[root] = c.kd_root[leaf0];
        }
        c.kd_root[node] = root;
        c.kd_word[node] = SDT_KD_LEAF_BIT | root;
    }
};

// ---- quadtrees --------------------------------------------------------------------
struct QRootItem {          // level 0: tree r copies the tree of root_src[r]; thr = E_root/100 (:519)
    RefineCtx c;
    SDT_HD void operator()(uint32_t r) const {
        DevHeader* H = c.H1;
        const uint32_t cap = H->quad_cap;
        if (r == 0u) {                                        // level bookkeeping of the forest being built
            uint32_t R = H->n_roots;
            if (R > cap) { H->error |= DEV_ERR_QUAD_CAPACITY; R = cap; }
            for (int l = 0; l < SDT_MAX_LEVELS + 2; ++l) { H->level_off[l] = R; H->level_cnt[l] = 0; }
            H->level_off[0] = 0;
            H->lvl_n[0] = R; H->lvl_n[1] = 0; H->lvl_trunc = 0;
            H->int_off[0] = 0;
        }
        if (r >= cap) return;
        const uint32_t src = c.root_src[r];
        c.s_src[r] = src;
        c.s_kind[r] = SDT_KIND_REACHED;
        const float e = c.e_cur[src];
        c.energy1[r] = e;
        c.pp1[r] = 1.0f;
        c.thr1[r] = c.no_quad ? c.thr0[src] : (c.thr_recip ? e * 0.01f : e / 100.0f);
    }
};

// what a node of the level being built becomes; the level's scan keeps it in registers between the ranking and the emit
struct QDecision { uint32_t old_cb; uint8_t nonleaf, virt, srem, reached; };
SDT_HD uint32_t sdt_scan_count(const QDecision& d) { return d.nonleaf; }

SDT_HD QDecision sdt_q_decide(const RefineCtx& c, uint32_t id, uint32_t level) {
    QDecision d;
    d.nonleaf = 0; d.virt = 0; d.srem = 0; d.reached = 0; d.old_cb = 0;
    // everything indexed by the node itself is requested up front (the rebuild is a chain of launches that pays latency:
    // one round of loads, then the one dependent gather)
    const uint32_t kind = c.s_kind[id];
    const uint32_t srem = c.s_srem[id], src = c.s_src[id];            // one of the two is stale scratch, and unused
    const float e = c.energy1[id], thr = c.thr1[id];
    if (kind & SDT_KIND_VIRTUAL) {
        const uint32_t s = srem;
        d.nonleaf = s > 0u; d.virt = 1; d.srem = (uint8_t)(s ? s - 1u : 0u);
        return d;
    }
    const uint32_t cb = c.child0[src];
    const bool reached = (kind & SDT_KIND_REACHED) != 0u;
    if (c.no_quad) { d.nonleaf = cb != 0u; d.old_cb = cb; d.reached = reached; return d; }
    // merge pass (:574-611): a reached non-leaf below the threshold becomes a leaf
    const bool merged = cb != 0u && reached && (e < thr);
    if (cb != 0u && !merged) {
        d.nonleaf = 1; d.old_cb = cb; d.reached = reached && (e >= thr);
        return d;
    }
    // split pass (:617-637): E > thr and depth < maxDepth, children get E/4
    uint32_t s = 0;
    float v = e;
    const uint32_t maxd = c.H1->quad_max_depth;
    while (v > thr && level + s < maxd) { v = v / 4.0f; ++s; }
    d.nonleaf = s > 0u; d.virt = 1; d.srem = (uint8_t)(s ? s - 1u : 0u);
    return d;
}

struct QLevelFlag {
    RefineCtx c; uint32_t level; uint32_t last;   // last: deepest level the arena layout allows
    SDT_HD QDecision operator()(uint32_t i) const {
        // lvl_trunc = 1 + the level whose children no longer fitted the arena: every deeper level is all leaves.  (It is
        // written by that level's own fin, possibly while blocks of the same scan still emit: this level must not see it.)
        if ((c.H1->lvl_trunc && level >= c.H1->lvl_trunc) || last) { QDecision d; d.old_cb = 0; d.nonleaf = d.virt = d.srem = d.reached = 0; return d; }
        return sdt_q_decide(c, c.H1->level_off[level] + i, level);
    }
};
// children of this level's non-leaf nodes that still fit the arena (the scan may run emit before fin, so both derive
// the cut from the same numbers: a node whose four children would not fit becomes a leaf)
SDT_HD uint32_t sdt_q_fit(const DevHeader* H, uint32_t level) {
    const uint32_t next_off = H->level_off[level + 1u];
    return next_off >= H->quad_cap ? 0u : (H->quad_cap - next_off) / 4u;
}
struct QLevelFin {
    RefineCtx c; uint32_t level;
    SDT_HD void operator()(uint32_t total) const {
        DevHeader* H = c.H1;
        const uint32_t next_off = H->level_off[level + 1u];
        const uint32_t fit = sdt_q_fit(H, level);
        if (total > fit) {                                        // arena exhausted: cut the forest here
            H->error |= DEV_ERR_QUAD_CAPACITY;
            if (!H->lvl_trunc) H->lvl_trunc = level + 1u;
            total = fit;
        }
        H->level_cnt[level] = H->level_off[level + 1u] - H->level_off[level];
        H->level_off[level + 2u] = next_off + 4u * total;
        H->lvl_n[(level + 1u) & 1u] = 4u * total;
        H->int_off[level + 1u] = H->int_off[level] + total;      // non-leaf nodes of the levels above the next one
    }
};
struct QLevelEmit {
    RefineCtx c; uint32_t level;
    SDT_HD void operator()(uint32_t i, uint32_t rank, const QDecision& d) const {
        const DevHeader* H = c.H1;
        const uint32_t id = H->level_off[level] + i;
        const uint32_t fit = sdt_q_fit(H, level);
        // record index = number of non-leaf nodes in front of this one (records follow node order, levels are contiguous;
        // int_off[level] was written by the previous level's fin)
        c.iidx1[id] = H->int_off[level] + (rank < fit ? rank : fit);
        if (!d.nonleaf || rank >= fit) { c.child1[id] = 0u; return; }
        const uint32_t cb = H->level_off[level + 1u] + 4u * rank;
        c.child1[id] = cb;
        const float thr = c.thr1[id];
        // the children's pdf products, parents before children like PpLevelItem: pp * ((4 * E_child) / E_node)  (:1084)
        const float own = c.energy1[id], p = c.pp1[id];
        if (d.virt) {
            const float e4 = own / 4.0f;                          // :133-134
            for (uint32_t k = 0; k < 4u; ++k) {
                c.s_kind[cb + k] = SDT_KIND_VIRTUAL;
                c.s_srem[cb + k] = (uint8_t)d.srem;
                c.energy1[cb + k] = e4;
                c.pp1[cb + k] = p * ((4.0f * e4) / own);
                c.thr1[cb + k] = thr;                             // :139-143
            }
        } else {
            for (uint32_t k = 0; k < 4u; ++k) {
                const uint32_t o = d.old_cb + k;
                const float ec = c.e_cur[o];
                c.s_src[cb + k] = o;
                c.s_kind[cb + k] = d.reached ? SDT_KIND_REACHED : 0u;
                c.energy1[cb + k] = ec;
                c.pp1[cb + k] = p * ((4.0f * ec) / own);
                c.thr1[cb + k] = c.no_quad ? c.thr0[o] : thr;
            }
        }
    }
};
struct QFinalize {
    RefineCtx c; uint32_t levels_bound;
    SDT_HD void operator()() const {
        DevHeader* H = c.H1;
        H->n_quad = H->level_off[levels_bound];
        H->n_interior = H->int_off[levels_bound];
        uint32_t nl = 0;
        for (uint32_t l = 0; l < SDT_MAX_LEVELS + 1u; ++l) {
            if (l >= levels_bound) { H->level_off[l + 1u] = H->n_quad; }
            H->level_cnt[l] = H->level_off[l + 1u] - H->level_off[l];
            if (H->level_cnt[l]) nl = l + 1u;
        }
        H->n_levels = nl;
        H->kd_leaves = H->n_roots;
        H->root_of_node0 = c.kd_root[0];
    }
};

// ---- records for the query kernels ----------------------------------------------
// (uploaded trees only; the refine numbers the records in its level scans)
struct RecFlag { const uint32_t* child; SDT_HD uint32_t operator()(uint32_t i) const { return child[i] ? 1u : 0u; } };
struct RecEmit { uint32_t* iidx; SDT_HD void operator()(uint32_t i, uint32_t rank, uint32_t) const { iidx[i] = rank; } };
struct RecFin { DevHeader* H; SDT_HD void operator()(uint32_t total) const { H->n_interior = total; } };
struct RecBuildItem {
    const DevHeader* H; const uint32_t* child; const float* energy; const uint32_t* iidx; QRec* rec; uint32_t* root_iidx;
    float* zero;             // refine: the statistics of `current`, consumed by the level build, start again from 0 (:586)
    SDT_HD void operator()(uint32_t i) const {
        if (zero) zero[i] = 0.0f;
        const uint32_t cb = child[i];
        if (i < H->n_roots) root_iidx[i] = cb ? iidx[i] : SDT_NONE;
        if (!cb) return;
        QRec r;
        r.child_base = cb;
        r.interior_base = iidx[cb];
        uint32_t leafmask = 0;
        for (uint32_t k = 0; k < 4u; ++k) {
            if (!child[cb + k]) leafmask |= 1u << k;
            r.e[k] = energy[cb + k];
        }
        r.cinfo = sdt_make_cinfo(leafmask);
        r.child_base |= sdt_rec_flags(r.e[0], r.e[1], r.e[2], r.e[3]);
        r.own = energy[i];
        rec[iidx[i]] = r;
    }
};

// spatial leaf word <- record of the root of the leaf's quadtree (see the layout note in sdt_core.h)
struct KdLeafWordItem {
    DevHeader* H; uint32_t* kd_word; const uint32_t* kd_root; const uint32_t* root_iidx;
    float* cnt; float* prev;  // refine: prev.vertCount <- current.vertCount; current <- 0 (:141-153, :401-432, :585)
    SDT_HD void operator()(uint32_t i) const {
        if (cnt) { prev[i] = cnt[i]; cnt[i] = 0.0f; }
        const uint32_t ri = root_iidx[kd_root[i]];
        if (i == 0u) H->rootrec_of_node0 = ri;
        if (kd_word[i] & SDT_KD_LEAF_BIT) kd_word[i] = SDT_KD_LEAF_BIT | (ri == SDT_NONE ? 0x7FFFFFFFu : ri);
    }
};

// per-node pdf product, one level per pass, parents before children (see sdt_core.h):
// pp[root] = 1; pp[child c] = pp[node] * ((4 * E[child c]) / E[node])   (src/quadtree.py:1084)
struct PpLevelItem {
    const DevHeader* H; const uint32_t* child; const float* energy; float* pp; uint32_t level;
    SDT_HD void operator()(uint32_t i) const {
        const uint32_t id = H->level_off[level] + i;
        if (level == 0u) pp[id] = 1.0f;
        const uint32_t cb = child[id];
        if (!cb) return;
        const float p = pp[id], own = energy[id];
        for (uint32_t k = 0; k < 4u; ++k) pp[cb + k] = p * ((4.0f * energy[cb + k]) / own);   // NaN is sticky
    }
};

// jump table: trees with a root record are records 0 .. iidx[n_roots]-1 (records follow node order)
struct JumpCountItem {
    DevHeader* H; const uint32_t* iidx; uint32_t cap;
    SDT_HD uint32_t rec_at(uint32_t node) const { return node < H->n_quad ? iidx[node] : H->n_interior; }
    SDT_HD void operator()() const {
        const uint32_t trees = H->n_roots < H->n_quad ? iidx[H->n_roots] : H->n_interior;
        H->jump_trees = trees <= cap ? trees : 0u;            // does not fit: table off, kernels take the level-by-level path
        H->lvl_n[0] = H->jump_trees * (SDT_JUMP_CELLS / 4u);   // one thread per 2 x 2 block of cells
        // the non-leaf nodes of level SDT_JUMP_LEVELS (records follow node order, levels are contiguous): candidates for a
        // second-stage table
        H->s2_rec_lo = rec_at(H->level_off[SDT_JUMP_LEVELS]);
        H->s2_rec_hi = rec_at(H->level_off[SDT_JUMP_LEVELS + 1]);
        H->lvl_n[1] = H->jump_trees ? H->s2_rec_hi - H->s2_rec_lo : 0u;
        H->s2_tables = 0;
        H->kd_round_new = 0;                                  // scratch: second-stage entries to build
    }
};
// second-stage tables go to the level-SDT_JUMP_LEVELS nodes that still have a non-leaf child (one more level below them
// is answered by their own record just as fast)
struct S2Flag {
    const DevHeader* H; const QRec* rec;
    SDT_HD uint32_t operator()(uint32_t i) const { return rec[H->s2_rec_lo + i].cinfo != 0xFFFFFFFFu ? 1u : 0u; }
};
struct S2Emit {
    const DevHeader* H; uint32_t* s2_of; uint32_t* s2_rec; uint32_t cap;
    SDT_HD void operator()(uint32_t i, uint32_t rank, uint32_t v) const {
        s2_of[i] = (v && rank < cap) ? rank : SDT_NONE;
        if (v && rank < cap) s2_rec[rank] = H->s2_rec_lo + i;
    }
};
struct S2Fin {
    DevHeader* H; uint32_t cap;
    SDT_HD void operator()(uint32_t total) const {
        H->s2_tables = total <= cap ? total : cap;            // (beyond the cap: those nodes simply keep their record)
        H->kd_round_new = H->s2_tables * (SDT_S2_CELLS / 4u);
    }
};
struct JumpBuildItem {
    const DevHeader* H; const QRec* rec; const float* pp; const uint32_t* s2_of; QJump* jump; uint32_t* jump_pp;
    SDT_HD void operator()(uint32_t i) const {
        // One thread per 2 x 2 block of cells (sdt_build_jump4); the threads of a warp take 32 blocks in Z order (16 x 8
        // cells): their descents share the upper records and mostly end in the same few leaves, and the warp writes whole
        // sectors of the row-major table.
        constexpr uint32_t Q = SDT_JUMP_CELLS / 4u;
        const uint32_t tr = i / Q, z = i % Q;
        const uint32_t qx = sdt_even_bits(z), qy = sdt_even_bits(z >> 1);
        QJump j[4];
        sdt_build_jump4(rec, tr, qx, qy, SDT_JUMP_LEVELS, j);
        for (uint32_t k = 0; k < 4u; ++k) {
            if (!(j[k] & SDT_JUMP_LEAF) && H->s2_tables && j[k] >= H->s2_rec_lo && j[k] < H->s2_rec_hi) {
                const uint32_t tid = s2_of[j[k] - H->s2_rec_lo];
                if (tid != SDT_NONE) j[k] = SDT_JUMP_TABLE | tid;
            }
            const size_t o = (size_t)tr * SDT_JUMP_CELLS + (2u * qy + (k >> 1)) * SDT_JUMP_SIDE + 2u * qx + (k & 1u);
            jump[o] = j[k];
            jump_pp[o] = sdt_jump_pp_entry(j[k], pp);
        }
    }
};
struct S2BuildItem {
    const QRec* rec; const float* pp; const uint32_t* s2_rec; uint32_t* s2; uint32_t* s2_pp;
    SDT_HD void operator()(uint32_t i) const {
        constexpr uint32_t Q = SDT_S2_CELLS / 4u;                                    // 2 x 2 blocks in Z order, as above
        const uint32_t tid = i / Q, z = i % Q;
        const uint32_t qx = sdt_even_bits(z), qy = sdt_even_bits(z >> 1);
        QJump j[4];
        sdt_build_jump4(rec, s2_rec[tid], qx, qy, SDT_S2_LEVELS, j);
        for (uint32_t k = 0; k < 4u; ++k) {
            const size_t o = (size_t)tid * SDT_S2_CELLS + (2u * qy + (k >> 1)) * SDT_S2_SIDE + 2u * qx + (k & 1u);
            s2[o] = j[k];
            s2_pp[o] = sdt_jump_pp_entry(j[k], pp);
        }
    }
};

struct QFinalizeAndCount {   // last step of the forest rebuild + the sizes of the tables that follow: one launch
    QFinalize fin; JumpCountItem cnt;
    SDT_HD void operator()() const { fin(); cnt(); }
};

struct KdGridItem { const uint32_t* kd_word; uint32_t* grid; SDT_HD void operator()(uint32_t c) const { grid[c] = sdt_kd_grid_node(kd_word, c); } };

// uploaded: the tree came through sdt_upload -- record indices (one scan over all nodes), per-node pdf products (one pass
// per level) and the table sizes are computed here; the refine has them already: it numbers the records and writes the
// products while it builds the levels, and sizes the tables in its finalize step
static void sdt_build_records(sdt_handle h, const ExecCtx& x, QuadSet& s, bool uploaded) {
    if (uploaded) launch_scan(x, &s.hdr->n_quad, 0, RecFlag{s.child}, RecEmit{s.iidx}, RecFin{s.hdr});
    launch_items(x, &s.hdr->n_quad, 0, RecBuildItem{s.hdr, s.child, s.energy, s.iidx, s.rec, s.root_iidx, uploaded ? nullptr : h->q_ecur});
    launch_items(x, &s.hdr->n_kd, 0, KdLeafWordItem{s.hdr, h->kd_word, h->kd_root, s.root_iidx,
                                                    uploaded ? nullptr : h->kd_count, uploaded ? nullptr : h->kd_prev_count});
    launch_items(x, nullptr, SDT_GRID_CELLS, KdGridItem{h->kd_word, h->kd_grid});
    for (uint32_t l = 0; uploaded && l < h->levels_hint && l < SDT_MAX_LEVELS; ++l)
        launch_items(x, &s.hdr->level_cnt[l], 0, PpLevelItem{s.hdr, s.child, s.energy, s.pp, l});
    if (uploaded) launch_single(x, JumpCountItem{s.hdr, s.iidx, h->jump_cap});
    launch_scan(x, &s.hdr->lvl_n[1], 0, S2Flag{s.hdr, s.rec}, S2Emit{s.hdr, s.s2_of, s.s2_rec, h->s2_cap}, S2Fin{s.hdr, h->s2_cap});
    launch_items(x, &s.hdr->lvl_n[0], 0, JumpBuildItem{s.hdr, s.rec, s.pp, s.s2_of, s.jump, s.jump_pp});
    launch_items(x, &s.hdr->kd_round_new, 0, S2BuildItem{s.rec, s.pp, s.s2_rec, s.s2, s.s2_pp});
}

struct ZeroItem { float* p; SDT_HD void operator()(uint32_t i) const { p[i] = 0.0f; } };

// the launch sequence of one refine on stream `st` (sweeps of the statistics included when they are due)
// split rounds the next refine can need at most (see sdt_tree_s::splat_bound): KDTree.refine's own loop (:346-347, :261-264)
// run on the largest count any leaf can hold
static uint32_t sdt_kd_rounds_bound(sdt_handle h) {
    const uint32_t maxd = (uint32_t)h->cfg.kd_max_depth;
    if (!h->splat_bound_valid) return maxd;
    float v = h->splat_bound > 16777216ull ? 16777216.0f : (float)h->splat_bound;
    const float T = h->max_leaf_host;
    uint32_t s = 0;
    while (v > T && s < maxd) { if (v > 0.0f) v = v / 2.0f; ++s; }
    return s;
}

static int sdt_refine_enqueue(sdt_handle h, cudaStream_t st, uint32_t flags, uint32_t levels_bound, uint32_t kd_rounds) {
    SDT_TRY(sdt_complete_stats(h, st, false));
    const ExecCtx x = exec_ctx(h, st);
    QuadSet& s0 = h->set[h->cur];
    QuadSet& s1 = h->set[1 - h->cur];
    RefineCtx c;
    c.H0 = s0.hdr; c.H1 = s1.hdr;
    c.kd_word = h->kd_word; c.kd_count = h->kd_count; c.kd_depth = h->kd_depth; c.kd_root = h->kd_root;
    c.kd_bmin = h->kd_bmin; c.kd_bmax = h->kd_bmax; c.kd_prev_count = h->kd_prev_count; c.kd_s = h->kd_s;
    c.kd_sel = h->kd_sel; c.kd_rank_cur = h->kd_rank[0]; c.kd_rank_prev = h->kd_rank[1]; c.root_src = h->root_src;
    c.child0 = s0.child; c.thr0 = s0.thr; c.e_cur = h->q_ecur;
    c.child1 = s1.child; c.energy1 = s1.energy; c.thr1 = s1.thr; c.iidx1 = s1.iidx; c.rec1 = s1.rec; c.root_iidx1 = s1.root_iidx; c.pp1 = s1.pp;
    c.s_src = h->s_src; c.s_kind = h->s_kind; c.s_srem = h->s_srem;
    c.no_quad = (flags & SDT_REFINE_NO_QUAD) ? 1u : 0u;
    c.kd_rounds = kd_rounds;
    c.thr_recip = h->quad_thr_reciprocal ? 1u : 0u;

    launch_items(x, nullptr, (uint32_t)(sizeof(DevHeader) / 4u), RefineInit{c});
    if (flags & SDT_REFINE_NO_KD) launch_items(x, &c.H1->n_roots_old, 0, RootIdentityItem{h->root_src});
    else {
        launch_items(x, &c.H1->kd_n_old, 0, KdLevelsItem{c});
        for (uint32_t r = 1; r <= kd_rounds; ++r) {
            RefineCtx cr = c;
            cr.kd_rank_cur = h->kd_rank[r & 1u]; cr.kd_rank_prev = h->kd_rank[(r & 1u) ^ 1u];
            launch_scan(x, &c.H1->kd_n_old, 0, KdRoundFlag{cr, r}, KdRoundEmit{cr}, KdRoundFin{cr, r});
            launch_items(x, &c.H1->kd_round_new, 0, KdMakeNodeItem{cr, r});
        }
    }
    launch_items(x, &c.H1->n_roots, 0, QRootItem{c});
    for (uint32_t l = 0; l < levels_bound; ++l)
        launch_scan(x, &c.H1->lvl_n[l & 1u], 0, QLevelFlag{c, l, (uint32_t)(l + 1u == levels_bound)}, QLevelEmit{c, l}, QLevelFin{c, l});
    launch_single(x, QFinalizeAndCount{QFinalize{c, levels_bound}, JumpCountItem{s1.hdr, s1.iidx, h->jump_cap}});
    h->levels_hint = levels_bound;
    sdt_build_records(h, x, s1, false);
    // (prev <- current and the reset of current, :582-586, ride on the record / leaf-word passes above)
    return sdt_post_launch(h, "sdt_refine");
}

extern "C" int sdt_refine(sdt_handle h, uint32_t flags, sdt_stream stream) {
    SDT_ENTER(h);
    cudaStream_t st = (cudaStream_t)stream;
    sdt_order_after_last(h, st);
    uint32_t levels_bound = (uint32_t)h->cfg.quad_max_depth + 1u;
    if (levels_bound < h->levels_hint) levels_bound = h->levels_hint;
    if (levels_bound > SDT_MAX_LEVELS) levels_bound = SDT_MAX_LEVELS;
    QuadSet& s1 = h->set[1 - h->cur];
    const uint32_t kd_rounds = sdt_kd_rounds_bound(h);
#ifndef SDT_HOSTEMU
    // The sequence is some 45 small dependent kernels whose arguments only depend on the buffer parity and a few settings:
    // it is captured once per such combination into a CUDA graph (programmatic-dependent-launch edges included) and
    // replayed with ONE launch per training iteration (the device-side chain is the bound: graph and plain launches time the same).
    bool done = false;
    if (h->use_graph) {
        const sdt_tree_s::RefineKey key{h->cur, flags & (SDT_REFINE_NO_KD | SDT_REFINE_NO_QUAD), h->levels_hint, levels_bound,
                                        (int)kd_rounds, h->quad_thr_reciprocal, h->use_pdl, h->stats_complete ? 1 : 0, sdt_sweep_levels(h)};
        auto it = h->refine_graphs.find(key);
        if (it == h->refine_graphs.end()) {
            sdt_tree_s::RefineGraph g;
            const uint64_t l0 = h->launches;
            const bool sc0 = h->stats_complete; const uint32_t lh0 = h->levels_hint;
            cudaGraph_t graph = nullptr;
            if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                const int rc = sdt_refine_enqueue(h, st, flags, levels_bound, kd_rounds);
                const cudaError_t ec = cudaStreamEndCapture(st, &graph);
                if (rc == SDT_OK && ec == cudaSuccess && graph && cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess) {
                    g.launches = h->launches - l0;
                    it = h->refine_graphs.emplace(key, g).first;
                }
                if (graph) cudaGraphDestroy(graph);
            }
            h->launches = l0; h->stats_complete = sc0; h->levels_hint = lh0;      // nothing has run yet
            if (it == h->refine_graphs.end()) { cudaGetLastError(); h->use_graph = 0; }   // capture not possible here: plain launches from now on
        }
        if (it != h->refine_graphs.end()) {
            SDT_CUDA(h, cudaGraphLaunch(it->second.exec, st));
            h->launches += it->second.launches;
            h->last_stream = st;
            done = true;
        }
    }
    if (!done) SDT_TRY(sdt_refine_enqueue(h, st, flags, levels_bound, kd_rounds));
#else
    SDT_TRY(sdt_refine_enqueue(h, st, flags, levels_bound, kd_rounds));
#endif
    h->prev_kd_dirty = !h->kd_complete;          // un-swept interior counts were rolled into prev
    h->cur = 1 - h->cur;
    h->jump_trees_known = 0;
    h->levels_known = 0;
    h->levels_hint = levels_bound;
    h->stats_complete = true; h->kd_complete = true;
    h->splat_bound = 0; h->splat_bound_valid = true;          // current's statistics are zero again
    // non-blocking read-back of the new sizes (only used to size the smem staging of later launches)
    if (cudaMemcpyAsync(h->h_hdr, s1.hdr, sizeof(DevHeader), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
        cudaEventRecord(h->hdr_event, st) == cudaSuccess) h->hdr_pending = true;
    if (flags & SDT_SYNC) SDT_CUDA(h, cudaStreamSynchronize(st));
    return SDT_OK;
}

extern "C" int sdt_hint_records(sdt_handle h, uint64_t records_all_ranks) {
    SDT_ENTER(h);
    h->splat_bound = records_all_ranks;
    h->splat_bound_valid = true;
    return SDT_OK;
}

extern "C" int sdt_reset_stats(sdt_handle h, sdt_stream stream) {
    SDT_ENTER(h);
    cudaStream_t st = (cudaStream_t)stream;
    sdt_order_after_last(h, st);
    const ExecCtx x = exec_ctx(h, st);
    QuadSet& s = h->set[h->cur];
    launch_items(x, &s.hdr->n_kd, 0, ZeroItem{h->kd_count});
    launch_items(x, &s.hdr->n_quad, 0, ZeroItem{h->q_ecur});
    h->stats_complete = true; h->kd_complete = true;
    h->splat_bound = 0; h->splat_bound_valid = true;
    return sdt_post_launch(h, "sdt_reset_stats");
}
