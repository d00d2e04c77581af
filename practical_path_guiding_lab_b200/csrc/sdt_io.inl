// Lifetime, tree exchange in the reference's npz schema, thresholds, introspection.

template <class T>
static int dev_alloc(sdt_handle h, T** p, size_t n) {
    SDT_CUDA(h, cudaMalloc((void**)p, (n ? n : 1) * sizeof(T)));
    SDT_CUDA(h, cudaMemsetAsync(*p, 0, (n ? n : 1) * sizeof(T), nullptr));
    return SDT_OK;
}

static int sdt_free_all(sdt_handle h) {
    void* ptrs[] = {h->kd_word, h->kd_count, h->kd_depth, h->kd_root, h->kd_bmin, h->kd_bmax, h->kd_prev_count, h->kd_s,
                    h->kd_sel, h->kd_rank[0], h->kd_rank[1], h->root_src, h->kd_grid, h->q_ecur, h->s_src, h->s_kind, h->s_srem, h->s_blk, h->stage, h->stage_o};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int k = 0; k < 2; ++k) {
        QuadSet& s = h->set[k];
        void* q[] = {s.child, s.energy, s.thr, s.pp, s.iidx, s.rec, s.jump, s.jump_pp, s.s2, s.s2_pp, s.s2_rec, s.s2_of, s.root_iidx, s.hdr};
        for (void* p : q) if (p) cudaFree(p);
    }
#ifndef SDT_HOSTEMU
    for (auto& kv : h->refine_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    h->refine_graphs.clear();
#endif
    if (h->h_hdr) cudaFreeHost(h->h_hdr);
    if (h->hdr_event) cudaEventDestroy(h->hdr_event);
#ifndef SDT_HOSTEMU
    if (h->order_event) cudaEventDestroy(h->order_event);
#endif
#ifndef SDT_HOSTEMU
    for (int k = 0; k < 2; ++k) { if (h->ev_in[k]) cudaEventDestroy(h->ev_in[k]); if (h->ev_comp[k]) cudaEventDestroy(h->ev_comp[k]); if (h->ev_out[k]) cudaEventDestroy(h->ev_out[k]); }
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
#endif
    return SDT_OK;
}

struct SetLeafSize { DevHeader* H; float v; SDT_HD void operator()() const { H->max_leaf_size = v; } };

extern "C" const char* sdt_last_error(sdt_handle h) { return h ? h->err.c_str() : g_create_err.c_str(); }

// header of a tree with one spatial leaf owning one single-node quadtree
static void sdt_initial_header(sdt_handle h, DevHeader& H) {
    memset(&H, 0, sizeof(H));
    H.n_kd = 1; H.n_quad = 1; H.n_roots = 1; H.n_interior = 0; H.n_levels = 1; H.kd_leaves = 1;
    H.kd_max_depth = (uint32_t)h->cfg.kd_max_depth; H.quad_max_depth = (uint32_t)h->cfg.quad_max_depth;
    H.store_nee = (uint32_t)(h->cfg.store_nee != 0);
    for (int a = 0; a < 3; ++a) { H.bbox_min[a] = h->cfg.bbox_min[a]; H.bbox_max[a] = h->cfg.bbox_max[a]; }
    H.max_leaf_size = 1.0f;                      // KDTree(max_leaf_size=1), src/kdtree.py:117
    H.kd_cap = h->kd_cap; H.quad_cap = h->quad_cap;
    H.rootrec_of_node0 = SDT_NONE;
    for (int l = 1; l < SDT_MAX_LEVELS + 2; ++l) H.level_off[l] = 1;
    H.level_cnt[0] = 1;
}

extern "C" int sdt_create(const sdt_config* cfg, sdt_handle* out) {
    if (!cfg || !out) return sdt_fail(nullptr, SDT_ERR_INVALID, "sdt_create: NULL argument");
    if (cfg->kd_max_depth < 0 || cfg->kd_max_depth > SDT_KD_MAX_DEPTH) return sdt_fail(nullptr, SDT_ERR_INVALID, "sdt_create: kd_max_depth must be in [0,40]");
    if (cfg->quad_max_depth < 0 || cfg->quad_max_depth > 32) return sdt_fail(nullptr, SDT_ERR_INVALID, "sdt_create: quad_max_depth must be in [0,32]");
    sdt_handle h = new sdt_tree_s();
    h->cfg = *cfg;
    h->kd_cap = cfg->kd_capacity ? cfg->kd_capacity : (1u << 21);
    h->quad_cap = cfg->quad_capacity ? cfg->quad_capacity : (1u << 26);      // ~4.5 GB of the B200's 180 GB: room for 3840x2160-scale forests
    if (h->kd_cap < 1) h->kd_cap = 1;
    if (h->quad_cap < 4) h->quad_cap = 4;
    h->rec_cap = h->quad_cap / 4u + 1u;
    h->jump_cap = h->kd_cap / 2u + 1u < 262144u ? h->kd_cap / 2u + 1u : 262144u;   // one table per tree = per spatial leaf
    if ((uint64_t)h->jump_cap * SDT_JUMP_CELLS > 4ull * h->quad_cap) h->jump_cap = (uint32_t)(4ull * h->quad_cap / SDT_JUMP_CELLS);
    h->s2_cap = h->rec_cap < (1u << 19) ? h->rec_cap : (1u << 19);       // second-stage tables: 256 B each, two kinds
#ifndef SDT_HOSTEMU
    {
        cudaError_t e = cudaSetDevice(cfg->device);
        if (e != cudaSuccess) { std::string m = std::string("cudaSetDevice: ") + cudaGetErrorString(e); delete h; return sdt_fail(nullptr, SDT_ERR_CUDA, m); }
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device) == cudaSuccess && sms > 0) h->num_sms = sms;
    }
#endif
    int st = SDT_OK;
#define A(p, n) if (st == SDT_OK) st = dev_alloc(h, &(p), (size_t)(n))
    A(h->kd_word, h->kd_cap); A(h->kd_count, h->kd_cap); A(h->kd_depth, h->kd_cap); A(h->kd_root, h->kd_cap);
    A(h->kd_bmin, 3ull * h->kd_cap); A(h->kd_bmax, 3ull * h->kd_cap); A(h->kd_prev_count, h->kd_cap); A(h->kd_s, h->kd_cap);
    A(h->kd_sel, h->kd_cap); A(h->kd_rank[0], h->kd_cap); A(h->kd_rank[1], h->kd_cap); A(h->root_src, h->kd_cap); A(h->kd_grid, SDT_GRID_CELLS);
    A(h->q_ecur, h->quad_cap); A(h->s_src, h->quad_cap); A(h->s_kind, h->quad_cap); A(h->s_srem, h->quad_cap);
    A(h->s_blk, SDT_SCAN_STATE_WORDS + 64);
    for (int k = 0; k < 2; ++k) {
        QuadSet& s = h->set[k];
        A(s.child, h->quad_cap); A(s.energy, h->quad_cap); A(s.thr, h->quad_cap); A(s.pp, h->quad_cap); A(s.iidx, h->quad_cap);
        A(s.rec, h->rec_cap); A(s.jump, (size_t)h->jump_cap * SDT_JUMP_CELLS); A(s.jump_pp, (size_t)h->jump_cap * SDT_JUMP_CELLS);
        A(s.s2, (size_t)h->s2_cap * SDT_S2_CELLS); A(s.s2_pp, (size_t)h->s2_cap * SDT_S2_CELLS); A(s.s2_rec, h->s2_cap); A(s.s2_of, h->rec_cap); A(s.root_iidx, h->kd_cap); A(s.hdr, 1);
    }
#undef A
    if (st == SDT_OK && cudaMallocHost((void**)&h->h_hdr, sizeof(DevHeader)) != cudaSuccess) st = sdt_fail(h, SDT_ERR_CUDA, "cudaMallocHost failed");
    if (st == SDT_OK && cudaEventCreate(&h->hdr_event) != cudaSuccess) st = sdt_fail(h, SDT_ERR_CUDA, "cudaEventCreate failed");
#ifndef SDT_HOSTEMU
    if (st == SDT_OK) {
        bool ok = cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking) == cudaSuccess;
        for (int k = 0; k < 2 && ok; ++k)
            ok = cudaEventCreateWithFlags(&h->ev_in[k], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&h->ev_comp[k], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&h->ev_out[k], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&h->order_event, cudaEventDisableTiming) == cudaSuccess;
        if (!ok) st = sdt_fail(h, SDT_ERR_CUDA, "staging pipeline streams/events could not be created");
    }
#endif
    if (st == SDT_OK) {
        // initial tree: src/kdtree.py:117-130, src/quadtree.py:350-362
        DevHeader H;
        sdt_initial_header(h, H);
        const uint32_t word0 = SDT_KD_LEAF_BIT | 0x7FFFFFFFu, none = SDT_NONE;
        const float inf = INFINITY;
        cudaMemcpy(h->set[0].hdr, &H, sizeof(H), cudaMemcpyHostToDevice);
        cudaMemcpy(h->set[1].hdr, &H, sizeof(H), cudaMemcpyHostToDevice);
        cudaMemcpy(h->kd_word, &word0, 4, cudaMemcpyHostToDevice);
        cudaMemcpy(h->kd_bmin, cfg->bbox_min, 12, cudaMemcpyHostToDevice);
        cudaMemcpy(h->kd_bmax, cfg->bbox_max, 12, cudaMemcpyHostToDevice);
        cudaMemcpy(h->set[0].thr, &inf, 4, cudaMemcpyHostToDevice);
        cudaMemcpy(h->set[0].root_iidx, &none, 4, cudaMemcpyHostToDevice);
        if (cudaDeviceSynchronize() != cudaSuccess) st = sdt_fail(h, SDT_ERR_CUDA, "sdt_create: initialisation failed");
        h->cur = 0; h->levels_hint = 1; h->stats_complete = true; h->kd_complete = true;
    }
    if (st != SDT_OK) { g_create_err = h->err; sdt_free_all(h); delete h; return st; }
    *out = h;
    return SDT_OK;
}

extern "C" int sdt_destroy(sdt_handle h) {
    SDT_ENTER(h);
    cudaDeviceSynchronize();
    sdt_nccl_destroy(h);
    sdt_free_all(h);
    delete h;
    return SDT_OK;
}

static int sdt_read_header(sdt_handle h, DevHeader& H) {
    SDT_CUDA(h, cudaStreamSynchronize(h->last_stream));
    SDT_CUDA(h, cudaMemcpy(h->h_hdr, h->set[h->cur].hdr, sizeof(DevHeader), cudaMemcpyDeviceToHost));
    H = *h->h_hdr;
#ifndef SDT_HOSTEMU
    uint32_t stall = 0;
    SDT_CUDA(h, cudaMemcpy(&stall, h->s_blk + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (stall) H.error |= DEV_ERR_SCAN_STALL;
#endif
    h->hdr_pending = false;
    h->kd_nodes_known = H.n_kd;
    h->n_quad_known = H.n_quad;
    h->jump_trees_known = H.jump_trees;
    h->dev_error_seen = H.error;
    h->levels_known = H.n_levels;
    return SDT_OK;
}

extern "C" int sdt_get_sizes(sdt_handle h, sdt_sizes* out) {
    if (!h || !out) return SDT_ERR_INVALID;
    SDT_ENTER(h);
    DevHeader H;
    SDT_TRY(sdt_read_header(h, H));
    out->n_kd = H.n_kd; out->n_quad = H.n_quad; out->n_roots = H.n_roots; out->n_interior = H.n_interior;
    out->n_levels = H.n_levels; out->kd_leaves = H.kd_leaves; out->error = H.error; out->refine_count = H.refine_count; out->jump_trees = H.jump_trees; out->jump2_tables = H.s2_tables;
    h->levels_hint = H.n_levels > 0 ? H.n_levels : 1;
    return SDT_OK;
}

// ---------------------------------------------------------------------------- upload
extern "C" int sdt_upload(sdt_handle h, const sdt_arrays* a) {
    if (!h || !a) return SDT_ERR_INVALID;
    SDT_ENTER(h);
    SDT_CHECK(h, a->n_kd >= 1 && a->n_quad >= 1 && a->n_roots >= 1, SDT_ERR_LAYOUT, "sdt_upload: empty tree");
    SDT_CHECK(h, a->n_kd <= h->kd_cap && a->n_roots <= h->kd_cap, SDT_ERR_CAPACITY, "sdt_upload: spatial arena too small");
    SDT_CHECK(h, a->n_quad <= h->quad_cap, SDT_ERR_CAPACITY, "sdt_upload: quadtree arena too small");
    SDT_CHECK(h, a->kd_max_depth >= 0 && a->kd_max_depth <= SDT_KD_MAX_DEPTH && a->quad_max_depth >= 0 && a->quad_max_depth <= 32,
              SDT_ERR_INVALID, "sdt_upload: max depth out of range");
    SDT_CHECK(h, a->kd_bbox_min && a->kd_bbox_max && a->kd_depth && a->kd_is_leaf && a->kd_quad_root && a->kd_child_left &&
                     a->kd_child_right && a->q_root_node && a->q_irradiance && a->q_is_leaf && a->q_child[0] && a->q_child[1] &&
                     a->q_child[2] && a->q_child[3], SDT_ERR_INVALID, "sdt_upload: NULL array");
    const uint32_t nk = a->n_kd, nq = a->n_quad, R = a->n_roots;
    // spatial tree: children must be adjacent (the reference's split appends them so)
    std::vector<uint32_t> word(nk);
    std::vector<uint8_t> root_used(R, 0);
    for (uint32_t i = 0; i < nk; ++i) {
        if (a->kd_is_leaf[i]) {
            SDT_CHECK(h, a->kd_quad_root[i] < R, SDT_ERR_LAYOUT, "sdt_upload: quadTreeRootIndex out of range");
            SDT_CHECK(h, !root_used[a->kd_quad_root[i]], SDT_ERR_LAYOUT, "sdt_upload: two spatial leaves share a quadtree");
            root_used[a->kd_quad_root[i]] = 1;
            word[i] = SDT_KD_LEAF_BIT | a->kd_quad_root[i];
        } else {
            const uint32_t l = a->kd_child_left[i], r = a->kd_child_right[i];
            SDT_CHECK(h, r == l + 1u && r < nk && l > i, SDT_ERR_LAYOUT, "sdt_upload: spatial children must be adjacent and after their parent");
            SDT_CHECK(h, a->kd_depth[l] == a->kd_depth[i] + 1u && a->kd_depth[r] == a->kd_depth[i] + 1u, SDT_ERR_LAYOUT, "sdt_upload: spatial depth array inconsistent");
            // the descent recomputes the split planes instead of reading child boxes: the stored boxes must be what
            // KDTree.split writes (src/kdtree.py:268-304) -- axis depth % 3 halved at fp32 (min + max) / 2
            const uint32_t ax = a->kd_depth[i] % 3u;
            const float* pmin = a->kd_bbox_min + 3ull * i; const float* pmax = a->kd_bbox_max + 3ull * i;
            const float mid = (pmin[ax] + pmax[ax]) / 2.0f;
            bool ok = true;
            for (uint32_t k = 0; k < 3u; ++k) {
                const float lmax = k == ax ? mid : pmax[k], rmin = k == ax ? mid : pmin[k];
                ok = ok && a->kd_bbox_min[3ull * l + k] == pmin[k] && a->kd_bbox_max[3ull * l + k] == lmax &&
                     a->kd_bbox_min[3ull * r + k] == rmin && a->kd_bbox_max[3ull * r + k] == pmax[k];
            }
            SDT_CHECK(h, ok, SDT_ERR_LAYOUT, "sdt_upload: spatial child boxes are not the midpoint split of their parent on axis depth % 3");
            word[i] = l;
        }
    }
    SDT_CHECK(h, a->kd_depth[0] == 0, SDT_ERR_LAYOUT, "sdt_upload: spatial root must have depth 0");
    // quadtrees: relabel to the canonical layout of clearTreeUnusedNode (multi-root BFS,
    // src/quadtree.py:695-828): roots first, then per level the 4 children of every non-leaf
    std::vector<uint32_t> order;            // new id -> old id
    order.reserve(nq);
    std::vector<uint32_t> child_new, level_off;
    std::vector<uint8_t> seen(nq, 0);
    for (uint32_t r = 0; r < R; ++r) {
        const uint32_t o = a->q_root_node[r];
        SDT_CHECK(h, o < nq && !seen[o], SDT_ERR_LAYOUT, "sdt_upload: rootNodeIndex invalid");
        seen[o] = 1; order.push_back(o);
    }
    level_off.push_back(0);
    size_t lo = 0;
    child_new.assign(order.size(), 0);
    while (lo < order.size()) {
        const size_t hi = order.size();
        level_off.push_back((uint32_t)hi);
        SDT_CHECK(h, level_off.size() <= SDT_MAX_LEVELS + 1, SDT_ERR_LAYOUT, "sdt_upload: quadtree deeper than 33 levels");
        for (size_t j = lo; j < hi; ++j) {
            const uint32_t o = order[j];
            if (a->q_is_leaf[o]) continue;
            child_new[j] = (uint32_t)order.size();
            for (int k = 0; k < 4; ++k) {
                const uint32_t c = a->q_child[k][o];
                SDT_CHECK(h, c < nq && !seen[c], SDT_ERR_LAYOUT, "sdt_upload: quadtree child index invalid or shared");
                seen[c] = 1; order.push_back(c);
            }
            child_new.resize(order.size(), 0);
        }
        lo = hi;
    }
    const uint32_t nq_new = (uint32_t)order.size();
    std::vector<float> energy(nq_new), thr(nq_new);
    for (uint32_t j = 0; j < nq_new; ++j) {
        energy[j] = a->q_irradiance[order[j]];
        thr[j] = a->q_threshold ? a->q_threshold[order[j]] : INFINITY;
    }
    h->cfg.kd_max_depth = a->kd_max_depth; h->cfg.quad_max_depth = a->quad_max_depth; h->cfg.store_nee = a->quad_store_nee;
    for (int k = 0; k < 3; ++k) { h->cfg.bbox_min[k] = a->kd_bbox_min[k]; h->cfg.bbox_max[k] = a->kd_bbox_max[k]; }
    DevHeader H;
    sdt_initial_header(h, H);
    H.n_kd = nk; H.n_quad = nq_new; H.n_roots = R; H.kd_leaves = R;
    H.root_of_node0 = a->kd_quad_root[0];
    H.max_leaf_size = a->kd_max_leaf_size;
    const uint32_t nlev = (uint32_t)level_off.size() - 1u;
    H.n_levels = nlev;
    for (uint32_t l = 0; l < SDT_MAX_LEVELS + 2; ++l) {
        H.level_off[l] = l < level_off.size() ? level_off[l] : nq_new;
        H.level_cnt[l] = 0;
    }
    for (uint32_t l = 0; l < nlev; ++l) H.level_cnt[l] = level_off[l + 1] - level_off[l];

    SDT_CUDA(h, cudaStreamSynchronize(h->last_stream));
    h->cur = 0;
    QuadSet& s = h->set[0];
    SDT_CUDA(h, cudaMemcpy(h->kd_word, word.data(), 4ull * nk, cudaMemcpyHostToDevice));
    SDT_CUDA(h, cudaMemcpy(h->kd_depth, a->kd_depth, 4ull * nk, cudaMemcpyHostToDevice));
    SDT_CUDA(h, cudaMemcpy(h->kd_root, a->kd_quad_root, 4ull * nk, cudaMemcpyHostToDevice));
    SDT_CUDA(h, cudaMemcpy(h->kd_bmin, a->kd_bbox_min, 12ull * nk, cudaMemcpyHostToDevice));
    SDT_CUDA(h, cudaMemcpy(h->kd_bmax, a->kd_bbox_max, 12ull * nk, cudaMemcpyHostToDevice));
    if (a->kd_vert_count) SDT_CUDA(h, cudaMemcpy(h->kd_prev_count, a->kd_vert_count, 4ull * nk, cudaMemcpyHostToDevice));
    else SDT_CUDA(h, cudaMemsetAsync(h->kd_prev_count, 0, 4ull * nk, nullptr));
    SDT_CUDA(h, cudaMemcpy(s.child, child_new.data(), 4ull * nq_new, cudaMemcpyHostToDevice));
    SDT_CUDA(h, cudaMemcpy(s.energy, energy.data(), 4ull * nq_new, cudaMemcpyHostToDevice));
    SDT_CUDA(h, cudaMemcpy(s.thr, thr.data(), 4ull * nq_new, cudaMemcpyHostToDevice));
    SDT_CUDA(h, cudaMemcpy(s.hdr, &H, sizeof(H), cudaMemcpyHostToDevice));
    // current <- same topology, zero statistics (src/path_guiding_integrator.py:603-608)
    SDT_CUDA(h, cudaMemsetAsync(h->kd_count, 0, 4ull * h->kd_cap, nullptr));
    SDT_CUDA(h, cudaMemsetAsync(h->q_ecur, 0, 4ull * h->quad_cap, nullptr));
    const ExecCtx x = exec_ctx(h, nullptr);
    h->levels_hint = nlev > 0 ? nlev : 1;          // the per-level passes of sdt_build_records cover the uploaded tree
    sdt_build_records(h, x, s, true);
    SDT_TRY(sdt_post_launch(h, "sdt_upload"));
    SDT_CUDA(h, cudaStreamSynchronize(nullptr));
    h->levels_hint = nlev > 0 ? nlev : 1;
    h->prev_kd_dirty = false;                      // the uploaded vertCount array is complete
    h->splat_bound = 0; h->splat_bound_valid = true; h->max_leaf_host = a->kd_max_leaf_size;
    h->kd_nodes_known = nk;
    h->hdr_pending = false;
    h->stats_complete = true; h->kd_complete = true;
    { DevHeader H2; SDT_TRY(sdt_read_header(h, H2)); }      // jump_trees of the uploaded tree
    return SDT_OK;
}

extern "C" int sdt_upload_stats(sdt_handle h, const float* q_irradiance, const float* kd_vert_count) {
    SDT_ENTER(h);
    SDT_TRY(sdt_complete_stats(h, h->last_stream));       // whatever is not overwritten below stays consistent
    DevHeader H;
    SDT_TRY(sdt_read_header(h, H));
    if (q_irradiance) SDT_CUDA(h, cudaMemcpy(h->q_ecur, q_irradiance, 4ull * H.n_quad, cudaMemcpyHostToDevice));
    if (kd_vert_count) SDT_CUDA(h, cudaMemcpy(h->kd_count, kd_vert_count, 4ull * H.n_kd, cudaMemcpyHostToDevice));
    h->stats_complete = true; h->kd_complete = true;       // the caller's interior values are taken as they are
    h->splat_bound_valid = false;
    return SDT_OK;
}

// ---------------------------------------------------------------------------- download
extern "C" int sdt_download(sdt_handle h, int which, sdt_arrays* out) {
    if (!h || !out) return SDT_ERR_INVALID;
    SDT_ENTER(h);
    SDT_CHECK(h, which == SDT_TREE_PREV || which == SDT_TREE_CURRENT, SDT_ERR_INVALID, "sdt_download: which must be 0 or 1");
    if (which == SDT_TREE_CURRENT) SDT_TRY(sdt_complete_stats(h, h->last_stream));
    else SDT_TRY(sdt_sweep_prev_counts(h, h->last_stream));
    DevHeader H;
    SDT_TRY(sdt_read_header(h, H));
    SDT_CHECK(h, out->n_kd >= H.n_kd && out->n_quad >= H.n_quad && out->n_roots >= H.n_roots, SDT_ERR_CAPACITY,
              "sdt_download: output arrays smaller than the tree (fill n_kd/n_quad/n_roots from sdt_get_sizes)");
    const uint32_t nk = H.n_kd, nq = H.n_quad, R = H.n_roots;
    const QuadSet& s = h->set[h->cur];
    out->n_kd = nk; out->n_quad = nq; out->n_roots = R;
    out->kd_max_leaf_size = H.max_leaf_size; out->kd_max_depth = (int32_t)H.kd_max_depth;
    out->quad_max_depth = (int32_t)H.quad_max_depth; out->quad_store_nee = (int32_t)H.store_nee;
    std::vector<uint32_t> word(nk), child(nq);
    SDT_CUDA(h, cudaMemcpy(word.data(), h->kd_word, 4ull * nk, cudaMemcpyDeviceToHost));
    SDT_CUDA(h, cudaMemcpy(child.data(), s.child, 4ull * nq, cudaMemcpyDeviceToHost));
    if (out->kd_bbox_min) SDT_CUDA(h, cudaMemcpy(out->kd_bbox_min, h->kd_bmin, 12ull * nk, cudaMemcpyDeviceToHost));
    if (out->kd_bbox_max) SDT_CUDA(h, cudaMemcpy(out->kd_bbox_max, h->kd_bmax, 12ull * nk, cudaMemcpyDeviceToHost));
    if (out->kd_depth) SDT_CUDA(h, cudaMemcpy(out->kd_depth, h->kd_depth, 4ull * nk, cudaMemcpyDeviceToHost));
    if (out->kd_quad_root) SDT_CUDA(h, cudaMemcpy(out->kd_quad_root, h->kd_root, 4ull * nk, cudaMemcpyDeviceToHost));
    if (out->kd_vert_count) SDT_CUDA(h, cudaMemcpy(out->kd_vert_count, which == SDT_TREE_PREV ? h->kd_prev_count : h->kd_count, 4ull * nk, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < nk; ++i) {
        const bool leaf = (word[i] & SDT_KD_LEAF_BIT) != 0;
        if (out->kd_is_leaf) out->kd_is_leaf[i] = leaf ? 1 : 0;
        if (out->kd_child_left) out->kd_child_left[i] = leaf ? 0u : word[i];
        if (out->kd_child_right) out->kd_child_right[i] = leaf ? 0u : word[i] + 1u;
    }
    if (out->q_irradiance) SDT_CUDA(h, cudaMemcpy(out->q_irradiance, which == SDT_TREE_PREV ? s.energy : h->q_ecur, 4ull * nq, cudaMemcpyDeviceToHost));
    if (out->q_threshold) SDT_CUDA(h, cudaMemcpy(out->q_threshold, s.thr, 4ull * nq, cudaMemcpyDeviceToHost));
    if (out->q_root_node) for (uint32_t r = 0; r < R; ++r) out->q_root_node[r] = r;
    // depth and boxes follow from the canonical order: parents precede their children
    std::vector<uint32_t> depth(nq, 0);
    std::vector<float> bmin(2ull * nq, 0.0f), bmax(2ull * nq, 1.0f);
    for (uint32_t i = 0; i < nq; ++i) {
        const uint32_t cb = child[i];
        if (out->q_is_leaf) out->q_is_leaf[i] = cb ? 0 : 1;
        for (uint32_t k = 0; k < 4; ++k) if (out->q_child[k]) out->q_child[k][i] = cb ? cb + k : 0u;
        if (!cb) continue;
        SDT_CHECK(h, cb + 3u < nq, SDT_ERR_LAYOUT, "sdt_download: corrupt child index");
        for (uint32_t k = 0; k < 4; ++k) {
            float lox = bmin[2 * i], loy = bmin[2 * i + 1], hix = bmax[2 * i], hiy = bmax[2 * i + 1];
            sdt_quadrant((int)k, lox, loy, hix, hiy);                  // src/quadtree.py:146-188
            bmin[2 * (cb + k)] = lox; bmin[2 * (cb + k) + 1] = loy; bmax[2 * (cb + k)] = hix; bmax[2 * (cb + k) + 1] = hiy;
            depth[cb + k] = depth[i] + 1u;
        }
    }
    if (out->q_depth) memcpy(out->q_depth, depth.data(), 4ull * nq);
    if (out->q_bbox_min) memcpy(out->q_bbox_min, bmin.data(), 8ull * nq);
    if (out->q_bbox_max) memcpy(out->q_bbox_max, bmax.data(), 8ull * nq);
    return SDT_OK;
}

// ---------------------------------------------------------------------------- thresholds
extern "C" int sdt_set_max_leaf_size(sdt_handle h, float max_leaf_size) {
    SDT_ENTER(h);
    h->max_leaf_host = max_leaf_size;
    launch_single(exec_ctx(h, h->last_stream), SetLeafSize{h->set[h->cur].hdr, max_leaf_size});
    return sdt_post_launch(h, "sdt_set_max_leaf_size");
}

extern "C" int sdt_set_iteration_threshold(sdt_handle h, int32_t iteration) {
    // c * sqrt(2^iteration) in double like the python, then the fp32 the comparison sees
    const double t = 12000.0 * sqrt(pow(2.0, (double)iteration));
    return sdt_set_max_leaf_size(h, (float)t);
}

extern "C" int sdt_stat_buffers(sdt_handle h, float** q_energy, uint32_t* n_quad, float** kd_count, uint32_t* n_kd) {
    SDT_ENTER(h);
    DevHeader H;
    SDT_TRY(sdt_read_header(h, H));
    if (q_energy) *q_energy = h->q_ecur;
    if (n_quad) *n_quad = H.n_quad;
    if (kd_count) *kd_count = h->kd_count;
    if (n_kd) *n_kd = H.n_kd;
    h->stats_complete = false; h->kd_complete = false;      // the caller may reduce into the buffers: interiors are re-swept from the leaves
    h->splat_bound_valid = false;
    return SDT_OK;
}

extern "C" int sdt_set_tuning(sdt_handle h, const char* key, int64_t value) {
    if (!h || !key) return SDT_ERR_INVALID;
    SDT_ENTER(h);
    const std::string k(key);
    if (k == "query_block") { SDT_CHECK(h, value == 0 || (value >= 64 && value <= 1024 && value % 32 == 0), SDT_ERR_INVALID, "query_block must be 0 (default) or 64..1024, multiple of 32"); h->query_block = (int)value; }
    else if (k == "query_ctas_per_sm") { SDT_CHECK(h, value >= 1 && value <= 32, SDT_ERR_INVALID, "query_ctas_per_sm must be 1..32"); h->query_ctas_per_sm = (int)value; }
    else if (k == "kd_smem_nodes") { SDT_CHECK(h, value >= 0 && value <= 49152, SDT_ERR_INVALID, "kd_smem_nodes must be 0..49152"); h->kd_smem_nodes = (int)value; }
    else if (k == "splat_stage_words") h->splat_stage_words = value != 0;
    else if (k == "kd_smem_count_nodes") { SDT_CHECK(h, value >= 0 && value <= 49152, SDT_ERR_INVALID, "kd_smem_count_nodes must be 0..49152"); h->kd_smem_count_nodes = (int)value; }
    else if (k == "splat_block") { SDT_CHECK(h, value == 0 || (value >= 64 && value <= 1024 && value % 32 == 0), SDT_ERR_INVALID, "splat_block must be 0 (default) or 64..1024, multiple of 32"); h->splat_block = (int)value; }
    else if (k == "splat_ctas_per_sm") { SDT_CHECK(h, value >= 1 && value <= 32, SDT_ERR_INVALID, "splat_ctas_per_sm must be 1..32"); h->splat_ctas_per_sm = (int)value; }
    else if (k == "fuse_sample_pdf") h->fuse_sample_pdf = value != 0;
    else if (k == "splat_aggregate") h->splat_aggregate = value != 0;
    else if (k == "use_jump") h->use_jump = value != 0;
    else if (k == "use_jump2") h->use_jump2 = value != 0;
    else if (k == "use_kd_grid") h->use_kd_grid = value != 0;
    else if (k == "use_int_cell") h->use_int_cell = value != 0;
    else if (k == "quad_thr_reciprocal") h->quad_thr_reciprocal = value != 0;
    else if (k == "use_pdl") h->use_pdl = value != 0;
    else if (k == "use_graph") h->use_graph = value != 0;
    else if (k == "helper_ctas_per_sm") {
        SDT_CHECK(h, value >= 1 && value <= 8, SDT_ERR_INVALID, "helper_ctas_per_sm must be 1..8");
        h->helper_ctas_per_sm = (int)value;
#ifndef SDT_HOSTEMU
        for (auto& kv : h->refine_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);     // captured with the old grids
        h->refine_graphs.clear();
#endif
    }
    else if (k == "use_compaction") h->use_compaction = value != 0;
    else if (k == "host_chunk") { SDT_CHECK(h, value >= 256, SDT_ERR_INVALID, "host_chunk must be >= 256 lanes"); h->host_chunk = (int)value; }
    else return sdt_fail(h, SDT_ERR_INVALID, "sdt_set_tuning: unknown key " + k);
    return SDT_OK;
}

extern "C" int sdt_synchronize(sdt_handle h, sdt_stream stream) {
    SDT_ENTER(h);
    SDT_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
    return SDT_OK;
}

extern "C" uint64_t sdt_kernel_launches(sdt_handle h) { return h ? h->launches : 0; }

// ---------------------------------------------------------------------------- L2 probe
#ifndef SDT_HOSTEMU
__global__ void __launch_bounds__(256) k_l2_read(const uint4* __restrict__ p, uint64_t n16, uint32_t passes, uint32_t* sink) {
    uint32_t acc = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint32_t r = 0; r < passes; ++r)
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
            const uint4 v = __ldcg(p + i);          // L2 only: an L1 hit would not measure L2
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x12345678u) *sink = acc;
}
#endif

#ifndef SDT_HOSTEMU
// What the descents actually do to the memory system: every lane of a warp reads ONE 32-byte sector at an unrelated
// address of an L2-resident set (the quadtree records), with one 256-bit load.  BATCH independent loads per lane and
// iteration, so the probe measures throughput, not latency.  via_l1 = ld.global.nc (the kernels' path), else ld.global.cg.
template <int BATCH, bool VIA_L1>
__global__ void __launch_bounds__(512) k_gather32(const char* __restrict__ p, uint32_t sectors_mask, uint32_t iters, uint32_t* sink) {
    uint32_t acc = 0;
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B1u + 0x7F4A7C15u;
    for (uint32_t it = 0; it < iters; ++it) {
        uint32_t v[BATCH][8];
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
            s = s * 747796405u + 2891336453u;
            const uint32_t sec = ((s >> 9) ^ (s << 7)) & sectors_mask;
            const char* q = p + (size_t)sec * 32u;
            if (VIA_L1) asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[b][0]), "=r"(v[b][1]), "=r"(v[b][2]), "=r"(v[b][3]), "=r"(v[b][4]), "=r"(v[b][5]), "=r"(v[b][6]), "=r"(v[b][7]) : "l"(q));
            else asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[b][0]), "=r"(v[b][1]), "=r"(v[b][2]), "=r"(v[b][3]), "=r"(v[b][4]), "=r"(v[b][5]), "=r"(v[b][6]), "=r"(v[b][7]) : "l"(q));
        }
#pragma unroll
        for (int b = 0; b < BATCH; ++b) acc += v[b][0] ^ v[b][3] ^ v[b][7];
    }
    if (acc == 0x12345678u) *sink = acc;
}
#endif

// Random 32-byte-sector gather bandwidth (GB/s of sectors delivered to the lanes) over an L2-resident set of `bytes`
// (rounded down to a power of two): the roof of a divergent tree descent, next to the sequential sweep of sdt_measure_l2.
extern "C" int sdt_measure_gather(sdt_handle h, uint64_t bytes, uint32_t iters, int32_t via_l1, float* gbps, sdt_stream stream) {
    if (!h || !gbps) return SDT_ERR_INVALID;
    SDT_ENTER(h);
#ifndef SDT_HOSTEMU
    cudaStream_t st = (cudaStream_t)stream;
    uint64_t sectors = 1;
    while (sectors * 2 * 32 <= bytes) sectors *= 2;
    SDT_CHECK(h, sectors >= 1024 && sectors <= (1ull << 31) && iters > 0, SDT_ERR_INVALID, "sdt_measure_gather: bytes must be 32 KiB .. 64 GiB, iters > 0");
    char* buf = nullptr;
    SDT_CUDA(h, cudaMalloc((void**)&buf, sectors * 32));
    SDT_CUDA(h, cudaMemsetAsync(buf, 1, sectors * 32, st));
    constexpr int BATCH = 8;
    const int grid = h->num_sms * 4, block = 512;
    uint32_t* sink = h->s_blk + SDT_SCAN_STATE_WORDS;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](uint32_t n_it) {
        if (via_l1) k_gather32<BATCH, true><<<grid, block, 0, st>>>(buf, (uint32_t)(sectors - 1), n_it, sink);
        else k_gather32<BATCH, false><<<grid, block, 0, st>>>(buf, (uint32_t)(sectors - 1), n_it, sink);
    };
    run(iters / 4 + 1);                         // warm: pull the set into L2
    cudaEventRecord(e0, st);
    run(iters);
    cudaEventRecord(e1, st);
    h->launches += 2;
    h->last_stream = st;
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf);
    if (e != cudaSuccess) return sdt_fail(h, SDT_ERR_CUDA, std::string("sdt_measure_gather: ") + cudaGetErrorString(e));
    *gbps = (float)((double)grid * block * (double)iters * BATCH * 32.0 / ((double)ms * 1e-3) / 1e9);
    return SDT_OK;
#else
    (void)bytes; (void)iters; (void)via_l1; (void)stream;
    *gbps = 0.0f;
    return sdt_fail(h, SDT_ERR_STATE, "sdt_measure_gather: not available in the host emulation");
#endif
}

extern "C" int sdt_measure_l2(sdt_handle h, uint64_t bytes, uint32_t passes, float* gbps, sdt_stream stream) {
    if (!h || !gbps) return SDT_ERR_INVALID;
    SDT_ENTER(h);
#ifndef SDT_HOSTEMU
    cudaStream_t st = (cudaStream_t)stream;
    uint4* buf = nullptr;
    const uint64_t n16 = bytes / 16;
    SDT_CHECK(h, n16 > 0 && passes > 0, SDT_ERR_INVALID, "sdt_measure_l2: empty probe");
    SDT_CUDA(h, cudaMalloc((void**)&buf, n16 * 16));
    SDT_CUDA(h, cudaMemsetAsync(buf, 1, n16 * 16, st));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = h->num_sms * 8;
    k_l2_read<<<grid, 256, 0, st>>>(buf, n16, 2, h->s_blk + SDT_SCAN_STATE_WORDS);      // warm: pull the set into L2
    cudaEventRecord(e0, st);
    k_l2_read<<<grid, 256, 0, st>>>(buf, n16, passes, h->s_blk + SDT_SCAN_STATE_WORDS);
    cudaEventRecord(e1, st);
    h->launches += 2;
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf);
    if (e != cudaSuccess) return sdt_fail(h, SDT_ERR_CUDA, std::string("sdt_measure_l2: ") + cudaGetErrorString(e));
    *gbps = (float)((double)n16 * 16.0 * passes / ((double)ms * 1e-3) / 1e9);
    return SDT_OK;
#else
    (void)bytes; (void)passes; (void)stream;
    *gbps = 0.0f;
    return sdt_fail(h, SDT_ERR_STATE, "sdt_measure_l2: not available in the host emulation");
#endif
}
