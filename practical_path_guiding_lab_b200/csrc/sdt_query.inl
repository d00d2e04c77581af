// Wavefront queries on the frozen `prev` tree: locate / sample / pdf / guided bounce,
// and the elementwise MIS helpers.  One thread per path vertex; the top of the spatial
// tree is staged in shared memory by every CTA, the quadtree records stay L2-resident.

// ---------------------------------------------------------------------------- launch
// A lane functor provides
//   kModes             1 (active / idle) or 2 (guided bounce: sample / pdf / idle)
//   mode_of(i)         0 = idle lane, 1..kModes = what the lane does
//   idle(i)            outputs of an idle lane
//   run_mode<KD>(k,i,m) the work of a non-idle lane; KD = spatial-descent variant (sdt_kd_descend)
//   kGrid, kSmemCounts, flush_count   see k_wavefront
#ifndef SDT_MIN_CTAS
#define SDT_MIN_CTAS 3
#endif
#ifndef SDT_TILE_MUL
#define SDT_TILE_MUL 8u       // lanes per thread and warp tile in the compacting kernel (tile = 256 lanes per warp)
#endif
#ifndef SDT_TILE_MUL_SMALL
#define SDT_TILE_MUL_SMALL 2u // ... on wavefronts too small to give every resident thread that many lanes (launch_wavefront)
#endif
template <class Lane, int KD>
SDT_HD void sdt_lane(const Lane& f, const KdCtx& k, uint32_t i) {
    const uint32_t m = f.mode_of(i);
    if (m) f.template run_mode<KD>(k, i, m);
    else f.idle(i);
}

#ifndef SDT_HOSTEMU
// COMPACT: the wavefront has idle lanes or lanes of different kinds.  Each warp sorts the lanes of its
// 256-lane tile by mode into shared-memory lists and then works through every list with all 32 threads:
// a warp never runs two code paths, and idle lanes (dead paths, masked-out vertices, inactive record
// slots) cost a classification, not a share of a descent.
template <class Lane, bool COMPACT, uint32_t TILE_MUL = SDT_TILE_MUL>
__global__ void __launch_bounds__(Lane::kMaxThreads, SDT_LB_CTAS) k_wavefront(Lane f, uint32_t n, uint32_t smem_cap, uint32_t cnt_cap, uint32_t use_grid, uint32_t aggregate) {
    extern __shared__ uint32_t kd_s[];
    const DevHeader* hdr = f.t.hdr;
    const uint32_t n_kd = hdr->n_kd;
    const uint32_t n_smem = n_kd < smem_cap ? n_kd : smem_cap;
    for (uint32_t j = threadIdx.x; j < n_smem; j += blockDim.x) kd_s[j] = __ldg(f.t.kd_word + j);
    KdCtx k = sdt_kd_ctx(kd_s, n_smem, f.t.kd_word, hdr);
    uint32_t* smem_next = kd_s + smem_cap;
    // splat kernels with the whole spatial tree staged: leaf counters live in shared memory for the
    // lifetime of the CTA and are flushed once (16M same-slice L2 atomics become a few per leaf and CTA)
    // (integer counters: a native shared-memory add, where a float add is a compare-and-swap loop; a CTA sees fewer
    // than 2^24 records, so the count is exact and its fp32 value is what a chain of +1.0f would have produced)
    uint32_t* cnt_s = nullptr;
    k.aggregate = aggregate;
    if (Lane::kSmemCounts) {
        if (n_kd <= cnt_cap) {               // (independent of whether the words are staged)
            cnt_s = smem_next;
            for (uint32_t j = threadIdx.x; j < n_kd; j += blockDim.x) cnt_s[j] = 0u;
            k.cnt_s = cnt_s;
        }
        smem_next += cnt_cap;
    }
    __syncthreads();
    // 16x16x8 grid over the first 11 spatial levels (built with the records, see sdt_kd_descend).  With
    // the whole tree staged it pays for the pdf / splat / locate kernels (measured -7 % / -4 %) but not
    // for the sampling kernels (+3 %: their long quadtree loop wants the registers); when the tree does
    // not fit in shared memory every kernel uses it, and only the levels below 11 go through L1/L2.
    const bool staged = n_smem == n_kd;
    if (use_grid && (Lane::kGrid || !staged)) {
        for (uint32_t c = threadIdx.x; c < SDT_GRID_CELLS; c += blockDim.x) smem_next[c] = __ldg(f.t.kd_grid + c);
        k.grid = smem_next;
    }
    smem_next += SDT_GRID_CELLS;
    __syncthreads();
    const int kd_mode = k.grid ? (staged ? 2 : 3) : (staged ? 1 : 0);
#define SDT_KD_DISPATCH(...)                                              \
    if (kd_mode == 2) { constexpr int KD = 2; __VA_ARGS__; }                \
    else if (kd_mode == 1) { constexpr int KD = 1; __VA_ARGS__; }           \
    else if (kd_mode == 3) { constexpr int KD = 3; __VA_ARGS__; }           \
    else { constexpr int KD = 0; __VA_ARGS__; }
    if (COMPACT) {
        // warp-local: every warp sorts its own tile of 32*SDT_TILE_MUL lanes into its slice of the
        // shared-memory lists (no CTA barrier: a warp never waits for the deepest descent of another warp)
        constexpr uint32_t TM = TILE_MUL;
        constexpr uint32_t tile_w = 32u * TM;
        const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
        uint16_t* list = reinterpret_cast<uint16_t*>(smem_next) + wib * tile_w * (uint32_t)Lane::kModes;
        const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
        const uint32_t warp_id = blockIdx.x * (blockDim.x >> 5) + wib;
        for (uint64_t tile64 = (uint64_t)warp_id * tile_w; tile64 < n; tile64 += (uint64_t)warps_total * tile_w) {
            const uint32_t tile = (uint32_t)tile64;
            uint32_t cnt[3] = {0u, 0u, 0u};
            uint32_t mq = 0;                 // modes of this thread's lanes, 2 bits each
#pragma unroll
            for (uint32_t q = 0; q < TM; ++q) {
                const uint32_t li = q * 32u + lane, i = tile + li;
                uint32_t m = 0;
                if (i < n) { m = f.mode_of(i); if (!m) f.idle(i); }
                mq |= m << (2u * q);
#pragma unroll
                for (uint32_t L = 0; L < (uint32_t)Lane::kModes; ++L) {
                    const uint32_t b = __ballot_sync(0xFFFFFFFFu, m == L + 1u);
                    if (m == L + 1u) list[L * tile_w + cnt[L] + __popc(b & ((1u << lane) - 1u))] = (uint16_t)li;
                    cnt[L] += (uint32_t)__popc(b);
                }
            }
            __syncwarp();
            if (Lane::kModes == 1 && cnt[0] * 2u > tile_w) {
                // a dense tile gains nothing from the indirection: every thread runs its own lanes in place (coalesced)
#pragma unroll
                for (uint32_t q = 0; q < TM; ++q) {
                    if (!((mq >> (2u * q)) & 3u)) continue;
                    const uint32_t i = tile + q * 32u + lane;
                    SDT_KD_DISPATCH(f.template run_mode<KD>(k, i, 1u))
                }
            } else {
#pragma unroll
                for (uint32_t L = 0; L < (uint32_t)Lane::kModes; ++L) {
                    for (uint32_t j = lane; j < cnt[L]; j += 32u) {
                        const uint32_t i = tile + list[L * tile_w + j];
                        SDT_KD_DISPATCH(f.template run_mode<KD>(k, i, L + 1u))
                    }
                }
            }
            __syncwarp();
        }
    } else {
        SDT_KD_DISPATCH(for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) sdt_lane<Lane, KD>(f, k, i))
    }
#undef SDT_KD_DISPATCH
    if (cnt_s) {
        __syncthreads();
        for (uint32_t j = threadIdx.x; j < n_kd; j += blockDim.x) {
            const uint32_t c = cnt_s[j];
            if (c != 0u) f.flush_count(j, (float)c);
        }
    }
}

// Persistent grid: 148 SMs x resident CTAs.  The smem staging is sized to the spatial tree the
// host last saw (exact after upload / get_sizes; after a refine a non-blocking header read-back
// refreshes it), capped by the "kd_smem_nodes" tuning; a larger tree falls back to global loads
// for the nodes beyond the staged prefix.
template <class Lane, bool COMPACT, uint32_t TILE_MUL = SDT_TILE_MUL>
static int launch_wavefront_c(sdt_handle h, cudaStream_t st, uint32_t n, const Lane& f, int block, int ctas_per_sm) {
    if (n == 0) return SDT_OK;
    if (block <= 0 || block > Lane::kMaxThreads) block = Lane::kMaxThreads;      // 0 = the kernel's own CTA size
    uint32_t want = h->hdr_pending ? 2u * h->kd_nodes_known + 2u : h->kd_nodes_known;   // a refine is in flight: be generous
    uint32_t smem_nodes = (want + 255u) & ~255u;
    if (smem_nodes > (uint32_t)h->kd_smem_nodes) smem_nodes = (uint32_t)h->kd_smem_nodes;
    if (Lane::kSmemCounts && !h->splat_stage_words) smem_nodes = 0;     // splat: counters + grid only, words through L1/L2
    uint32_t cnt_nodes = 0;                  // shared-memory leaf counters of the splat kernels
    if (Lane::kSmemCounts) {
        cnt_nodes = (want + 255u) & ~255u;
        if (cnt_nodes > (uint32_t)h->kd_smem_count_nodes) cnt_nodes = (uint32_t)h->kd_smem_count_nodes;
    }
    const size_t smem = (size_t)smem_nodes * 4u + (size_t)cnt_nodes * 4u + SDT_GRID_CELLS * 4u +
                        (COMPACT ? (size_t)block * TILE_MUL * 2u * (size_t)Lane::kModes : 0u);   // uint16 lists: 32*TM entries per warp and mode
    sdt_tree_s::LaunchCache& lc = h->launch_cache[(const void*)k_wavefront<Lane, COMPACT, TILE_MUL>];
    constexpr size_t kSmemMax = 227u * 1024u;            // the sm_100 limit per CTA
    if (smem > kSmemMax) return sdt_fail(h, SDT_ERR_INVALID, "k_wavefront: staging tunings ask for more than 227 KB of shared memory per CTA");
    if (smem > 48u * 1024u && smem > lc.attr_smem) {
        if (cudaFuncSetAttribute(k_wavefront<Lane, COMPACT, TILE_MUL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax) != cudaSuccess)
            return sdt_fail(h, SDT_ERR_CUDA, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");
        lc.attr_smem = kSmemMax;
    }
    if (lc.occ_smem != smem || lc.occ_block != block) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&lc.occ, k_wavefront<Lane, COMPACT, TILE_MUL>, block, smem) != cudaSuccess || lc.occ < 1) lc.occ = 1;
        lc.occ_smem = smem; lc.occ_block = block;
    }
    const int occ_cache = lc.occ;
    int per_sm = ctas_per_sm < occ_cache ? ctas_per_sm : occ_cache;
    if (per_sm < 1) per_sm = 1;
    const uint32_t per_cta = (uint32_t)block * (COMPACT ? TILE_MUL : 1u);
    uint32_t grid = (n + per_cta - 1u) / per_cta;
    const uint32_t cap = (uint32_t)(h->num_sms * per_sm);
    if (grid > cap) grid = cap;
    k_wavefront<Lane, COMPACT, TILE_MUL><<<grid, block, smem, st>>>(f, n, smem_nodes, cnt_nodes, (uint32_t)h->use_kd_grid, (uint32_t)h->splat_aggregate);
    ++h->launches;
    h->last_stream = st;
    return sdt_post_launch(h, "k_wavefront");
}
// compact = the wavefront may contain idle lanes / several kinds of lanes
template <class Lane>
static int launch_wavefront(sdt_handle h, cudaStream_t st, uint32_t n, const Lane& f, int block, int ctas_per_sm, bool compact) {
    if (compact && h->use_compaction) {
        // A 256-lane tile sorts the lanes of several rounds into dense warps, but a thread then owns 8 lanes one after the
        // other.  A wavefront that cannot give every resident thread two lanes anyway (a 256 x 256 or 512 x 512 pass: 65 k /
        // 262 k lanes on 227 k thread slots) pays latency, not issue slots: 64-lane tiles put its lanes side by side -- the
        // first bounces of such a pass take 2 rounds of descents per warp instead of up to 8.
        const uint64_t slots = (uint64_t)h->num_sms * (uint64_t)(ctas_per_sm > 0 ? ctas_per_sm : 1) *
                               (uint64_t)((block > 0 && block <= Lane::kMaxThreads) ? block : Lane::kMaxThreads);
        if ((uint64_t)n <= 2u * slots) return launch_wavefront_c<Lane, true, SDT_TILE_MUL_SMALL>(h, st, n, f, block, ctas_per_sm);
        return launch_wavefront_c<Lane, true>(h, st, n, f, block, ctas_per_sm);
    }
    return launch_wavefront_c<Lane, false>(h, st, n, f, block, ctas_per_sm);
}
#else
// host emulation: serial lanes, but the same choice of spatial-descent variant as the kernel makes
// ("staged" prefix = kd_smem_nodes, grid on / off), so that all four variants run on the CPU too
template <class Lane>
static int launch_wavefront(sdt_handle h, cudaStream_t st, uint32_t n, const Lane& f, int, int, bool) {
    const uint32_t n_kd = f.t.hdr->n_kd;
    const uint32_t n_smem = n_kd < (uint32_t)h->kd_smem_nodes ? n_kd : (uint32_t)h->kd_smem_nodes;
    KdCtx k = sdt_kd_ctx(f.t.kd_word, n_smem, f.t.kd_word, f.t.hdr);
    const bool staged = n_smem == n_kd;
    if (h->use_kd_grid && (Lane::kGrid || !staged)) k.grid = f.t.kd_grid;
    const int kd_mode = k.grid ? (staged ? 2 : 3) : (staged ? 1 : 0);
    for (uint32_t i = 0; i < n; ++i) {
        if (kd_mode == 2) sdt_lane<Lane, 2>(f, k, i);
        else if (kd_mode == 1) sdt_lane<Lane, 1>(f, k, i);
        else if (kd_mode == 3) sdt_lane<Lane, 3>(f, k, i);
        else sdt_lane<Lane, 0>(f, k, i);
    }
    ++h->launches;
    h->last_stream = st;
    return SDT_OK;
}
#endif

// ---------------------------------------------------------------------------- lanes
struct LocateLane {
    static constexpr bool kSmemCounts = false, kGrid = true;
    static constexpr int kModes = 1, kMaxThreads = SDT_QUERY_THREADS;
    SDT_HD void flush_count(uint32_t, float) const {}
    TreeView t;
    sdt_vec3 pos; const uint8_t* active; uint32_t* leaf; uint32_t* root;
    SDT_HD uint32_t mode_of(uint32_t i) const { return active ? (SDT_LDG(active + i) != 0 ? 1u : 0u) : 1u; }
    SDT_HD void idle(uint32_t i) const {                 // inactive: node 0, masked gather -> 0
        if (leaf) leaf[i] = 0u;
        if (root) root[i] = 0u;
    }
    template <int KD>
    SDT_HD void run_mode(const KdCtx& k, uint32_t i, uint32_t) const {
        const KdResult r = sdt_kd_descend<KD>(k, sdt_ld(pos.x, pos.stride, i), sdt_ld(pos.y, pos.stride, i), sdt_ld(pos.z, pos.stride, i));
        if (leaf) leaf[i] = r.leaf;
        if (root) root[i] = SDT_LDG(t.kd_root + r.leaf);
    }
};

template <bool EXPLICIT_U>
struct SampleLane {
    static constexpr bool kSmemCounts = false, kGrid = SDT_SAMPLE_GRID;
    static constexpr int kModes = 1, kMaxThreads = SDT_SAMPLE_THREADS;
    SDT_HD void flush_count(uint32_t, float) const {}
    TreeView t;
    sdt_vec3 pos; const uint8_t* active;
    const float* u; uint32_t u_stride, seed, lane_offset;
    sdt_vec3_out dir; float* pdf; uint32_t* dbg; int fuse;
    SDT_HD uint32_t mode_of(uint32_t i) const { return active ? (SDT_LDG(active + i) != 0 ? 1u : 0u) : 1u; }
    SDT_HD void idle(uint32_t i) const {                 // inactive lanes: pos (0,0) -> (0,0,-1), pdf 1
        const int64_t o = (int64_t)i * dir.stride;
        dir.x[o] = 0.0f; dir.y[o] = 0.0f; dir.z[o] = -1.0f;
        pdf[i] = 1.0f;
        if (dbg) { dbg[4u * i] = 0u; dbg[4u * i + 1u] = 0u; dbg[4u * i + 2u] = 0u; dbg[4u * i + 3u] = 0u; }
    }
    template <int KD>
    SDT_HD void run_mode(const KdCtx& k, uint32_t i, uint32_t) const {
        const KdResult r = sdt_kd_descend<KD>(k, sdt_ld(pos.x, pos.stride, i), sdt_ld(pos.y, pos.stride, i), sdt_ld(pos.z, pos.stride, i));
        const uint32_t root = dbg ? SDT_LDG(t.kd_root + r.leaf) : 0u;
        GuidedSample g;
        if (EXPLICIT_U) g = sdt_sample_tree(t, r.rootrec, root, ExplicitRng(u, u_stride, i), fuse != 0);
        else g = sdt_sample_tree(t, r.rootrec, root, CounterRng(seed, lane_offset + i), fuse != 0);
        const int64_t o = (int64_t)i * dir.stride;
        dir.x[o] = g.dx; dir.y[o] = g.dy; dir.z[o] = g.dz;
        pdf[i] = g.pdf;
        if (dbg) { dbg[4u * i] = r.leaf; dbg[4u * i + 1u] = root; dbg[4u * i + 2u] = g.sample_node; dbg[4u * i + 3u] = g.pdf_node; }
    }
};

// KDTree.sample + KDTree.pdf of a second, given direction on the same vertices: ONE spatial descent, then the sampling
// descent and the pdf descent in the same quadtree (a path vertex asks both: the guided sample, src/path_guiding_integrator.py:301,
// and the tree's pdf of the emitter direction for the NEE MIS weight, :244)
template <bool EXPLICIT_U>
struct SamplePdfLane {
    static constexpr bool kSmemCounts = false, kGrid = SDT_SAMPLE_GRID;
    static constexpr int kModes = 1, kMaxThreads = SDT_SAMPLE_THREADS;
    SDT_HD void flush_count(uint32_t, float) const {}
    TreeView t;
    sdt_vec3 pos; const uint8_t* active;
    const float* u; uint32_t u_stride, seed, lane_offset;
    sdt_vec3_out dir; float* pdf; sdt_vec3 qdir; float* qpdf; int fuse;
    SDT_HD uint32_t mode_of(uint32_t i) const { return active ? (SDT_LDG(active + i) != 0 ? 1u : 0u) : 1u; }
    SDT_HD void idle(uint32_t i) const {                 // as sdt_sample / sdt_pdf leave their inactive lanes
        const int64_t o = (int64_t)i * dir.stride;
        dir.x[o] = 0.0f; dir.y[o] = 0.0f; dir.z[o] = -1.0f;
        pdf[i] = 1.0f;
        qpdf[i] = 1.0f;
    }
    template <int KD>
    SDT_HD void run_mode(const KdCtx& k, uint32_t i, uint32_t) const {
        const KdResult r = sdt_kd_descend<KD>(k, sdt_ld(pos.x, pos.stride, i), sdt_ld(pos.y, pos.stride, i), sdt_ld(pos.z, pos.stride, i));
        // the given direction: its pdf descent is issued first (short: jump table), the long sampling descent runs after it
        float x, y;
        sdt_dir_to_canonical(sdt_ld(qdir.x, qdir.stride, i), sdt_ld(qdir.y, qdir.stride, i), sdt_ld(qdir.z, qdir.stride, i), x, y);
        uint32_t nd;
        qpdf[i] = sdt_quad_pdf(t, r.rootrec, 0u, x, y, nd, false);
        GuidedSample g;
        if (EXPLICIT_U) g = sdt_sample_tree(t, r.rootrec, 0u, ExplicitRng(u, u_stride, i), fuse != 0);
        else g = sdt_sample_tree(t, r.rootrec, 0u, CounterRng(seed, lane_offset + i), fuse != 0);
        const int64_t o = (int64_t)i * dir.stride;
        dir.x[o] = g.dx; dir.y[o] = g.dy; dir.z[o] = g.dz;
        pdf[i] = g.pdf;
    }
};

struct PdfLane {
    static constexpr bool kSmemCounts = false, kGrid = true;
    static constexpr int kModes = 1, kMaxThreads = SDT_QUERY_THREADS;
    SDT_HD void flush_count(uint32_t, float) const {}
    TreeView t;
    sdt_vec3 pos; sdt_vec3 dir; const uint8_t* active; float* pdf; uint32_t* dbg;
    SDT_HD uint32_t mode_of(uint32_t i) const { return active ? (SDT_LDG(active + i) != 0 ? 1u : 0u) : 1u; }
    SDT_HD void idle(uint32_t i) const {
        pdf[i] = 1.0f;
        if (dbg) { dbg[3u * i] = 0u; dbg[3u * i + 1u] = 0u; dbg[3u * i + 2u] = 0u; }
    }
    template <int KD>
    SDT_HD void run_mode(const KdCtx& k, uint32_t i, uint32_t) const {
        // the direction is fetched before the spatial descent, so that its (DRAM) latency runs under it
        const float wx = sdt_ld(dir.x, dir.stride, i), wy = sdt_ld(dir.y, dir.stride, i), wz = sdt_ld(dir.z, dir.stride, i);
        const KdResult r = sdt_kd_descend<KD>(k, sdt_ld(pos.x, pos.stride, i), sdt_ld(pos.y, pos.stride, i), sdt_ld(pos.z, pos.stride, i));
        float x, y;
        sdt_dir_to_canonical(wx, wy, wz, x, y);
        uint32_t nd;
        const uint32_t root = dbg ? SDT_LDG(t.kd_root + r.leaf) : 0u;
        pdf[i] = sdt_quad_pdf(t, r.rootrec, root, x, y, nd, dbg != nullptr);
        if (dbg) { dbg[3u * i] = r.leaf; dbg[3u * i + 1u] = root; dbg[3u * i + 2u] = nd; }
    }
};

// one bounce: mode 1 = sample the tree (src/path_guiding_integrator.py:301),
// mode 2 = tree pdf of the BSDF-sampled direction (:307) + fused mixture (:310-311);
// with em_dir given, every lane whose em_active is set also gets the tree's pdf of the emitter direction (:244) from the
// SAME spatial descent (lanes that do nothing else run as mode 3)
template <bool EXPLICIT_U, bool WITH_EM>
struct GuidedLane {
    static constexpr bool kSmemCounts = false, kGrid = SDT_SAMPLE_GRID;
    static constexpr int kModes = WITH_EM ? 3 : 2, kMaxThreads = SDT_SAMPLE_THREADS;
    SDT_HD void flush_count(uint32_t, float) const {}
    TreeView t;
    sdt_guided_args a; int fuse;
    float f, omf;           // fp32(bsdfSamplingFraction), fp32(1 - bsdfSamplingFraction) with the difference formed in double
    SDT_HD bool em_of(uint32_t i) const { return WITH_EM && (a.em_active ? SDT_LDG(a.em_active + i) != 0 : true); }
    SDT_HD uint32_t mode_of(uint32_t i) const {
        uint32_t m = SDT_LDG(a.mode + i);
        if (m > 2u) m = 0u;
        return (m == 0u && em_of(i)) ? 3u : m;
    }
    SDT_HD void idle(uint32_t) const {}
    template <int KD>
    SDT_HD void run_mode(const KdCtx& k, uint32_t i, uint32_t m) const {
        const KdResult r = sdt_kd_descend<KD>(k, sdt_ld(a.pos.x, a.pos.stride, i), sdt_ld(a.pos.y, a.pos.stride, i), sdt_ld(a.pos.z, a.pos.stride, i));
        if (WITH_EM && (m == 3u || em_of(i))) {
            float x, y;
            sdt_dir_to_canonical(sdt_ld(a.em_dir.x, a.em_dir.stride, i), sdt_ld(a.em_dir.y, a.em_dir.stride, i), sdt_ld(a.em_dir.z, a.em_dir.stride, i), x, y);
            uint32_t nd;
            a.sdtree_pdf_em[i] = sdt_quad_pdf(t, r.rootrec, 0u, x, y, nd, false);
            if (m == 3u) return;
        }
        if (m == 1u) {
            GuidedSample g;
            if (EXPLICIT_U) g = sdt_sample_tree(t, r.rootrec, 0u, ExplicitRng(a.u, a.u_stride, i), fuse != 0);
            else g = sdt_sample_tree(t, r.rootrec, 0u, CounterRng(a.seed, a.lane_offset + i), fuse != 0);
            const int64_t o = (int64_t)i * a.dir.stride;
            a.dir.x[o] = g.dx; a.dir.y[o] = g.dy; a.dir.z[o] = g.dz;
            a.sdtree_pdf[i] = g.pdf;
        } else {
            float x, y;
            sdt_dir_to_canonical(sdt_ld(a.wo.x, a.wo.stride, i), sdt_ld(a.wo.y, a.wo.stride, i), sdt_ld(a.wo.z, a.wo.stride, i), x, y);
            uint32_t nd;
            const float p = sdt_quad_pdf(t, r.rootrec, 0u, x, y, nd, false);
            a.sdtree_pdf[i] = p;
            if (a.bsdf_pdf && a.wo_pdf) {
                const float wp = (f * SDT_LDG(a.bsdf_pdf + i)) + omf * p;              // :310
                a.wo_pdf[i] = wp;
                if (a.bsdf_value.x && a.weight.x) {
                    const int64_t o = (int64_t)i * a.weight.stride;
                    a.weight.x[o] = sdt_ld(a.bsdf_value.x, a.bsdf_value.stride, i) / wp;   // :311
                    a.weight.y[o] = sdt_ld(a.bsdf_value.y, a.bsdf_value.stride, i) / wp;
                    a.weight.z[o] = sdt_ld(a.bsdf_value.z, a.bsdf_value.stride, i) / wp;
                }
            }
        }
    }
};

struct MisNeeItem {
    const float* bsdf_pdf_em; const float* sdtree_pdf_em; const float* pdf_with_delta; const float* pdf_without_delta;
    const float* ds_pdf; const uint8_t* ds_delta; float f, omf; int32_t iteration; float* surface_pdf_em; float* mis_em;
    SDT_HD void operator()(uint32_t i) const {
        const float bp = SDT_LDG(bsdf_pdf_em + i);
        float surface = bp;
        if (iteration > 1) {                                                            // :250
            const float eps = 0.00001f;
            const float pdf_diffuse = (SDT_LDG(pdf_with_delta + i) + eps) / (SDT_LDG(pdf_without_delta + i) + eps);   // :241
            surface = f * bp + (omf * SDT_LDG(sdtree_pdf_em + i)) * pdf_diffuse;           // :247
        }
        if (surface_pdf_em) surface_pdf_em[i] = surface;
        if (mis_em) {
            const bool delta = ds_delta ? SDT_LDG(ds_delta + i) != 0 : false;
            mis_em[i] = delta ? 1.0f : sdt_mis_weight(SDT_LDG(ds_pdf + i), surface);    // :253
        }
    }
};

struct MisMixtureItem {
    const float* bsdf_pdf; const float* sdtree_pdf; sdt_vec3 bsdf_value; const uint8_t* do_mis; float f, omf;
    float* wo_pdf; sdt_vec3_out weight;
    SDT_HD void operator()(uint32_t i) const {
        const float bp = SDT_LDG(bsdf_pdf + i);
        const bool mis = do_mis ? SDT_LDG(do_mis + i) != 0 : true;
        const float wp = mis ? (f * bp) + omf * SDT_LDG(sdtree_pdf + i) : bp;
        if (wo_pdf) wo_pdf[i] = wp;
        if (weight.x && bsdf_value.x) {
            const int64_t o = (int64_t)i * weight.stride;
            weight.x[o] = sdt_ld(bsdf_value.x, bsdf_value.stride, i) / wp;
            weight.y[o] = sdt_ld(bsdf_value.y, bsdf_value.stride, i) / wp;
            weight.z[o] = sdt_ld(bsdf_value.z, bsdf_value.stride, i) / wp;
        }
    }
};

// ---------------------------------------------------------------------------- entry points
extern "C" int sdt_locate(sdt_handle h, const sdt_vec3* pos, const uint8_t* active, uint32_t n,
                          uint32_t* leaf, uint32_t* root, uint32_t flags, sdt_stream stream) {
    SDT_ENTER(h);
    if (n == 0) return SDT_OK;                          // empty wavefront: nothing to do (pointers may be NULL)
    SDT_CHECK(h, pos && pos->x, SDT_ERR_INVALID, "sdt_locate: pos is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    Stager sg(h, st, flags);
    SDT_TRY(sg.reserve((size_t)n * (12 + 1 + 8) + 4096));
    LocateLane f{tree_view(h), sg.in3(*pos, n), sg.in_t(active, n), sg.out_t(leaf, n), sg.out_t(root, n)};
    if (sg.status != SDT_OK) return sg.status;
    SDT_TRY(launch_wavefront(h, st, n, f, h->query_block, h->query_ctas_per_sm, false));
    return sg.finish(flags);
}

extern "C" int sdt_sample(sdt_handle h, const sdt_vec3* pos, const uint8_t* active, uint32_t n,
                          const float* u, uint32_t u_stride, uint32_t seed, uint32_t lane_offset,
                          const sdt_vec3_out* dir, float* pdf, uint32_t* dbg, uint32_t flags, sdt_stream stream) {
    SDT_ENTER(h);
    if (n == 0) return SDT_OK;                          // empty wavefront: nothing to do (pointers may be NULL)
    SDT_CHECK(h, pos && pos->x && dir && dir->x && pdf, SDT_ERR_INVALID, "sdt_sample: pos / dir / pdf is NULL");
    SDT_CHECK(h, !u || u_stride >= 3, SDT_ERR_INVALID, "sdt_sample: u_stride must be >= 3");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t per_lane = 12 + 1 + 12 + 4 + (dbg ? 16 : 0) + (u ? 4ull * u_stride : 0) + 8;
    return sdt_run_chunked(h, st, flags, n, per_lane, true, [&](Stager& sg, uint32_t off, uint32_t cnt) -> int {
        SampleLane<false> f{tree_view(h), sg.in3(sdt_off3(*pos, off), cnt), sg.in_t(sdt_offp(active, off), cnt),
                            sg.in_t(sdt_offp(u, (size_t)off * u_stride), (size_t)cnt * u_stride), u_stride, seed, lane_offset + off,
                            sg.out3(sdt_off3o(*dir, off), cnt), sg.out_t(sdt_offp(pdf, off), cnt),
                            sg.out_t(sdt_offp(dbg, (size_t)off * 4), (size_t)cnt * 4), h->fuse_sample_pdf};
        if (sg.status != SDT_OK) return sg.status;
        sg.before_launch();
        if (u) {
            SampleLane<true> fe{f.t, f.pos, f.active, f.u, f.u_stride, f.seed, f.lane_offset, f.dir, f.pdf, f.dbg, f.fuse};
            return launch_wavefront(h, st, cnt, fe, h->query_block, h->query_ctas_per_sm, active != nullptr);
        }
        return launch_wavefront(h, st, cnt, f, h->query_block, h->query_ctas_per_sm, active != nullptr);
    });
}

extern "C" int sdt_pdf(sdt_handle h, const sdt_vec3* pos, const sdt_vec3* dir, const uint8_t* active,
                       uint32_t n, float* pdf, uint32_t* dbg, uint32_t flags, sdt_stream stream) {
    SDT_ENTER(h);
    if (n == 0) return SDT_OK;                          // empty wavefront: nothing to do (pointers may be NULL)
    SDT_CHECK(h, pos && pos->x && dir && dir->x && pdf, SDT_ERR_INVALID, "sdt_pdf: pos / dir / pdf is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t per_lane = 24 + 1 + 4 + (dbg ? 12 : 0) + 8;
    return sdt_run_chunked(h, st, flags, n, per_lane, true, [&](Stager& sg, uint32_t off, uint32_t cnt) -> int {
        PdfLane f{tree_view(h), sg.in3(sdt_off3(*pos, off), cnt), sg.in3(sdt_off3(*dir, off), cnt), sg.in_t(sdt_offp(active, off), cnt),
                  sg.out_t(sdt_offp(pdf, off), cnt), sg.out_t(sdt_offp(dbg, (size_t)off * 3), (size_t)cnt * 3)};
        if (sg.status != SDT_OK) return sg.status;
        sg.before_launch();
        return launch_wavefront(h, st, cnt, f, h->query_block, h->query_ctas_per_sm, active != nullptr);
    });
}

extern "C" int sdt_sample_pdf(sdt_handle h, const sdt_vec3* pos, const uint8_t* active, uint32_t n,
                              const float* u, uint32_t u_stride, uint32_t seed, uint32_t lane_offset,
                              const sdt_vec3_out* dir, float* pdf, const sdt_vec3* qdir, float* qpdf,
                              uint32_t flags, sdt_stream stream) {
    SDT_ENTER(h);
    if (n == 0) return SDT_OK;                          // empty wavefront: nothing to do (pointers may be NULL)
    SDT_CHECK(h, pos && pos->x && dir && dir->x && pdf && qdir && qdir->x && qpdf, SDT_ERR_INVALID, "sdt_sample_pdf: pos / dir / pdf / qdir / qpdf is NULL");
    SDT_CHECK(h, !u || u_stride >= 3, SDT_ERR_INVALID, "sdt_sample_pdf: u_stride must be >= 3");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t per_lane = 12 + 12 + 1 + 12 + 4 + 4 + (u ? 4ull * u_stride : 0) + 8;
    return sdt_run_chunked(h, st, flags, n, per_lane, true, [&](Stager& sg, uint32_t off, uint32_t cnt) -> int {
        SamplePdfLane<false> f{tree_view(h), sg.in3(sdt_off3(*pos, off), cnt), sg.in_t(sdt_offp(active, off), cnt),
                               sg.in_t(sdt_offp(u, (size_t)off * u_stride), (size_t)cnt * u_stride), u_stride, seed, lane_offset + off,
                               sg.out3(sdt_off3o(*dir, off), cnt), sg.out_t(sdt_offp(pdf, off), cnt),
                               sg.in3(sdt_off3(*qdir, off), cnt), sg.out_t(sdt_offp(qpdf, off), cnt), h->fuse_sample_pdf};
        if (sg.status != SDT_OK) return sg.status;
        sg.before_launch();
        if (u) {
            SamplePdfLane<true> fe{f.t, f.pos, f.active, f.u, f.u_stride, f.seed, f.lane_offset, f.dir, f.pdf, f.qdir, f.qpdf, f.fuse};
            return launch_wavefront(h, st, cnt, fe, h->query_block, h->query_ctas_per_sm, active != nullptr);
        }
        return launch_wavefront(h, st, cnt, f, h->query_block, h->query_ctas_per_sm, active != nullptr);
    });
}

extern "C" int sdt_guided(sdt_handle h, const sdt_guided_args* a, uint32_t n, uint32_t flags, sdt_stream stream) {
    SDT_ENTER(h);
    if (n == 0) return SDT_OK;                          // empty wavefront: nothing to do (pointers may be NULL)
    SDT_CHECK(h, a && a->pos.x && a->mode && a->sdtree_pdf && a->dir.x, SDT_ERR_INVALID, "sdt_guided: pos / mode / dir / sdtree_pdf is NULL");
    SDT_CHECK(h, !a->u || a->u_stride >= 3, SDT_ERR_INVALID, "sdt_guided: u_stride must be >= 3");
    SDT_CHECK(h, !a->em_dir.x || a->sdtree_pdf_em, SDT_ERR_INVALID, "sdt_guided: em_dir needs sdtree_pdf_em");
    cudaStream_t st = (cudaStream_t)stream;
    Stager sg(h, st, flags);
    SDT_TRY(sg.reserve((size_t)n * (12 + 12 + 1 + 4 + 12 + 12 + 4 + 4 + 12 + 12 + 1 + 4 + (a->u ? 4ull * a->u_stride : 0)) + 32768));
    sdt_guided_args d = *a;
    d.pos = sg.in3(a->pos, n);
    d.wo = sg.in3(a->wo, n);
    d.mode = sg.in_t(a->mode, n);
    d.u = sg.in_t(a->u, (size_t)n * a->u_stride);
    d.bsdf_pdf = sg.in_t(a->bsdf_pdf, n);
    d.bsdf_value = sg.in3(a->bsdf_value, n);
    d.em_dir = sg.in3(a->em_dir, n);
    d.em_active = sg.in_t(a->em_active, n);
    if (sg.host) {
        // outputs are partial (mode-dependent): stage the caller's current contents first
        SDT_CHECK(h, a->dir.stride == 3 || a->dir.stride == 1, SDT_ERR_INVALID, "sdt_guided: host dir must have stride 3 or 1");
        sdt_vec3 cur_dir{a->dir.x, a->dir.y, a->dir.z, a->dir.stride};
        sdt_vec3 dd = sg.in3(cur_dir, n);
        d.dir = sdt_vec3_out{(float*)dd.x, (float*)dd.y, (float*)dd.z, dd.stride};
        if (dd.stride == 3) sg.outs.push_back(Stager::Out{a->dir.x, (void*)dd.x, (size_t)n * 12});
        else { sg.outs.push_back(Stager::Out{a->dir.x, (void*)dd.x, (size_t)n * 4}); sg.outs.push_back(Stager::Out{a->dir.y, (void*)dd.y, (size_t)n * 4}); sg.outs.push_back(Stager::Out{a->dir.z, (void*)dd.z, (size_t)n * 4}); }
        const float* sp = sg.in_t((const float*)a->sdtree_pdf, n);
        d.sdtree_pdf = (float*)sp; sg.outs.push_back(Stager::Out{a->sdtree_pdf, (void*)sp, (size_t)n * 4});
        if (a->wo_pdf) { const float* wp = sg.in_t((const float*)a->wo_pdf, n); d.wo_pdf = (float*)wp; sg.outs.push_back(Stager::Out{a->wo_pdf, (void*)wp, (size_t)n * 4}); }
        if (a->em_dir.x) { const float* ep = sg.in_t((const float*)a->sdtree_pdf_em, n); d.sdtree_pdf_em = (float*)ep; sg.outs.push_back(Stager::Out{a->sdtree_pdf_em, (void*)ep, (size_t)n * 4}); }
        if (a->weight.x) {
            SDT_CHECK(h, a->weight.stride == 3, SDT_ERR_INVALID, "sdt_guided: host weight must be interleaved (stride 3)");
            const float* ww = sg.in_t((const float*)a->weight.x, (size_t)n * 3);
            d.weight = sdt_vec3_out{(float*)ww, (float*)ww + 1, (float*)ww + 2, 3};
            sg.outs.push_back(Stager::Out{a->weight.x, (void*)ww, (size_t)n * 12});
        }
    }
    if (sg.status != SDT_OK) return sg.status;
    const float fr = (float)a->bsdf_sampling_fraction, omf = (float)(1.0 - a->bsdf_sampling_fraction);
#define SDT_GUIDED_LAUNCH(EXPL, EM)                                                                              \
    {                                                                                                            \
        GuidedLane<EXPL, EM> f{tree_view(h), d, h->fuse_sample_pdf, fr, omf};                                    \
        SDT_TRY(launch_wavefront(h, st, n, f, h->query_block, h->query_ctas_per_sm, true));                      \
    }
    if (d.u) { if (d.em_dir.x) SDT_GUIDED_LAUNCH(true, true) else SDT_GUIDED_LAUNCH(true, false) }
    else { if (d.em_dir.x) SDT_GUIDED_LAUNCH(false, true) else SDT_GUIDED_LAUNCH(false, false) }
#undef SDT_GUIDED_LAUNCH
    return sg.finish(flags);
}

extern "C" int sdt_mis_nee(sdt_handle h, uint32_t n, const float* bsdf_pdf_em, const float* sdtree_pdf_em,
                           const float* pdf_with_delta, const float* pdf_without_delta, const float* ds_pdf,
                           const uint8_t* ds_delta, double bsdf_sampling_fraction, int32_t iteration,
                           float* surface_pdf_em, float* mis_em, uint32_t flags, sdt_stream stream) {
    SDT_ENTER(h);
    if (n == 0) return SDT_OK;                          // empty wavefront: nothing to do (pointers may be NULL)
    SDT_CHECK(h, bsdf_pdf_em && (iteration <= 1 || (sdtree_pdf_em && pdf_with_delta && pdf_without_delta)) && (!mis_em || ds_pdf),
              SDT_ERR_INVALID, "sdt_mis_nee: missing input");
    cudaStream_t st = (cudaStream_t)stream;
    Stager sg(h, st, flags);
    SDT_TRY(sg.reserve((size_t)n * 32 + 16384));
    MisNeeItem f{sg.in_t(bsdf_pdf_em, n), sg.in_t(sdtree_pdf_em, n), sg.in_t(pdf_with_delta, n), sg.in_t(pdf_without_delta, n),
                 sg.in_t(ds_pdf, n), sg.in_t(ds_delta, n), (float)bsdf_sampling_fraction, (float)(1.0 - bsdf_sampling_fraction), iteration,
                 sg.out_t(surface_pdf_em, n), sg.out_t(mis_em, n)};
    if (sg.status != SDT_OK) return sg.status;
    launch_items(exec_ctx(h, st), nullptr, n, f);
    SDT_TRY(sdt_post_launch(h, "sdt_mis_nee"));
    return sg.finish(flags);
}

extern "C" int sdt_mis_mixture(sdt_handle h, uint32_t n, const float* bsdf_pdf, const float* sdtree_pdf,
                               const sdt_vec3* bsdf_value, const uint8_t* do_mis, double bsdf_sampling_fraction,
                               float* wo_pdf, const sdt_vec3_out* weight, uint32_t flags, sdt_stream stream) {
    SDT_ENTER(h);
    if (n == 0) return SDT_OK;                          // empty wavefront: nothing to do (pointers may be NULL)
    SDT_CHECK(h, bsdf_pdf && sdtree_pdf, SDT_ERR_INVALID, "sdt_mis_mixture: missing input");
    cudaStream_t st = (cudaStream_t)stream;
    Stager sg(h, st, flags);
    SDT_TRY(sg.reserve((size_t)n * 40 + 16384));
    sdt_vec3 bv{nullptr, nullptr, nullptr, 0};
    sdt_vec3_out wv{nullptr, nullptr, nullptr, 0};
    if (bsdf_value) bv = sg.in3(*bsdf_value, n);
    if (weight) wv = sg.out3(*weight, n);
    MisMixtureItem f{sg.in_t(bsdf_pdf, n), sg.in_t(sdtree_pdf, n), bv, sg.in_t(do_mis, n), (float)bsdf_sampling_fraction, (float)(1.0 - bsdf_sampling_fraction), sg.out_t(wo_pdf, n), wv};
    if (sg.status != SDT_OK) return sg.status;
    launch_items(exec_ctx(h, st), nullptr, n, f);
    SDT_TRY(sdt_post_launch(h, "sdt_mis_mixture"));
    return sg.finish(flags);
}

struct DirToCanonicalItem {
    sdt_vec3 d; float* out;
    SDT_HD void operator()(uint32_t i) const {
        float x, y;
        sdt_dir_to_canonical(sdt_ld(d.x, d.stride, i), sdt_ld(d.y, d.stride, i), sdt_ld(d.z, d.stride, i), x, y);
        out[2u * i] = x; out[2u * i + 1u] = y;
    }
};
struct CanonicalToDirItem {
    sdt_vec2 p; sdt_vec3_out d;
    SDT_HD void operator()(uint32_t i) const {
        float x, y, z;
        sdt_canonical_to_dir(sdt_ld(p.x, p.stride, i), sdt_ld(p.y, p.stride, i), x, y, z);
        const int64_t o = (int64_t)i * d.stride;
        d.x[o] = x; d.y[o] = y; d.z[o] = z;
    }
};

extern "C" int sdt_dir_to_canonical(sdt_handle h, const sdt_vec3* dir, uint32_t n, float* out_xy, uint32_t flags, sdt_stream stream) {
    SDT_ENTER(h);
    if (n == 0) return SDT_OK;                          // empty wavefront: nothing to do (pointers may be NULL)
    SDT_CHECK(h, dir && dir->x && out_xy, SDT_ERR_INVALID, "sdt_dir_to_canonical: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    Stager sg(h, st, flags);
    SDT_TRY(sg.reserve((size_t)n * 20 + 8192));
    DirToCanonicalItem f{sg.in3(*dir, n), sg.out_t(out_xy, (size_t)n * 2)};
    if (sg.status != SDT_OK) return sg.status;
    launch_items(exec_ctx(h, st), nullptr, n, f);
    SDT_TRY(sdt_post_launch(h, "sdt_dir_to_canonical"));
    return sg.finish(flags);
}

extern "C" int sdt_canonical_to_dir(sdt_handle h, const sdt_vec2* pos, uint32_t n, const sdt_vec3_out* dir, uint32_t flags, sdt_stream stream) {
    SDT_ENTER(h);
    if (n == 0) return SDT_OK;                          // empty wavefront: nothing to do (pointers may be NULL)
    SDT_CHECK(h, pos && pos->x && dir && dir->x, SDT_ERR_INVALID, "sdt_canonical_to_dir: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    Stager sg(h, st, flags);
    SDT_TRY(sg.reserve((size_t)n * 20 + 8192));
    CanonicalToDirItem f{sg.in2(*pos, n), sg.out3(*dir, n)};
    if (sg.status != SDT_OK) return sg.status;
    launch_items(exec_ctx(h, st), nullptr, n, f);
    SDT_TRY(sdt_post_launch(h, "sdt_canonical_to_dir"));
    return sg.finish(flags);
}
