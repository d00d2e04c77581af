// Host-side handle of libsdtree and small helpers shared by the .inl sections.
#pragma once

#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "sdt_exec.h"

struct QuadSet {
    uint32_t* child = nullptr;      // child_base per canonical node (0 = leaf)
    float* energy = nullptr;        // prev energies per canonical node
    float* thr = nullptr;           // refinementThreshold per node
    float* pp = nullptr;            // per node: pdf product of the root->node path (see sdt_core.h)
    uint32_t* iidx = nullptr;       // per node: number of non-leaf nodes before it (= record index if non-leaf)
    QRec* rec = nullptr;
    QJump* jump = nullptr;          // [root record][cell] jump table over the top SDT_JUMP_LEVELS levels
    uint32_t* jump_pp = nullptr;    // the pdf descents' table of the same shape (path products, SDT_JUMP_NEXT)
    uint32_t* s2 = nullptr;         // second-stage tables [table][8x8] of `jump` (SDT_JUMP_TABLE), ...
    uint32_t* s2_pp = nullptr;      // ... of `jump_pp`
    uint32_t* s2_rec = nullptr;     // [table] -> record of its node
    uint32_t* s2_of = nullptr;      // [level-5 record - s2_rec_lo] -> table id or SDT_NONE
    uint32_t* root_iidx = nullptr;
    DevHeader* hdr = nullptr;
};

struct sdt_tree_s {
    sdt_config cfg{};
    std::string err;
    int num_sms = 148;
    uint32_t kd_cap = 0, quad_cap = 0, rec_cap = 0;

    // spatial tree (append-only, in place)
    uint32_t* kd_word = nullptr;
    float* kd_count = nullptr;      // current.vertCount (fp32 like the reference, src/kdtree.py:21)
    uint32_t* kd_depth = nullptr;
    uint32_t* kd_root = nullptr;    // quadTreeRootIndex of every node
    float* kd_bmin = nullptr;       // 3 per node
    float* kd_bmax = nullptr;
    float* kd_prev_count = nullptr; // prev.vertCount as left by the last refine / upload
    // spatial refine scratch
    uint8_t* kd_s = nullptr;        // split levels per old leaf
    uint32_t* kd_sel = nullptr;     // round: selected old leaves in ascending id
    uint32_t* kd_rank[2] = {nullptr, nullptr};  // round: rank of an old leaf among the selected
    uint32_t* root_src = nullptr;   // new root id -> root id whose tree it copies
    uint32_t* kd_grid = nullptr;    // 16x16x8 cells -> node reached after the first 11 levels (rebuilt with the records)

    QuadSet set[2];
    int cur = 0;                    // set[cur] = topology + prev energies in use
    float* q_ecur = nullptr;        // current: energy accumulators per canonical node

    // quadtree refine scratch (indexed by NEW node id)
    uint32_t* s_src = nullptr;
    uint8_t* s_kind = nullptr;
    uint8_t* s_srem = nullptr;
    uint32_t* s_blk = nullptr;

    bool stats_complete = true;     // interior quadtree energies of `current` are valid
    bool kd_complete = true;        // interior spatial counts of `current` are valid
    // upper bound of the records splatted into `current` since its statistics were last zero, and the host's copy of
    // KDTree.maxLeafSize: together they bound the number of split rounds the next refine can need (no leaf count exceeds
    // the number of records), so the rounds nobody can reach are not launched.  Invalid (all rounds run) once statistics
    // come from elsewhere: an all-reduce, sdt_upload_stats, a caller writing through sdt_stat_buffers.
    uint64_t splat_bound = 0;
    bool splat_bound_valid = true;
    float max_leaf_host = 1.0f;
    bool prev_kd_dirty = false;     // prev.vertCount holds leaf counts only (the refine rolled un-swept counts): sweep before showing them
    DevHeader* h_hdr = nullptr;     // pinned mirror
    cudaStream_t last_stream = nullptr;
    uint64_t launches = 0;
    uint32_t levels_hint = 1;       // upper bound of quadtree levels in use
    uint32_t jump_cap = 0;          // trees the jump table can hold
    uint32_t s2_cap = 0;            // second-stage tables the arena can hold
    int use_jump2 = 1;
    uint32_t jump_trees_known = 0;  // trees covered, as last seen by the host (0 until known: slow path)
    int use_jump = 1;
    int use_int_cell = 1;
    int use_pdl = 1;                // programmatic dependent launch for the refine / sweep helper kernels
    int helper_ctas_per_sm = 8;     // CTAs (256 threads) per SM of the helper launches whose item count lives on the device (measured 1 / 2 / 3 / 4 / 8: refine 0.74 / 0.51 / 0.47 / 0.42 / 0.39 ms on a 3.2 M-node forest)
    int quad_thr_reciprocal = 0;    // semantics switch, see sdt_set_tuning in sdtree.h
    int use_compaction = 1;         // sort the lanes of a tile by mode when a wavefront has idle / mixed lanes
    int use_kd_grid = 1;            // per-CTA 16x16x8 grid over the first 11 spatial levels
    uint32_t kd_nodes_known = 1;    // last spatial node count seen by the host (sizes the smem staging)
    cudaEvent_t hdr_event = nullptr;
    cudaEvent_t order_event = nullptr;   // orders a call on a new stream after the work enqueued on last_stream
    uint32_t n_quad_known = 1;      // quadtree node count as last seen by the host
    uint32_t levels_known = 0;      // quadtree levels of the `prev` tree as last seen by the host (0: not known, use levels_hint)
    bool hdr_pending = false;       // an async header read-back (after refine) is in flight

    // staging for SDT_HOST_PTRS: one arena; large host calls run as a 2-slot pipeline
    // (H2D of chunk k+1 | kernel of chunk k | D2H of chunk k-1 on three streams) whose outputs
    // live in a second arena of the same size, so that H2D never has to wait for a D2H
    char* stage = nullptr;
    char* stage_o = nullptr;
    size_t stage_cap = 0, stage_off = 0;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    int host_chunk = 1 << 20;       // lanes per pipeline chunk
    bool arena_on_st = false;       // a non-pipelined host call staged through the arena on `arena_stream`
    cudaStream_t arena_stream = nullptr;

    // tuning
    int query_block = 0;            // 0 = each kernel's own CTA size (SDT_SAMPLE_THREADS / SDT_QUERY_THREADS / SDT_SPLAT_THREADS);
    int query_ctas_per_sm = 2;      // 2 big CTAs per SM rather than 3-4 small ones: fewer staged copies of the spatial tree -> more L1 (measured)
    int splat_stage_words = 1;
    int kd_smem_count_nodes = 24576; // cap of the shared-memory leaf counters of the splat kernels
    int kd_smem_nodes = 24576;      // cap of the smem-staged prefix of the spatial tree (96 KB)
    int splat_block = 0;
    int splat_ctas_per_sm = 2;
    int fuse_sample_pdf = 1;
    int splat_aggregate = 0;        // warp-aggregate same-address adds of the splat (match_any); pays on coherent wavefronts only

    // per-HANDLE (= per device) launch state of the wavefront kernels: the >48 KB dynamic shared-memory opt-in is a
    // per-device function attribute and the occupancy answer depends on the device, so neither may be cached per process
    struct LaunchCache { size_t attr_smem = 0; int occ = 0, occ_block = 0; size_t occ_smem = ~(size_t)0; };
    std::map<const void*, LaunchCache> launch_cache;
    uint32_t dev_error_seen = 0;    // DevHeader.error as last read back (sticky device-side flag, see sdt_get_sizes)

    // the refine's launch sequence, captured once per (buffer parity, flags, level bounds, settings) and replayed as one graph
    typedef std::tuple<int, uint32_t, uint32_t, uint32_t, int, int, int, int, uint32_t> RefineKey;
#ifndef SDT_HOSTEMU
    struct RefineGraph { cudaGraphExec_t exec = nullptr; uint64_t launches = 0; };
    std::map<RefineKey, RefineGraph> refine_graphs;
#endif
    int use_graph = 1;

    // NCCL
    void* nccl_comm = nullptr;
    int rank = 0, nranks = 1;
};

static thread_local std::string g_create_err;

static int sdt_fail(sdt_handle h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_err = msg;
    return code;
}

#define SDT_CUDA(h, expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return sdt_fail(h, SDT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
// Every entry point runs on the handle's device, whatever device the calling thread had current.
static inline int sdt_enter(sdt_handle h) {
#ifndef SDT_HOSTEMU
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != h->cfg.device) {
        cudaError_t e = cudaSetDevice(h->cfg.device);
        if (e != cudaSuccess) return sdt_fail(h, SDT_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    }
#endif
    return SDT_OK;
}
#define SDT_ENTER(h) do { if (!(h)) return SDT_ERR_INVALID; int _d = sdt_enter(h); if (_d != SDT_OK) return _d; } while (0)
#define SDT_CHECK(h, cond, code, msg) do { if (!(cond)) return sdt_fail(h, code, msg); } while (0)
#define SDT_TRY(expr) do { int _s = (expr); if (_s != SDT_OK) return _s; } while (0)

static inline ExecCtx exec_ctx(sdt_handle h, cudaStream_t st) {
    h->last_stream = st;
    return ExecCtx{st, h->num_sms, h->s_blk, &h->launches, h->use_pdl != 0, h->helper_ctas_per_sm};
}

static inline TreeView tree_view(sdt_tree_s* h) {
    // a refine leaves a non-blocking read-back of the new sizes in flight: pick it up once it landed
    if (h->hdr_pending && cudaEventQuery(h->hdr_event) == cudaSuccess) {
        h->hdr_pending = false;
        h->kd_nodes_known = h->h_hdr->n_kd;
        h->n_quad_known = h->h_hdr->n_quad;
        h->jump_trees_known = h->h_hdr->jump_trees;
        h->dev_error_seen = h->h_hdr->error;
        h->levels_known = h->h_hdr->n_levels;
    }
    const QuadSet& s = h->set[h->cur];
    // deepest quadtree level the sampler can meet: exact once the header has been read back, else the refine's bound
    const uint32_t levels = h->levels_known ? h->levels_known : h->levels_hint;
    const uint32_t cell_mode = !h->use_int_cell ? 0u : (levels <= 17u ? 1u : (levels <= 24u && h->cfg.quad_max_depth <= 23 ? 2u : 0u));
    return TreeView{s.hdr, h->kd_word, h->kd_root, h->kd_grid, s.rec, s.jump, s.jump_pp, s.s2, s.s2_pp, s.s2_rec,
                    (uint32_t)(h->use_jump2 != 0), s.pp, h->use_jump ? h->jump_trees_known : 0u, cell_mode};
}

static int sdt_read_header(sdt_handle h, DevHeader& H);

// refine / allreduce / reset / download work on the statistics the splats wrote: when the caller hands a different
// stream than the one the last work was enqueued on, order the new stream after it (an event, no host wait)
static inline void sdt_order_after_last(sdt_handle h, cudaStream_t st) {
#ifndef SDT_HOSTEMU
    if (h->last_stream != st && h->order_event) {
        if (cudaEventRecord(h->order_event, h->last_stream) == cudaSuccess) cudaStreamWaitEvent(st, h->order_event, 0);
    }
#else
    (void)h; (void)st;
#endif
}

static inline int sdt_post_launch(sdt_handle h, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sdt_fail(h, SDT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return SDT_OK;
}

// Per-call staging of host buffers (SDT_HOST_PTRS): inputs are copied H2D into the
// handle's staging arena before the kernels, outputs D2H after them.
struct Stager {
    sdt_handle h;
    cudaStream_t st;
    bool host;
    int status = SDT_OK;
    int slot = -1;                  // >= 0: pipelined chunk using half `slot` of the arena
    size_t off = 0, lim = 0;
    size_t off_o = 0, lim_o = 0;    // pipelined chunks: outputs in the second arena
    struct Out { void* host; void* dev; size_t bytes; };
    std::vector<Out> outs;

    Stager(sdt_handle h_, cudaStream_t st_, uint32_t flags, int slot_ = -1)
        : h(h_), st(st_), host((flags & SDT_HOST_PTRS) != 0), slot(slot_) {
        if (slot >= 0) {
            const size_t half = (h->stage_cap / 2) & ~(size_t)255;
            off = (size_t)slot * half; lim = off + half;
            off_o = off; lim_o = lim;
            // the slot's input buffers are free once the kernels of the chunk that used it have run
            cudaStreamWaitEvent(h->s_in, h->ev_comp[slot], 0);
        } else {
            off = 0; lim = h->stage_cap;
            if (host && h->s_in) {
                // whole-arena call: chunks of an earlier pipelined call (possibly SDT_NO_WAIT, possibly on another
                // stream) may still be reading / writing their slots
                for (int k = 0; k < 2; ++k) { cudaStreamWaitEvent(st, h->ev_comp[k], 0); cudaStreamWaitEvent(st, h->ev_out[k], 0); }
            }
        }
    }
    // total bytes this call will stage; grows the arena once (pointers stay valid)
    int reserve(size_t bytes) {
        if (!host) return SDT_OK;
        bytes += 4096;
        if (bytes <= h->stage_cap) { if (slot < 0) lim = h->stage_cap; return SDT_OK; }
        if (h->stage) { cudaDeviceSynchronize(); cudaFree(h->stage); cudaFree(h->stage_o); h->stage = h->stage_o = nullptr; h->stage_cap = 0; }
        size_t cap = bytes + bytes / 4;
        if (cudaMalloc((void**)&h->stage, cap) != cudaSuccess || (h->s_in && cudaMalloc((void**)&h->stage_o, cap) != cudaSuccess)) {
            if (h->stage) { cudaFree(h->stage); h->stage = nullptr; }
            status = sdt_fail(h, SDT_ERR_CUDA, "staging arena cudaMalloc failed"); return status;
        }
        h->stage_cap = cap;
        if (slot < 0) lim = cap;
        return SDT_OK;
    }
    void* alloc(size_t bytes) {
        size_t o = (off + 255) & ~(size_t)255;
        if (o + bytes > lim) { status = sdt_fail(h, SDT_ERR_INVALID, "staging arena overflow (reserve too small)"); return nullptr; }
        off = o + bytes;
        if (slot < 0) { h->arena_on_st = true; h->arena_stream = st; }
        return h->stage + o;
    }
    const void* in(const void* p, size_t bytes) {
        if (!host || !p) return p;
        void* d = alloc(bytes);
        if (!d) return nullptr;
        if (cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, slot >= 0 ? h->s_in : st) != cudaSuccess) { status = sdt_fail(h, SDT_ERR_CUDA, "H2D staging copy failed"); return nullptr; }
        return d;
    }
    void* out(void* p, size_t bytes) {
        if (!host || !p) return p;
        void* d;
        if (slot >= 0) {
            const size_t o = (off_o + 255) & ~(size_t)255;
            if (o + bytes > lim_o) { status = sdt_fail(h, SDT_ERR_INVALID, "staging arena overflow (reserve too small)"); return nullptr; }
            off_o = o + bytes;
            d = h->stage_o + o;
        } else d = alloc(bytes);
        if (!d) return nullptr;
        outs.push_back(Out{p, d, bytes});
        return d;
    }
    // pipelined chunk: the kernels wait for this chunk's inputs and for the D2H of the chunk that
    // used the slot's output buffers before
    void before_launch() {
        if (slot < 0) return;
        cudaEventRecord(h->ev_in[slot], h->s_in);
        cudaStreamWaitEvent(st, h->ev_in[slot], 0);
        cudaStreamWaitEvent(st, h->ev_out[slot], 0);
    }
    template <class T> const T* in_t(const T* p, size_t n) { return (const T*)in(p, n * sizeof(T)); }
    template <class T> T* out_t(T* p, size_t n) { return (T*)out(p, n * sizeof(T)); }
    // host vectors must be either one interleaved (n,3) block or three SoA planes
    sdt_vec3 in3(const sdt_vec3& v, size_t n) {
        if (!host || !v.x) return v;
        sdt_vec3 r = v;
        if (v.stride == 3 && v.y == v.x + 1 && v.z == v.x + 2) {
            const float* d = in_t(v.x, 3 * n);
            r.x = d; r.y = d + 1; r.z = d + 2;
        } else if (v.stride == 1) {
            r.x = in_t(v.x, n); r.y = in_t(v.y, n); r.z = in_t(v.z, n);
        } else status = sdt_fail(h, SDT_ERR_INVALID, "host vec3 must be interleaved (stride 3) or SoA (stride 1)");
        return r;
    }
    sdt_vec2 in2(const sdt_vec2& v, size_t n) {
        if (!host || !v.x) return v;
        sdt_vec2 r = v;
        if (v.stride == 2 && v.y == v.x + 1) {
            const float* d = in_t(v.x, 2 * n);
            r.x = d; r.y = d + 1;
        } else if (v.stride == 1) {
            r.x = in_t(v.x, n); r.y = in_t(v.y, n);
        } else status = sdt_fail(h, SDT_ERR_INVALID, "host vec2 must be interleaved (stride 2) or SoA (stride 1)");
        return r;
    }
    sdt_vec3_out out3(const sdt_vec3_out& v, size_t n) {
        if (!host || !v.x) return v;
        sdt_vec3_out r = v;
        if (v.stride == 3 && v.y == v.x + 1 && v.z == v.x + 2) {
            float* d = out_t(v.x, 3 * n);
            r.x = d; r.y = d + 1; r.z = d + 2;
        } else if (v.stride == 1) {
            r.x = out_t(v.x, n); r.y = out_t(v.y, n); r.z = out_t(v.z, n);
        } else status = sdt_fail(h, SDT_ERR_INVALID, "host vec3 must be interleaved (stride 3) or SoA (stride 1)");
        return r;
    }
    int finish(uint32_t flags) {
        if (status != SDT_OK) return status;
        if (slot >= 0) {            // pipelined chunk: D2H on the output stream, no synchronisation here
            cudaEventRecord(h->ev_comp[slot], st);
            cudaStreamWaitEvent(h->s_out, h->ev_comp[slot], 0);
            for (const Out& o : outs)
                if (cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, h->s_out) != cudaSuccess)
                    return sdt_fail(h, SDT_ERR_CUDA, "D2H staging copy failed");
            cudaEventRecord(h->ev_out[slot], h->s_out);
            return SDT_OK;
        }
        for (const Out& o : outs) {
            if (cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess)
                return sdt_fail(h, SDT_ERR_CUDA, "D2H staging copy failed");
        }
        // results in pageable/pinned host memory are only valid after the stream drains; a host-pointer call without
        // outputs (a splat) waits too, so that the caller may reuse its input arrays as soon as the call returns
        if ((flags & SDT_SYNC) || (host && !(flags & SDT_NO_WAIT))) {
            if (cudaStreamSynchronize(st) != cudaSuccess) return sdt_fail(h, SDT_ERR_CUDA, "stream synchronize failed");
        }
        return SDT_OK;
    }
};

// Runs body(stager, first_lane, lane_count) once (device pointers / small host calls) or as a
// 2-slot pipeline over chunks of h->host_chunk lanes (large SDT_HOST_PTRS calls): the H2D copy of
// chunk k+1, the kernels of chunk k and the D2H copy of chunk k-1 overlap on three streams.
template <class Body>
static int sdt_run_chunked(sdt_handle h, cudaStream_t st, uint32_t flags, uint32_t n, size_t bytes_per_lane, bool has_outputs, Body body) {
    const bool host = (flags & SDT_HOST_PTRS) != 0;
    const uint32_t chunk = (uint32_t)h->host_chunk;
    if (!host || !h->s_in || n <= chunk + chunk / 2) {
        Stager sg(h, st, flags);
        SDT_TRY(sg.reserve((size_t)n * bytes_per_lane + 65536));
        SDT_TRY(body(sg, 0u, n));
        return sg.finish(flags);
    }
    {
        Stager probe(h, st, flags);
        SDT_TRY(probe.reserve(2 * ((size_t)chunk * bytes_per_lane + 65536)));
    }
    // a small host call stages through the whole arena on its own stream: the input stream starts after it.
    // Chunks of an earlier pipelined call (SDT_NO_WAIT) are covered slot by slot: ev_comp (Stager) frees the slot's
    // inputs for the H2D, ev_out (before_launch) its outputs for the kernels; inputs and outputs live in separate
    // arenas, so the D2H tail of the earlier call runs under this call's H2D head.
    if (h->arena_on_st) {
        cudaEventRecord(h->ev_in[0], h->arena_stream);
        cudaStreamWaitEvent(h->s_in, h->ev_in[0], 0);
        h->arena_on_st = false;
    }
    int k = 0;
    for (uint32_t off = 0; off < n; off += chunk, ++k) {
        const uint32_t cnt = n - off < chunk ? n - off : chunk;
        Stager sg(h, st, flags, k & 1);
        SDT_TRY(body(sg, off, cnt));
        SDT_TRY(sg.finish(flags));
    }
    // stream order for the caller: everything, D2H included, is complete when `st` drains
    cudaStreamWaitEvent(st, h->ev_out[0], 0);
    cudaStreamWaitEvent(st, h->ev_out[1], 0);
    if ((flags & SDT_SYNC) || (has_outputs && !(flags & SDT_NO_WAIT))) {
        if (cudaStreamSynchronize(st) != cudaSuccess) return sdt_fail(h, SDT_ERR_CUDA, "stream synchronize failed");
    } else if (!(flags & SDT_NO_WAIT)) {
        // no outputs (a splat): the caller may reuse its host arrays once the last H2D copy has been issued AND has run
        if (cudaStreamSynchronize(h->s_in) != cudaSuccess) return sdt_fail(h, SDT_ERR_CUDA, "input stream synchronize failed");
    }
    return SDT_OK;
}

static inline sdt_vec3 sdt_off3(const sdt_vec3& v, uint32_t off) {
    sdt_vec3 r = v;
    if (v.x) { r.x = v.x + (int64_t)off * v.stride; r.y = v.y + (int64_t)off * v.stride; r.z = v.z + (int64_t)off * v.stride; }
    return r;
}
static inline sdt_vec3_out sdt_off3o(const sdt_vec3_out& v, uint32_t off) {
    sdt_vec3_out r = v;
    if (v.x) { r.x = v.x + (int64_t)off * v.stride; r.y = v.y + (int64_t)off * v.stride; r.z = v.z + (int64_t)off * v.stride; }
    return r;
}
static inline sdt_vec2 sdt_off2(const sdt_vec2& v, uint32_t off) {
    sdt_vec2 r = v;
    if (v.x) { r.x = v.x + (int64_t)off * v.stride; r.y = v.y + (int64_t)off * v.stride; }
    return r;
}
template <class T> static inline T* sdt_offp(T* p, size_t off) { return p ? p + off : p; }
