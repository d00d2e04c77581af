// Host-side handle of libsdtree and small helpers shared by the .inl sections.
#pragma once

#include <string>
#include <vector>

#include "sdt_exec.h"

struct QuadSet {
    uint32_t* child = nullptr;      // child_base per canonical node (0 = leaf)
    float* energy = nullptr;        // prev energies per canonical node
    float* thr = nullptr;           // refinementThreshold per node
    uint32_t* iidx = nullptr;       // per node: number of non-leaf nodes before it (= record index if non-leaf)
    QRec* rec = nullptr;
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

    QuadSet set[2];
    int cur = 0;                    // set[cur] = topology + prev energies in use
    float* q_ecur = nullptr;        // current: energy accumulators per canonical node

    // quadtree refine scratch (indexed by NEW node id)
    uint32_t* s_src = nullptr;
    uint8_t* s_kind = nullptr;
    uint8_t* s_srem = nullptr;
    uint32_t* s_blk = nullptr;

    bool stats_complete = true;     // interior statistics of `current` are valid
    DevHeader* h_hdr = nullptr;     // pinned mirror
    cudaStream_t last_stream = nullptr;
    uint64_t launches = 0;
    uint32_t levels_hint = 1;       // upper bound of quadtree levels in use
    uint32_t kd_nodes_known = 1;    // last spatial node count seen by the host (sizes the smem staging)
    cudaEvent_t hdr_event = nullptr;
    bool hdr_pending = false;       // an async header read-back (after refine) is in flight

    // staging for SDT_HOST_PTRS
    char* stage = nullptr;
    size_t stage_cap = 0, stage_off = 0;

    // tuning
    int query_block = 512;
    int query_ctas_per_sm = 3;
    int kd_smem_nodes = 24576;      // cap of the smem-staged prefix of the spatial tree (96 KB)
    int splat_block = 512;
    int splat_ctas_per_sm = 3;
    int fuse_sample_pdf = 1;
    int splat_all_levels = 0;       // 1: atomics at every level like the reference (no sweep)

    // NCCL
    void* nccl_lib = nullptr;
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
#define SDT_CHECK(h, cond, code, msg) do { if (!(cond)) return sdt_fail(h, code, msg); } while (0)
#define SDT_TRY(expr) do { int _s = (expr); if (_s != SDT_OK) return _s; } while (0)

static inline ExecCtx exec_ctx(sdt_handle h, cudaStream_t st) {
    h->last_stream = st;
    return ExecCtx{st, h->num_sms, h->s_blk, &h->launches};
}

static inline TreeView tree_view(const sdt_tree_s* h) {
    const QuadSet& s = h->set[h->cur];
    return TreeView{s.hdr, h->kd_word, h->kd_root, s.rec};
}

static int sdt_read_header(sdt_handle h, DevHeader& H);

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
    struct Out { void* host; void* dev; size_t bytes; };
    std::vector<Out> outs;
    size_t h2d = 0, d2h = 0;

    Stager(sdt_handle h_, cudaStream_t st_, uint32_t flags) : h(h_), st(st_), host((flags & SDT_HOST_PTRS) != 0) {
        h->stage_off = 0;
    }
    // total bytes this call will stage; grows the arena once (pointers stay valid)
    int reserve(size_t bytes) {
        if (!host) return SDT_OK;
        bytes += 4096;
        if (bytes <= h->stage_cap) return SDT_OK;
        if (h->stage) { cudaStreamSynchronize(st); cudaFree(h->stage); h->stage = nullptr; h->stage_cap = 0; }
        size_t cap = bytes + bytes / 4;
        if (cudaMalloc((void**)&h->stage, cap) != cudaSuccess) { status = sdt_fail(h, SDT_ERR_CUDA, "staging arena cudaMalloc failed"); return status; }
        h->stage_cap = cap;
        return SDT_OK;
    }
    void* alloc(size_t bytes) {
        size_t off = (h->stage_off + 255) & ~(size_t)255;
        if (off + bytes > h->stage_cap) { status = sdt_fail(h, SDT_ERR_INVALID, "staging arena overflow (reserve too small)"); return nullptr; }
        h->stage_off = off + bytes;
        return h->stage + off;
    }
    const void* in(const void* p, size_t bytes) {
        if (!host || !p) return p;
        void* d = alloc(bytes);
        if (!d) return nullptr;
        if (cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) { status = sdt_fail(h, SDT_ERR_CUDA, "H2D staging copy failed"); return nullptr; }
        h2d += bytes;
        return d;
    }
    void* out(void* p, size_t bytes) {
        if (!host || !p) return p;
        void* d = alloc(bytes);
        if (!d) return nullptr;
        outs.push_back(Out{p, d, bytes});
        return d;
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
        for (const Out& o : outs) {
            if (cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess)
                return sdt_fail(h, SDT_ERR_CUDA, "D2H staging copy failed");
            d2h += o.bytes;
        }
        // results in pageable/pinned host memory are only valid after the stream drains
        if ((flags & SDT_SYNC) || (host && !outs.empty())) {
            if (cudaStreamSynchronize(st) != cudaSuccess) return sdt_fail(h, SDT_ERR_CUDA, "stream synchronize failed");
        }
        return SDT_OK;
    }
};
