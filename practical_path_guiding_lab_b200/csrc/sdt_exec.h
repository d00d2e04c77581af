// Launch helpers: "for every item i < n" and "exclusive scan over items", where n may
// live in device memory (so that a chain of dependent passes -- the refine -- needs no
// host round-trip).  CUDA: fixed persistent grids (a multiple of the SM count) with
// grid-stride / chunked loops.  SDT_HOSTEMU: serial loops (test infrastructure only,
// see sdt_platform.h).
#pragma once

#include "sdt_core.h"

#define SDT_SCAN_MAX_BLOCKS 1024

struct ExecCtx {
    cudaStream_t st;
    int num_sms;
    uint32_t* blk;        // SDT_SCAN_MAX_BLOCKS words of device scratch for the scans
    uint64_t* launches;   // kernel launch counter of the handle
    bool pdl;             // programmatic dependent launch for the helper kernels (see sdt_launch)
};

#ifndef SDT_HOSTEMU
// ---------------------------------------------------------------------------- CUDA
// The refine is a chain of ~200 tiny dependent kernels: launch latency is all it costs.  Every helper
// kernel therefore starts with sdt_grid_dep(): it lets the NEXT kernel of the stream be scheduled right
// away (griddepcontrol.launch_dependents) and then waits until the PREVIOUS one has completed and
// flushed (griddepcontrol.wait) -- the dependency is kept, the launch latencies overlap.  Both are no-ops
// for a kernel launched without the programmatic-serialization attribute.
__device__ __forceinline__ void sdt_grid_dep() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <class... KArgs, class... Args>
static inline void sdt_launch(const ExecCtx& x, void (*kernel)(KArgs...), uint32_t grid, uint32_t block, Args... args) {
    if (!x.pdl) { kernel<<<grid, block, 0, x.st>>>(args...); return; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = x.st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <class F>
__global__ void __launch_bounds__(256) k_items(F f, const uint32_t* n_ptr, uint32_t n_imm) {
    sdt_grid_dep();
    const uint32_t n = n_ptr ? *n_ptr : n_imm;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) f(i);
}

// n = *n_ptr (device) when n_ptr != NULL, else n_imm
template <class F>
static inline void launch_items(const ExecCtx& x, const uint32_t* n_ptr, uint32_t n_imm, F f) {
    uint32_t grid;
    if (n_ptr) grid = (uint32_t)x.num_sms * 4u;
    else {
        if (n_imm == 0) return;
        grid = (n_imm + 255u) / 256u;
        const uint32_t cap = (uint32_t)x.num_sms * 8u;
        if (grid > cap) grid = cap;
    }
    sdt_launch(x, k_items<F>, grid, 256u, f, n_ptr, n_imm);
    ++*x.launches;
}

__device__ __forceinline__ uint32_t sdt_block_excl_scan(uint32_t v, uint32_t* warp_sums, uint32_t& block_total) {
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31u) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const uint32_t nw = blockDim.x >> 5;
        uint32_t w = lane < nw ? warp_sums[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, o);
            if (lane >= (uint32_t)o) wi += t;
        }
        if (lane < nw) warp_sums[lane] = wi - w;          // exclusive warp offsets
        if (lane == nw - 1u) warp_sums[32] = wi;          // block total
    }
    __syncthreads();
    const uint32_t r = warp_sums[wid] + incl - v;
    block_total = warp_sums[32];
    __syncthreads();
    return r;
}

__device__ __forceinline__ void sdt_chunk(uint32_t n, uint32_t& lo, uint32_t& hi) {
    uint32_t chunk = (n + gridDim.x - 1u) / gridDim.x;
    chunk = (chunk + blockDim.x - 1u) / blockDim.x * blockDim.x;
    const uint64_t a = (uint64_t)blockIdx.x * chunk, b = a + chunk;
    lo = a < n ? (uint32_t)a : n;
    hi = b < n ? (uint32_t)b : n;
}

template <class Flag>
__global__ void __launch_bounds__(256) k_scan_reduce(Flag flag, const uint32_t* n_ptr, uint32_t n_imm, uint32_t* blk) {
    __shared__ uint32_t ws[33];
    sdt_grid_dep();
    const uint32_t n = n_ptr ? *n_ptr : n_imm;
    uint32_t lo, hi;
    sdt_chunk(n, lo, hi);
    uint32_t acc = 0;
    for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += flag(i);
    uint32_t tot;
    sdt_block_excl_scan(acc, ws, tot);
    if (threadIdx.x == 0) blk[blockIdx.x] = tot;
}

template <class Fin>
__global__ void __launch_bounds__(SDT_SCAN_MAX_BLOCKS) k_scan_blocks(uint32_t* blk, uint32_t nblk, Fin fin) {
    __shared__ uint32_t ws[33];
    sdt_grid_dep();
    const uint32_t v = threadIdx.x < nblk ? blk[threadIdx.x] : 0u;
    uint32_t tot;
    const uint32_t ex = sdt_block_excl_scan(v, ws, tot);
    if (threadIdx.x < nblk) blk[threadIdx.x] = ex;
    if (threadIdx.x == 0) fin(tot);
}

template <class Flag, class Emit>
__global__ void __launch_bounds__(256) k_scan_emit(Flag flag, Emit emit, const uint32_t* n_ptr, uint32_t n_imm, const uint32_t* blk) {
    __shared__ uint32_t ws[33];
    sdt_grid_dep();
    const uint32_t n = n_ptr ? *n_ptr : n_imm;
    uint32_t lo, hi;
    sdt_chunk(n, lo, hi);
    uint32_t carry = blk[blockIdx.x];
    for (uint32_t base = lo; base < hi; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < hi ? flag(i) : 0u;
        uint32_t tot;
        const uint32_t ex = sdt_block_excl_scan(v, ws, tot);
        if (i < hi) emit(i, carry + ex, v);
        carry += tot;
    }
}

// for i < n: emit(i, sum_{j<i} flag(j), flag(i)); then fin(sum over all) on one thread.
// flag must be pure and must not read anything emit writes.
template <class Flag, class Emit, class Fin>
static inline void launch_scan(const ExecCtx& x, const uint32_t* n_ptr, uint32_t n_imm, Flag flag, Emit emit, Fin fin) {
    uint32_t grid = (uint32_t)x.num_sms * 4u;
    if (grid > SDT_SCAN_MAX_BLOCKS) grid = SDT_SCAN_MAX_BLOCKS;
    if (!n_ptr) {
        const uint32_t g = (n_imm + 255u) / 256u;
        if (g < grid) grid = g ? g : 1u;
    }
    sdt_launch(x, k_scan_reduce<Flag>, grid, 256u, flag, n_ptr, n_imm, x.blk);
    sdt_launch(x, k_scan_blocks<Fin>, 1u, (uint32_t)SDT_SCAN_MAX_BLOCKS, x.blk, grid, fin);
    sdt_launch(x, k_scan_emit<Flag, Emit>, grid, 256u, flag, emit, n_ptr, n_imm, (const uint32_t*)x.blk);
    *x.launches += 3;
}

template <class F>
__global__ void k_single(F f) { sdt_grid_dep(); f(); }
template <class F>
static inline void launch_single(const ExecCtx& x, F f) {
    sdt_launch(x, k_single<F>, 1u, 1u, f);
    ++*x.launches;
}

#else
// ---------------------------------------------------------------------------- host emulation
template <class F>
static inline void launch_items(const ExecCtx& x, const uint32_t* n_ptr, uint32_t n_imm, F f) {
    const uint32_t n = n_ptr ? *n_ptr : n_imm;
    for (uint32_t i = 0; i < n; ++i) f(i);
    ++*x.launches;
}
template <class Flag, class Emit, class Fin>
static inline void launch_scan(const ExecCtx& x, const uint32_t* n_ptr, uint32_t n_imm, Flag flag, Emit emit, Fin fin) {
    const uint32_t n = n_ptr ? *n_ptr : n_imm;
    // like the device version: ranks from a first evaluation of flag, fin, then emit with
    // flag evaluated AGAIN (fin may have changed what it reads, e.g. a capacity cut-off)
    uint32_t run = 0;
    uint32_t* ranks = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    uint32_t* vals = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    for (uint32_t i = 0; i < n; ++i) { vals[i] = flag(i); ranks[i] = run; run += vals[i]; }
    fin(run);
    for (uint32_t i = 0; i < n; ++i) emit(i, ranks[i], flag(i));
    free(ranks); free(vals);
    *x.launches += 3;
}
template <class F>
static inline void launch_single(const ExecCtx& x, F f) { f(); ++*x.launches; }
#endif

SDT_HD void sdt_atomic_add_f32(float* p, float v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
