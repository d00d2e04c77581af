// Launch helpers: "for every item i < n" and "exclusive scan over items", where n may
// live in device memory (so that a chain of dependent passes -- the refine -- needs no
// host round-trip).  CUDA: fixed persistent grids (a multiple of the SM count) with
// grid-stride / chunked loops.  SDT_HOSTEMU: serial loops (test infrastructure only,
// see sdt_platform.h).
#pragma once

#include "sdt_core.h"

#define SDT_SCAN_MAX_BLOCKS 1024
// scratch of one scan: [0] ticket counter, [1] finished-blocks counter, then one 64-bit status word per block
#define SDT_SCAN_STATE_WORDS (4 + 2 * SDT_SCAN_MAX_BLOCKS)

struct ExecCtx {
    cudaStream_t st;
    int num_sms;
    uint32_t* blk;        // device scratch of the scans: SDT_SCAN_STATE_WORDS words, all zero between scans
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

// Exclusive scan + emit + fin in ONE kernel (the refine is a chain of dependent launches, so a scan that costs one launch
// instead of three shortens it by two launch latencies per level): single pass with decoupled look-back.  Blocks take
// their chunk in the order they START (a ticket), so a block only ever waits for blocks that are already running -- no
// co-residency assumption, safe under programmatic dependent launch.  Status word of a block: (kind << 32) | value with
// kind 1 = its own total, 2 = inclusive prefix.  The last block to finish zeroes the scratch for the next scan.
// Order of side effects: emit of every block may run BEFORE fin (fin runs on the block that holds the last chunk, after
// its look-back); emit therefore must not depend on what fin writes.
template <class Flag, class Emit, class Fin>
__global__ void __launch_bounds__(256) k_scan_fused(Flag flag, Emit emit, Fin fin, const uint32_t* n_ptr, uint32_t n_imm, uint32_t* state) {
    __shared__ uint32_t ws[33];
    __shared__ uint32_t s_b, s_prefix;
    sdt_grid_dep();
    unsigned long long* status = reinterpret_cast<unsigned long long*>(state + 4);
    if (threadIdx.x == 0) s_b = atomicAdd(state, 1u);
    __syncthreads();
    const uint32_t b = s_b;
    const uint32_t n = n_ptr ? *n_ptr : n_imm;
    uint32_t chunk = (n + gridDim.x - 1u) / gridDim.x;
    chunk = (chunk + blockDim.x - 1u) / blockDim.x * blockDim.x;
    const uint64_t a64 = (uint64_t)b * chunk, b64 = a64 + chunk;
    const uint32_t lo = a64 < n ? (uint32_t)a64 : n, hi = b64 < n ? (uint32_t)b64 : n;
    uint32_t acc = 0;
    for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += flag(i);
    uint32_t total;
    sdt_block_excl_scan(acc, ws, total);
    if (threadIdx.x < 32u) {
        const uint32_t lane = threadIdx.x;
        if (lane == 0u) {
            atomicExch(&status[b], ((unsigned long long)(b == 0u ? 2u : 1u) << 32) | total);
        }
        uint32_t prefix = 0;
        if (b > 0u) {
            int j = (int)b - 1;                       // look back, 32 predecessors at a time
            for (;;) {
                const int idx = j - (int)lane;
                unsigned long long sv = 0;
                if (idx >= 0) {
                    do { sv = *reinterpret_cast<volatile unsigned long long*>(&status[idx]); } while ((sv >> 32) == 0ull);
                } else sv = 2ull << 32;               // before the first block: an inclusive prefix of 0
                const uint32_t kind = (uint32_t)(sv >> 32), val = (uint32_t)sv;
                const uint32_t incl = __ballot_sync(0xFFFFFFFFu, kind == 2u);
                // lanes up to and including the nearest inclusive prefix contribute
                const uint32_t first = incl ? (uint32_t)__ffs(incl) - 1u : 32u;
                uint32_t v = lane <= first ? val : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
                prefix += v;
                if (incl) break;
                j -= 32;
            }
            if (lane == 0u) atomicExch(&status[b], (2ull << 32) | (unsigned long long)(prefix + total));
        }
        if (lane == 0u) s_prefix = prefix;
    }
    __syncthreads();
    uint32_t carry = s_prefix;
    if (threadIdx.x == 0 && b == gridDim.x - 1u) fin(carry + total);
    for (uint32_t base = lo; base < hi; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < hi ? flag(i) : 0u;
        uint32_t tot;
        const uint32_t ex = sdt_block_excl_scan(v, ws, tot);
        if (i < hi) emit(i, carry + ex, v);
        carry += tot;
    }
    // the last block out resets the scratch (every block is past its look-back by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_b = atomicAdd(state + 1, 1u);
    }
    __syncthreads();
    if (s_b == gridDim.x - 1u) {
        for (uint32_t k = threadIdx.x; k < 2u * gridDim.x; k += blockDim.x) state[4u + k] = 0u;
        if (threadIdx.x == 0) { state[0] = 0u; state[1] = 0u; }
    }
}

// for i < n: emit(i, sum_{j<i} flag(j), flag(i)); fin(sum over all) on one thread, possibly AFTER some emits.
// flag must be pure and must not read anything emit or fin writes.
template <class Flag, class Emit, class Fin>
static inline void launch_scan(const ExecCtx& x, const uint32_t* n_ptr, uint32_t n_imm, Flag flag, Emit emit, Fin fin) {
    uint32_t grid = (uint32_t)x.num_sms * 4u;
    if (grid > SDT_SCAN_MAX_BLOCKS) grid = SDT_SCAN_MAX_BLOCKS;
    if (!n_ptr) {
        const uint32_t g = (n_imm + 255u) / 256u;
        if (g < grid) grid = g ? g : 1u;
    }
    sdt_launch(x, k_scan_fused<Flag, Emit, Fin>, grid, 256u, flag, emit, fin, n_ptr, n_imm, x.blk);
    *x.launches += 1;
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
    // like the device version: ranks from a first evaluation of flag, emit with flag evaluated AGAIN, and fin LAST (on
    // the device fin runs on one block while others may already have emitted: emit must not depend on it)
    uint32_t run = 0;
    uint32_t* ranks = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    for (uint32_t i = 0; i < n; ++i) { ranks[i] = run; run += flag(i); }
    for (uint32_t i = 0; i < n; ++i) emit(i, ranks[i], flag(i));
    fin(run);
    free(ranks);
    *x.launches += 1;
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
