// Launch helpers: "for every item i < n" and "exclusive scan over items", where n may
// live in device memory (so that a chain of dependent passes -- the refine -- needs no
// host round-trip).  CUDA: fixed persistent grids (a multiple of the SM count) with
// grid-stride / chunked loops.  SDT_HOSTEMU: serial loops (test infrastructure only,
// see sdt_platform.h).
#pragma once

#include "sdt_core.h"

#define SDT_SCAN_MAX_BLOCKS 1024
// scratch of the scans: [0] scan counter (the epoch that tags the status words), [1] stall flag, then one 64-bit status word per block
#define SDT_SCAN_STATE_WORDS (4 + 2 * SDT_SCAN_MAX_BLOCKS)

struct ExecCtx {
    cudaStream_t st;
    int num_sms;
    uint32_t* blk;        // device scratch of the scans: SDT_SCAN_STATE_WORDS words (zero at creation, never reset)
    uint64_t* launches;   // kernel launch counter of the handle
    bool pdl;             // programmatic dependent launch for the helper kernels (see sdt_launch)
    int ctas_per_sm;      // grid of the device-sized helper launches ("helper_ctas_per_sm")
};

// what a scan adds up: the flag functor returns the 0/1 count itself, or a struct that carries it together with whatever
// the emit wants back (overload sdt_scan_count next to the struct)
SDT_HD uint32_t sdt_scan_count(uint32_t v) { return v; }

#ifndef SDT_HOSTEMU
// ---------------------------------------------------------------------------- CUDA
// The refine is a chain of some 45 small dependent kernels: latency is what it costs.  Every helper
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
    if (n_ptr) grid = (uint32_t)x.num_sms * (uint32_t)x.ctas_per_sm;
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
// instead of three shortens it by two launch latencies per level): single pass with decoupled look-back.
//
// What the chain pays per scan is LATENCY, and the first version of this kernel spent most of it on two same-address
// atomics per block (a start-order ticket and a finished-blocks counter for the scratch reset: ~4 ns each, 8-11 us for an
// EMPTY level at 592 blocks).  Now there is none:
//  * the grid is fixed (n lives on the device) at 4 blocks of 256 threads per SM, which __launch_bounds__(256, 4) makes
//    co-resident by construction -- so a block may wait for any lower-numbered block whatever order they start in (blocks
//    of the previous kernel, still around under programmatic dependent launch, leave on their own);
//  * only ceil(n / chunk) blocks take part, the others return at once: on a level of a few thousand nodes the look-back
//    chain is as short as the data;
//  * status words are tagged with the scan's number: (epoch << 34) | (kind << 32) | value, kind 1 = the block's own total,
//    2 = inclusive prefix.  Words left by earlier scans never match, so nothing is reset; the epoch lives in state[0] and is
//    advanced by the block that holds the last chunk once its look-back is through -- every other participant has
//    published by then, hence read the epoch.
//  * a thread takes as few items as the grid allows (one up to 151 k items on 148 SMs); while a chunk fits SDT_SCAN_ITEMS
//    items per thread each thread evaluates flag ONCE, keeps the results in registers -- flag may return a struct
//    (sdt_scan_count names its 0/1 count), emit gets it back -- and ONE block scan yields the block total and the ranks;
//    larger inputs loop over the chunk twice and evaluate flag again for the emit.
// Order of side effects: emit of every block may run BEFORE fin (fin runs on the block that holds the last chunk, after
// its look-back); emit therefore must not depend on what fin writes.
#define SDT_SCAN_ITEMS 4
#define SDT_SCAN_BLOCKS_PER_SM 4
template <class Flag, class Emit, class Fin>
__global__ void __launch_bounds__(256, SDT_SCAN_BLOCKS_PER_SM) k_scan_fused(Flag flag, Emit emit, Fin fin, const uint32_t* n_ptr, uint32_t n_imm, uint32_t* state) {
    typedef decltype(flag(0u)) V;
    __shared__ uint32_t ws[33];
    __shared__ uint32_t s_prefix;
    sdt_grid_dep();
    unsigned long long* status = reinterpret_cast<unsigned long long*>(state + 4);
    const uint32_t n = n_ptr ? *n_ptr : n_imm;
    uint32_t chunk = (n + gridDim.x - 1u) / gridDim.x;
    chunk = (chunk + blockDim.x - 1u) / blockDim.x * blockDim.x;
    if (chunk < blockDim.x) chunk = blockDim.x;
    uint32_t eff = (uint32_t)(((uint64_t)n + chunk - 1u) / chunk);   // blocks that take part (block 0 always does: fin)
    if (eff == 0u) eff = 1u;
    const uint32_t b = blockIdx.x;
    if (b >= eff) return;
    const uint64_t a64 = (uint64_t)b * chunk, b64 = a64 + chunk;
    const uint32_t lo = a64 < n ? (uint32_t)a64 : n, hi = b64 < n ? (uint32_t)b64 : n;
    const bool cached = chunk <= SDT_SCAN_ITEMS * blockDim.x;
    const uint32_t per = chunk / blockDim.x;                          // items per thread (thread-contiguous)
    const uint32_t t0 = lo + threadIdx.x * per;
    uint32_t e_raw = 0;
    if (threadIdx.x == 0) e_raw = *reinterpret_cast<volatile uint32_t*>(state);
    V vals[SDT_SCAN_ITEMS];
    uint32_t acc = 0;
    if (cached) {
#pragma unroll
        for (uint32_t k = 0; k < SDT_SCAN_ITEMS; ++k)
            if (k < per && t0 + k < hi) { vals[k] = flag(t0 + k); acc += sdt_scan_count(vals[k]); }
    } else {
        for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += sdt_scan_count(flag(i));
    }
    uint32_t total;
    const uint32_t ex0 = sdt_block_excl_scan(acc, ws, total);
    if (threadIdx.x < 32u) {
        const uint32_t lane = threadIdx.x;
        e_raw = __shfl_sync(0xFFFFFFFFu, e_raw, 0);
        const unsigned long long epoch = (unsigned long long)(e_raw % 0x3FFFFFFFu) + 1ull;     // 1 .. 2^30 - 1: zeroed scratch never matches
        if (lane == 0u) atomicExch(&status[b], (epoch << 34) | ((unsigned long long)(b == 0u ? 2u : 1u) << 32) | total);
        uint32_t prefix = 0;
        if (b > 0u) {
            int j = (int)b - 1;                           // look back, 32 predecessors at a time
            for (;;) {
                const int idx = j - (int)lane;
                unsigned long long sv = 0;
                if (idx >= 0) {
                    // (the bound only keeps a broken co-residency assumption from hanging the GPU: state[1] is raised and
                    // surfaces as DEV_ERR_SCAN_STALL with the next header read-back)
                    uint32_t spins = 0;
                    do { sv = *reinterpret_cast<volatile unsigned long long*>(&status[idx]); } while ((sv >> 34) != epoch && ++spins < (1u << 20));
                    if ((sv >> 34) != epoch) { state[1] = 1u; sv = 2ull << 32; }
                } else sv = 2ull << 32;                   // before the first block: an inclusive prefix of 0
                const uint32_t kind = (uint32_t)(sv >> 32) & 3u, val = (uint32_t)sv;
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
            if (lane == 0u) atomicExch(&status[b], (epoch << 34) | (2ull << 32) | (unsigned long long)(prefix + total));
        }
        if (lane == 0u) {
            s_prefix = prefix;
            if (b == eff - 1u) {
                fin(prefix + total);
                *reinterpret_cast<volatile uint32_t*>(state) = e_raw + 1u;    // the next scan's epoch
            }
        }
    }
    __syncthreads();
    uint32_t carry = s_prefix;
    if (cached) {
        uint32_t r = carry + ex0;
#pragma unroll
        for (uint32_t k = 0; k < SDT_SCAN_ITEMS; ++k)
            if (k < per && t0 + k < hi) { emit(t0 + k, r, vals[k]); r += sdt_scan_count(vals[k]); }
    } else {
        for (uint32_t base = lo; base < hi; base += blockDim.x) {
            const uint32_t i = base + threadIdx.x;
            V v = V();
            if (i < hi) v = flag(i);
            uint32_t tot;
            const uint32_t ex = sdt_block_excl_scan(i < hi ? sdt_scan_count(v) : 0u, ws, tot);
            if (i < hi) emit(i, carry + ex, v);
            carry += tot;
        }
    }
}

// for i < n: emit(i, sum_{j<i} flag(j), flag(i)); fin(sum over all) on one thread, possibly AFTER some emits.
// flag must be pure and must not read anything emit or fin writes.
template <class Flag, class Emit, class Fin>
static inline void launch_scan(const ExecCtx& x, const uint32_t* n_ptr, uint32_t n_imm, Flag flag, Emit emit, Fin fin) {
    // co-resident by the kernel's launch bounds
    uint32_t grid = (uint32_t)x.num_sms * (uint32_t)(x.ctas_per_sm < SDT_SCAN_BLOCKS_PER_SM ? x.ctas_per_sm : SDT_SCAN_BLOCKS_PER_SM);
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
    // like the device version: flag evaluated for every item first (ranks), emit with those values afterwards, and fin
    // LAST (on the device fin runs on one block while others may already have emitted: emit must not depend on it)
    typedef decltype(flag(0u)) V;
    uint32_t run = 0;
    uint32_t* ranks = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    V* vals = (V*)malloc(sizeof(V) * (n ? n : 1));
    for (uint32_t i = 0; i < n; ++i) { vals[i] = flag(i); ranks[i] = run; run += sdt_scan_count(vals[i]); }
    for (uint32_t i = 0; i < n; ++i) emit(i, ranks[i], vals[i]);
    fin(run);
    free(ranks);
    free(vals);
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
