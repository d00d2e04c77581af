// Device layout, fp32 math and per-lane descent rules of the SD-tree hot path.
// Everything here is per-lane logic (SDT_HD); the kernels that stage shared memory
// and walk the wavefront are in sdtree.cu.  Reference citations are relative to
// /root/reference.
#pragma once

#include "sdt_platform.h"
#include "../../include/sdtree.h"

// compile-time shape of the wavefront kernels (kernel experiments: tools/build_variant.sh)
// __launch_bounds__ of k_wavefront per kind of lane: 2 CTAs per SM (one staged copy of the spatial tree each).
// The sampling kernels keep 768 threads (up to 42 registers: their descent loop wants them); the pdf / locate / splat
// kernels fit in 32 registers and run 1024 threads = all 64 warps of an SM (measured: pdf -10 %, splat -5 %).
#ifndef SDT_SAMPLE_THREADS
#define SDT_SAMPLE_THREADS 768
#endif
#ifndef SDT_QUERY_THREADS
#define SDT_QUERY_THREADS 1024
#endif
#ifndef SDT_SPLAT_THREADS
#define SDT_SPLAT_THREADS 1024
#endif
#ifndef SDT_LB_CTAS
#define SDT_LB_CTAS 2
#endif
#ifndef SDT_SAMPLE_GRID
#define SDT_SAMPLE_GRID true       // sampling kernels use the spatial grid even when the whole tree is staged (measured: sample -5 %)
#endif

#ifndef SDT_L1_HINTS
#define SDT_L1_HINTS 1             // L1 cache hints on the sampler's record loads (sdt_load_rec)
#endif

#define SDT_MAX_LEVELS 34          // quadtree levels 0..33 (QuadTree.maxDepth <= 32)
#define SDT_KD_MAX_DEPTH 40        // KDTree.maxDepth upper bound
#define SDT_KD_LEAF_BIT 0x80000000u
#define SDT_NONE 0xFFFFFFFFu

// ---------------------------------------------------------------------------
// Device layout
// ---------------------------------------------------------------------------
// Spatial binary tree: ONE 32-bit word per node (4 B per descent level):
//   interior: child_base (left = base, right = base+1; the reference appends both
//             children adjacently, src/kdtree.py:243-245)
//   leaf    : SDT_KD_LEAF_BIT | record index of its quadtree's root (0x7FFFFFFF: single-leaf
//             tree) -- the descent lands directly on the first quadtree record; the root ID
//             (= canonical node id of the root) is in kd_root[] for the callers that need it
// The split plane is never stored: children partition axis depth%3 at
// (min+max)/2 (src/kdtree.py:268-304), so the running box reproduces the
// reference's stored child boxes bit for bit.
//
// Directional quadtrees, canonical layout of clearTreeUnusedNode
// (src/quadtree.py:695-828,844-851): node ids 0..R-1 are the roots (node id ==
// root id), then level by level the four adjacent children of every non-leaf node.
// Per NON-LEAF node one 32-byte record (= one L2 sector per descent level):
struct __align__(32) QRec {
    uint32_t child_base;     // canonical node id of child_1 (children are base..base+3); bit 31: SDT_REC_IRREGULAR
    uint32_t interior_base;  // record index of the first non-leaf child
    uint32_t cinfo;          // byte c: rank of child c among the non-leaf children, 0xFF when child c is a leaf (sdt_make_cinfo)
    float own;               // this node's stored energy (pdf denominator, :1056)
    float e[4];              // the four children's stored energies (:969-972, :1057-1060)
};
// Set in child_base when the four child energies are not all finite and >= 0 (negative / NaN / infinite radiance can
// produce such sums): only then can the reference's four masked bin assignments (src/quadtree.py:983-991) overlap or
// all fail, and the sampler runs them literally.  For every other record the bins partition [0, e4) and the selected
// child is simply the number of cumulative energies <= s.  Node ids therefore stay below 2^31 (sdt_create).
#define SDT_REC_IRREGULAR 0x80000000u
#define SDT_NODE_MASK 0x7FFFFFFFu

// Device-resident description of the tree; kernels read sizes from here so that
// refine needs no host round-trip.
struct DevHeader {
    uint32_t n_kd, n_quad, n_roots, n_interior, n_levels, kd_leaves, error, refine_count;
    uint32_t root_of_node0;                 // quadTreeRootIndex[0] (out-of-box lanes, :224,:482)
    uint32_t kd_max_depth, quad_max_depth, store_nee;
    uint32_t rootrec_of_node0;              // record index of that tree's root (SDT_NONE: single-leaf tree)
    uint32_t jump_trees;                    // trees covered by the jump table (0: none)
    uint32_t s2_tables;                     // second-stage tables in use (0: none), see SDT_JUMP_TABLE
    uint32_t s2_rec_lo;                     // records [s2_rec_lo, s2_rec_hi) = the non-leaf nodes of level SDT_JUMP_LEVELS
    float bbox_min[3], bbox_max[3];         // spatial root box
    float max_leaf_size;                    // KDTree.maxLeafSize as fp32
    uint32_t s2_rec_hi;
    uint32_t level_off[SDT_MAX_LEVELS + 2];  // quadtree level l = nodes [level_off[l], level_off[l+1])
    uint32_t level_cnt[SDT_MAX_LEVELS + 2];  // level_off[l+1] - level_off[l]
    // ---- refine scratch ----
    uint32_t kd_n_old;                      // spatial nodes before this refine
    uint32_t kd_sel;                        // leaves selected in the current split round
    uint32_t kd_round_new;                  // nodes appended in the current round (2 * kd_sel << (r-1))
    uint32_t kd_round_base;                 // first node id appended in the current round
    uint32_t kd_prev_round_base;            // ... in the previous round
    uint32_t kd_round_root_base;            // first root id created in the current round
    uint32_t kd_stop;                       // capacity exhausted: later rounds select nothing
    uint32_t kd_cap, quad_cap;              // arena sizes
    uint32_t lvl_n[2];                      // nodes of the quadtree level being built (ping-pong)
    uint32_t lvl_trunc;                     // capacity exhausted: the level being built becomes all leaves
    uint32_t n_roots_old, n_quad_old;
    uint32_t pad2[2];
    uint32_t int_off[SDT_MAX_LEVELS + 2];    // non-leaf nodes in the levels above level l of the forest being built
};

enum DevError : uint32_t {
    DEV_OK = 0, DEV_ERR_KD_CAPACITY = 1, DEV_ERR_QUAD_CAPACITY = 2,
    DEV_ERR_SCAN_STALL = 4,      // a scan block gave up waiting for a lower-numbered block (see k_scan_fused): results are invalid
    DEV_ERR_KD_ROUNDS = 8        // a leaf wanted more split rounds than the host launched (sdt_hint_records bound too small)
};

// What the query kernels need (by value).
// Per-node pdf product `pp[node]`: pdfQuadTree multiplies, level by level, pdf *= (4*childE)/nodeE
// along the root->node path (src/quadtree.py:1084) and gives up with 0 when the product goes NaN
// (:1090-1092).  The path to a node is unique, so that running product is a property of the NODE:
// it is computed once per refine, top-down, with exactly those operations in exactly that order
// (NaN is sticky), and a descent only has to find its leaf: pdf = isnan(pp[leaf]) ? 0 : pp[leaf]/(4 pi).
// No division, multiply or NaN test per level in the query kernels -- and bit-identical results.
// The one case where the reference's pdf is NOT a function of the reached leaf is a point exactly
// on a split line (its first-match energy child and last-match descent child differ): those points
// take the level-by-level path.
//
// Jump table over the top SDT_JUMP_LEVELS levels of every non-single-leaf quadtree: for each of the
// SIDE x SIDE cells of [0,1]^2 the node a pdf / splat descent reaches after that many levels (its record
// index), or the leaf it meets earlier.  Points on a 1/SIDE grid line take the level-by-level path.
// (measured on the config-2 tree: 4 levels / 16x16 cells: pdf 0.273, splat 0.412 ms; 5 / 32x32: 0.257, 0.386; 6 / 64x64:
// 0.265, 0.368 at 16 KB per tree -- 5 is the default)
#ifndef SDT_JUMP_LEVELS
#define SDT_JUMP_LEVELS 5
#endif
#define SDT_JUMP_SIDE (1u << SDT_JUMP_LEVELS)                  // 32
#define SDT_JUMP_CELLS (SDT_JUMP_SIDE * SDT_JUMP_SIDE)         // 1024
#define SDT_JUMP_SIDE_F ((float)SDT_JUMP_SIDE)
#define SDT_JUMP_INV_F (1.0f / (float)SDT_JUMP_SIDE)           // exact: a power of two
#define SDT_JUMP_LEAF 0x80000000u      // entry = LEAF | node id: a leaf was reached
typedef uint32_t QJump;
// Second table of the same shape for the pdf descents, which want the leaf's path product and not its id: an entry is
// the fp32 bits of pp[leaf] (sign bit clear), or SDT_JUMP_NEXT | record to continue from.  A product that is negative
// or NaN is stored as SDT_JUMP_PP_SLOW (a NaN): those lanes take the level-by-level path.  One gather instead of two
// (table entry, then pp[leaf]) for every direction whose leaf lies in the top SDT_JUMP_LEVELS levels.
#define SDT_JUMP_NEXT 0x80000000u
#define SDT_JUMP_PP_SLOW 0x7FC00000u
// Second stage: a node of level SDT_JUMP_LEVELS that still has a non-leaf child owns an 8x8 table over its next
// SDT_S2_LEVELS = 3 levels (256 B; on the config-2 forest 23 k of the 60 k level-5 nodes, 5.9 MB per table kind).  The root
// table's continuation entry for such a node is SDT_JUMP_TABLE | table id instead of its record, so a descent that goes
// deep pays one more 4-byte gather for three levels instead of three 16 / 32-byte ones.  Entries: as in the first stage
// (leaf id / path product, or the level-8 record to continue from).  Points on a 1/256 grid line, and every point when the
// stage is switched off ("use_jump2"), continue level by level from the node's record, s2_rec[table id].
#define SDT_S2_LEVELS 3
#define SDT_S2_SIDE (1u << SDT_S2_LEVELS)
#define SDT_S2_CELLS (SDT_S2_SIDE * SDT_S2_SIDE)
#define SDT_JUMP_TABLE 0x40000000u     // in a continuation entry: the rest is a second-stage table id, not a record
#define SDT_S2_FINE_F ((float)(SDT_JUMP_SIDE * SDT_S2_SIDE))               // 256
#define SDT_S2_FINE_INV_F (1.0f / (float)(SDT_JUMP_SIDE * SDT_S2_SIDE))

struct TreeView {
    const DevHeader* hdr;
    const uint32_t* kd_word;
    const uint32_t* kd_root;    // per spatial node: quadTreeRootIndex (= canonical node id of the tree's root)
    const uint32_t* kd_grid;    // SDT_GRID_CELLS cells -> spatial node after the first 11 levels
    const QRec* rec;
    const QJump* jump;          // [root record][cell]
    const uint32_t* jump_pp;    // [root record][cell], see SDT_JUMP_NEXT
    const uint32_t* s2;         // [table][8x8] second stage of `jump` (SDT_JUMP_TABLE)
    const uint32_t* s2_pp;      // ... of `jump_pp`
    const uint32_t* s2_rec;     // [table] record of the node the table belongs to
    uint32_t use_s2;            // 0: walk level by level from s2_rec instead
    const float* pp;            // per node: pdf product of the root->node path (NaN: it went NaN)
    uint32_t jump_trees;        // 0: table not in use
    uint32_t int_cell;          // sampler's cell tracking (sdt_quad_sample CELL): 1 = depth <= 16, 2 = depth <= 23, 0 = deeper
};

// ---------------------------------------------------------------------------
// fp32 math: every operation separately rounded (nvcc -fmad=false, IEEE div / sqrt;
// g++ -ffp-contract=off for the host emulation).  Same operation order as
// oracle/drjit_math.py, which restates Dr.Jit's CEPHES-based sincos / atan2
// (third-party, absent from the reference).
// ---------------------------------------------------------------------------
#define SDT_PI 3.14159265358979323846f
#define SDT_TWO_PI 6.28318530717958647692f
#define SDT_HALF_PI 1.57079632679489661923f
#define SDT_QUARTER_PI 0.78539816339744830962f
#define SDT_INV_FOUR_PI 0.07957747154594766788f

SDT_HD uint32_t sdt_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
SDT_HD float sdt_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

SDT_HD void sdt_sincos(float x, float& s, float& c) {
    const float xa = fabsf(x);
    uint32_t j = (uint32_t)(xa * 1.27323954473516f);
    j = (j + 1u) & 0xFFFFFFFEu;
    const float y = (float)j;
    const float xr = ((xa - y * 0.78515625f) - y * 2.4187564849853515625e-4f) - y * 3.77489497744594108e-8f;
    const float z = xr * xr;
    const float ps = (((-1.9515295891e-4f * z + 8.3321608736e-3f) * z + -1.6666654611e-1f) * z) * xr + xr;
    const float pc = ((((2.443315711809948e-5f * z + -1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z) * z - 0.5f * z) + 1.0f;
    const bool swap = (j & 2u) != 0u;
    s = swap ? pc : ps;
    c = swap ? ps : pc;
    const bool neg_s = ((j & 4u) != 0u) != (x < 0.0f);
    const bool neg_c = ((j + 2u) & 4u) != 0u;
    s = neg_s ? -s : s;
    c = neg_c ? -c : c;
}

SDT_HD float sdt_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mn = fminf(ax, ay), mx = fmaxf(ax, ay);
    const float a = mn / mx;
    const bool big = a > 0.4142135623730950f;
    const float t = big ? (a - 1.0f) / (a + 1.0f) : a;
    const float base = big ? SDT_QUARTER_PI : 0.0f;
    const float z = t * t;
    const float p = ((((8.05374449538e-2f * z + -1.38776856032e-1f) * z + 1.99777106478e-1f) * z + -3.33329491539e-1f) * z) * t + t;
    float r = base + p;
    r = (ay > ax) ? SDT_HALF_PI - r : r;
    r = (x < 0.0f) ? SDT_PI - r : r;
    r = (y < 0.0f) ? -r : r;
    r = (mx == 0.0f) ? 0.0f : r;
    return r;
}

// canonicalToDir, src/common.py:100-129
SDT_HD void sdt_canonical_to_dir(float px, float py, float& dx, float& dy, float& dz) {
    const float cos_theta = 2.0f * py - 1.0f;
    const float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    const float phi = SDT_TWO_PI * px;
    float sp, cp;
    sdt_sincos(phi, sp, cp);
    dx = sin_theta * cp;
    dy = sin_theta * sp;
    dz = cos_theta;
}

// dirToCanonical, src/common.py:132-158
SDT_HD void sdt_dir_to_canonical(float dx, float dy, float dz, float& px, float& py) {
    const float cos_theta = fminf(fmaxf(dz, -1.0f), 1.0f);
    float phi = sdt_atan2(dy, dx);
    while (phi < 0.0f) phi += SDT_TWO_PI;     // loop "rotate phi"
    px = phi / SDT_TWO_PI;
    py = (cos_theta + 1.0f) / 2.0f;
    // isfinite on all three components, else (0,0)
    const bool fin = (fabsf(dx) <= 3.402823466e38f) && (fabsf(dy) <= 3.402823466e38f) && (fabsf(dz) <= 3.402823466e38f);
    if (!fin) { px = 0.0f; py = 0.0f; }
}

// mi.luminance(Color3f): Rec.709 weights (Mitsuba constant)
SDT_HD float sdt_luminance(float r, float g, float b) {
    return (r * 0.212671f + g * 0.715160f) + b * 0.072169f;
}

// Perf-mode uniforms in [0,1), keyed (seed, lane, idx) and restated identically in the oracle
// (counter_uniform).  The lane key h0 is a murmur3 finaliser of (seed, lane).  idx = 3*level + {0,1,2}
// as the reference consumes them (src/quadtree.py:956,980):
//   idx % 3 == 2 (the child-selection uniform, one per visited level -- the hot one): the top 24 bits of
//                a 32-bit LCG stream started at h0, t_{level+1} = t_level * M + C, so that a descent
//                pays one multiply-add per level instead of a hash;
//   idx % 3 != 2 (the leaf position, used once per sample): a second finaliser round of h0 ^ f(idx).
// M, C: the 32-bit multiplier / an odd increment of the PCG family (full period 2^32).
#define SDT_LCG_M 747796405u
#define SDT_LCG_C 2891336453u
SDT_HD uint32_t sdt_fmix(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
SDT_HD float sdt_u24(uint32_t h) { return (float)(h >> 8) * 5.9604644775390625e-08f; }

// Uniform stream of one lane.  select(level) must be called for level = 0, 1, 2, ... in turn (a descent
// does); pos(idx) is random access.  ExplicitRng: u[lane*stride + idx], clamped like the oracle's
// ExplicitSampler.
struct CounterRng {
    static constexpr bool kUnitInterval = true;      // every uniform is in [0, 1)
    uint32_t h0, t;
    SDT_HD CounterRng(uint32_t seed, uint32_t lane_id) : h0(sdt_fmix(seed + lane_id * 0x9E3779B1u)), t(h0) {}
    SDT_HD float select(uint32_t) { t = t * SDT_LCG_M + SDT_LCG_C; return sdt_u24(t); }
    SDT_HD float pos(uint32_t idx) const { return sdt_u24(sdt_fmix(h0 ^ (idx * 0x85EBCA77u + 0x165667B1u))); }
};
struct ExplicitRng {
    static constexpr bool kUnitInterval = false;     // caller-provided numbers: may be anything, NaN included
    const float* row; uint32_t u_stride;
    SDT_HD ExplicitRng(const float* u, uint32_t stride, uint32_t lane_index) : row(u + (size_t)lane_index * stride), u_stride(stride) {}
    SDT_HD float pos(uint32_t idx) const { return SDT_LDG(row + (idx < u_stride ? idx : u_stride - 1u)); }
    SDT_HD float select(uint32_t level) const { return pos(3u * level + 2u); }
};

SDT_HD float sdt_ld(const float* p, int64_t stride, uint32_t i) { return SDT_LDG(p + (int64_t)i * stride); }

// ---- spatial descent: KDTree.getLeafNodeIndex, src/kdtree.py:435-470 --------
// Returns the leaf node id (0 for lanes outside the root box, like the reference).
// `kd` is the smem-staged prefix of kd_word (n_smem words), `kdg` the full array;
// ALL_SMEM: the whole tree is staged (no range check, no global path in the loop).
struct KdResult { uint32_t leaf; uint32_t rootrec; bool inbox; };   // rootrec: record index of the quadtree root or SDT_NONE

template <bool ALL_SMEM>
SDT_HD uint32_t sdt_kd_load(const uint32_t* __restrict__ kd, uint32_t n_smem, const uint32_t* __restrict__ kdg, uint32_t node) {
    if (ALL_SMEM) return kd[node];
    return (node < n_smem) ? kd[node] : SDT_LDG(kdg + node);
}

// one level on one axis: children are left [lo, mid] and right [mid, hi]; the reference assigns
// left first, right second, so the right child wins on the plane (:462-468)
#define SDT_KD_STEP(P, LO, HI)                                              \
    {                                                                       \
        const float mid = (LO + HI) / 2.0f;                                 \
        const bool right = P >= mid;                                        \
        LO = right ? mid : LO;                                              \
        HI = right ? HI : mid;                                              \
        node = w + (right ? 1u : 0u);                                       \
        w = sdt_kd_load<ALL_SMEM>(kd, n_smem, kdg, node);                   \
        if (w & SDT_KD_LEAF_BIT) break;                                     \
    }

// The first SDT_GRID_LEVELS = 11 levels split x,y,z,x,y,z,x,y,z,x,y: 4 halvings of x, 4 of y, 3 of z.
// The split planes of one axis do not depend on the other axes, so the three chains of exact fp32
// midpoints run independently (no memory access, ILP 3) and give the cell of a 16x16x8 grid; the
// grid (built per CTA from the staged tree) names the node the level-by-level walk would have
// reached -- a leaf shallower than 11 levels fills all its cells.
// (measured, config-2 tree: 12 levels / 16 KB: pdf -1.5 %, splat -0.5 %, guided +2 %; 13 levels / 32 KB: splat +33 % --
// the shared-memory footprint eats the L1; 11 stays)
#ifndef SDT_GRID_LEVELS
#define SDT_GRID_LEVELS 11
#endif
#define SDT_GRID_NX ((SDT_GRID_LEVELS + 2) / 3)       // halvings of x, y, z among the first SDT_GRID_LEVELS levels
#define SDT_GRID_NY ((SDT_GRID_LEVELS + 1) / 3)
#define SDT_GRID_NZ (SDT_GRID_LEVELS / 3)
#define SDT_GRID_CELLS (1u << SDT_GRID_LEVELS)
#define SDT_GRID_CELL(cx, cy, cz) (((cx) << (SDT_GRID_NY + SDT_GRID_NZ)) | ((cy) << SDT_GRID_NZ) | (cz))
// per-thread constants of the spatial descent (loaded once per thread, not per vertex)
struct KdCtx {
    const uint32_t* kd;       // smem-staged prefix of kd_word
    uint32_t n_smem;
    const uint32_t* kdg;      // full array
    float lo[3], hi[3];       // root box
    uint32_t rootrec0;        // root record of the tree owned by node 0 (out-of-box lanes, :224,:482)
    uint32_t* cnt_s;          // splat kernels: per-CTA shared-memory leaf counters (NULL: count in global memory)
    uint32_t aggregate;       // splat kernels: combine lanes of a warp that hit the same address before the atomic
    const uint32_t* grid;     // 16x16x8 cell -> node reached after the first 11 levels (NULL: descend from the root)
};
SDT_HD KdCtx sdt_kd_ctx(const uint32_t* kd, uint32_t n_smem, const uint32_t* kdg, const DevHeader* hdr) {
    KdCtx k;
    k.kd = kd; k.n_smem = n_smem; k.kdg = kdg;
    for (int a = 0; a < 3; ++a) { k.lo[a] = hdr->bbox_min[a]; k.hi[a] = hdr->bbox_max[a]; }
    k.rootrec0 = hdr->rootrec_of_node0;
    k.cnt_s = nullptr;
    k.aggregate = 0u;
    k.grid = nullptr;
    return k;
}

#define SDT_AXIS_STEP(P, LO, HI, C)                                         \
    {                                                                       \
        const float mid = (LO + HI) / 2.0f;                                 \
        const bool right = P >= mid;                                        \
        LO = right ? mid : LO;                                              \
        HI = right ? HI : mid;                                              \
        C = (C << 1) | (right ? 1u : 0u);                                   \
    }
// (tried in round 2: the cell from a scaled guess verified against a shared-memory table of the exact cell boundaries
// instead of the 11 chained halvings -- fewer instructions, but the dependent LDS + F2I chain and the rare fix-up
// branches made all three kernels 3 % slower; the arithmetic chains stay)
// node reached from the root by the path bits of cell (cx NX bits, cy NY bits, cz NZ bits)
SDT_HD uint32_t sdt_kd_grid_node(const uint32_t* __restrict__ kd, uint32_t cell) {
    const uint32_t cx = cell >> (SDT_GRID_NY + SDT_GRID_NZ), cy = (cell >> SDT_GRID_NZ) & ((1u << SDT_GRID_NY) - 1u), cz = cell & ((1u << SDT_GRID_NZ) - 1u);
    uint32_t node = 0;
    for (uint32_t l = 0; l < SDT_GRID_LEVELS; ++l) {
        const uint32_t w = kd[node];
        if (w & SDT_KD_LEAF_BIT) break;
        const uint32_t a = l % 3u, j = l / 3u;
        const uint32_t bit = a == 0u ? (cx >> (SDT_GRID_NX - 1u - j)) & 1u : (a == 1u ? (cy >> (SDT_GRID_NY - 1u - j)) & 1u : (cz >> (SDT_GRID_NZ - 1u - j)) & 1u);
        node = w + bit;
    }
    return node;
}

// MODE 0: partly staged tree (global loads beyond the prefix), walk from the root; 1: whole tree
// staged, walk from the root; 2: whole tree staged + grid over the first 11 levels; 3: grid + partly
// staged tree (large trees: 11 levels by arithmetic, the few levels below through L1/L2)
template <int MODE>
SDT_HD KdResult sdt_kd_descend(const KdCtx& k, float px, float py, float pz) {
    constexpr bool ALL_SMEM = MODE == 1 || MODE == 2;
    const uint32_t* __restrict__ kd = k.kd;
    const uint32_t* __restrict__ kdg = k.kdg;
    const uint32_t n_smem = k.n_smem;
    KdResult r;
    float lo0 = k.lo[0], lo1 = k.lo[1], lo2 = k.lo[2];
    float hi0 = k.hi[0], hi1 = k.hi[1], hi2 = k.hi[2];
    // BoundingBox3f.contains: inclusive; NaN fails every comparison
    r.inbox = (px >= lo0) && (px <= hi0) && (py >= lo1) && (py <= hi1) && (pz >= lo2) && (pz <= hi2);
    r.leaf = 0;
    r.rootrec = k.rootrec0;
    if (!r.inbox) return r;
    uint32_t node = 0;
    uint32_t w;
    if (MODE >= 2) {
        uint32_t cx = 0, cy = 0, cz = 0;
#pragma unroll
        for (int l = 0; l < SDT_GRID_LEVELS; ++l) {          // fully unrolled: three independent chains
            if (l % 3 == 0) SDT_AXIS_STEP(px, lo0, hi0, cx)
            else if (l % 3 == 1) SDT_AXIS_STEP(py, lo1, hi1, cy)
            else SDT_AXIS_STEP(pz, lo2, hi2, cz)
        }
        node = k.grid[SDT_GRID_CELL(cx, cy, cz)];
        w = sdt_kd_load<ALL_SMEM>(kd, n_smem, kdg, node);
        if (!(w & SDT_KD_LEAF_BIT)) {
            for (int guard = SDT_GRID_LEVELS; guard < SDT_KD_MAX_DEPTH; guard += 3) {   // level SDT_GRID_LEVELS splits axis LEVELS % 3, ...
#if SDT_GRID_LEVELS % 3 == 2
                SDT_KD_STEP(pz, lo2, hi2)
                SDT_KD_STEP(px, lo0, hi0)
                SDT_KD_STEP(py, lo1, hi1)
#elif SDT_GRID_LEVELS % 3 == 0
                SDT_KD_STEP(px, lo0, hi0)
                SDT_KD_STEP(py, lo1, hi1)
                SDT_KD_STEP(pz, lo2, hi2)
#else
                SDT_KD_STEP(py, lo1, hi1)
                SDT_KD_STEP(pz, lo2, hi2)
                SDT_KD_STEP(px, lo0, hi0)
#endif
            }
        }
    } else {
        w = sdt_kd_load<ALL_SMEM>(kd, n_smem, kdg, 0u);
        if (!(w & SDT_KD_LEAF_BIT)) {
            for (int guard = 0; guard < SDT_KD_MAX_DEPTH; guard += 3) {   // axis = depth % 3 (:277)
                SDT_KD_STEP(px, lo0, hi0)
                SDT_KD_STEP(py, lo1, hi1)
                SDT_KD_STEP(pz, lo2, hi2)
            }
        }
    }
    r.leaf = node;
    r.rootrec = w & ~SDT_KD_LEAF_BIT;
    if (r.rootrec == 0x7FFFFFFFu) r.rootrec = SDT_NONE;
    return r;
}

struct __align__(16) SdtF4 { float x, y, z, w; };

// record words 0..3 {child_base, interior_base, cinfo, own}
struct QHead { uint32_t child_base, interior_base, cinfo; float own; };

SDT_HD QHead sdt_load_head(const QRec* __restrict__ rec, uint32_t i) {
    QHead h;
#if defined(__CUDA_ARCH__)
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(rec + i));
    h.child_base = a.x; h.interior_base = a.y; h.cinfo = a.z; h.own = __uint_as_float(a.w);
#else
    h.child_base = rec[i].child_base; h.interior_base = rec[i].interior_base; h.cinfo = rec[i].cinfo; h.own = rec[i].own;
#endif
    return h;
}
// the whole 32-byte record with ONE load instruction (sm_100 256-bit global load): for a
// divergent gather the L1 tag stage is charged per instruction and lane, so one LDG.256
// costs half of two LDG.128 on the same sector.
// HINT: 0 = plain, 1 = keep in L1 (evict last), 2 = do not allocate in L1.  The sampler loads the ROOT record of a
// quadtree with 1 and every deeper record with 2: the root records of the forest (32 B x trees = 131 KB on the config-2
// tree) are the only level small enough to live in an SM's L1 next to the staged spatial tree, the deeper ones are
// never re-used before they are evicted and only push the roots out.  Every L1 hit is one slot less on the L1 -> L2
// request port the sampling kernels are bound by (measured: sample 0.651 -> 0.623 ms, fused query 0.745 -> 0.724 ms;
// also tried: two levels kept, the leaf's path product / the jump tables / the pdf and splat record loads not
// allocated -- each slower, the last by 7 % on the splat, which does re-use them).
template <int HINT = 0>
SDT_HD void sdt_load_rec(const QRec* __restrict__ rec, uint32_t i, QHead& h, SdtF4& e) {
#if defined(__CUDA_ARCH__)
    uint32_t a, b, c, d;
    if (HINT == 1 && SDT_L1_HINTS)
        asm("ld.global.nc.L1::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=f"(e.x), "=f"(e.y), "=f"(e.z), "=f"(e.w) : "l"(rec + i));
    else if (HINT == 2 && SDT_L1_HINTS)
        asm("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=f"(e.x), "=f"(e.y), "=f"(e.z), "=f"(e.w) : "l"(rec + i));
    else
        asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=f"(e.x), "=f"(e.y), "=f"(e.z), "=f"(e.w) : "l"(rec + i));
    h.child_base = a; h.interior_base = b; h.cinfo = c; h.own = __uint_as_float(d);
#else
    h.child_base = rec[i].child_base; h.interior_base = rec[i].interior_base; h.cinfo = rec[i].cinfo; h.own = rec[i].own;
    e.x = rec[i].e[0]; e.y = rec[i].e[1]; e.z = rec[i].e[2]; e.w = rec[i].e[3];
#endif
}
// cinfo: byte c = rank of child c among the non-leaf children (its record is interior_base + rank),
// or 0xFF when child c is a leaf
SDT_HD uint32_t sdt_make_cinfo(uint32_t leafmask) {
    uint32_t ci = 0, run = 0;
    for (uint32_t c = 0; c < 4u; ++c) {
        if ((leafmask >> c) & 1u) ci |= 0xFFu << (8u * c);
        else { ci |= run << (8u * c); ++run; }
    }
    return ci;
}
// byte c of cinfo, sign-extended: the rank (0..3) of a non-leaf child, -1 for a leaf
SDT_HD int32_t sdt_child_code(uint32_t cinfo, uint32_t c) {
    return (int32_t)(int8_t)(uint8_t)(cinfo >> (8u * c));
}
// the same with the child given as a byte-permute selector SDT_SEL(c): byte 0 <- byte c, bytes 1..3 <- its sign
// (selector nibbles c, 8|c, 8|c, 8|c) -- one PRMT, and the sampler's comparison chain yields the selector directly
#define SDT_SEL(c) ((c) * 0x1111u + 0x8880u)
SDT_HD int32_t sdt_child_code_sel(uint32_t cinfo, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(d) : "r"(cinfo), "r"(sel));
    return (int32_t)d;
#else
    return sdt_child_code(cinfo, sel & 3u);
#endif
}
// record index of child c, or SDT_NONE when it is a leaf
SDT_HD uint32_t sdt_child_rec(uint32_t cinfo, uint32_t interior_base, uint32_t c) {
    const int32_t code = sdt_child_code(cinfo, c);
    return code < 0 ? SDT_NONE : interior_base + (uint32_t)code;
}
// flag of a record whose child energies need the literal bin assignments (see SDT_REC_IRREGULAR)
SDT_HD uint32_t sdt_rec_flags(float e0, float e1, float e2, float e3) {
    const float big = 3.402823466e38f;
    const bool ok = (e0 >= 0.0f) && (e1 >= 0.0f) && (e2 >= 0.0f) && (e3 >= 0.0f) &&
                    ((((e0 + e1) + e2) + e3) <= big);
    return ok ? 0u : SDT_REC_IRREGULAR;
}

// quadrant c (0..3 = child_1..child_4) of [lo,hi], src/quadtree.py:153-175:
// c1 = [mid,max] upper-right, c2 upper-left, c3 = [min,mid] lower-left, c4 lower-right
SDT_HD void sdt_quadrant_m(uint32_t c, float mx, float my, float& lox, float& loy, float& hix, float& hiy) {
    const bool right = (c == 0u) || (c == 3u);
    const bool top = c < 2u;
    lox = right ? mx : lox;
    hix = right ? hix : mx;
    loy = top ? my : loy;
    hiy = top ? hiy : my;
}
SDT_HD void sdt_quadrant(int c, float& lox, float& loy, float& hix, float& hiy) {
    sdt_quadrant_m((uint32_t)c, (lox + hix) / 2.0f, (loy + hiy) / 2.0f, lox, loy, hix, hiy);
}

// child that the DESCENT follows for a point of the node box: the reference assigns
// c1..c4 in turn, so the LAST matching (inclusive) child box wins
// (src/quadtree.py:1095-1098 for pdf, :424-438 for splat)
SDT_HD uint32_t sdt_descend_child(float x, float y, float mx, float my) {
    return (y <= my) ? ((x >= mx) ? 3u : 2u) : ((x <= mx) ? 1u : 0u);
}
// child whose ENERGY enters the pdf: FIRST matching child box (src/quadtree.py:1063-1075)
SDT_HD uint32_t sdt_energy_child(float x, float y, float mx, float my) {
    return (y >= my) ? ((x >= mx) ? 0u : 1u) : ((x <= mx) ? 2u : 3u);
}
SDT_HD float sdt_pick4(const SdtF4& v, uint32_t c) {
    const float a = (c & 1u) ? v.y : v.x;
    const float b = (c & 1u) ? v.w : v.z;
    return (c & 2u) ? b : a;
}

// cell of a canonical position; false when it lies on a grid line or outside [0,1)
SDT_HD bool sdt_jump_cell(float x, float y, uint32_t& cx, uint32_t& cy) {
    const float fx = x * SDT_JUMP_SIDE_F, fy = y * SDT_JUMP_SIDE_F;          // exact scalings
    if (!(fx >= 0.0f && fx < SDT_JUMP_SIDE_F && fy >= 0.0f && fy < SDT_JUMP_SIDE_F)) return false;
    cx = (uint32_t)fx; cy = (uint32_t)fy;
    return fx != (float)cx && fy != (float)cy;
}

// QuadTree.pdfQuadTree, src/quadtree.py:1001-1101, level by level with the reference's own
// arithmetic: the path for points on split lines, and for naming the node a NaN stop happens at.
SDT_HD float sdt_quad_pdf_levels(const QRec* __restrict__ rec, uint32_t ri, uint32_t root_node,
                                 float x, float y, uint32_t& node_out) {
    float pdf = 1.0f;
    float lox = 0.0f, loy = 0.0f, hix = 1.0f, hiy = 1.0f;
    uint32_t node = root_node;
    bool dead = false;
    for (int level = 0; level < SDT_MAX_LEVELS; ++level) {
        if (ri == SDT_NONE) break;
        QHead h; SdtF4 e;
        sdt_load_rec(rec, ri, h, e);
        const float mx = (lox + hix) / 2.0f, my = (loy + hiy) / 2.0f;
        const uint32_t ce = sdt_energy_child(x, y, mx, my);
        const uint32_t cd = sdt_descend_child(x, y, mx, my);
        pdf = pdf * ((4.0f * sdt_pick4(e, ce)) / h.own);                // :1084
        if (pdf != pdf) { dead = true; break; }                         // :1090-1092
        node = h.child_base + cd;
        sdt_quadrant_m(cd, mx, my, lox, loy, hix, hiy);
        ri = sdt_child_rec(h.cinfo, h.interior_base, cd);
    }
    pdf = dead ? 0.0f : pdf * SDT_INV_FOUR_PI;                          // :1030
    node_out = node & SDT_NODE_MASK;
    return pdf;
}

// Follows a SDT_JUMP_TABLE continuation of the first stage.  In: table id.  Out: `entry` = the second-stage entry when
// the point lies strictly inside a 1/256 cell (then cx8 / cy8 name it), else false with ri = the node's own record.
SDT_HD bool sdt_s2_lookup(const TreeView& t, const uint32_t* __restrict__ tab, uint32_t tid, float x, float y,
                          uint32_t& entry, uint32_t& cx8, uint32_t& cy8, uint32_t& ri) {
    const float fx = x * SDT_S2_FINE_F, fy = y * SDT_S2_FINE_F;             // exact scalings; the point is inside [0,1)^2 here
    cx8 = (uint32_t)fx; cy8 = (uint32_t)fy;
    if (t.use_s2 && fx != (float)cx8 && fy != (float)cy8) {
        entry = SDT_LDG(tab + (size_t)tid * SDT_S2_CELLS + (cy8 & (SDT_S2_SIDE - 1u)) * SDT_S2_SIDE + (cx8 & (SDT_S2_SIDE - 1u)));
        return true;
    }
    ri = SDT_LDG(t.s2_rec + tid);
    return false;
}

// pdfQuadTree from canonical position (x,y) in [0,1]^2.  ri = record index of the root (SDT_NONE:
// single-leaf tree).  Finds the leaf (jump table, then 16 B of each record per level) and reads its
// path product; split-line points and NaN stops go through sdt_quad_pdf_levels.
SDT_HD float sdt_quad_pdf(const TreeView& t, uint32_t ri, uint32_t root_node,
                          float x, float y, uint32_t& node_out, bool want_node = true) {
    if (ri == SDT_NONE) { node_out = root_node; return SDT_INV_FOUR_PI; }     // 1 * 1/(4 pi)
    const QRec* __restrict__ rec = t.rec;
    const uint32_t ri0 = ri;
    float lox = 0.0f, loy = 0.0f, hix = 1.0f, hiy = 1.0f;
    uint32_t node = root_node;
    int level = 0;
    bool tie = false;
    uint32_t cx, cy;
    if (ri < t.jump_trees && sdt_jump_cell(x, y, cx, cy)) {
        uint32_t cont;                           // continuation: a record, or SDT_JUMP_TABLE | table id
        if (!want_node) {
            // the caller only wants the pdf: the table holds the leaf's path product itself
            const uint32_t j = SDT_LDG(t.jump_pp + (size_t)ri * SDT_JUMP_CELLS + cy * SDT_JUMP_SIDE + cx);
            if (!(j & SDT_JUMP_NEXT)) {
                const float pp = sdt_u2f(j);
                node_out = 0u;
                if (pp == pp) return pp * SDT_INV_FOUR_PI;
                return sdt_quad_pdf_levels(rec, ri0, root_node, x, y, node_out);
            }
            cont = j & ~SDT_JUMP_NEXT;
        } else {
            const QJump j = SDT_LDG(t.jump + (size_t)ri * SDT_JUMP_CELLS + cy * SDT_JUMP_SIDE + cx);
            if (j & SDT_JUMP_LEAF) { node = j & ~SDT_JUMP_LEAF; cont = SDT_NONE; }
            else cont = j;
        }
        ri = cont;
        float inv = SDT_JUMP_INV_F;
        level = SDT_JUMP_LEVELS;
        if (cont != SDT_NONE && (cont & SDT_JUMP_TABLE)) {
            uint32_t e, cx8, cy8;
            if (sdt_s2_lookup(t, want_node ? t.s2 : t.s2_pp, cont & ~SDT_JUMP_TABLE, x, y, e, cx8, cy8, ri)) {
                if (!want_node) {
                    if (!(e & SDT_JUMP_NEXT)) {
                        const float pp = sdt_u2f(e);
                        node_out = 0u;
                        if (pp == pp) return pp * SDT_INV_FOUR_PI;
                        return sdt_quad_pdf_levels(rec, ri0, root_node, x, y, node_out);
                    }
                    ri = e & ~SDT_JUMP_NEXT;
                } else if (e & SDT_JUMP_LEAF) { node = e & ~SDT_JUMP_LEAF; ri = SDT_NONE; }
                else ri = e;
                cx = cx8; cy = cy8; inv = SDT_S2_FINE_INV_F;
                level = SDT_JUMP_LEVELS + SDT_S2_LEVELS;
            }
        }
        if (ri != SDT_NONE) {
            lox = (float)cx * inv; hix = (float)(cx + 1u) * inv;
            loy = (float)cy * inv; hiy = (float)(cy + 1u) * inv;
        } else level = 0;
    }
    for (; level < SDT_MAX_LEVELS && ri != SDT_NONE; ++level) {
        const QHead h = sdt_load_head(rec, ri);
        const float mx = (lox + hix) / 2.0f, my = (loy + hiy) / 2.0f;
        tie = tie || (x == mx) || (y == my);
        const uint32_t cd = sdt_descend_child(x, y, mx, my);
        node = h.child_base + cd;
        sdt_quadrant_m(cd, mx, my, lox, loy, hix, hiy);
        ri = sdt_child_rec(h.cinfo, h.interior_base, cd);
    }
    node &= SDT_NODE_MASK;
    const float pp = SDT_LDG(t.pp + node);
    if (tie || pp != pp) return sdt_quad_pdf_levels(rec, ri0, root_node, x, y, node_out);
    node_out = node;
    return pp * SDT_INV_FOUR_PI;
}

// Jump-table entries: where the descent arrives for any interior point of a cell (strict interior: every tie rule agrees).
// the four entries of the 2 x 2 block of cells (2qx + bx, 2qy + by), out[by * 2 + bx]: the block shares every level of the
// descent but the last, so a table costs a quarter of the dependent loads of one descent per cell
SDT_HD void sdt_build_jump4(const QRec* __restrict__ rec, uint32_t root_rec, uint32_t qx, uint32_t qy, int levels, QJump out[4]) {
    uint32_t ri = root_rec;
    for (int l = 0; l + 1 < levels; ++l) {
        const QRec r = rec[ri];
        const uint32_t bx = (qx >> (levels - 2 - l)) & 1u, by = (qy >> (levels - 2 - l)) & 1u;
        const uint32_t c = by ? (bx ? 0u : 1u) : (bx ? 3u : 2u);
        ri = sdt_child_rec(r.cinfo, r.interior_base, c);
        if (ri == SDT_NONE) {
            out[0] = out[1] = out[2] = out[3] = SDT_JUMP_LEAF | ((r.child_base & SDT_NODE_MASK) + c);
            return;
        }
    }
    const QRec r = rec[ri];
    for (uint32_t k = 0; k < 4u; ++k) {
        const uint32_t bx = k & 1u, by = k >> 1;
        const uint32_t c = by ? (bx ? 0u : 1u) : (bx ? 3u : 2u);
        const uint32_t nri = sdt_child_rec(r.cinfo, r.interior_base, c);
        out[k] = nri == SDT_NONE ? (SDT_JUMP_LEAF | ((r.child_base & SDT_NODE_MASK) + c)) : nri;
    }
}

// the pdf descents' entry for the same cell (see SDT_JUMP_NEXT)
SDT_HD uint32_t sdt_jump_pp_entry(QJump j, const float* __restrict__ pp) {
    if (!(j & SDT_JUMP_LEAF)) return SDT_JUMP_NEXT | j;
    const uint32_t b = sdt_f2u(pp[j & ~SDT_JUMP_LEAF]);
    const float v = sdt_u2f(b);
    return ((b & 0x80000000u) || v != v) ? SDT_JUMP_PP_SLOW : b;
}

// QuadTree.sampleQuadTree, src/quadtree.py:931-998.  Consumes 3 uniforms per visited
// node: (u_x, u_y) at index 3*level, 3*level+1 (used at the leaf only, :956-962) and
// u_select at 3*level+2 (:980).  While descending it also accumulates the pdf of the
// path, with exactly the operations of pdfQuadTree, so that KDTree.sample's second
// descent (src/kdtree.py:483-484) can be skipped whenever the canonical->dir->canonical
// round trip stays strictly inside the sampled leaf cell (then both descents take the
// same path and every tie rule is moot).
struct QSample {
    float x, y;                 // canonical position
    uint32_t node;              // leaf node reached
    bool moved;                 // the tree has a record (node is a real node id)
    bool stuck;                 // NaN child energies: no bin matched (oracle: stop at (0,0))
    float lox, loy, hix, hiy;   // leaf cell
};

// CELL: how the sampler knows the leaf cell it ended in.
//   0  four floats halved per level with the reference's own (min+max)/2 (any depth; QuadTree.maxDepth > 23)
//   1  the path as 2 bits per level in 32 bits (trees of at most 16 levels below the root)
//   2  the same in 64 bits (at most 23 levels)
// For 1 and 2 the cell (ix, iy) / 2^level is rebuilt once at the leaf, where the warp has reconverged.  Cell corners
// k/2^level are exact in fp32 up to level 23, where the reference's repeated (min+max)/2 is exact too, so the box is
// bit-identical.  child_1 / child_4 are the right half, child_1 / child_2 the upper half (:153-175): for the child
// number cu = b1 b0 the x bit is ~(b1 ^ b0) and the y bit ~b1.
SDT_HD uint32_t sdt_even_bits(uint32_t x) {          // bits 0,2,4,.. of x packed into the low half
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    x = (x | (x >> 8)) & 0x0000FFFFu;
    return x;
}
SDT_HD void sdt_path_cell(uint32_t path, uint32_t level, uint32_t& ix, uint32_t& iy) {
    const uint32_t odd = path >> 1;
    const uint32_t m = level >= 32u ? 0xFFFFFFFFu : (1u << level) - 1u;
    ix = sdt_even_bits(~(odd ^ path)) & m;
    iy = sdt_even_bits(~odd) & m;
}
SDT_HD void sdt_path_cell64(uint64_t path, uint32_t level, uint32_t& ix, uint32_t& iy) {
    uint32_t xl, yl, xh, yh;
    sdt_path_cell((uint32_t)path, 16u, xl, yl);
    sdt_path_cell((uint32_t)(path >> 32), 16u, xh, yh);
    const uint32_t m = level >= 32u ? 0xFFFFFFFFu : (1u << level) - 1u;
    ix = ((xh << 16) | xl) & m;
    iy = ((yh << 16) | yl) & m;
}

template <class Rng, int CELL>
SDT_HD QSample sdt_quad_sample(const QRec* __restrict__ rec, uint32_t ri, uint32_t root_node, Rng rng) {
    QSample q;
    q.x = 0.0f; q.y = 0.0f; q.node = root_node; q.moved = ri != SDT_NONE; q.stuck = false;
    q.lox = 0.0f; q.loy = 0.0f; q.hix = 1.0f; q.hiy = 1.0f;
    uint32_t level = 0, path32 = 0;
    uint64_t path64 = 0;
#define SDT_SAMPLE_STEP(CU, SEL)                                                                                  \
    {                                                                                                         \
        q.node = h.child_base + (CU);                                                                         \
        if (CELL == 1) path32 = (path32 << 2) | (CU);                                                         \
        else if (CELL == 2) path64 = (path64 << 2) | (CU);                                                    \
        else sdt_quadrant_m((CU), (q.lox + q.hix) / 2.0f, (q.loy + q.hiy) / 2.0f, q.lox, q.loy, q.hix, q.hiy); \
        ++level;                                                                                              \
        const int32_t code = sdt_child_code_sel(h.cinfo, (SEL));                                              \
        if (code < 0) break;                                                                                  \
        ri = h.interior_base + (uint32_t)code;                                                                \
        if (level >= SDT_MAX_LEVELS) break;                          /* impossible for a valid tree */         \
    }
    // the loops only descend; the leaf position is drawn after them, where the warp has reconverged
    if (ri != SDT_NONE) {
        // fast loop: records whose child energies are finite and >= 0 -- the bins partition [0, e4), e1 <= e2 <= e3,
        // and the child is the number of cumulative energies <= s.  The first record that is not (or a NaN s: an
        // explicit uniform can be anything) hands the rest of the descent to the literal loop below.
        float u_pending = 0.0f;
        bool literal = false;
        for (;;) {
            QHead h; SdtF4 e;
            if (level == 0u) sdt_load_rec<1>(rec, ri, h, e);             // the root record: keep it in L1
            else sdt_load_rec<2>(rec, ri, h, e);                         // deeper records: do not displace the roots
            const float e1 = e.x;
            const float e2 = e.y + e1;                                   // :975-977
            const float e3 = e.z + e2;
            const float e4 = e.w + e3;
            const float u = rng.select(level);
            const float s = u * e4;                                      // :980
            if (((int32_t)h.child_base < 0) || (!Rng::kUnitInterval && (s != s))) { literal = true; u_pending = u; break; }
            const uint32_t sel = (s >= e3) ? SDT_SEL(3u) : ((s >= e2) ? SDT_SEL(2u) : ((s >= e1) ? SDT_SEL(1u) : SDT_SEL(0u)));
            const uint32_t cu = sel & 3u;
            SDT_SAMPLE_STEP(cu, sel)
        }
        if (literal) {
            bool first = true;
            for (;;) {
                QHead h; SdtF4 e;
                sdt_load_rec(rec, ri, h, e);
                h.child_base &= SDT_NODE_MASK;
                const float e1 = e.x;
                const float e2 = e.y + e1;
                const float e3 = e.z + e2;
                const float e4 = e.w + e3;
                const float u = first ? u_pending : rng.select(level);
                first = false;
                const float s = u * e4;
                // :983-991 literally: four masked assignments in turn, a later bin overrides an earlier one (they
                // only overlap for negative energies); no bin (NaN) -> the lane is stuck
                const bool b1 = e1 <= s, b2 = e2 <= s, b3 = e3 <= s;
                const bool m0 = s < e1, m1 = b1 && (s < e2), m2 = b2 && (s < e3);
                if (!(m0 || m1 || m2 || b3)) { q.stuck = true; break; }
                const uint32_t cu = b3 ? 3u : (m2 ? 2u : (m1 ? 1u : 0u));
                SDT_SAMPLE_STEP(cu, SDT_SEL(cu))
            }
        }
    }
#undef SDT_SAMPLE_STEP
    if (level >= SDT_MAX_LEVELS) q.stuck = true;   // deeper than SDT_MAX_LEVELS: impossible for a valid tree
    if (CELL != 0) {
        uint32_t ix, iy;
        if (CELL == 1) sdt_path_cell(path32, level, ix, iy);
        else sdt_path_cell64(path64, level, ix, iy);
        const float sc = sdt_u2f((127u - level) << 23);                  // 2^-level
        q.lox = (float)ix * sc; q.hix = (float)(ix + 1u) * sc;
        q.loy = (float)iy * sc; q.hiy = (float)(iy + 1u) * sc;
    }
    if (!q.stuck) {
        const float ux = rng.pos(3u * level), uy = rng.pos(3u * level + 1u);
        q.x = q.lox + ux * (q.hix - q.lox);                              // :960-962
        q.y = q.loy + uy * (q.hiy - q.loy);
    }
    return q;
}

// KDTree.sample after the spatial descent (src/kdtree.py:482-485).  The pdf of the sampled direction
// (second descent of the reference, :483-484) is the path product of the sampled leaf whenever the
// canonical->dir->canonical round trip stays strictly inside the leaf cell -- then both descents
// arrive at the same leaf and every tie rule is moot; otherwise the pdf descent runs.
struct GuidedSample { float dx, dy, dz, pdf; uint32_t sample_node, pdf_node; };

// ri = record of the tree's root (from the spatial leaf word); root = its node id, only used for
// the node ids reported to dbg (pass 0 when not needed)
template <class Rng>
SDT_HD GuidedSample sdt_sample_tree(const TreeView& t, uint32_t ri, uint32_t root, const Rng& rng, bool fuse) {
    GuidedSample g;
    const QSample q = t.int_cell == 1u ? sdt_quad_sample<Rng, 1>(t.rec, ri, root, rng)
                    : (t.int_cell == 2u ? sdt_quad_sample<Rng, 2>(t.rec, ri, root, rng) : sdt_quad_sample<Rng, 0>(t.rec, ri, root, rng));
    sdt_canonical_to_dir(q.x, q.y, g.dx, g.dy, g.dz);                    // :996
    float px, py;
    sdt_dir_to_canonical(g.dx, g.dy, g.dz, px, py);                      // :1016
    g.sample_node = q.node;
    // (the one gather of the sampler that is not a record: taking it away altogether -- a timing experiment with a made-up
    // product -- is worth 3.7 % of the sample kernel / 2 % of the step, `profiles/r03h_kbench_no_pp_gather.log`; the product could
    // ride in the leaf's parent record if child_base were derived from (record, level) instead of stored)
    const float pp = q.moved ? SDT_LDG(t.pp + q.node) : 1.0f;
    if (fuse && !q.stuck && pp == pp && px > q.lox && px < q.hix && py > q.loy && py < q.hiy) {
        g.pdf = pp * SDT_INV_FOUR_PI;
        g.pdf_node = q.node;
    } else {
        uint32_t nd;
        g.pdf = sdt_quad_pdf(t, ri, root, px, py, nd);
        g.pdf_node = nd;
    }
    return g;
}

// Leaf reached by QuadTree.addDataPropagate's descent (src/quadtree.py:401-441) for a
// canonical direction; SDT_NONE when the root box does not contain it (:405).
SDT_HD uint32_t sdt_quad_leaf(const TreeView& t, uint32_t ri, uint32_t root_node, float x, float y) {
    if (!(x >= 0.0f && x <= 1.0f && y >= 0.0f && y <= 1.0f)) return SDT_NONE;
    const QRec* __restrict__ rec = t.rec;
    float lox = 0.0f, loy = 0.0f, hix = 1.0f, hiy = 1.0f;
    uint32_t node = root_node;
    int level = 0;
    uint32_t cx, cy;
    if (ri < t.jump_trees && sdt_jump_cell(x, y, cx, cy)) {
        const QJump j = SDT_LDG(t.jump + (size_t)ri * SDT_JUMP_CELLS + cy * SDT_JUMP_SIDE + cx);
        if (j & SDT_JUMP_LEAF) return j & ~SDT_JUMP_LEAF;
        ri = j;
        float inv = SDT_JUMP_INV_F;
        level = SDT_JUMP_LEVELS;
        if (j & SDT_JUMP_TABLE) {
            uint32_t e, cx8, cy8;
            if (sdt_s2_lookup(t, t.s2, j & ~SDT_JUMP_TABLE, x, y, e, cx8, cy8, ri)) {
                if (e & SDT_JUMP_LEAF) return e & ~SDT_JUMP_LEAF;
                ri = e;
                cx = cx8; cy = cy8; inv = SDT_S2_FINE_INV_F;
                level = SDT_JUMP_LEVELS + SDT_S2_LEVELS;
            }
        }
        lox = (float)cx * inv; hix = (float)(cx + 1u) * inv;
        loy = (float)cy * inv; hiy = (float)(cy + 1u) * inv;
    }
    for (; level < SDT_MAX_LEVELS && ri != SDT_NONE; ++level) {
        const QHead h = sdt_load_head(rec, ri);      // 16 of the record's 32 bytes
        const float mx = (lox + hix) / 2.0f, my = (loy + hiy) / 2.0f;
        const uint32_t cd = sdt_descend_child(x, y, mx, my);
        node = h.child_base + cd;
        sdt_quadrant_m(cd, mx, my, lox, loy, hix, hiy);
        ri = sdt_child_rec(h.cinfo, h.interior_base, cd);
    }
    return node & SDT_NODE_MASK;
}

// power heuristic, src/path_guiding_integrator.py:16-24 (dr.fma(b,b,a*a), NaN -> 0)
SDT_HD float sdt_mis_weight(float a, float b) {
    const float a2 = a * a;
    const float den = fmaf(b, b, a2);
    float w = (a > 0.0f) ? a2 / den : 0.0f;
    if (w != w) w = 0.0f;
    return w;
}
