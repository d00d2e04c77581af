/* ORACLE / CPU BASELINE (test infrastructure only -- never linked or loaded by the product).
 *
 * Plain-C, OpenMP restatement of the reference's per-vertex SD-tree operations on the
 * reference's OWN data layout (the SoA arrays of KDTreeNode / QuadTreeNode, i.e. the 23 arrays of
 * KDTree.saveToFile, /root/reference/src/kdtree.py:539-602): every level gathers the child ids
 * and tests the stored child bounding boxes exactly as the Dr.Jit code does.  It exists so that
 * bench.py can time "the reference's algorithm on all host cores" (Mitsuba 3 / Dr.Jit are not
 * installable in this image); tests/test_oracle_port.py holds it against the numpy
 * oracle (bit-exact directions / pdfs / node ids), which itself is pinned to the reference's own
 * source run on numpy stand-ins for Dr.Jit (oracle/refshim, tests/test_reference_on_shim.py); the
 * semantics of the Dr.Jit primitives stay assumed.  Build: oracle/Makefile -> oracle/_build/libsdtree_port.so
 *   gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math (fp32 ops separately rounded, like numpy)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct {
    /* spatial tree, src/kdtree.py:16-26 */
    uint32_t n_kd;
    const float* kd_bmin; const float* kd_bmax;            /* (n,3) */
    const uint8_t* kd_leaf; const uint32_t* kd_root; const uint32_t* kd_left; const uint32_t* kd_right;
    float* kd_count;
    /* quadtree forest, src/quadtree.py:12-25 */
    uint32_t n_q;
    const uint32_t* q_rootnode;
    const float* q_bmin; const float* q_bmax;              /* (n,2) */
    const uint8_t* q_leaf; const uint32_t* q_child[4];
    float* q_energy;
} port_tree;

static const float TWO_PI = 6.28318530717958647692f, PI_F = 3.14159265358979323846f;
static const float HALF_PI = 1.57079632679489661923f, QUARTER_PI = 0.78539816339744830962f;
static const float INV_FOUR_PI = 0.07957747154594766788f;

/* Dr.Jit sincos / atan2 (CEPHES), same operation order as oracle/drjit_math.py */
static void p_sincos(float x, float* s, float* c) {
    float xa = fabsf(x);
    uint32_t j = (uint32_t)(xa * 1.27323954473516f);
    j = (j + 1u) & 0xFFFFFFFEu;
    float y = (float)j;
    float xr = ((xa - y * 0.78515625f) - y * 2.4187564849853515625e-4f) - y * 3.77489497744594108e-8f;
    float z = xr * xr;
    float ps = (((-1.9515295891e-4f * z + 8.3321608736e-3f) * z + -1.6666654611e-1f) * z) * xr + xr;
    float pc = ((((2.443315711809948e-5f * z + -1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z) * z - 0.5f * z) + 1.0f;
    int swap = (j & 2u) != 0u;
    float ss = swap ? pc : ps, cc = swap ? ps : pc;
    int neg_s = ((j & 4u) != 0u) != (x < 0.0f);
    int neg_c = ((j + 2u) & 4u) != 0u;
    *s = neg_s ? -ss : ss;
    *c = neg_c ? -cc : cc;
}
static float p_atan2(float y, float x) {
    float ax = fabsf(x), ay = fabsf(y);
    float mn = fminf(ax, ay), mx = fmaxf(ax, ay);
    float a = mn / mx;
    int big = a > 0.4142135623730950f;
    float t = big ? (a - 1.0f) / (a + 1.0f) : a;
    float base = big ? QUARTER_PI : 0.0f;
    float z = t * t;
    float p = ((((8.05374449538e-2f * z + -1.38776856032e-1f) * z + 1.99777106478e-1f) * z + -3.33329491539e-1f) * z) * t + t;
    float r = base + p;
    r = (ay > ax) ? HALF_PI - r : r;
    r = (x < 0.0f) ? PI_F - r : r;
    r = (y < 0.0f) ? -r : r;
    r = (mx == 0.0f) ? 0.0f : r;
    return r;
}
/* src/common.py:100-129 */
static void canonical_to_dir(float px, float py, float* d) {
    float ct = 2.0f * py - 1.0f;
    float st = sqrtf(1.0f - ct * ct);
    float sp, cp;
    p_sincos(TWO_PI * px, &sp, &cp);
    d[0] = st * cp; d[1] = st * sp; d[2] = ct;
}
/* src/common.py:132-158 */
static void dir_to_canonical(const float* d, float* px, float* py) {
    float ct = fminf(fmaxf(d[2], -1.0f), 1.0f);
    float phi = p_atan2(d[1], d[0]);
    while (phi < 0.0f) phi += TWO_PI;
    *px = phi / TWO_PI;
    *py = (ct + 1.0f) / 2.0f;
    if (!(isfinite(d[0]) && isfinite(d[1]) && isfinite(d[2]))) { *px = 0.0f; *py = 0.0f; }
}

static int box3(const float* mn, const float* mx, const float* p) {
    return p[0] >= mn[0] && p[0] <= mx[0] && p[1] >= mn[1] && p[1] <= mx[1] && p[2] >= mn[2] && p[2] <= mx[2];
}
static int box2(const float* mn, const float* mx, float x, float y) {
    return x >= mn[0] && x <= mx[0] && y >= mn[1] && y <= mx[1];
}

/* KDTree.getLeafNodeIndex, src/kdtree.py:435-470 */
static uint32_t kd_leaf_of(const port_tree* t, const float* p, int active, int* inbox) {
    uint32_t node = 0;
    int act = active && box3(t->kd_bmin, t->kd_bmax, p);
    if (inbox) *inbox = act;
    while (act) {
        if (t->kd_leaf[node]) break;
        uint32_t l = t->kd_left[node], r = t->kd_right[node];
        uint32_t next = node;
        if (box3(t->kd_bmin + 3 * l, t->kd_bmax + 3 * l, p)) next = l;       /* left first */
        if (box3(t->kd_bmin + 3 * r, t->kd_bmax + 3 * r, p)) next = r;       /* right second: wins on the plane */
        if (next == node) break;
        node = next;
    }
    return node;
}

static uint32_t fmix(uint32_t h) { h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16; return h; }
/* the library's perf-mode generator (sdt_core.h CounterRng, oracle counter_uniform): position
 * uniforms by hash, the selection uniform of level L = top 24 bits of the (L+1)-th LCG state */
static float counter_pos(uint32_t h0, uint32_t idx) {
    uint32_t h = fmix(h0 ^ (idx * 0x85EBCA77u + 0x165667B1u));
    return (float)(h >> 8) * 5.9604644775390625e-08f;
}

/* QuadTree.pdfQuadTree, src/quadtree.py:1001-1101 */
static float quad_pdf(const port_tree* t, uint32_t root, const float* dir, uint32_t* node_out) {
    float x, y;
    dir_to_canonical(dir, &x, &y);
    uint32_t node = t->q_rootnode[root];
    float pdf = 1.0f;
    for (int guard = 0; guard < 64; ++guard) {
        if (t->q_leaf[node]) { pdf = pdf * INV_FOUR_PI; break; }
        uint32_t c[4];
        int in[4];
        for (int k = 0; k < 4; ++k) { c[k] = t->q_child[k][node]; in[k] = box2(t->q_bmin + 2 * c[k], t->q_bmax + 2 * c[k], x, y); }
        float ce = in[0] ? t->q_energy[c[0]] : in[1] ? t->q_energy[c[1]] : in[2] ? t->q_energy[c[2]] : in[3] ? t->q_energy[c[3]] : 0.0f;
        pdf = pdf * ((4.0f * ce) / t->q_energy[node]);
        if (pdf != pdf) { pdf = 0.0f; break; }
        uint32_t next = node;
        for (int k = 0; k < 4; ++k) if (in[k]) next = c[k];                    /* last match wins */
        if (next == node) break;
        node = next;
    }
    if (node_out) *node_out = node;
    return pdf;
}

/* KDTree.sample (src/kdtree.py:473-486): descent, QuadTree.sampleQuadTree (:931-998), then the pdf
 * of the sampled direction by a second descent */
void port_sample(const port_tree* t, uint32_t n, const float* pos, const uint8_t* active, uint32_t seed, uint32_t lane_offset,
                 float* dir, float* pdf, uint32_t* dbg) {
#pragma omp parallel for schedule(static, 1024)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        int act = active ? active[i] != 0 : 1;
        float px = 0.0f, py = 0.0f;
        uint32_t leaf = 0, root = 0, node = 0, pnode = 0;
        float p = 1.0f;
        if (act) {
            leaf = kd_leaf_of(t, pos + 3 * i, 1, 0);
            root = t->kd_root[leaf];
            node = t->q_rootnode[root];
            const uint32_t h0 = fmix(seed + (lane_offset + (uint32_t)i) * 0x9E3779B1u);
            uint32_t lcg = h0;
            for (uint32_t level = 0; level < 64; ++level) {
                float ux = counter_pos(h0, 3 * level);       /* drawn at every level like the reference (:956) */
                float uy = counter_pos(h0, 3 * level + 1);
                if (t->q_leaf[node]) {
                    const float* mn = t->q_bmin + 2 * node; const float* mx = t->q_bmax + 2 * node;
                    px = mn[0] + ux * (mx[0] - mn[0]);
                    py = mn[1] + uy * (mx[1] - mn[1]);
                    break;
                }
                uint32_t c[4];
                for (int k = 0; k < 4; ++k) c[k] = t->q_child[k][node];
                float e1 = t->q_energy[c[0]];
                float e2 = t->q_energy[c[1]] + e1;
                float e3 = t->q_energy[c[2]] + e2;
                float e4 = t->q_energy[c[3]] + e3;
                lcg = lcg * 747796405u + 2891336453u;
                float s = ((float)(lcg >> 8) * 5.9604644775390625e-08f) * e4;
                int pick = -1;
                if (s < e1) pick = 0;
                if (e1 <= s && s < e2) pick = 1;
                if (e2 <= s && s < e3) pick = 2;
                if (e3 <= s) pick = 3;
                if (pick < 0) break;
                node = c[pick];
            }
        }
        float d[3];
        float cpos[2] = {px, py};
        canonical_to_dir(cpos[0], cpos[1], d);
        if (act) p = quad_pdf(t, root, d, &pnode);
        dir[3 * i] = d[0]; dir[3 * i + 1] = d[1]; dir[3 * i + 2] = d[2];
        pdf[i] = p;
        if (dbg) { dbg[4 * i] = leaf; dbg[4 * i + 1] = root; dbg[4 * i + 2] = node; dbg[4 * i + 3] = pnode; }
    }
}

/* KDTree.pdf, src/kdtree.py:489-496 */
void port_pdf(const port_tree* t, uint32_t n, const float* pos, const float* dir, const uint8_t* active, float* pdf, uint32_t* dbg) {
#pragma omp parallel for schedule(static, 1024)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        int act = active ? active[i] != 0 : 1;
        float p = 1.0f;
        uint32_t leaf = 0, root = 0, node = 0;
        if (act) {
            leaf = kd_leaf_of(t, pos + 3 * i, 1, 0);
            root = t->kd_root[leaf];
            p = quad_pdf(t, root, dir + 3 * i, &node);
        }
        pdf[i] = p;
        if (dbg) { dbg[3 * i] = leaf; dbg[3 * i + 1] = root; dbg[3 * i + 2] = node; }
    }
}

static void quad_splat(port_tree* t, uint32_t root, float x, float y, float irr) {
    uint32_t node = t->q_rootnode[root];
    if (!box2(t->q_bmin + 2 * node, t->q_bmax + 2 * node, x, y)) return;       /* src/quadtree.py:405 */
    for (int guard = 0; guard < 64; ++guard) {
#pragma omp atomic
        t->q_energy[node] += irr;                                             /* :411, every visited node */
        if (t->q_leaf[node]) break;
        uint32_t next = node;
        for (int k = 0; k < 4; ++k) {
            uint32_t c = t->q_child[k][node];
            if (box2(t->q_bmin + 2 * c, t->q_bmax + 2 * c, x, y)) next = c;    /* :424-438 */
        }
        if (next == node) break;
        node = next;
    }
}

/* KDTree.addDataPropagate + QuadTree.addDataPropagate, src/kdtree.py:180-225, src/quadtree.py:389-464
 * (no NEE radiance) */
void port_splat(port_tree* t, uint32_t n, const float* pos, const float* dir2, const float* radiance, const float* wo_pdf) {
#pragma omp parallel for schedule(static, 1024)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const float* p = pos + 3 * i;
        uint32_t node = 0;
        int act = box3(t->kd_bmin, t->kd_bmax, p);
        while (act) {
#pragma omp atomic
            t->kd_count[node] += 1.0f;                                        /* src/kdtree.py:199 */
            if (t->kd_leaf[node]) break;
            uint32_t l = t->kd_left[node], r = t->kd_right[node];
            uint32_t next = node;
            if (box3(t->kd_bmin + 3 * l, t->kd_bmax + 3 * l, p)) next = l;
            if (box3(t->kd_bmin + 3 * r, t->kd_bmax + 3 * r, p)) next = r;
            if (next == node) break;
            node = next;
        }
        uint32_t root = t->kd_root[node];                                      /* :224 unmasked */
        float irr = wo_pdf[i] > 0.0f ? radiance[i] / wo_pdf[i] : 0.0f;          /* src/quadtree.py:451 */
        quad_splat(t, root, dir2[2 * i], dir2[2 * i + 1], irr);
    }
}

/* torchrun exports OMP_NUM_THREADS=1 to every worker: the reference arm of bench.py asks for the host's cores explicitly */
void port_set_threads(int n) {
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int port_max_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
