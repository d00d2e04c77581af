/* sdtree.h -- C ABI of libsdtree.so: the SD-tree hot path of the Mitsuba 3
 * "Practical Path Guiding" lab, as hand-written sm_100a CUDA kernels.
 *
 * This is the drop-in boundary.  The reference has no FFI (it is pure
 * Python traced by Dr.Jit); the cut is the interface its integrator uses on its two
 * KDTree objects (sdTree_prev / sdTree_current,
 * /root/reference/src/path_guiding_integrator.py:68-69).  One sdt_handle owns that
 * PAIR: an immutable "prev" tree that sample/pdf read, and the "current" statistics
 * (same topology) that splat accumulates into, exactly as the reference uses them.
 * Every entry point cites the reference method it replaces (paths relative to
 * /root/reference).
 *
 * Conventions
 *  - every function returns 0 (SDT_OK) or a negative sdt_status; the message is at
 *    sdt_last_error(handle).  No exception crosses the boundary.
 *  - one handle per GPU, not thread-safe.  All device work is enqueued on the
 *    `stream` argument (a cudaStream_t; NULL = default stream) and is asynchronous
 *    unless SDT_SYNC is given.
 *  - query / record buffers belong to the caller.  They are DEVICE pointers unless
 *    SDT_HOST_PTRS is given, in which case they are host pointers (pinned for speed)
 *    and the library stages them through its own device buffers inside the call;
 *    such a call returns when its host outputs are complete, unless SDT_NO_WAIT.
 *  - vectors are strided component views (sdt_vec3 / sdt_vec2): Dr.Jit's SoA
 *    Vector3f is {x,y,z,stride=1}; a C-contiguous (n,3) array is {p,p+1,p+2,stride=3}.
 *  - no torch / Dr.Jit / Mitsuba type appears in any signature.
 */
#ifndef SDTREE_H
#define SDTREE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDT_VERSION 1

typedef struct sdt_tree_s* sdt_handle;
typedef void* sdt_stream; /* cudaStream_t */

typedef enum sdt_status {
    SDT_OK = 0,
    SDT_ERR_INVALID = -1,  /* bad argument */
    SDT_ERR_CUDA = -2,     /* CUDA runtime error (message has the cudaError string) */
    SDT_ERR_CAPACITY = -3, /* node arena too small for the requested tree */
    SDT_ERR_LAYOUT = -4,   /* uploaded arrays are not a valid SD-tree */
    SDT_ERR_NCCL = -5,     /* NCCL missing or failed */
    SDT_ERR_STATE = -6     /* call not valid in the current state */
} sdt_status;

/* flags for the query / splat / refine entry points */
#define SDT_HOST_PTRS 1u     /* buffers are host memory; H2D/D2H staging inside the call */
#define SDT_SYNC 2u          /* cudaStreamSynchronize(stream) before returning */
#define SDT_REFINE_NO_KD 4u   /* sdt_refine: skip the spatial split (KDTree.refine) */
#define SDT_REFINE_NO_QUAD 8u /* sdt_refine: skip threshold + merge/split of the quadtrees */
#define SDT_NO_WAIT 16u      /* with SDT_HOST_PTRS: return once everything is enqueued instead of waiting for the
                              * host outputs; they are valid (and the host inputs may be reused) after the caller
                              * synchronises `stream`.  Back-to-back calls then overlap: the D2H tail of one call
                              * runs under the H2D head of the next */

typedef struct sdt_vec3 { const float* x; const float* y; const float* z; int64_t stride; } sdt_vec3;
typedef struct sdt_vec2 { const float* x; const float* y; int64_t stride; } sdt_vec2;
typedef struct sdt_vec3_out { float* x; float* y; float* z; int64_t stride; } sdt_vec3_out;

/* PathGuidingIntegrator.setup() arguments that reach the trees
 * (src/path_guiding_integrator.py:77-105) plus arena sizes. */
typedef struct sdt_config {
    float bbox_min[3];
    float bbox_max[3];
    int32_t kd_max_depth;    /* KDTree.maxDepth            (sdTreeMaxDepth)   */
    int32_t quad_max_depth;  /* QuadTree.maxDepth          (quadTreeMaxDepth, <= 32) */
    int32_t store_nee;       /* QuadTree.isStoreNEERadiance                    */
    int32_t device;          /* CUDA ordinal                                   */
    uint32_t kd_capacity;    /* spatial-node arena (0 = 1<<21)                 */
    uint32_t quad_capacity;  /* quadtree-node arena per buffer (0 = 1<<26)     */
} sdt_config;

typedef struct sdt_sizes {
    uint32_t n_kd;       /* spatial nodes  (KDTreeNode.getWidth)   */
    uint32_t n_quad;     /* quadtree nodes (QuadTreeNode.getWidth) */
    uint32_t n_roots;    /* quadtrees      (len(rootNodeIndex))    */
    uint32_t n_interior; /* non-leaf quadtree nodes                */
    uint32_t n_levels;   /* quadtree levels in use                 */
    uint32_t kd_leaves;
    uint32_t error;      /* sticky device-side error flags (0 = none; 1 spatial arena, 2 quadtree arena exhausted: tree truncated;
                            4 a refine scan stalled: internal error, tree invalid;
                            8 sdt_hint_records was given less than the records splatted: spatial splits are missing) */
    uint32_t refine_count;
    uint32_t jump_trees; /* quadtrees covered by the 32x32 jump table over their top 5 levels */
    uint32_t jump2_tables; /* level-5 nodes that own a second-stage 8x8 table over their next 3 levels */
} sdt_sizes;

/* The reference's on-disk contract: the 23 arrays of KDTree.saveToFile
 * (src/kdtree.py:539-602), host memory, dtypes float32 / uint32 / uint8(bool). */
typedef struct sdt_arrays {
    uint32_t n_kd, n_quad, n_roots;
    float kd_max_leaf_size;
    int32_t kd_max_depth;
    int32_t quad_max_depth;
    int32_t quad_store_nee;
    float* kd_bbox_min;        /* n_kd*3 */
    float* kd_bbox_max;        /* n_kd*3 */
    uint32_t* kd_depth;        /* n_kd   */
    float* kd_vert_count;      /* n_kd   */
    uint8_t* kd_is_leaf;       /* n_kd   */
    uint32_t* kd_quad_root;    /* n_kd   quadTreeRootIndex */
    uint32_t* kd_child_left;   /* n_kd   */
    uint32_t* kd_child_right;  /* n_kd   */
    uint32_t* q_root_node;     /* n_roots rootNodeIndex */
    float* q_bbox_min;         /* n_quad*2 */
    float* q_bbox_max;         /* n_quad*2 */
    uint32_t* q_depth;         /* n_quad */
    float* q_irradiance;       /* n_quad */
    uint8_t* q_is_leaf;        /* n_quad */
    float* q_threshold;        /* n_quad refinementThreshold */
    uint32_t* q_child[4];      /* n_quad child_1..4_index */
} sdt_arrays;

#define SDT_TREE_PREV 0    /* sdTree_prev   : what sample/pdf read   */
#define SDT_TREE_CURRENT 1 /* sdTree_current: what splat writes into */

/* ---- lifetime -------------------------------------------------------------- */
/* KDTree() x2 + PathGuidingIntegrator.setup(): src/path_guiding_integrator.py:68-69,
 * 98-105; KDTree.setup src/kdtree.py:133-138.  Starts as one spatial leaf that owns
 * one single-node quadtree with threshold +inf (src/quadtree.py:350-362). */
int sdt_create(const sdt_config* cfg, sdt_handle* out);
int sdt_destroy(sdt_handle h);
/* handle may be NULL: message of the last failed sdt_create in this thread */
const char* sdt_last_error(sdt_handle h);

/* ---- tree exchange in the reference schema --------------------------------- */
/* loadSDTreeFromFile (src/path_guiding_integrator.py:597-608 -> KDTree.loadFromFile
 * src/kdtree.py:156-170): prev <- arrays; current <- same topology, zero statistics.
 * Spatial children must be adjacent and follow their parent (the reference's split appends them
 * so, src/kdtree.py:243-245); quadtree nodes may come in any order and are re-labelled to the
 * canonical layout of clearTreeUnusedNode (src/quadtree.py:844-851), which is also the numbering
 * of every node id the library reports. */
int sdt_upload(sdt_handle h, const sdt_arrays* host);
/* Overwrite the statistics of `current` with frozen buffers (all nodes, interior
 * included; same node numbering as the last upload / refine): q_irradiance[n_quad],
 * kd_vert_count[n_kd].  Used to refine from identical stat buffers. */
int sdt_upload_stats(sdt_handle h, const float* q_irradiance, const float* kd_vert_count);
/* blocks until pending work on the handle's last stream is done */
int sdt_get_sizes(sdt_handle h, sdt_sizes* out);
/* saveSDTreeToFile (src/path_guiding_integrator.py:589-594 / src/kdtree.py:539-602).
 * `out` arrays are caller-allocated with the sizes from sdt_get_sizes. */
int sdt_download(sdt_handle h, int which, sdt_arrays* out);

/* ---- queries on `prev` ------------------------------------------------------ */
/* KDTree.getLeafNodeIndex (src/kdtree.py:435-470) + the masked quadTreeRootIndex
 * gather of KDTree.sample/pdf (:482,:493).  active may be NULL (= all active). */
int sdt_locate(sdt_handle h, const sdt_vec3* pos, const uint8_t* active, uint32_t n,
               uint32_t* leaf, uint32_t* root, uint32_t flags, sdt_stream stream);

/* KDTree.sample (src/kdtree.py:473-486): descent -> QuadTree.sampleQuadTree
 * (src/quadtree.py:931-998) -> QuadTree.pdfQuadTree of the sampled direction
 * (:1001-1101).  Uniforms: if `u` != NULL lane i uses u[i*u_stride + 3*level + k],
 * k = 0,1,2 = (u_x, u_y, u_select), consumed as the reference consumes its sampler;
 * else the library's own generator keyed (seed, lane_offset + i, 3*level + k): a murmur3
 * finaliser of (seed, lane) starts the lane; the per-level u_select is the top 24 bits of a
 * 32-bit LCG stream from that key, the leaf position (u_x, u_y) a second hash round
 * (csrc/sdt_core.h CounterRng; restated in oracle/sdtree_oracle.py counter_uniform).
 * dbg (optional, 4*n uint32): per lane {kd leaf, quadtree root id, quadtree node
 * reached by the sample, quadtree node reached by the pdf}. */
int sdt_sample(sdt_handle h, const sdt_vec3* pos, const uint8_t* active, uint32_t n,
               const float* u, uint32_t u_stride, uint32_t seed, uint32_t lane_offset,
               const sdt_vec3_out* dir, float* pdf, uint32_t* dbg,
               uint32_t flags, sdt_stream stream);

/* KDTree.pdf (src/kdtree.py:489-496).  dbg (optional, 3*n): {kd leaf, root id, node}. */
int sdt_pdf(sdt_handle h, const sdt_vec3* pos, const sdt_vec3* dir, const uint8_t* active,
            uint32_t n, float* pdf, uint32_t* dbg, uint32_t flags, sdt_stream stream);

/* KDTree.sample (src/kdtree.py:473-486) and KDTree.pdf (:489-496) of a second, GIVEN direction `qdir` on the same
 * vertices with one spatial descent: what a path vertex asks of the tree when it draws the guided direction
 * (src/path_guiding_integrator.py:301) and also needs the tree's pdf of the emitter direction for the NEE MIS weight
 * (:244).  dir / pdf are exactly sdt_sample's outputs, qpdf exactly sdt_pdf's; the position crosses the bus once. */
int sdt_sample_pdf(sdt_handle h, const sdt_vec3* pos, const uint8_t* active, uint32_t n,
                   const float* u, uint32_t u_stride, uint32_t seed, uint32_t lane_offset,
                   const sdt_vec3_out* dir, float* pdf, const sdt_vec3* qdir, float* qpdf,
                   uint32_t flags, sdt_stream stream);

/* One bounce of the integrator's guided/BSDF choice in ONE pass over the wavefront
 * (src/path_guiding_integrator.py:283-311): lanes with mode[i]==1 are sampled from
 * the tree (:301), lanes with mode[i]==2 get the tree pdf of the BSDF-sampled
 * direction `wo` (:307) and, when bsdf_pdf/bsdf_value are given, the fused one-sample
 * mixture woPdf = f*bsdf_pdf + (1-f)*sdtree_pdf, weight = bsdf_value/woPdf (:310-311);
 * mode 0 lanes are untouched.  Outputs: dir (mode 1), sdtree_pdf (mode 1,2), wo_pdf and
 * weight (mode 2 when fused).  The reference runs two full descents for this.
 * Optional third query of the same vertex: with em_dir given, every lane whose em_active is set (all lanes when it is
 * NULL) -- whatever its mode -- also gets the tree's pdf of the emitter direction, sdtree_pdf_em = sdTree_prev.pdf(si.p,
 * ds.d) of the NEE MIS weight (:244), from the same spatial descent; other lanes' sdtree_pdf_em stays untouched. */
typedef struct sdt_guided_args {
    sdt_vec3 pos;
    sdt_vec3 wo;               /* BSDF-sampled world direction (mode 2) */
    const uint8_t* mode;
    const float* u; uint32_t u_stride; uint32_t seed; uint32_t lane_offset;
    const float* bsdf_pdf;     /* optional */
    sdt_vec3 bsdf_value;       /* optional (x == NULL -> none) */
    double bsdf_sampling_fraction; /* the reference's Python float: the library uses fp32(f) and fp32(1 - f), the
                                    * two constants Dr.Jit sees (1 - f is formed in double, src/path_guiding_integrator.py:247,310) */
    sdt_vec3_out dir;          /* in: unused; out: sampled dir on mode 1 lanes */
    float* sdtree_pdf;
    float* wo_pdf;             /* optional */
    sdt_vec3_out weight;       /* optional */
    sdt_vec3 em_dir;           /* optional (x == NULL -> none): world direction towards the emitter sample */
    const uint8_t* em_active;  /* optional mask of em_dir */
    float* sdtree_pdf_em;      /* out, required with em_dir */
} sdt_guided_args;
int sdt_guided(sdt_handle h, const sdt_guided_args* a, uint32_t n, uint32_t flags, sdt_stream stream);

/* mis_weight + NEE surface pdf (src/path_guiding_integrator.py:16-24, 241-253).
 * iteration <= 1 -> surface_pdf_em = bsdf_pdf_em.  Outputs may be NULL. */
int sdt_mis_nee(sdt_handle h, uint32_t n, const float* bsdf_pdf_em, const float* sdtree_pdf_em,
                const float* pdf_with_delta, const float* pdf_without_delta, const float* ds_pdf,
                const uint8_t* ds_delta, double bsdf_sampling_fraction, int32_t iteration,
                float* surface_pdf_em, float* mis_em, uint32_t flags, sdt_stream stream);
/* one-sample mixture (src/path_guiding_integrator.py:310-311) on lanes with do_mis != 0 */
int sdt_mis_mixture(sdt_handle h, uint32_t n, const float* bsdf_pdf, const float* sdtree_pdf,
                    const sdt_vec3* bsdf_value, const uint8_t* do_mis, double bsdf_sampling_fraction,
                    float* wo_pdf, const sdt_vec3_out* weight, uint32_t flags, sdt_stream stream);

/* dirToCanonical / canonicalToDir (src/common.py:100-158) on n vectors: what the integrator
 * stores in SurfaceInteractionRecord.direction / direction_nee
 * (src/path_guiding_integrator.py:325,338). */
int sdt_dir_to_canonical(sdt_handle h, const sdt_vec3* dir, uint32_t n, float* out_xy /* n*2 */,
                         uint32_t flags, sdt_stream stream);
int sdt_canonical_to_dir(sdt_handle h, const sdt_vec2* pos, uint32_t n, const sdt_vec3_out* dir,
                         uint32_t flags, sdt_stream stream);

/* ---- splat into `current` ---------------------------------------------------- */
/* KDTree.addDataPropagate + QuadTree.addDataPropagate (src/kdtree.py:180-225,
 * src/quadtree.py:389-464) on already-filtered records.  radiance_nee / direction_nee
 * are read only when the tree was created with store_nee.  active may be NULL. */
typedef struct sdt_records {
    sdt_vec3 position;
    sdt_vec2 direction;       /* canonical (phi/2pi, (cos theta + 1)/2) */
    const float* radiance;
    const float* wo_pdf;
    sdt_vec3 radiance_nee;
    sdt_vec2 direction_nee;
    const uint8_t* active;
} sdt_records;
int sdt_splat_records(sdt_handle h, const sdt_records* rec, uint32_t n, uint32_t flags, sdt_stream stream);

/* processPathData + scatterDataIntoSDTree + addDataPropagate in one pass
 * (src/path_guiding_integrator.py:434-500): slot s belongs to ray s / max_depth;
 * radiance = luminance(((Lfinal[ray]-throughputRadiance)/throughputBsdf)/bsdf) with the
 * reference's NaN scrubbing; the reference's filter replaces its compaction. */
typedef struct sdt_path_data {
    uint32_t slots;           /* numRays * max_depth */
    uint32_t max_depth;
    sdt_vec3 l_final;         /* numRays */
    sdt_vec3 throughput_radiance, throughput_bsdf, bsdf; /* slots */
    sdt_vec3 position;
    sdt_vec2 direction;
    const float* wo_pdf;
    sdt_vec3 radiance_nee;
    sdt_vec2 direction_nee;
    const uint8_t* active;
    float* radiance_out;      /* optional: SurfaceInteractionRecord.radiance (slots) */
} sdt_path_data;
int sdt_splat_path_data(sdt_handle h, const sdt_path_data* pd, uint32_t flags, sdt_stream stream);

/* ---- per-iteration refine ---------------------------------------------------- */
/* KDTree.setRefinementThreshold (src/kdtree.py:327-330): maxLeafSize = 12000*sqrt(2^it) */
int sdt_set_iteration_threshold(sdt_handle h, int32_t iteration);
/* direct write of KDTree.maxLeafSize */
int sdt_set_max_leaf_size(sdt_handle h, float max_leaf_size);
/* refineAndPrepareSDTreeForNextIteration (src/path_guiding_integrator.py:566-586):
 * KDTree.refine/split, setQuadTreeRefinementThreshold, refineAllQuadTree,
 * cleanUnusedQuadTree, prev <- current, reset of current.  Entirely on the device,
 * no host round-trip. */
int sdt_refine(sdt_handle h, uint32_t flags, sdt_stream stream);
/* resetTreeVertCount + resetAllQuadTreeIrradiance (src/kdtree.py:401-432,531-532) */
int sdt_reset_stats(sdt_handle h, sdt_stream stream);

/* ---- multi-GPU: replicated tree, one exchange per training iteration ---------- */
/* libnccl.so.2 is dlopen'ed on first use.  id = 128-byte ncclUniqueId made by rank 0
 * (sdt_comm_unique_id) and broadcast by the host program. */
int sdt_comm_unique_id(void* id128);
int sdt_comm_init(sdt_handle h, const void* id128, int32_t rank, int32_t nranks);
/* ncclAllReduce(sum) over current's [quadtree energies | spatial leaf counts] */
int sdt_allreduce(sdt_handle h, sdt_stream stream);
/* Optional, after sdt_allreduce (or after writing through sdt_stat_buffers / sdt_upload_stats): an UPPER BOUND of the number
 * of records that were splatted into `current`, over all ranks, since its statistics were last zero -- e.g. passes x rays x
 * max_depth of the iteration, which every rank can compute without communication.  KDTree.refine's loop
 * (src/kdtree.py:346-347: split while vertCount > maxLeafSize, children get half) cannot run more rounds than that many
 * records allow, so the refine does not launch the split rounds nobody can reach (a single rank keeps this bound by itself;
 * the counts of other ranks' records arrive unannounced, and without the hint all kd_max_depth rounds are launched).  No
 * reference counterpart: its refine is a host loop.  A bound that is too small does not go unnoticed: a leaf that wants
 * more rounds than were launched raises device error flag 8 (sdt_sizes.error). */
int sdt_hint_records(sdt_handle h, uint64_t records_all_ranks);
/* device pointers + element counts of those two buffers (for callers that bring
 * their own collective, e.g. torch.distributed) */
int sdt_stat_buffers(sdt_handle h, float** q_energy, uint32_t* n_quad, float** kd_count, uint32_t* n_kd);

/* ---- tuning / introspection --------------------------------------------------- */
/* Speed keys (results do not depend on them):
 *   "query_block", "query_ctas_per_sm", "splat_block", "splat_ctas_per_sm"   CTA shape of the wavefront kernels (block 0 = the
 *                       kernel's own size: 768 threads for the sampling kernels, 1024 for pdf / locate / splat; larger values are clamped)
 *   "kd_smem_nodes", "kd_smem_count_nodes", "splat_stage_words"              what of the spatial tree is staged in shared memory
 *   "use_kd_grid", "use_jump", "use_jump2", "use_int_cell", "fuse_sample_pdf", "use_compaction"   fast paths on / off (each has an exact slow path)
 *   "splat_aggregate"   combine the lanes of a warp that splat into the same node before the atomic (pays on pixel-coherent
 *                       wavefronts; off by default: on incoherent records it finds no peers and only costs instructions)
 *   "use_pdl"                                                                programmatic dependent launch of the helper kernels
 *   "use_graph"         replay the refine's launch chain as one captured CUDA graph per buffer parity
 *   "helper_ctas_per_sm"   grid of the refine / sweep helper launches whose item count lives on the device: that many CTAs of
 *                       256 threads per SM (1..8, default 8 = every thread slot of an SM; the scans use at most 4)
 *   "host_chunk"   lanes per chunk of the pipelined SDT_HOST_PTRS staging (H2D of chunk k+1 | kernels of chunk k | D2H of chunk k-1).
 * One key switches semantics rather than speed: "quad_thr_reciprocal" = 1 computes the
 * quadtree refinement threshold (src/quadtree.py:519, `E / 100`) as E * fp32(0.01), the
 * form a Dr.Jit build that lowers division by a literal to a reciprocal multiply would
 * produce (first uncertainty); default 0 = IEEE fp32 division. */
int sdt_set_tuning(sdt_handle h, const char* key, int64_t value);
/* cudaStreamSynchronize(stream) on the handle's device: completes SDT_NO_WAIT calls for callers
 * that have no CUDA runtime of their own (ctypes / cgo hosts) */
int sdt_synchronize(sdt_handle h, sdt_stream stream);
/* number of kernels this handle has launched since creation */
uint64_t sdt_kernel_launches(sdt_handle h);
/* L2-resident read bandwidth probe (GB/s): `bytes` working set read `passes` times */
int sdt_measure_l2(sdt_handle h, uint64_t bytes, uint32_t passes, float* gbps, sdt_stream stream);
/* random 32-byte-sector gather probe (GB/s of sectors delivered): every lane of a warp reads one unrelated sector of an
 * L2-resident set of `bytes` per load, 8 independent 256-bit loads per lane and iteration -- the access pattern of a
 * divergent quadtree descent; via_l1 != 0: ld.global.nc (through L1, the kernels' path), 0: ld.global.cg (L2 only) */
int sdt_measure_gather(sdt_handle h, uint64_t bytes, uint32_t iters, int32_t via_l1, float* gbps, sdt_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* SDTREE_H */
