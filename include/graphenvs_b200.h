/* graphenvs_b200.h -- C ABI of the B200-native batched GraphEnvs engine.
 *
 * One `ge_batch` describes B independent environment instances of ONE env kind with uniform
 * (N nodes, M = 2*n_edges directed edges).  Every pointer in it is a DEVICE pointer owned by the
 * caller (the Python host allocates them as torch tensors, a C host with cudaMalloc); the
 * library allocates nothing and keeps no state between calls, so a descriptor can be copied,
 * sliced per rank, or rebuilt freely.  All calls enqueue work on `stream` (a cudaStream_t passed
 * as void*), return 0 on success / a negative ge_status on error (text via ge_last_error()),
 * never throw, and are not thread-safe per descriptor (two host threads may work on DIFFERENT descriptors
 * concurrently: the library's only process-wide state, its kernel-attribute table and its CUDA-graph cache,
 * is mutex-guarded).  No torch types appear here.
 *
 * What each entry point replaces in the reference (paths relative to graph_envs/):
 *   ge_reset          tail of every Env.reset(): state init + first info['mask']
 *                     (shortest_path.py:74-98, longest_path.py:82-122, steiner_tree.py:89-112,
 *                      tsp.py:119-162, max_independent_set.py:70-89, densest_subgraph.py:68-101,
 *                      multicast_routing.py:118-152, distribution_center.py:94-127)
 *   ge_step           Env.step() + Env._get_mask() for the 8 envs
 *                     (shortest_path.py:105-141, longest_path.py:125-196, steiner_tree.py:116-157,
 *                      tsp.py:174-258, max_independent_set.py:92-124, densest_subgraph.py:105-196,
 *                      multicast_routing.py:155-266, distribution_center.py:129-174,
 *                      perishable_product_delivery.py:175-271)
 *   ge_obs_flat       utils.vectorize_graph (utils.py:87-88), layout of utils.devectorize_graph
 *   ge_obs_graph      utils.devectorize_graph (utils.py:14-23) applied on the device: (x, edge_features, edge_index)
 *   ge_features       feature_extraction.generate_features (feature_extraction.py:6-37)
 *   ge_prepare        reset-time derived data: eval heuristics (shortest_path.py:88-90,
 *                     longest_path.py:103-106, steiner_tree.py:77-85, multicast_routing.py:107-115), Multicast max_distance
 *                     (multicast_routing.py:98-103), DistributionCenter in-range tables
 *                     (distribution_center.py:25-26,113-116)
 *   ge_generate       instance generation of reset() (shortest_path.py:54-75 and peers) --
 *                     distribution parity only (device RNG), see DESIGN.md
 *   ge_sample_actions README.md:54-68 "random valid action" loop (policy stand-in for benches)
 *   ge_step_sampled   the same loop body fused: action draw + Env.step() in one launch
 */
#ifndef GRAPHENVS_B200_H
#define GRAPHENVS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GE_ABI_VERSION 3

typedef enum {
    GE_SHORTEST_PATH = 0,      /* ShortestPath-v0        shortest_path.py        node actions */
    GE_LONGEST_PATH = 1,       /* LongestPath-v0         longest_path.py         node actions */
    GE_STEINER_TREE = 2,       /* SteinerTree-v0 / MST   steiner_tree.py         edge actions */
    GE_TSP = 3,                /* TSP-v0                 tsp.py                  node actions */
    GE_MAX_INDEPENDENT_SET = 4,/* MaxIndependentSet-v0   max_independent_set.py  node actions */
    GE_DENSEST_SUBGRAPH = 5,   /* DensestSubgraph-v0     densest_subgraph.py     node actions */
    GE_MULTICAST_ROUTING = 6,  /* MulticastRouting-v0    multicast_routing.py    edge actions */
    GE_DISTRIBUTION_CENTER = 7,/* DistributionCenter-v0  distribution_center.py  node actions */
    GE_PERISHABLE_DELIVERY = 8 /* PerishableProductDelivery-v0  perishable_product_delivery.py  node actions (action == head: pick up).
                                  n_dests = n_products (<= 5), targets = [B, 2 * n_products] pickups then dropoffs,
                                  max_dist32 = delivery time; counters[b] = {2-bit status per product, moves made} */
} ge_kind;

typedef enum {
    GE_OK = 0,
    GE_ERR_ARG = -1,      /* bad descriptor / unsupported size */
    GE_ERR_CUDA = -2,     /* CUDA runtime error, see ge_last_error() */
    GE_ERR_UNSUPPORTED = -3
} ge_status;

/* ge_batch.flags */
#define GE_FLAG_AUTO_RESET 1u   /* on done: re-init the env's state on its own graph inside ge_step */
#define GE_FLAG_WEIGHTED_PR 2u  /* pagerank uses edge weights (TSP stores them as 'weight', tsp.py:90) */
#define GE_FLAG_UNWEIGHTED 4u   /* ge_generate: weighted=False (all edge weights / MIS costs 1.0) */
#define GE_FLAG_FORCE_WARP 8u   /* testing: use the warp-per-env kernels even where the lane-per-env ones apply */
#define GE_FLAG_PDL 16u         /* ge_step / ge_step_sampled launch with programmatic stream serialization: the step kernel may become
                                   resident -- and prefetch STATIC instance data (adjacency tiles, automaton tables) -- while the previous
                                   launch of the stream is still running; it waits for that launch (griddepcontrol.wait) before it touches
                                   any env state.  The caller promises that the preceding work in the stream does not write this batch's
                                   static arrays (i.e. it is another step, not ge_generate / ge_build_adjacency / ge_pool_refill). */

/* per-env status written by ge_step into flags[b].status */
#define GE_STEP_OK 0
#define GE_STEP_INVALID 1     /* the reference would raise AssertionError; state unchanged */
#define GE_STEP_AFTER_DONE 2  /* env already done and auto-reset off; state unchanged */

/* One 4-byte record per env per step (written with a single 32-bit store). */
typedef struct {
    uint8_t done;     /* 1 if this step ended the episode */
    int8_t solved;    /* -1 = key absent from info, 0 / 1 = info['solved'] */
    uint8_t status;   /* GE_STEP_* */
    uint8_t has_mask; /* 0 only for LongestPath's invalid-move early return (longest_path.py:169-173) */
} ge_step_flags;

typedef struct ge_batch {
    /* ---- shape / parameters ---- */
    int32_t kind, B, N, M;        /* M = 2 * n_edges directed edges */
    int32_t parenting;            /* per-env rules of SURVEY 8(b) are enforced by the host */
    int32_t n_dests;              /* SteinerTree / MulticastRouting */
    int32_t n_choices;            /* DensestSubgraph (already floor(N/e) when defaulted) */
    int32_t n_targets;            /* DistributionCenter target_count (row count of in_range) */
    uint32_t flags;
    int32_t env_id0;              /* global id of env 0 of this slice (rank offset; feeds the action sampler) */
    int32_t NW, MW;               /* ceil(N/32), ceil(M/32) */
    int32_t A, AW, AP;            /* mask length (N or M), its words, byte stride (A rounded up to 16) */
    int32_t RP, MP, ADJS;         /* strides: row_ptr (ints), col/w (elements), adj_bits (words) per env */
    int32_t acc_stride;           /* elements between the 4 components of `acc` (= B of the batch the array was allocated for;
                                     ge_fill_layout sets it to B, ge_batch_slice keeps the parent's) */
    int32_t dfa_bytes;            /* size of the `dfa` array in bytes (0 = unknown) */
    double max_distance;          /* DistributionCenter cutoff */

    /* ---- graph store (static per instance) ---- */
    const int32_t *row_ptr;       /* [B, RP]   CSR offsets, reference edge order (source-sorted) */
    const int32_t *col;           /* [B, MP]   destination of every directed edge */
    const float *w32;             /* [B, MP]   edge feature column 0 (float32), kinds stepping in fp32 */
    const double *w64;            /* [B, MP]   float64 edge attribute, kinds stepping in fp64 / prepare */
    uint32_t *adj_bits;           /* [B, ADJS] N rows of NW words: adjacency bit-matrix (derived by ge_build_adjacency).
                                              For N <= 64 node-action kinds without GE_FLAG_FORCE_WARP the library stores it in
                                              tiles of 32 envs, [ceil(B/32)][N][32 envs] of NW words (bank-conflict-free for the
                                              lane-per-env kernels): allocate (B rounded up to 32) * ADJS words */
    int32_t *rev;                 /* [B, MP]   index of the reverse edge (v->u) of every edge (u->v) (derived), or NULL */
    int32_t *esrc;                /* [B, MP]   source node of every edge (derived), or NULL */
    double *wsort;                /* [B, MP]   w64 permuted so that every row is in ascending destination order (derived), or
                                              NULL: adj[u, v] = wsort[row_ptr[u] + rank of v in the adjacency bit-row of u] */
    const uint8_t *wcode;         /* [B, MP]   index of every edge weight in the batch's small set of distinct weights, or NULL */
    const uint8_t *dfa;           /* exact fp64 distance automaton for the cutoff SSSP, or NULL: [S, W, T[S*W], expand[S], cmax[S]]
                                              (cmax[i] = largest weight code j with T[i*W+j] != 255, or 255).
                                              State i = the i-th smallest value reachable as a left-fold fp64 sum of the W
                                              distinct weights without exceeding max_distance; T[i*W+j] = state of
                                              fl(value_i + weight_j) or 255 when it exceeds the cutoff; expand[i] = value_i +
                                              smallest weight <= cutoff.  Built on the host with the same IEEE additions,
                                              so state order == distance order and every comparison is exact. */
    uint32_t *dc_edges;           /* [B, MP]   DistributionCenter with an automaton: every CSR row re-ordered by weight code,
                                              entry = col | code << 16 (derived by ge_prepare bit 4), or NULL: the cutoff search then
                                              visits only the row prefix that can stay within the cutoff (csrc/ge_dc.cu) */
    double *wmin;                 /* [B]       smallest edge weight of the instance (derived), or NULL: lets the cutoff
                                              SSSP skip nodes that cannot relax anything within the cutoff */
    double *wmat;                 /* [B, N, N] dense float64 weight matrix = the reference's self.adj (derived by
                                              ge_build_adjacency), used when N <= 64 by the kinds stepping in fp64; or NULL */

    /* ---- instance parameters (static) ---- */
    int32_t *src, *dest;          /* [B] */
    uint32_t *target_bits;        /* [B, NW]  IS_TARGET column as a bitset */
    float *node_cost;             /* [B, N]   MIS weight / DistCenter cost column */
    float *node_xy;               /* [B, N, 2] TSP spatial coordinates or NULL */
    float *max_dist32;            /* [B]      Multicast MAX_DISTANCE column value */
    int32_t *targets;             /* [B, n_targets] DistributionCenter target node ids */
    uint32_t *in_range;           /* [B, n_targets, NW] nodes within max_distance of each target */
    uint32_t *in_range_t;         /* [B, N, 4] the same table transposed: per node a 128-bit set of the targets that have it in
                                              range (n_targets <= 128; derived by ge_prepare bit 2 when non-NULL), or NULL */
    double *heuristic;            /* [B]      info['heuristic_solution'] */
    double *heuristic_alt;        /* [B]      info['heuristic_device']: labelled alternative where the reference's value is defined by
                                              networkx iteration order (Steiner shortest-path heuristic, TSP nearest neighbour, greedy
                                              MIS; csrc/ge_heuristics.cu), or NULL */
    float *features;              /* [B, N, 5] structural features or NULL (zeros in obs) */

    /* ---- dynamic state ---- */
    int32_t *head;                /* [B] */
    uint32_t *node_bits;          /* [B, NW]  HAS_MSG / TAKEN column as a bitset */
    uint32_t *node_bits2;         /* [B, NW]  DistCenter IS_COVERED; Densest neighbour-union */
    uint32_t *edge_bits;          /* [B, MW]  Multicast EDGE_IS_TAKEN */
    float *dist32;                /* [B, N]   Multicast DISTANCE_FROM_SOURCE column */
    uint64_t *bestkey;            /* [B, N]   Multicast parenting >= 3: running argmin per frontier vertex,
                                              (float32 bits of dist[src e] + delay[e]) << 32 | e, ~0 = none; or NULL */
    double *cost;                 /* [B]      running solution_cost (fp32 kinds keep a float value in it) */
    int32_t *counters;            /* [B, 4]   Densest: k_taken, edge_cnt.  Incremental kernels (ge_incr.cu): targets in the
                                              tree / nodes taken, popcount of the mask, constraints satisfied */
    uint8_t *done;                /* [B] */
    uint32_t *mask_bits;          /* [B, AW]  current valid-action mask, packed */
    uint32_t *mask_cnt;           /* [B, 8]   incremental-mask kernels with large masks: popcounts of the 16 chunks of ceil(AW/16) words
                                              of the packed mask, 16 bits each (state, maintained with the mask), or NULL */
    uint8_t *mask_bytes;          /* [B, AP]  same mask as bytes (torch.bool view) or NULL.  Kept current by the kernels that
                                              rewrite the whole mask; the INCREMENTAL-mask kernels (SteinerTree, Multicast p >= 2,
                                              MaxIndependentSet N > 64: ge_mask_bytes_current() == 0) update only the packed mask and
                                              the byte view is produced on demand by ge_mask_bytes */
    uint32_t *mask_mirror;        /* [B, AW]  optional second destination of every packed-mask write, e.g. PINNED HOST memory
                                              (zero-copy results, see ge_step_host); honoured by the kernels that rewrite
                                              the whole mask, not by the incremental ones (ge_mask_mirror_supported) */
    uint32_t *mask0_bits;         /* [B, AW]  mask right after reset(), written by ge_reset, or NULL.  It depends only on the
                                              instance, so auto-reset inside ge_step copies it instead of recomputing it */
    double *acc;                  /* [4, acc_stride] per-env statistics: episodes, solved, sum reward, sum final cost */
    uint64_t *traj;               /* [B]      rolling checksum of (action, done, solved, status) per env, or NULL;
                                              same recurrence as oracle/graphenvs_oracle.c oenv_rollout */
    uint32_t *env_steps;          /* [B]      per-env count of accepted steps, or NULL.  ge_step increments it and the
                                              samplers add it to `t`, so a captured CUDA graph (frozen kernel
                                              arguments) draws fresh actions on every replay */
    float *obs_x;                 /* [B, N, F] optional DEVICE buffer, or NULL: ge_step_host_pipelined rewrites the node columns of the
                                              observation (ge_obs_nodes; utils.py:14-23 `x`) of every slice on its write-back lane, while
                                              the next slice steps -- the dynamic part of the observation for a device-resident consumer */
    uint32_t *dc_rows;            /* [B, N, 32] DistributionCenter with an automaton, optional (derived by ge_prepare bit 4 next to dc_edges):
                                              the first 32 entries of every weight-sorted row at a FIXED stride of 128 bytes, padded with
                                              0xffffffff.  The cutoff search then needs no row_ptr lookup (one dependent memory round
                                              less per search level), reads sector-aligned rows, and prefetches a row when its node is
                                              queued; a prefix longer than 32 entries continues in dc_edges */
    uint32_t *progress;           /* [ceil(B / 1024)] optional DEVICE counters, or NULL: a step kernel that supports it (ge_progress_supported)
                                              adds the number of envs it has finished -- every store of those envs made visible first -- to
                                              progress[env >> 10].  ge_step_host_pipelined with chunks = 0 uses it to stream results to the host
                                              from a concurrent write-back kernel while the step kernel is still running */
} ge_batch;

/* step outputs (device pointers) */
typedef struct {
    float *reward;            /* [B] */
    ge_step_flags *flags;     /* [B] */
    double *solution_cost;    /* [B] info['solution_cost'] (NaN = key absent) */
} ge_step_out;

int ge_abi_version(void);
const char *ge_last_error(void);

/* Fills the derived size fields (NW, MW, A, AW, AP, RP, MP, ADJS) from kind/N/M. */
int ge_fill_layout(ge_batch *batch);
/* Descriptor of the sub-batch [lo, lo + count) of `batch`: every per-env pointer advanced, B = count, env_id0 += lo.
 * All arrays are SoA with fixed per-env strides, so a slice is a contiguous range of each of them; it can be stepped,
 * reset and observed on its own (other streams, other host threads).  `lo` must be a multiple of 32. */
int ge_batch_slice(const ge_batch *batch, int lo, int count, ge_batch *out);
/* Bytes of dynamic shared memory one step launch uses (for diagnostics / occupancy reports). */
int ge_step_smem_bytes(const ge_batch *batch);

/* Fills every DERIVED graph array whose pointer is set: adj_bits, wmat, wsort, rev, esrc, wmin. */
int ge_build_adjacency(const ge_batch *batch, void *stream);
/* what: bit0 heuristics with the reference's value: SSSP / MST (tie-independent) and Multicast's union of first-found
 *            shortest paths (networkx's pop order restated on the device),
 *       bit3 labelled alternative heuristics -> heuristic_alt (SteinerTree 1 < n_dests < N-1, TSP, MaxIndependentSet),
 *       bit1 Multicast max_distance from u01[B] (the reference's np.random.rand() draw); u01 == NULL takes the draw
 *            ge_generate left in max_dist32 (a pure function of seed and global env id),
 *       bit2 DistributionCenter in-range tables,
 *       bit4 DistributionCenter weight-sorted edge list dc_edges (needs wcode). */
int ge_prepare(const ge_batch *batch, int what, const double *u01, void *stream);
int ge_features(const ge_batch *batch, void *stream);
int ge_generate(const ge_batch *batch, uint64_t seed, int32_t *row_ptr, int32_t *col, double *w64, float *w32,
                void *stream);

/* Envs of the most recent ge_generate on `stream` whose rejection loop (4096 draws) found no valid graph and that were
 * emitted connected-by-construction instead; synchronises the stream.  Negative = error. */
int ge_generate_fallbacks(void *stream);

/* Instance turnover (every reference reset() builds a new graph): gives each env whose episode has ended (done[b] != 0) the
 * next instance of its slot from a pool of `n_banks` prepared batches -- bank order[episode[b] % n_active], episode[b]++ --
 * by copying that instance's static arrays over the env's own, and writes select[b] = 1 for those envs (0 for the others);
 * follow with ge_reset(batch, select).  banks: HOST array of descriptors shaped like `batch`; order (int32[n_active]),
 * episode (uint32[B]), select (uint8[B]): device arrays.  A bank that is being regenerated is left out of `order`. */
int ge_pool_refill(const ge_batch *batch, const ge_batch *banks, int n_banks, const int32_t *order, int n_active,
                   uint32_t *episode, uint8_t *select, void *stream);

/* select: device uint8[B] (1 = reset this env) or NULL for all. */
int ge_reset(const ge_batch *batch, const uint8_t *select, void *stream);
int ge_step(const ge_batch *batch, const int32_t *actions, const ge_step_out *out, void *stream);
int ge_sample_actions(const ge_batch *batch, uint64_t seed, uint32_t t, int32_t *actions, void *stream);
/* ge_sample_actions + ge_step in ONE launch (random-rollout mode: at these batch sizes a launch costs
 * as much as the work).  The chosen actions are written to `actions` (may not be NULL). */
int ge_step_sampled(const ge_batch *batch, uint64_t seed, uint32_t t, int32_t *actions, const ge_step_out *out, void *stream);

/* Reference wire format (utils.py:87-88): out is float32[count, N*F + M*Fe + 2*M]. */
int ge_obs_len(const ge_batch *batch);
int ge_obs_flat(const ge_batch *batch, int env_lo, int count, float *out, void *stream);
/* The same observation as three tensors (what utils.devectorize_graph, utils.py:14-23, slices out of the flat vector):
 * x float32[count, N, F], edge_attr float32[count, M, Fe], edge_index int64[count, M, 2] -- no float round trip of indices. */
int ge_obs_graph(const ge_batch *batch, int env_lo, int count, float *x, float *edge_attr, int64_t *edge_index, void *stream);
/* x only: the node columns are the part of the observation a step changes (utils.py:14-23 `x`). */
int ge_obs_nodes(const ge_batch *batch, int env_lo, int count, float *x, void *stream);
/* Name of the CUDA kernel ge_step / ge_step_sampled dispatches this descriptor to (reports, profiles). */
const char *ge_step_kernel_name(const ge_batch *batch, int sampled);

/* End-to-end entry with HOST buffers: copies actions H2D, steps, copies reward / flags /
 * solution_cost (and the byte mask [B, AP] when h_mask != NULL, the packed mask [B, AW] when
 * h_mask_bits != NULL) D2H, and synchronises the stream.  If reward | flags | solution_cost |
 * mask_bits are laid out back to back in that order on the device AND on the host (B even), the
 * results come back in one copy.
 * d_actions / out are device staging buffers owned by the caller. */
/* ZERO-COPY mode (d_actions == NULL): h_actions / h_reward / h_flags / h_solution_cost must be pinned host
 * memory (device-accessible under UVA); the step kernel reads the actions from it and writes its results
 * to it directly over PCIe -- no copy calls, one launch, one stream sync.  The packed mask arrives the same
 * way when batch->mask_mirror == h_mask_bits and ge_mask_mirror_supported(batch), otherwise by one copy. */
int ge_mask_mirror_supported(const ge_batch *batch);
/* 1 when every ge_step / ge_reset leaves mask_bytes equal to the packed mask; 0 when the byte view must be refreshed with
 * ge_mask_bytes before it is read (incremental-mask kernels: a scattered byte store per changed edge was 2.3 KB of HBM
 * traffic per Multicast step against 1.8 KB of useful bytes). */
int ge_mask_bytes_current(const ge_batch *batch);
/* Expands the packed mask of envs [env_lo, env_lo + count) into mask_bytes (one coalesced pass). */
int ge_mask_bytes(const ge_batch *batch, int env_lo, int count, void *stream);
int ge_step_host(const ge_batch *batch, const int32_t *h_actions, int32_t *d_actions, const ge_step_out *out,
                 float *h_reward, ge_step_flags *h_flags, double *h_solution_cost, uint8_t *h_mask,
                 uint32_t *h_mask_bits, void *stream);

/* PIPELINED end-to-end step with pinned (device-mapped) HOST buffers.  The batch is cut into `chunks` slices; each slice
 * runs copy-in -> step kernel -> write-back on its own branch of ONE CUDA graph, so slice i's results cross PCIe while
 * slice i+1 is stepping and slice i+2's actions arrive.  Results are written back by a small copy kernel straight into
 * the caller's four host arrays (coalesced 128-byte PCIe writes; no staging layout imposed on the host side), and the step
 * kernels read the actions straight from h_actions (zero-copy; GE_PIPE_ZC=0 copies them to d_actions first).  Blocks
 * until the results are visible to the host (spin on stream completion).  h_solution_cost / h_mask_bits may be NULL. */
int ge_step_host_pipelined(const ge_batch *batch, const int32_t *h_actions, int32_t *d_actions, const ge_step_out *out,
                           float *h_reward, ge_step_flags *h_flags, double *h_solution_cost, uint32_t *h_mask_bits,
                           int chunks, void *stream);
/* ge_step_host_pipelined with COMPACT host results: 9 instead of 16 bytes per env next to the packed mask (PCIe write time is what a
 * host step waits for).  h_flags8[b] = done | (solved + 1) << 1 | status << 3 | has_mask << 5 (GE_FLAGS8_* below); h_solution_cost32 =
 * (float) solution_cost, NaN = key absent (within the 1e-5 relative tolerance of the parity contract; costs of the fp32 kinds are
 * float values already).  Same slices, lanes and completion rule. */
#define GE_FLAGS8_DONE(f) ((f) & 1)
#define GE_FLAGS8_SOLVED(f) ((int)(((f) >> 1) & 3) - 1)
#define GE_FLAGS8_STATUS(f) (((f) >> 3) & 3)
#define GE_FLAGS8_HAS_MASK(f) (((f) >> 5) & 1)
int ge_step_host_compact(const ge_batch *batch, const int32_t *h_actions, int32_t *d_actions, const ge_step_out *out,
                         float *h_reward, uint8_t *h_flags8, float *h_solution_cost32, uint32_t *h_mask_bits, int chunks, void *stream);
/* 1 when the step kernel this batch dispatches to signals ge_batch.progress (the lane-per-env families, DistributionCenter). */
int ge_progress_supported(const ge_batch *batch);
/* Drops the cached CUDA graphs ge_step_host / ge_step_host_pipelined built for this batch (call before freeing its memory). */
int ge_step_host_release(const ge_batch *batch);

/* Reduces acc[4,B] to out[4] (device double[4]): episodes, solved, sum reward, sum final cost. */
int ge_stats(const ge_batch *batch, double *out4, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GRAPHENVS_B200_H */
