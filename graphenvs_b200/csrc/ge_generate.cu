// ge_generate.cu -- device-side instance generation (the graph / weight / terminal part of every
// reference reset(), e.g. shortest_path.py:54-75).  One warp per environment; the candidate
// graph lives as an N x NW adjacency bit-matrix in that warp's shared-memory slice.
//
// DISTRIBUTION parity with the reference, not stream parity: connected G(n,m) by rejection
// (nx.gnm_random_graph: i.i.d. endpoint pairs, self-loops/duplicates rejected, until m edges;
// whole graph rejected until connected; TSP additionally rejects degree-1 nodes and graphs whose
// removal of node 0 disconnects them, tsp.py:60-71), weights k/10 with k~U{3..9} per undirected
// edge, terminals uniform without replacement.  RNG is a counter-based hash, so instance b of a
// batch is a pure function of (seed, env_id0 + b) -- rank-sliced batches equal the single-GPU batch.
// Rows are emitted in INSERTION order like the reference's list(DiGraph.edges) (which edge wins np.argmin ties under
// Multicast parenting >= 3 depends on it).  A draw whose rejection loop runs out of attempts is replaced by a
// connected-by-construction graph and counted (ge_generate_fallbacks); spatial TSP draws coordinates (tsp.py:80-86).
#include <cstdio>

#include "ge_common.cuh"

using namespace ge;

extern "C" int ge_set_error(int code, const char *fmt, ...);
int ge_grant_smem(const void *kernel, size_t smem);  // ge_api.cu

namespace {

__device__ inline uint64_t mix64(uint64_t seed, uint64_t a, uint64_t b) {
    uint64_t z = seed ^ (a * 0x9E3779B97F4A7C15ull) ^ (b * 0xC2B2AE3D27D4EB4Full + 0x165667B19E3779F9ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ inline uint32_t bounded(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * n) >> 32); }

// reachability inside `allowed` over a shared-memory bit-matrix
__device__ inline int reach_count(const uint32_t *mat, int NW, int lane, const uint32_t *allowed, uint32_t *reach,
                                  uint32_t *frontier, uint32_t *next, int seed_node) {
    for (int w = lane; w < NW; w += 32) { reach[w] = 0; frontier[w] = 0; }
    __syncwarp();
    if (lane == 0) { reach[seed_node >> 5] = 1u << (seed_node & 31); frontier[seed_node >> 5] = 1u << (seed_node & 31); }
    __syncwarp();
    for (;;) {
        for (int w = lane; w < NW; w += 32) next[w] = 0;
        for (int fw = 0; fw < NW; ++fw) {
            uint32_t bits = frontier[fw];
            while (bits) {
                int v = (fw << 5) + __ffs(bits) - 1;
                bits &= bits - 1;
                for (int w = lane; w < NW; w += 32) next[w] |= mat[(size_t)v * NW + w];
            }
        }
        __syncwarp();
        uint32_t any = 0;
        for (int w = lane; w < NW; w += 32) {
            uint32_t n = next[w] & allowed[w] & ~reach[w];
            reach[w] |= n;
            frontier[w] = n;
            any |= n;
        }
        __syncwarp();
        if (!__any_sync(GE_FULL, any != 0)) break;
    }
    int c = 0;
    for (int w = lane; w < NW; w += 32) c += __popc(reach[w]);
    return __reduce_add_sync(GE_FULL, c);
}

__device__ unsigned int g_generate_fallbacks;   // envs of the most recent ge_generate that exhausted the rejection loop

#define GE_GEN_MAX_ATTEMPTS 4096

__global__ void generate_kernel(ge_batch d, uint64_t seed, int32_t *__restrict__ row_ptr, int32_t *__restrict__ col,
                                double *__restrict__ w64, float *__restrict__ w32, int words_per_warp, int wpb, int weighted) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * wpb + warp;
    if (b >= d.B) return;
    const int N = d.N, NW = d.NW, E = d.M / 2;
    const int n = (d.kind == GE_DENSEST_SUBGRAPH) ? N - 1 : N;  // densest_subgraph.py:59-65: node N-1 isolated
    uint32_t *mat = smem + (size_t)warp * words_per_warp;
    uint32_t *allowed = mat + (size_t)N * NW, *reach = allowed + NW, *frontier = reach + NW, *next = frontier + NW;
    int *cursor = reinterpret_cast<int *>(next + NW);                       // [N]  next free CSR slot of every row
    uint32_t *elist = reinterpret_cast<uint32_t *>(cursor + N);             // [E]  accepted edges in INSERTION order, (u << 16) | v
    const uint64_t gid = (uint64_t)(uint32_t)(d.env_id0 + b);
    const long long max_edges = (long long)n * (n - 1) / 2;
    const bool complete = E >= max_edges;

    // i.i.d. endpoint pairs, self-loops / duplicates rejected, until E edges (nx.gnm_random_graph); 32 proposals per round,
    // accepted ones appended in (round, lane) order.  `count` edges are already present.
    auto sample_edges = [&](int count, uint32_t attempt) {
        for (uint32_t it = 0; count < E; ++it) {
            const int need = min(32, E - count);
            bool isnew = false;
            int u = 0, v = 0;
            if (lane < need) {
                uint64_t r = mix64(seed, gid, ((uint64_t)attempt << 40) | ((uint64_t)it << 5) | (uint64_t)lane);
                u = (int)bounded((uint32_t)r, (uint32_t)n); v = (int)bounded((uint32_t)(r >> 32), (uint32_t)n);
                if (u != v) {
                    int a = min(u, v), c = max(u, v);
                    uint32_t bit = 1u << (c & 31);
                    uint32_t old = atomicOr(&mat[(size_t)a * NW + (c >> 5)], bit);
                    if (!(old & bit)) {
                        isnew = true;
                        atomicOr(&mat[(size_t)c * NW + (a >> 5)], 1u << (a & 31));
                    }
                }
            }
            const unsigned acc = __ballot_sync(GE_FULL, isnew);
            if (isnew) elist[count + __popc(acc & ((1u << lane) - 1u))] = ((uint32_t)u << 16) | (uint32_t)v;
            count += __popc(acc);
        }
        __syncwarp();
    };
    auto add_edge_lane0 = [&](int u, int v, int k) {   // called by one lane
        mat[(size_t)u * NW + (v >> 5)] |= 1u << (v & 31);
        mat[(size_t)v * NW + (u >> 5)] |= 1u << (u & 31);
        elist[k] = ((uint32_t)u << 16) | (uint32_t)v;
    };

    bool accepted = false;
    for (uint32_t attempt = 0; attempt < GE_GEN_MAX_ATTEMPTS && !accepted; ++attempt) {
        for (int i = lane; i < N * NW; i += 32) mat[i] = 0;
        __syncwarp();
        if (complete) {  // complete_graph, no randomness (nx:generators/random_graphs.py:294-296)
            for (int u = lane; u < n; u += 32)
                for (int w = 0; w < NW; ++w) {
                    uint32_t m = tail_mask(n, w);
                    if (w == (u >> 5)) m &= ~(1u << (u & 31));
                    mat[(size_t)u * NW + w] = m;
                }
        } else {
            sample_edges(0, attempt);
        }
        __syncwarp();
        for (int w = lane; w < NW; w += 32) allowed[w] = tail_mask(n, w);
        __syncwarp();
        if (reach_count(mat, NW, lane, allowed, reach, frontier, next, 0) != n) continue;
        if (d.kind == GE_TSP) {
            int bad = 0;
            for (int u = lane; u < n; u += 32) {
                int deg = 0;
                for (int w = 0; w < NW; ++w) deg += __popc(mat[(size_t)u * NW + w]);
                bad |= (deg == 1);
            }
            if (__any_sync(GE_FULL, bad)) continue;
            if (lane == 0) allowed[0] &= ~1u;
            __syncwarp();
            if (n > 1 && reach_count(mat, NW, lane, allowed, reach, frontier, next, 1) != n - 1) continue;
        }
        accepted = true;
    }
    if (!accepted) {
        // The reference would keep drawing; for (n, m) where a connected draw is this rare the loop would never end
        // here either.  Emit a graph that is valid BY CONSTRUCTION instead and count it (ge_generate_fallbacks): a
        // random recursive tree (TSP: a Hamiltonian ring, which also has no degree-1 node and stays connected without
        // the start) plus uniformly drawn extra edges.
        for (int i = lane; i < N * NW; i += 32) mat[i] = 0;
        __syncwarp();
        int base = 0;
        if (lane == 0) {
            atomicAdd(&g_generate_fallbacks, 1u);
            if (d.kind == GE_TSP) {
                for (int v = 0; v < n; ++v) add_edge_lane0(v, (v + 1) % n, v);
            } else {
                for (int v = 1; v < n; ++v) add_edge_lane0((int)bounded((uint32_t)(mix64(seed ^ 0xA24BAED4963EE407ull, gid, (uint64_t)v) >> 32), (uint32_t)v), v, v - 1);
            }
        }
        base = d.kind == GE_TSP ? n : n - 1;
        __syncwarp();
        sample_edges(min(base, E), GE_GEN_MAX_ATTEMPTS);
    }

    // ---- CSR + weights.  Row u lists its neighbours in the order the undirected edges touching u were INSERTED
    //      (list(G.to_directed().edges) of the reference: source-sorted, insertion order inside a row); a complete
    //      graph's insertion order (itertools.combinations) is ascending neighbour id.
    int32_t *rp = row_ptr + (size_t)b * d.RP;
    int32_t *cl = col + (size_t)b * d.MP;
    double *wd = w64 ? w64 + (size_t)b * d.MP : nullptr;
    float *wf = w32 ? w32 + (size_t)b * d.MP : nullptr;
    // spatial TSP (tsp.py:80-86): coordinates U(0,10)^2, weight = Euclidean distance in fp64
    const bool spatial = d.kind == GE_TSP && d.node_xy != nullptr;
    const uint64_t xyseed = seed ^ 0x2545F4914F6CDD1Dull;
    auto coord = [&](int v, int axis) {
        return (double)(mix64(xyseed, gid, ((uint64_t)v << 1) | (uint64_t)axis) >> 11) * (1.0 / 9007199254740992.0) * 10.0;
    };
    auto edge_weight = [&](int u, int v) {
        if (spatial) {
            const double dx = coord(u, 0) - coord(v, 0), dy = coord(u, 1) - coord(v, 1);
            return sqrt(dx * dx + dy * dy);
        }
        if (!weighted || d.kind == GE_MAX_INDEPENDENT_SET || d.kind == GE_DENSEST_SUBGRAPH) return 1.0;
        uint64_t r = mix64(seed ^ 0x5851F42D4C957F2Dull, gid, ((uint64_t)min(u, v) << 32) | (uint64_t)max(u, v));
        return (double)(3 + (int)bounded((uint32_t)(r >> 32), 7u)) / 10.0;  // randint(3,10)/10.0
    };
    int running = 0;
    if (lane == 0) rp[0] = 0;
    for (int base = 0; base < N; base += 32) {
        int u = base + lane, deg = 0;
        if (u < N)
            for (int w = 0; w < NW; ++w) deg += __popc(mat[(size_t)u * NW + w]);
        int inc = deg;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int x = __shfl_up_sync(GE_FULL, inc, o);
            if (lane >= o) inc += x;
        }
        if (u < N) {
            rp[u + 1] = running + inc;
            int e = running + inc - deg;
            cursor[u] = e;
            if (complete)
                for (int w = 0; w < NW; ++w) {
                    uint32_t bits = mat[(size_t)u * NW + w];
                    while (bits) {
                        int v = (w << 5) + __ffs(bits) - 1;
                        bits &= bits - 1;
                        cl[e] = v;
                        double wt = edge_weight(u, v);
                        if (wd) wd[e] = wt;
                        if (wf) wf[e] = (float)wt;
                        ++e;
                    }
                }
        }
        running += __shfl_sync(GE_FULL, inc, 31);
    }
    __syncwarp();
    if (!complete) {
        for (int k0 = 0; k0 < E; k0 += 32) {   // 32 edges per trip; a node met several times in one trip is ranked by slot order
            const int k = k0 + lane;
            const bool valid = k < E;
            const uint32_t pr = valid ? elist[k] : 0xffffffffu;
            const int a = (int)(pr >> 16), c = (int)(pr & 0xffffu);
            int ra = 0, rc = 0;
            for (int j = 0; j < 32; ++j) {
                const uint32_t pj = __shfl_sync(GE_FULL, pr, j);
                if (j < lane && pj != 0xffffffffu) {
                    const int aj = (int)(pj >> 16), cj = (int)(pj & 0xffffu);
                    ra += (aj == a) + (cj == a);
                    rc += (aj == c) + (cj == c);
                }
            }
            int pa = 0, pc = 0;
            if (valid) { pa = cursor[a] + ra; pc = cursor[c] + rc; }
            __syncwarp();
            if (valid) {
                const double wt = edge_weight(a, c);
                cl[pa] = c; cl[pc] = a;
                if (wd) { wd[pa] = wt; wd[pc] = wt; }
                if (wf) { wf[pa] = (float)wt; wf[pc] = (float)wt; }
                atomicAdd(&cursor[a], 1);
                atomicAdd(&cursor[c], 1);
            }
            __syncwarp();
        }
    }
    for (int e = d.M + lane; e < d.MP; e += 32) { cl[e] = 0; if (wd) wd[e] = 0; if (wf) wf[e] = 0; }

    // ---- terminals / node parameters
    const uint64_t tseed = seed ^ 0xD6E8FEB86659FD93ull;
    if (d.kind == GE_SHORTEST_PATH || d.kind == GE_LONGEST_PATH) {
        if (lane == 0) {
            uint64_t r = mix64(tseed, gid, 1);
            int s = (int)bounded((uint32_t)r, (uint32_t)N), t = (int)bounded((uint32_t)(r >> 32), (uint32_t)(N - 1));
            if (t >= s) t++;
            d.src[b] = s;
            d.dest[b] = t;
        }
    } else if (d.kind == GE_STEINER_TREE || d.kind == GE_MULTICAST_ROUTING || d.kind == GE_DISTRIBUTION_CENTER) {
        uint32_t *tb = d.target_bits + (size_t)b * NW;
        for (int w = lane; w < NW; w += 32) reach[w] = 0;  // chosen set
        __syncwarp();
        if (lane == 0) {
            int src = 0, lo = 0;
            if (d.kind == GE_STEINER_TREE) { src = (int)bounded((uint32_t)mix64(tseed, gid, 2), (uint32_t)N); d.src[b] = src; reach[src >> 5] |= 1u << (src & 31); }
            if (d.kind == GE_MULTICAST_ROUTING) { d.src[b] = 0; reach[0] |= 1u; lo = 1; }
            int want = d.kind == GE_DISTRIBUTION_CENTER ? d.n_targets : d.n_dests;
            int avail = (d.kind == GE_DISTRIBUTION_CENTER) ? N : N - 1;
            if (want > avail) want = avail;
            int got = 0;
            for (uint64_t ctr = 16; got < want; ++ctr) {
                int v = lo + (int)bounded((uint32_t)mix64(tseed, gid, ctr), (uint32_t)(N - lo));
                if ((reach[v >> 5] >> (v & 31)) & 1u) continue;
                reach[v >> 5] |= 1u << (v & 31);
                if (d.kind == GE_DISTRIBUTION_CENTER) d.targets[(size_t)b * d.n_targets + got] = v;
                ++got;
            }
            if (d.kind == GE_STEINER_TREE) reach[src >> 5] &= ~(1u << (src & 31));
            if (d.kind == GE_MULTICAST_ROUTING) reach[0] &= ~1u;
        }
        __syncwarp();
        for (int w = lane; w < NW; w += 32) tb[w] = reach[w];
    }
    if (d.node_cost) {
        for (int v = lane; v < N; v += 32) {
            uint32_t r = (uint32_t)(mix64(tseed, gid, 0x100000000ull + (uint64_t)v) >> 32);
            float c;
            if (d.kind == GE_DISTRIBUTION_CENTER) c = (float)(1 + (int)bounded(r, 3u));          // randint(1,4)
            else c = weighted ? (float)((double)(3 + (int)bounded(r, 7u)) / 10.0) : 1.0f;       // randint(3,10)/10
            d.node_cost[(size_t)b * N + v] = c;
        }
    }
    if (d.node_xy)
        for (int i = lane; i < 2 * N; i += 32) d.node_xy[(size_t)b * N * 2 + i] = spatial ? (float)coord(i >> 1, i & 1) : 0.f;
    // Multicast max_distance draw (multicast_routing.py:103 np.random.rand()): a pure function of (seed, global env id),
    // parked in max_dist32 until ge_prepare(bit 1, u01 = NULL) turns it into the distance
    if (d.kind == GE_MULTICAST_ROUTING && d.max_dist32 && lane == 0)
        d.max_dist32[b] = (float)((double)(mix64(tseed, gid, 3) >> 40) * (1.0 / 16777216.0));
}

__global__ void clear_fallbacks_kernel() { g_generate_fallbacks = 0; }

// PerishableProductDelivery terminals (perishable_product_delivery.py:96-114), distribution parity: delivery time
// U(dt_mn, dt_mx) with the constructor's range (:51-58); per product a pickup uniform among the unused nodes and a dropoff
// uniform among the unused nodes within the delivery time of it (Dijkstra distances; the reference reads them from
// Floyd-Warshall).  Where the reference would redraw the whole graph (no dropoff in range) this kernel redraws the pickup,
// and after 32 tries takes the nearest unused node (counted in ge_generate_fallbacks).  Runs after the CSR exists.
__global__ void __launch_bounds__(GE_WPB * 32) ppd_terminals_kernel(ge_batch d, uint64_t seed, int words_per_warp, int weighted) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * GE_WPB + warp;
    if (b >= d.B) return;
    Scr s = carve(smem + (size_t)warp * words_per_warp, d);
    const int N = d.N, NW = d.NW, P = d.n_dests;
    const uint64_t gid = (uint64_t)(uint32_t)(d.env_id0 + b), tseed = seed ^ 0x9FB21C651E98DF25ull;
    const double avg_degree = (double)d.M / (double)N;
    double avg_dist = log((double)N) / log(avg_degree);
    if (weighted) avg_dist = avg_dist * (0.3 + 1.0) / 2.0;
    const double dt_mn = avg_dist * 0.6, dt_mx = avg_dist * 1.4;
    const double u = (double)(mix64(tseed, gid, 1) >> 11) * (1.0 / 9007199254740992.0);
    const double dt = u * (dt_mx - dt_mn) + dt_mn;
    if (lane == 0) d.max_dist32[b] = (float)dt;
    int32_t *tg = d.targets + (size_t)b * d.n_targets;
    for (int w = lane; w < NW; w += 32) s.aux[w] = 0;                       // used nodes
    __syncwarp();
    uint64_t ctr = 16;
    for (int i = 0; i < P; ++i) {
        int pickup = -1, dropoff = -1;
        for (int attempt = 0; attempt < 33 && dropoff < 0; ++attempt) {
            do { pickup = (int)bounded((uint32_t)(mix64(tseed, gid, ctr++) >> 32), (uint32_t)N); } while (tbit(s.aux, pickup));
            sssp_warp(d, b, lane, s, pickup, 0.0, false);
            __syncwarp();
            const bool last = attempt == 32;
            // candidates: unused nodes other than the pickup within the delivery time (last try: the nearest unused node)
            int cnt = 0;
            for (int v0 = 0; v0 < N; v0 += 32) {
                const int v = v0 + lane;
                const bool c = v < N && v != pickup && !tbit(s.aux, v) && __longlong_as_double((long long)s.q[v]) < dt + 1e-6;
                cnt += __popc(__ballot_sync(GE_FULL, c));
            }
            if (cnt > 0) {
                int r = (int)bounded((uint32_t)(mix64(tseed, gid, ctr++) >> 32), (uint32_t)cnt);
                for (int v0 = 0; v0 < N && dropoff < 0; v0 += 32) {
                    const int v = v0 + lane;
                    const bool c = v < N && v != pickup && !tbit(s.aux, v) && __longlong_as_double((long long)s.q[v]) < dt + 1e-6;
                    const unsigned bal = __ballot_sync(GE_FULL, c);
                    const int k = __popc(bal);
                    if (r < k) dropoff = v0 + nth_set_bit(bal, r); else r -= k;
                }
            } else if (last) {
                double bd = __longlong_as_double(0x7ff0000000000000ll);
                int bv = -1;
                for (int v = lane; v < N; v += 32)
                    if (v != pickup && !tbit(s.aux, v)) { const double dv = __longlong_as_double((long long)s.q[v]); if (dv < bd) { bd = dv; bv = v; } }
                for (int o = 16; o > 0; o >>= 1) {
                    const double od = __shfl_xor_sync(GE_FULL, bd, o);
                    const int ov = __shfl_xor_sync(GE_FULL, bv, o);
                    if (ov >= 0 && (bv < 0 || od < bd || (od == bd && ov < bv))) { bd = od; bv = ov; }
                }
                dropoff = bv;
                if (lane == 0) atomicAdd(&g_generate_fallbacks, 1u);
            }
            __syncwarp();
        }
        if (lane == 0) {
            tg[i] = pickup; tg[P + i] = dropoff;
            s.aux[pickup >> 5] |= 1u << (pickup & 31);
            if (dropoff >= 0) s.aux[dropoff >> 5] |= 1u << (dropoff & 31);
        }
        __syncwarp();
    }
}

}  // namespace

extern "C" int ge_generate(const ge_batch *d, uint64_t seed, int32_t *row_ptr, int32_t *col, double *w64, float *w32, void *stream) {
    GE_NVTX("ge_generate");
    if (!d || !row_ptr || !col) return ge_set_error(GE_ERR_ARG, "ge_generate: null buffers");
    if (d->N > 1024) return ge_set_error(GE_ERR_UNSUPPORTED, "ge_generate: N=%d > 1024 (bit-matrix must fit one warp's shared-memory slice)", d->N);
    const int E = d->M / 2;
    const int n = d->kind == GE_DENSEST_SUBGRAPH ? d->N - 1 : d->N;
    if (E < n - 1) return ge_set_error(GE_ERR_ARG, "ge_generate: n_edges=%d < n-1, graph cannot be connected", E);
    if (d->kind == GE_TSP && E < n && n > 2) return ge_set_error(GE_ERR_ARG, "ge_generate: TSP needs n_edges >= n_nodes (no degree-1 node, tsp.py:64-65)");
    const bool complete = (long long)E >= (long long)n * (n - 1) / 2;      // no edge list needed: rows are ascending
    int wpw = (d->N * d->NW + 4 * d->NW + d->N + (complete ? 0 : E) + 3) & ~3;
    size_t per_warp = (size_t)wpw * sizeof(uint32_t);
    int wpb = (int)((200 * 1024) / per_warp);
    if (wpb < 1) return ge_set_error(GE_ERR_UNSUPPORTED, "ge_generate: graph too large");
    if (wpb > GE_WPB) wpb = GE_WPB;
    size_t smem = per_warp * wpb;
    int rc = ge_grant_smem((const void *)generate_kernel, smem);
    if (rc) return rc;
    int weighted = (d->flags & GE_FLAG_UNWEIGHTED) ? 0 : 1;
    clear_fallbacks_kernel<<<1, 1, 0, (cudaStream_t)stream>>>();
    generate_kernel<<<(d->B + wpb - 1) / wpb, wpb * 32, smem, (cudaStream_t)stream>>>(*d, seed, row_ptr, col, w64, w32, wpw, wpb, weighted);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return ge_set_error(GE_ERR_CUDA, "generate_kernel launch: %s", cudaGetErrorString(e));
    if (d->kind == GE_PERISHABLE_DELIVERY) {
        if (!w64 || !d->targets || !d->max_dist32 || d->n_targets != 2 * d->n_dests || 2 * d->n_dests > d->N)
            return ge_set_error(GE_ERR_ARG, "ge_generate: PerishableProductDelivery needs w64, targets[2 * n_products], max_dist32 and 2 * n_products <= n_nodes");
        ge_batch dd = *d;                                   // the terminals kernel reads the CSR just written
        dd.row_ptr = row_ptr; dd.col = col; dd.w64 = w64;
        const int sw = scratch_words(dd);
        const size_t sm = (size_t)sw * GE_WPB * sizeof(uint32_t);
        if ((rc = ge_grant_smem((const void *)ppd_terminals_kernel, sm))) return rc;
        ppd_terminals_kernel<<<(d->B + GE_WPB - 1) / GE_WPB, GE_WPB * 32, sm, (cudaStream_t)stream>>>(dd, seed, sw, weighted);
        e = cudaGetLastError();
        if (e != cudaSuccess) return ge_set_error(GE_ERR_CUDA, "ppd_terminals_kernel launch: %s", cudaGetErrorString(e));
    }
    return GE_OK;
}

// Number of envs of the most recent ge_generate on `stream` whose rejection loop ran out of attempts and that were
// emitted as connected-by-construction graphs instead (synchronises the stream).  0 for every BASELINE configuration.
extern "C" int ge_generate_fallbacks(void *stream) {
    unsigned int n = 0;
    cudaError_t e = cudaMemcpyFromSymbolAsync(&n, g_generate_fallbacks, sizeof(n), 0, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return ge_set_error(GE_ERR_CUDA, "ge_generate_fallbacks: %s", cudaGetErrorString(e));
    return (int)n;
}
