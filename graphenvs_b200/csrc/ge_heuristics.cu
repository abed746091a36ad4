// ge_heuristics.cu -- eval-mode heuristic solvers as batched warp-synchronous kernels (one warp per env).
//
//   Multicast    union of the FIRST-FOUND shortest paths src -> every destination (multicast_routing.py:107-115).
//                The value depends on networkx's tie order, so the search restates the pop order of nx
//                `_dijkstra_multisource` (nx:algorithms/shortest_paths/weighted.py:853-881) exactly: the heap holds
//                (dist, insertion counter, node); a node is re-pushed only on a STRICT improvement, so of all heap
//                entries of a node the newest has the smallest distance and is popped first, the older ones are
//                skipped as "already final".  The pop order is therefore argmin over the non-final seen nodes of
//                (seen[u], counter of u's last improvement) -- a warp argmin, no heap.  Neighbours are relaxed in
//                adjacency insertion order (= CSR row order), counters advance in that order, a predecessor is replaced
//                only by a strictly shorter path.  -> ge_batch.heuristic (info['heuristic_solution']), pinned on the
//                recorded reference values in tests/golden/heuristics.json.
//
//   The three heuristics whose reference VALUE is defined by Python set / dict iteration order inside networkx
//   (Kou Steiner steiner_tree.py:84-85, Christofides tsp.py:114-117, Ramsey max_independent_set.py:62-67) are not
//   restated; SURVEY 8(f2) allows clearly labelled alternatives under a NEW info key.  -> ge_batch.heuristic_alt
//   (info['heuristic_device']):
//   SteinerTree  shortest-path heuristic (Takahashi-Matsuyama): grow the tree from the source, repeatedly attach the
//                terminal closest to the tree by a shortest path (multi-source ordered Dijkstra).  2-approximation like Kou.
//   TSP          nearest neighbour: from the head take the cheapest edge to an unvisited node (lowest id on ties); when the
//                head has no unvisited neighbour, walk the shortest path to the nearest unvisited node; close the walk
//                back to the start.  A closed walk over real edges, like the reference's expanded Christofides cycle.
//   MIS          greedy minimum-degree independent set size (lowest id on ties).
#include "ge_common.cuh"

using namespace ge;

extern "C" int ge_set_error(int code, const char *fmt, ...);
int ge_grant_smem(const void *kernel, size_t smem);  // ge_api.cu

namespace {

constexpr u64 H_INF = 0x7ff0000000000000ull;

struct HScr {
    u64 *seen;       // [N] fp64 distance bit patterns (non-negative => order preserving), H_INF = not seen
    double *predw;   // [N] weight of the edge (pred[v], v)
    uint32_t *cnt;   // [N] insertion counter of the node's last improvement
    int32_t *pred;   // [N]
    uint32_t *fin;   // [NW] final (popped) nodes
    uint32_t *set0;  // [NW] scratch set (tree / visited / on-path)
    uint32_t *set1;  // [NW] scratch set (stop set)
};

__host__ __device__ inline int hscr_words(int N, int NW) { return (2 * N + 2 * N + N + N + 3 * NW + 3) & ~3; }

__device__ inline HScr hcarve(uint32_t *base, int N, int NW) {
    HScr s;
    s.seen = reinterpret_cast<u64 *>(base);
    s.predw = reinterpret_cast<double *>(base + 2 * N);
    s.cnt = base + 4 * N;
    s.pred = reinterpret_cast<int32_t *>(base + 5 * N);
    s.fin = base + 6 * N;
    s.set0 = s.fin + NW;
    s.set1 = s.set0 + NW;
    return s;
}

// Runs the ordered search on the state in `s` (seen / cnt / fin initialised by the caller) until the heap is empty
// (returns -1) or a node of `stop` is popped (returns it; its distance is final).  `counter` = next insertion counter.
__device__ inline int dijkstra_ordered(const int32_t *rp, const int32_t *col, const double *w64, int N, int lane, HScr &s,
                                       uint32_t &counter, const uint32_t *stop) {
    for (;;) {
        u64 best = ~0ull;
        uint32_t bc = 0xffffffffu;
        int bv = -1;
        for (int v = lane; v < N; v += 32) {
            if (tbit(s.fin, v)) continue;
            const u64 k = s.seen[v];
            if (k == H_INF) continue;
            const uint32_t c = s.cnt[v];
            if (k < best || (k == best && c < bc)) { best = k; bc = c; bv = v; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const u64 ob = __shfl_xor_sync(GE_FULL, best, o);
            const uint32_t oc = __shfl_xor_sync(GE_FULL, bc, o);
            const int ov = __shfl_xor_sync(GE_FULL, bv, o);
            if (ov >= 0 && (bv < 0 || ob < best || (ob == best && oc < bc))) { best = ob; bc = oc; bv = ov; }
        }
        if (bv < 0) return -1;
        if (lane == 0) s.fin[bv >> 5] |= 1u << (bv & 31);
        __syncwarp();
        if (stop && tbit(stop, bv)) return bv;
        const double dv = __longlong_as_double((long long)best);
        const int lo = rp[bv], hi = rp[bv + 1];
        for (int e0 = lo; e0 < hi; e0 += 32) {   // neighbours in adjacency insertion order; counters advance in that order
            const int e = e0 + lane;
            bool improves = false;
            int u = 0;
            double vu = 0.0, w = 0.0;
            if (e < hi) {
                u = col[e];
                w = w64[e];
                vu = dv + w;
                if (!tbit(s.fin, u)) {
                    const u64 old = s.seen[u];
                    improves = old == H_INF || (u64)__double_as_longlong(vu) < old;
                }
            }
            const unsigned bal = __ballot_sync(GE_FULL, improves);
            if (improves) {
                s.seen[u] = (u64)__double_as_longlong(vu);
                s.cnt[u] = counter + __popc(bal & ((1u << lane) - 1u));
                s.pred[u] = bv;
                s.predw[u] = w;
            }
            counter += __popc(bal);
        }
        __syncwarp();
    }
}

__device__ inline void hreset(HScr &s, int N, int NW, int lane) {
    for (int v = lane; v < N; v += 32) s.seen[v] = H_INF;
    for (int w = lane; w < NW; w += 32) s.fin[w] = 0;
    __syncwarp();
}

struct HEnv {
    const int32_t *rp, *col;
    const double *w64;
    HScr s;
    int b, lane;
    bool live;
};

__device__ inline HEnv henv(const ge_batch &d, uint32_t *smem, int words_per_warp, int wpb) {
    HEnv h;
    const int warp = threadIdx.x >> 5;
    h.lane = threadIdx.x & 31;
    h.b = blockIdx.x * wpb + warp;
    h.live = h.b < d.B;
    const int b = h.live ? h.b : 0;
    h.rp = d.row_ptr + (size_t)b * d.RP;
    h.col = d.col + (size_t)b * d.MP;
    h.w64 = d.w64 + (size_t)b * d.MP;
    h.s = hcarve(smem + (size_t)warp * words_per_warp, d.N, d.NW);
    return h;
}

// ---- MulticastRouting: union of first-found shortest paths (multicast_routing.py:107-115)
__global__ void __launch_bounds__(GE_WPB * 32) heur_multicast_kernel(ge_batch d, int words_per_warp, int wpb) {
    extern __shared__ __align__(16) uint32_t smem[];
    HEnv h = henv(d, smem, words_per_warp, wpb);
    if (!h.live) return;
    const int N = d.N, NW = d.NW, lane = h.lane;
    HScr &s = h.s;
    hreset(s, N, NW, lane);
    if (lane == 0) { s.seen[0] = 0ull; s.cnt[0] = 0u; s.pred[0] = -1; }   // src = 0 (multicast_routing.py:94)
    for (int w = lane; w < NW; w += 32) s.set0[w] = 0;
    __syncwarp();
    uint32_t counter = 1;
    dijkstra_ordered(h.rp, h.col, h.w64, N, lane, s, counter, nullptr);
    const uint32_t *tg = d.target_bits + (size_t)h.b * NW;
    if (lane == 0)                                  // edges of the paths = tree edges (pred[v], v) of the nodes on them
        for (int w = 0; w < NW; ++w) {
            uint32_t bits = tg[w];
            while (bits) {
                int v = (w << 5) + __ffs(bits) - 1;
                bits &= bits - 1;
                while (v > 0 && !tbit(s.set0, v) && s.seen[v] != H_INF) {
                    s.set0[v >> 5] |= 1u << (v & 31);
                    v = s.pred[v];
                }
            }
        }
    __syncwarp();
    double total = 0.0;
    for (int v = lane; v < N; v += 32)
        if (tbit(s.set0, v)) total += s.predw[v];
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(GE_FULL, total, o);
    if (lane == 0) d.heuristic[h.b] = total;
}

// ---- SteinerTree: shortest-path heuristic (labelled alternative to Kou, steiner_tree.py:84-85)
__global__ void __launch_bounds__(GE_WPB * 32) heur_steiner_kernel(ge_batch d, int words_per_warp, int wpb) {
    extern __shared__ __align__(16) uint32_t smem[];
    HEnv h = henv(d, smem, words_per_warp, wpb);
    if (!h.live) return;
    const int N = d.N, NW = d.NW, lane = h.lane;
    HScr &s = h.s;
    const uint32_t *tg = d.target_bits + (size_t)h.b * NW;
    const int src = d.src[h.b];
    for (int w = lane; w < NW; w += 32) { s.set0[w] = (w == (src >> 5)) ? (1u << (src & 31)) : 0u; s.set1[w] = tg[w]; }  // tree, open terminals
    __syncwarp();
    if (lane == 0) s.set1[src >> 5] &= ~(1u << (src & 31));
    __syncwarp();
    double total = 0.0;
    for (int round = 0; round < N; ++round) {
        uint32_t open = 0;
        for (int w = lane; w < NW; w += 32) open |= s.set1[w];
        if (!__any_sync(GE_FULL, open != 0u)) break;
        hreset(s, N, NW, lane);
        for (int v = lane; v < N; v += 32)
            if (tbit(s.set0, v)) { s.seen[v] = 0ull; s.cnt[v] = (uint32_t)v; s.pred[v] = -1; }
        __syncwarp();
        uint32_t counter = (uint32_t)N;
        const int t = dijkstra_ordered(h.rp, h.col, h.w64, N, lane, s, counter, s.set1);
        if (t < 0) break;                                           // unreachable terminal: disconnected instance
        total += __longlong_as_double((long long)s.seen[t]);
        if (lane == 0) {
            int v = t;
            while (v >= 0 && !tbit(s.set0, v)) {                    // attach the path; terminals met on the way are closed too
                s.set0[v >> 5] |= 1u << (v & 31);
                s.set1[v >> 5] &= ~(1u << (v & 31));
                v = s.pred[v];
            }
        }
        __syncwarp();
    }
    if (lane == 0) d.heuristic_alt[h.b] = total;
}

// ---- TSP: nearest neighbour closed walk (labelled alternative to Christofides, tsp.py:114-117)
__global__ void __launch_bounds__(GE_WPB * 32) heur_tsp_kernel(ge_batch d, int words_per_warp, int wpb) {
    extern __shared__ __align__(16) uint32_t smem[];
    HEnv h = henv(d, smem, words_per_warp, wpb);
    if (!h.live) return;
    const int N = d.N, NW = d.NW, lane = h.lane;
    HScr &s = h.s;
    for (int w = lane; w < NW; w += 32) { s.set0[w] = (w == 0) ? 1u : 0u; s.set1[w] = tail_mask(N, w) & ~((w == 0) ? 1u : 0u); }  // visited, unvisited
    __syncwarp();
    int head = 0;
    double total = 0.0;
    bool ok = true;
    for (int step = 1; step <= N; ++step) {
        const bool closing = step == N;
        if (closing && lane == 0) s.set1[0] |= 1u;                  // the only node left to reach is the start
        __syncwarp();
        u64 best = ~0ull;                                           // (weight bits, node id): cheapest edge, lowest id on ties
        int bv = -1;
        for (int e = h.rp[head] + lane; e < h.rp[head + 1]; e += 32) {
            const int v = h.col[e];
            if (tbit(s.set1, v)) {
                const u64 k = (u64)__double_as_longlong(h.w64[e]);
                if (k < best || (k == best && v < bv)) { best = k; bv = v; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const u64 ob = __shfl_xor_sync(GE_FULL, best, o);
            const int ov = __shfl_xor_sync(GE_FULL, bv, o);
            if (ov >= 0 && (bv < 0 || ob < best || (ob == best && ov < bv))) { best = ob; bv = ov; }
        }
        if (bv < 0) {                                               // stuck: shortest path to the nearest unvisited node
            hreset(s, N, NW, lane);
            if (lane == 0) { s.seen[head] = 0ull; s.cnt[head] = 0u; s.pred[head] = -1; }
            __syncwarp();
            uint32_t counter = 1;
            bv = dijkstra_ordered(h.rp, h.col, h.w64, N, lane, s, counter, s.set1);
            if (bv < 0) { ok = false; break; }
            best = s.seen[bv];
        }
        total += __longlong_as_double((long long)best);
        head = bv;
        if (lane == 0) { s.set0[bv >> 5] |= 1u << (bv & 31); s.set1[bv >> 5] &= ~(1u << (bv & 31)); }
        __syncwarp();
    }
    if (lane == 0) d.heuristic_alt[h.b] = ok ? total : -1.0;
}

// ---- MaxIndependentSet: greedy minimum-degree independent set size (labelled alternative to Ramsey, max_independent_set.py:62-67)
__global__ void __launch_bounds__(GE_WPB * 32) heur_mis_kernel(ge_batch d, int words_per_warp, int wpb) {
    extern __shared__ __align__(16) uint32_t smem[];
    HEnv h = henv(d, smem, words_per_warp, wpb);
    if (!h.live) return;
    const int N = d.N, NW = d.NW, lane = h.lane;
    HScr &s = h.s;
    int *deg = reinterpret_cast<int *>(s.cnt);
    for (int v = lane; v < N; v += 32) deg[v] = h.rp[v + 1] - h.rp[v];
    for (int w = lane; w < NW; w += 32) s.set0[w] = tail_mask(N, w);   // alive
    __syncwarp();
    int size = 0;
    for (;;) {
        int bd = 0x7fffffff, bv = -1;
        for (int v = lane; v < N; v += 32)
            if (tbit(s.set0, v) && deg[v] < bd) { bd = deg[v]; bv = v; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int od = __shfl_xor_sync(GE_FULL, bd, o), ov = __shfl_xor_sync(GE_FULL, bv, o);
            if (ov >= 0 && (bv < 0 || od < bd || (od == bd && ov < bv))) { bd = od; bv = ov; }
        }
        if (bv < 0) break;
        ++size;
        if (lane == 0) s.set0[bv >> 5] &= ~(1u << (bv & 31));
        __syncwarp();
        // remove the alive neighbours of bv; every removal lowers the degree of ITS alive neighbours
        for (int e0 = h.rp[bv]; e0 < h.rp[bv + 1]; ++e0) {
            const int u = h.col[e0];
            if (!tbit(s.set0, u)) continue;                         // warp-uniform
            if (lane == 0) s.set0[u >> 5] &= ~(1u << (u & 31));
            __syncwarp();
            for (int e = h.rp[u] + lane; e < h.rp[u + 1]; e += 32) {
                const int x = h.col[e];
                if (tbit(s.set0, x)) atomicSub(&deg[x], 1);
            }
            __syncwarp();
        }
    }
    if (lane == 0) d.heuristic_alt[h.b] = (double)size;
}

}  // namespace

// what: 1 = Multicast union-of-paths -> heuristic; 8 = labelled alternatives -> heuristic_alt.  Called from ge_prepare.
int ge_heuristics_launch(const ge_batch *d, int what, cudaStream_t st) {
    if (!d->w64 && d->kind != GE_MAX_INDEPENDENT_SET) return ge_set_error(GE_ERR_ARG, "eval heuristics need w64");
    const int wpw = hscr_words(d->N, d->NW);
    const size_t per_warp = (size_t)wpw * sizeof(uint32_t);
    int wpb = (int)((200 * 1024) / per_warp);
    if (wpb < 1) return ge_set_error(GE_ERR_UNSUPPORTED, "eval heuristic: N=%d too large for shared scratch", d->N);
    if (wpb > GE_WPB) wpb = GE_WPB;
    const size_t smem = per_warp * wpb;
    const int blocks = (d->B + wpb - 1) / wpb;
    const void *kernel = nullptr;
    if ((what & 1) && d->kind == GE_MULTICAST_ROUTING) kernel = (const void *)heur_multicast_kernel;
    else if ((what & 8) && d->kind == GE_STEINER_TREE) kernel = (const void *)heur_steiner_kernel;
    else if ((what & 8) && d->kind == GE_TSP) kernel = (const void *)heur_tsp_kernel;
    else if ((what & 8) && d->kind == GE_MAX_INDEPENDENT_SET) kernel = (const void *)heur_mis_kernel;
    if (!kernel) return GE_OK;
    if (kernel != (const void *)heur_multicast_kernel && !d->heuristic_alt) return ge_set_error(GE_ERR_ARG, "heuristic_alt buffer is null");
    if (kernel == (const void *)heur_multicast_kernel && (!d->heuristic || !d->target_bits)) return ge_set_error(GE_ERR_ARG, "heuristic / target_bits buffer is null");
    int rc = ge_grant_smem(kernel, smem);
    if (rc) return rc;
    ge_batch dd = *d;
    int a1 = wpw, a2 = wpb;
    void *args[] = {&dd, &a1, &a2};
    cudaError_t e = cudaLaunchKernel(kernel, dim3((unsigned)blocks), dim3((unsigned)(wpb * 32)), args, smem, st);
    return e == cudaSuccess ? GE_OK : ge_set_error(GE_ERR_CUDA, "eval heuristic kernel launch: %s", cudaGetErrorString(e));
}
