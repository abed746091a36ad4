// ge_envs.cuh -- per-environment transition + mask rules, one warp per env.
// Every function cites the reference lines it reproduces (paths relative to graph_envs/).
#pragma once
#include "ge_common.cuh"

namespace ge {

struct EnvPtrs {  // per-env slices of the batch arrays
    const int32_t *rp, *col;
    const float *w32;
    const double *w64;
    const uint32_t *adj;
    const uint32_t *tgt;
};

__device__ inline EnvPtrs env_ptrs(const ge_batch &d, int b) {
    EnvPtrs p;
    p.rp = d.row_ptr + (size_t)b * d.RP;
    p.col = d.col + (size_t)b * d.MP;
    p.w32 = d.w32 ? d.w32 + (size_t)b * d.MP : nullptr;
    p.w64 = d.w64 ? d.w64 + (size_t)b * d.MP : nullptr;
    p.adj = d.adj_bits ? d.adj_bits + (size_t)b * d.ADJS : nullptr;
    p.tgt = d.target_bits ? d.target_bits + (size_t)b * d.NW : nullptr;
    return p;
}

struct StepRes {
    double reward, sol;
    int done, solved, has_mask, status;
};

// ------------------------------------------------------------------ masks
// All mask builders read the CURRENT state from shared memory (s.vis, s.aux) and leave the
// mask in s.msk; emit_mask() publishes it.

// ShortestPath (shortest_path.py:105-109), LongestPath p=1, TSP base: N(head) & ~visited.
__device__ inline void mask_head_row(const ge_batch &d, const EnvPtrs &p, Scr &s, int lane, int head) {
    for (int w = lane; w < d.NW; w += 32) s.msk[w] = p.adj[(size_t)head * d.NW + w] & ~s.vis[w];
    __syncwarp();
}

// LongestPath (longest_path.py:125-145).
__device__ inline void mask_longest_path(const ge_batch &d, const EnvPtrs &p, Scr &s, int lane, int head, int dest) {
    if (d.parenting == 0) {
        for (int w = lane; w < d.NW; w += 32) s.msk[w] = 0xffffffffu;
        __syncwarp();
        return;
    }
    mask_head_row(d, p, s, lane, head);
    if (d.parenting < 2) return;
    if (tbit(s.vis, dest)) return;  // dest not in alt_G (:135-136)
    // alt_G == graph induced on unvisited nodes; symmetric => has_path(k, dest) <=> k in reach(dest)
    if (d.N <= 64) {
        // register fast path: lane l holds adjacency rows l and l+32 as 64-bit sets
        uint64_t r0 = 0, r1 = 0, visw = 0;
        if (d.NW == 1) {
            if (lane < d.N) r0 = p.adj[lane];
            visw = s.vis[0];
        } else {  // env base is 16-byte aligned (ADJS % 4 == 0): one 64-bit load per row
            const uint2 *rows = reinterpret_cast<const uint2 *>(p.adj);
            if (lane < d.N) { uint2 t = rows[lane]; r0 = (uint64_t)t.x | ((uint64_t)t.y << 32); }
            if (lane + 32 < d.N) { uint2 t = rows[lane + 32]; r1 = (uint64_t)t.x | ((uint64_t)t.y << 32); }
            visw = (uint64_t)s.vis[0] | ((uint64_t)s.vis[1] << 32);
        }
        uint64_t allowed = ~visw & (d.N == 64 ? ~0ull : ((1ull << d.N) - 1ull));
        uint64_t reach = 1ull << dest, frontier = reach;
        while (frontier) {
            uint64_t c = (((frontier >> lane) & 1ull) ? r0 : 0ull) | (((frontier >> (lane + 32)) & 1ull) ? r1 : 0ull);
            uint32_t lo = __reduce_or_sync(GE_FULL, (uint32_t)c);
            uint32_t hi = __reduce_or_sync(GE_FULL, (uint32_t)(c >> 32));
            uint64_t nx = (((uint64_t)hi << 32) | lo) & allowed & ~reach;
            reach |= nx;
            frontier = nx;
        }
        if (lane == 0) {
            s.msk[0] &= (uint32_t)reach;
            if (d.NW > 1) s.msk[1] &= (uint32_t)(reach >> 32);
        }
        __syncwarp();
        if (d.parenting == 3 && __popcll(allowed) <= d.N / 3) {  // :141-143
            if (lane == 0) {
                s.msk[0] |= (uint32_t)allowed;
                if (d.NW > 1) s.msk[1] |= (uint32_t)(allowed >> 32);
            }
            __syncwarp();
        }
        return;
    }
    int n_alt = 0;
    for (int w = lane; w < d.NW; w += 32) {
        uint32_t a = ~s.vis[w] & tail_mask(d.N, w);
        s.t2[w] = a;  // allowed
        s.t3[w] = 0;  // reach
        n_alt += __popc(a);
    }
    n_alt = __reduce_add_sync(GE_FULL, n_alt);
    __syncwarp();
    if (lane == 0) s.t3[dest >> 5] = 1u << (dest & 31);
    __syncwarp();
    bfs_bits(p.adj, d.NW, lane, s.t2, s.t3, s.t0, s.t1);
    for (int w = lane; w < d.NW; w += 32) {
        uint32_t m = s.msk[w] & s.t3[w];
        if (d.parenting == 3 && n_alt <= d.N / 3) m |= s.t2[w];
        s.msk[w] = m;
    }
    __syncwarp();
}

// TSP (tsp.py:174-199).  s.vis = TAKEN.
__device__ inline void mask_tsp(const ge_batch &d, const EnvPtrs &p, Scr &s, int lane, int head) {
    const int start = 0;
    int taken = 0;
    for (int w = lane; w < d.NW; w += 32) {
        s.msk[w] = p.adj[(size_t)head * d.NW + w] & ~s.vis[w];
        taken += __popc(s.vis[w]);
    }
    taken = __reduce_add_sync(GE_FULL, taken);
    __syncwarp();
    if (taken < d.N - 1 && lane == 0) s.msk[0] &= ~1u;  // :178-179
    __syncwarp();
    if (d.parenting < 2) return;
    // residual alt_G = all nodes minus start minus taken (:121-123, :230-232)
    int n_res = 0;
    for (int w = lane; w < d.NW; w += 32) {
        uint32_t r = ~s.vis[w] & tail_mask(d.N, w);
        if (w == 0) r &= ~1u;
        s.aux[w] = r;
        n_res += __popc(r);
    }
    n_res = __reduce_add_sync(GE_FULL, n_res);
    __syncwarp();
    // One spanning search of alt_G first (sequential worklist, lane w owns word w of every set): a
    // node whose expansion discovered nothing is a LEAF of a spanning tree, and removing a leaf
    // cannot disconnect the graph -- only the tree's internal nodes still need the literal
    // "remove v, test connectivity" of tsp.py:186-194.  The internal set lives in the (unused by TSP)
    // distance scratch s.q, because t0..t3 are recycled by the per-candidate searches below.
    uint32_t *internal = reinterpret_cast<uint32_t *>(s.q);
    bool have_internal = false;
    if (n_res >= 2) {
        int first = 0x7fffffff;
        for (int w = lane; w < d.NW; w += 32) {
            internal[w] = 0; s.t1[w] = 0; s.t3[w] = 0;   // internal, frontier, reach
            if (s.aux[w]) first = min(first, (w << 5) + __ffs(s.aux[w]) - 1);
        }
        first = __reduce_min_sync(GE_FULL, first);
        __syncwarp();
        if (lane == 0) { s.t1[first >> 5] = 1u << (first & 31); s.t3[first >> 5] = 1u << (first & 31); }
        __syncwarp();
        int reached = 1;
        while (reached < n_res) {
            int r = 0x7fffffff;  // next frontier node (lowest id)
            for (int w = lane; w < d.NW; w += 32)
                if (s.t1[w]) r = min(r, (w << 5) + __ffs(s.t1[w]) - 1);
            r = __reduce_min_sync(GE_FULL, r);
            if (r == 0x7fffffff) break;  // frontier empty: alt_G is disconnected
            int added = 0;
            for (int w = lane; w < d.NW; w += 32) {
                uint32_t nx = p.adj[(size_t)r * d.NW + w] & s.aux[w] & ~s.t3[w];
                uint32_t f = s.t1[w] | nx;
                if (w == (r >> 5)) f &= ~(1u << (r & 31));
                s.t1[w] = f;
                s.t3[w] |= nx;
                added += __popc(nx);
            }
            added = __reduce_add_sync(GE_FULL, added);
            if (added && lane == 0) internal[r >> 5] |= 1u << (r & 31);
            reached += added;
            __syncwarp();
        }
        have_internal = reached == n_res;
    }
    for (int cw = 0; cw < d.NW; ++cw) {
        uint32_t cand = s.msk[cw];  // snapshot of valid_nodes for this word (entries only clear themselves)
        if (have_internal) cand &= internal[cw] | (cw == 0 ? 1u : 0u);
        while (cand) {
            int v = (cw << 5) + __ffs(cand) - 1;
            cand &= cand - 1;
            if (v == start) continue;
            // G_copy = alt_G - v ; v is unvisited and != start, hence in alt_G
            if (n_res - 1 == 0) return;  // number_of_nodes()==0 -> break (:191-192)
            for (int w = lane; w < d.NW; w += 32) {
                uint32_t r = s.aux[w];
                if (w == (v >> 5)) r &= ~(1u << (v & 31));
                s.t2[w] = r;
                s.t3[w] = 0;
            }
            __syncwarp();
            // seed = lowest node of G_copy
            int first = 0x7fffffff;
            for (int w = lane; w < d.NW; w += 32)
                if (s.t2[w]) first = min(first, (w << 5) + __ffs(s.t2[w]) - 1);
            first = __reduce_min_sync(GE_FULL, first);
            if (lane == 0) s.t3[first >> 5] = 1u << (first & 31);
            __syncwarp();
            bfs_bits(p.adj, d.NW, lane, s.t2, s.t3, s.t0, s.t1, n_res - 1);
            int reached = 0;
            for (int w = lane; w < d.NW; w += 32) reached += __popc(s.t3[w]);
            reached = __reduce_add_sync(GE_FULL, reached);
            if (reached != n_res - 1 && lane == 0) s.msk[v >> 5] &= ~(1u << (v & 31));  // not connected (:193-194)
            __syncwarp();
        }
    }
}

// SteinerTree (steiner_tree.py:116-120) and Multicast parenting 2 (multicast_routing.py:162-164):
// has[src(e)] & !has[dst(e)].  Only rows of tree nodes are read.
__device__ inline void mask_tree_frontier_edges(const ge_batch &d, const EnvPtrs &p, Scr &s, int lane) {
    for (int w = lane; w < d.AW; w += 32) s.msk[w] = 0;
    __syncwarp();
    expand_set(
        p.rp, s.vis, d.NW, lane, [](int) {},
        [&](int, int e, bool active) {
            if (active) {
                int v = p.col[e];
                if (!tbit(s.vis, v)) atomicOr(&s.msk[e >> 5], 1u << (e & 31));
            }
        });
    __syncwarp();
}

// MulticastRouting (multicast_routing.py:155-188).
__device__ inline void mask_multicast(const ge_batch &d, const EnvPtrs &p, Scr &s, int lane, int b) {
    if (d.parenting == 1) {  // not taken
        const uint32_t *eb = d.edge_bits + (size_t)b * d.MW;
        for (int w = lane; w < d.AW; w += 32) s.msk[w] = ~eb[w];
        __syncwarp();
        return;
    }
    if (d.parenting == 2) {
        // a taken edge always has has[dst]=1, so "& ~taken" is implied
        mask_tree_frontier_edges(d, p, s, lane);
        return;
    }
    // parenting >= 3: one edge per frontier vertex v: argmin_e { dist[src e] + delay[e] } in fp32,
    // lowest edge index on ties (np.argmin over all M edges, :179-185).
    const float *dist = d.dist32 + (size_t)b * d.N;
    for (int v = lane; v < d.N; v += 32) s.q[v] = ~0ull;
    for (int w = lane; w < d.AW; w += 32) s.msk[w] = 0;
    __syncwarp();
    float du = 0.f;
    expand_set(
        p.rp, s.vis, d.NW, lane, [&](int u) { du = u >= 0 ? dist[u] : 0.f; },
        [&](int owner, int e, bool active) {
            float dsrc = __shfl_sync(GE_FULL, du, owner);
            if (active) {
                int v = p.col[e];
                if (!tbit(s.vis, v)) {
                    float c = __fadd_rn(dsrc, p.w32[e]);
                    u64 key = ((u64)__float_as_uint(c) << 32) | (uint32_t)e;
                    atomicMin(&s.q[v], key);
                }
            }
        });
    __syncwarp();
    for (int v = lane; v < d.N; v += 32) {
        u64 k = s.q[v];
        if (k != ~0ull) {
            uint32_t e = (uint32_t)k;
            atomicOr(&s.msk[e >> 5], 1u << (e & 31));
        }
    }
    __syncwarp();
}

// DensestSubgraph (densest_subgraph.py:105-129).  s.vis = TAKEN, s.aux = union of N(u), u taken.
__device__ inline void mask_densest(const ge_batch &d, Scr &s, int lane, int k_taken) {
    for (int w = lane; w < d.NW; w += 32) {
        uint32_t m;
        if (k_taken == 0) m = 0xffffffffu;
        else if (d.parenting == 0) m = ~s.vis[w];
        else m = s.aux[w] & ~s.vis[w];
        s.msk[w] = m;
    }
    __syncwarp();
}

// DistributionCenter (distribution_center.py:129-141).  s.vis = TAKEN, s.aux = COVERED.
__device__ inline void mask_distribution_center(const ge_batch &d, const EnvPtrs &p, Scr &s, int lane, int b) {
    if (d.parenting != 2) {
        for (int w = lane; w < d.NW; w += 32) s.msk[w] = ~s.vis[w];
        __syncwarp();
        return;
    }
    const int32_t *tg = d.targets + (size_t)b * d.n_targets;
    const uint32_t *ir = d.in_range + (size_t)b * d.n_targets * d.NW;
    // 1. compact the uncovered targets into a list (s.lst: free again once the step's SSSP is done)
    uint16_t *live = s.lst;
    int nlive = 0;
    for (int t0 = 0; t0 < d.n_targets; t0 += 32) {
        int t = t0 + lane;
        bool unc = t < d.n_targets && !tbit(s.aux, tg[t]);
        unsigned bal = __ballot_sync(GE_FULL, unc);
        if (unc) live[nlive + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)t;
        nlive += __popc(bal);
    }
    __syncwarp();
    // 2. OR their in-range rows.  A row is NW words: 32/L rows are read per pass by groups of L lanes, four
    //    passes are issued back to back (independent loads) -- a one-row-at-a-time loop serialised
    //    ~100 dependent-latency loads per step and was the largest single cost of this kernel.
    const int L = d.NW <= 1 ? 1 : d.NW <= 2 ? 2 : d.NW <= 4 ? 4 : d.NW <= 8 ? 8 : d.NW <= 16 ? 16 : 32;
    const int rpp = 32 / L, grp = lane / L, wl = lane % L;
    uint32_t acc[4] = {0, 0, 0, 0};  // lane owns words wl, wl+L, ... (N <= 4096)
    for (int i0 = 0; i0 < nlive; i0 += 4 * rpp) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int idx = i0 + j * rpp + grp;
            if (idx < nlive) {
                const uint32_t *row = ir + (size_t)live[idx] * d.NW;
                int k = 0;
                for (int w = wl; w < d.NW; w += L, ++k) acc[k] |= __ldg(row + w);
            }
        }
    }
    for (int o = L; o < 32; o <<= 1)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] |= __shfl_xor_sync(GE_FULL, acc[k], o);
    if (grp == 0) {
        int k = 0;
        for (int w = wl; w < d.NW; w += L, ++k) s.msk[w] = acc[k] & ~s.vis[w];
    }
    __syncwarp();
}

// ------------------------------------------------------------------ state init (tail of reset())
__device__ inline void init_state(const ge_batch &d, int b, int lane, Scr &s) {
    const int kind = d.kind;
    int src = 0;
    if (kind == GE_SHORTEST_PATH || kind == GE_LONGEST_PATH || kind == GE_STEINER_TREE) src = d.src[b];
    bool seeded = (kind == GE_SHORTEST_PATH || kind == GE_LONGEST_PATH || kind == GE_STEINER_TREE ||
                   kind == GE_MULTICAST_ROUTING);
    for (int w = lane; w < d.NW; w += 32) {
        uint32_t v = (seeded && w == (src >> 5)) ? (1u << (src & 31)) : 0u;
        s.vis[w] = v;
        s.aux[w] = 0;
        d.node_bits[(size_t)b * d.NW + w] = v;
        if (d.node_bits2) d.node_bits2[(size_t)b * d.NW + w] = 0;
    }
    if (kind == GE_MULTICAST_ROUTING) {
        for (int w = lane; w < d.MW; w += 32) d.edge_bits[(size_t)b * d.MW + w] = 0;
        for (int v = lane; v < d.N; v += 32) d.dist32[(size_t)b * d.N + v] = (v == 0) ? 0.f : -1.f;
    }
    if (lane == 0) {
        d.head[b] = src;  // TSP: start = 0
        d.cost[b] = 0.0;
        d.done[b] = 0;
        int4 z = make_int4(0, 0, 0, 0);
        *reinterpret_cast<int4 *>(d.counters + (size_t)b * 4) = z;
    }
    __syncwarp();
}

// Builds the mask for the current state (s.vis / s.aux populated) into s.msk.
__device__ inline void build_mask(const ge_batch &d, const EnvPtrs &p, Scr &s, int lane, int b, int head, int k_taken) {
    switch (d.kind) {
    case GE_SHORTEST_PATH: mask_head_row(d, p, s, lane, head); break;
    case GE_LONGEST_PATH: mask_longest_path(d, p, s, lane, head, d.dest[b]); break;
    case GE_STEINER_TREE: mask_tree_frontier_edges(d, p, s, lane); break;
    case GE_TSP: mask_tsp(d, p, s, lane, head); break;
    case GE_MAX_INDEPENDENT_SET:
        for (int w = lane; w < d.NW; w += 32) s.msk[w] = ~s.vis[w];
        __syncwarp();
        break;
    case GE_DENSEST_SUBGRAPH: mask_densest(d, s, lane, k_taken); break;
    case GE_MULTICAST_ROUTING: mask_multicast(d, p, s, lane, b); break;
    case GE_DISTRIBUTION_CENTER: mask_distribution_center(d, p, s, lane, b); break;
    }
}

// reset(): state init + first mask; TSP patches an empty first mask to {start} (tsp.py:154-155).
__device__ inline void reset_env(const ge_batch &d, const EnvPtrs &p, Scr &s, int lane, int b) {
    init_state(d, b, lane, s);
    int head = (d.kind == GE_SHORTEST_PATH || d.kind == GE_LONGEST_PATH) ? d.src[b] : 0;
    build_mask(d, p, s, lane, b, head, 0);
    if (d.kind == GE_TSP) {
        int c = 0;
        for (int w = lane; w < d.AW; w += 32) c += __popc(s.msk[w] & tail_mask(d.A, w));
        c = __reduce_add_sync(GE_FULL, c);
        if (c == 0 && lane == 0) s.msk[0] |= 1u;
        __syncwarp();
    }
    emit_mask(d, b, lane, s.msk);
}

__device__ inline void store_node_bits(const ge_batch &d, int b, int lane, const Scr &s, bool second) {
    for (int w = lane; w < d.NW; w += 32) {
        d.node_bits[(size_t)b * d.NW + w] = s.vis[w];
        if (second) d.node_bits2[(size_t)b * d.NW + w] = s.aux[w];
    }
}

__device__ inline int popc_and_not(const uint32_t *a, const uint32_t *nb, int NW, int lane) {  // |a & ~nb|
    int c = 0;
    for (int w = lane; w < NW; w += 32) c += __popc(a[w] & ~nb[w]);
    return __reduce_add_sync(GE_FULL, c);
}

// ------------------------------------------------------------------ step()
// Returns with r filled; state written back.  `a` is warp-uniform.
__device__ inline void step_env(const ge_batch &d, const EnvPtrs &p, Scr &s, int lane, int b, int a, StepRes &r) {
    const int N = d.N, kind = d.kind;
    r.reward = 0.0; r.sol = __longlong_as_double(0x7ff8000000000000ll); r.done = 0; r.solved = -1; r.has_mask = 1; r.status = GE_STEP_OK;
    for (int w = lane; w < d.NW; w += 32) {
        s.vis[w] = d.node_bits[(size_t)b * d.NW + w];
        if (kind == GE_DENSEST_SUBGRAPH || kind == GE_DISTRIBUTION_CENTER) s.aux[w] = d.node_bits2[(size_t)b * d.NW + w];
    }
    __syncwarp();
    int head = d.head[b];
    double cost = d.cost[b];
    const uint32_t *mb = d.mask_bits + (size_t)b * d.AW;
    bool in_range = a >= 0 && a < d.A;
    bool mask_ok = in_range && ((mb[a >> 5] >> (a & 31)) & 1u);

    if (kind == GE_TSP && a == 0 && head == 0) {  // tsp.py:203-211 (precedes the asserts)
        r.done = 1; r.reward = -(double)N; r.solved = 0; r.sol = -1.0;
        mask_tsp(d, p, s, lane, head);
        emit_mask(d, b, lane, s.msk);
        if (lane == 0) d.done[b] = 1;
        return;
    }
    if (!mask_ok) { r.status = GE_STEP_INVALID; r.has_mask = 0; return; }

    int cnt = 0;
    switch (kind) {
    case GE_SHORTEST_PATH: {  // shortest_path.py:111-141
        double w;
        if (d.wsort) w = edge_weight_ranked(d, b, p.adj, head, a, lane);
        else { int e = find_edge(p.rp, p.col, head, a, lane); w = e >= 0 ? p.w64[e] : 0.0; }
        r.reward = -w;
        cost = cost + w;
        if (a == d.dest[b]) { r.done = 1; r.solved = 1; }
        if (lane == 0) s.vis[a >> 5] |= 1u << (a & 31);
        __syncwarp();
        head = a;
        mask_head_row(d, p, s, lane, head);
        cnt = emit_mask(d, b, lane, s.msk);
        if (!r.done && cnt == 0) { r.done = 1; r.reward = -(double)N; r.solved = 0; }
        if (r.done) r.sol = cost;
        store_node_bits(d, b, lane, s, false);
        break; }
    case GE_LONGEST_PATH: {  // longest_path.py:147-196
        bool nb = (p.adj[(size_t)head * d.NW + (a >> 5)] >> (a & 31)) & 1u;
        bool vis = tbit(s.vis, a);
        if (d.parenting >= 1 && (!nb || vis)) { r.status = GE_STEP_INVALID; r.has_mask = 0; return; }
        double w = 0.0;
        if (nb) {
            if (d.wsort) w = edge_weight_ranked(d, b, p.adj, head, a, lane);
            else { int e = find_edge(p.rp, p.col, head, a, lane); w = e >= 0 ? p.w64[e] : 0.0; }
        }
        r.reward = w;
        cost = cost - w;
        r.sol = cost;  // info['solution_cost'] on every step (:163-165)
        if (!nb || vis) {  // :169-173, early return without info['mask']
            r.done = 1; r.solved = 0; r.reward = -2.0 * N; r.has_mask = 0;
            break;
        }
        head = a;
        if (lane == 0) s.vis[a >> 5] |= 1u << (a & 31);
        __syncwarp();
        if (a == d.dest[b]) { r.done = 1; r.solved = 1; }
        mask_longest_path(d, p, s, lane, head, d.dest[b]);
        cnt = emit_mask(d, b, lane, s.msk);
        if (!r.done && cnt == 0) { r.done = 1; r.reward = -2.0 * N; r.solved = 0; }
        store_node_bits(d, b, lane, s, false);
        break; }
    case GE_STEINER_TREE: {  // steiner_tree.py:123-157
        int v = p.col[a];
        float w = p.w32[a];
        float c32 = __fadd_rn((float)cost, w);
        cost = (double)c32;
        r.reward = -(double)w;
        if (lane == 0) s.vis[v >> 5] |= 1u << (v & 31);
        __syncwarp();
        if (popc_and_not(p.tgt, s.vis, d.NW, lane) == 0) r.done = 1;
        mask_tree_frontier_edges(d, p, s, lane);
        emit_mask(d, b, lane, s.msk);
        if (r.done) { r.solved = 1; r.sol = cost; }
        store_node_bits(d, b, lane, s, false);
        break; }
    case GE_TSP: {  // tsp.py:213-258
        double w;
        if (d.wsort) w = edge_weight_ranked(d, b, p.adj, head, a, lane);
        else { int e = find_edge(p.rp, p.col, head, a, lane); w = e >= 0 ? p.w64[e] : 0.0; }
        r.reward = 0.0 - w;
        cost = cost + w;
        if (lane == 0) s.vis[a >> 5] |= 1u << (a & 31);
        __syncwarp();
        head = a;
        int taken = 0;
        for (int wi = lane; wi < d.NW; wi += 32) taken += __popc(s.vis[wi]);
        taken = __reduce_add_sync(GE_FULL, taken);
        if (taken == N && a == 0) { r.done = 1; r.solved = 1; }
        mask_tsp(d, p, s, lane, head);
        cnt = emit_mask(d, b, lane, s.msk);
        if (!r.done && cnt == 0) { r.done = 1; r.reward -= 2.0 * N; r.solved = 0; }
        if (r.done) r.sol = cost;
        store_node_bits(d, b, lane, s, false);
        break; }
    case GE_MAX_INDEPENDENT_SET: {  // max_independent_set.py:102-124
        float w = d.node_cost[(size_t)b * N + a];
        float c32 = __fadd_rn((float)cost, w);
        cost = (double)c32;
        r.reward = -(double)w;
        if (lane == 0) s.vis[a >> 5] |= 1u << (a & 31);
        __syncwarp();
        for (int wi = lane; wi < d.NW; wi += 32) s.msk[wi] = ~s.vis[wi];
        __syncwarp();
        cnt = emit_mask(d, b, lane, s.msk);
        if (cnt == 0) { r.done = 1; r.solved = 1; r.sol = cost; }
        store_node_bits(d, b, lane, s, false);
        break; }
    case GE_DENSEST_SUBGRAPH: {  // densest_subgraph.py:135-196
        r.solved = 1;
        int4 c = *reinterpret_cast<const int4 *>(d.counters + (size_t)b * 4);
        int k = c.x, ecnt = c.y;
        if (a == N - 1) {  // stop action (:148-154); state and mask unchanged
            r.reward = 0.0; r.done = 1; r.sol = cost;
            break;
        }
        int ne = 0;
        for (int wi = lane; wi < d.NW; wi += 32) {
            uint32_t row = p.adj[(size_t)a * d.NW + wi];
            ne += __popc(row & s.vis[wi]);
            s.aux[wi] |= row;
        }
        ne = __reduce_add_sync(GE_FULL, ne);
        if (k == 0) r.reward = 0.0;
        else r.reward = ((double)(ecnt + ne) / (double)(k + 1)) - ((double)ecnt / (double)k);
        ecnt += ne; k += 1;
        if (lane == 0) s.vis[a >> 5] |= 1u << (a & 31);
        __syncwarp();
        cost = (double)ecnt / (double)k;
        mask_densest(d, s, lane, k);
        emit_mask(d, b, lane, s.msk);
        if (k == d.n_choices) { r.done = 1; r.sol = cost; }
        if (lane == 0) { c.x = k; c.y = ecnt; *reinterpret_cast<int4 *>(d.counters + (size_t)b * 4) = c; }
        store_node_bits(d, b, lane, s, true);
        break; }
    case GE_MULTICAST_ROUTING: {  // multicast_routing.py:191-266
        int u = edge_src(p.rp, N, a), v = p.col[a];
        float w = p.w32[a];
        float penalty = (float)(-2 * N * d.n_dests);
        float rew = -w;
        float c32 = __fadd_rn((float)cost, w);
        cost = (double)c32;
        r.sol = -1.0;
        bool hasu = tbit(s.vis, u), hasv = tbit(s.vis, v);
        if (!hasu || hasv) {  // :213-219 (parenting <= 1 only); state unchanged except the cost
            r.reward = penalty; r.done = 1; r.solved = 0;
            break;
        }
        float *dist = d.dist32 + (size_t)b * N;
        float dv = __fadd_rn(dist[u], w);
        __syncwarp();
        if (lane == 0) {
            s.vis[v >> 5] |= 1u << (v & 31);
            dist[v] = dv;
            d.edge_bits[(size_t)b * d.MW + (a >> 5)] |= 1u << (a & 31);
        }
        __syncwarp();
        __threadfence_block();
        bool violated = false;
        if (tbit(p.tgt, v)) {
            float lim = __fadd_rn(d.max_dist32[b], 1e-4f);  // float32 compare under numpy 2 (:232)
            if (dv > lim) violated = true;
            else { rew = __fadd_rn(rew, 1.0f); if (lane == 0) d.counters[(size_t)b * 4 + 2] += 1; }
        }
        int left = popc_and_not(p.tgt, s.vis, d.NW, lane);
        mask_multicast(d, p, s, lane, b);
        cnt = emit_mask(d, b, lane, s.msk);
        store_node_bits(d, b, lane, s, false);
        if (violated) { r.reward = penalty; r.done = 1; r.solved = 0; break; }
        r.reward = rew;
        if (left == 0) { r.done = 1; r.solved = 1; r.sol = cost; }
        else if (cnt == 0) { r.reward = penalty; r.done = 1; r.solved = 0; }
        break; }
    case GE_DISTRIBUTION_CENTER: {  // distribution_center.py:144-174
        float w = d.node_cost[(size_t)b * N + a];
        float c32 = __fadd_rn((float)cost, w);
        cost = (double)c32;
        float rew = -w;
        if (lane == 0) s.vis[a >> 5] |= 1u << (a & 31);
        __syncwarp();
        cutoff_reach(d, b, lane, s, a);  // find_nodes_in_range (:25-26,155) -> s.t2
        int gained = 0;
        for (int wi = lane; wi < d.NW; wi += 32) {
            uint32_t reach = s.t2[wi];
            uint32_t newly = reach & ~s.aux[wi];
            s.aux[wi] |= reach;
            gained += __popc(newly & p.tgt[wi]);
        }
        gained = __reduce_add_sync(GE_FULL, gained);
        __syncwarp();
        rew += (float)gained;
        r.reward = rew;
        mask_distribution_center(d, p, s, lane, b);
        emit_mask(d, b, lane, s.msk);
        if (popc_and_not(p.tgt, s.aux, d.NW, lane) == 0) { r.done = 1; r.solved = 1; r.sol = cost; }
        store_node_bits(d, b, lane, s, true);
        break; }
    }
    if (lane == 0) {
        d.head[b] = head;
        d.cost[b] = cost;
        if (r.done) d.done[b] = 1;
    }
}

}  // namespace ge
