/* graphenvs_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the GraphEnvs hot path (step / _get_mask / vectorize_graph /
 * feature_extraction / SSSP+MST heuristics).  It keeps the SAME state the reference keeps
 * (nodes float32[N,F], edges float32[M,Fe], edge_links int[M,2]) and evaluates the same rules
 * in the same order and the same arithmetic width, so it is the checker the CUDA path is
 * compared against.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  Parity of THIS file is pinned against runs of
 * the unmodified reference recorded in tests/golden/*.npz (oracle/gen_golden.py), because the
 * reference's own tests pin nothing for this path.
 *
 * Reference citations are relative to /root/reference/graph_envs/ ; "nx:" = networkx 3.6.1.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { K_SP = 0, K_LP = 1, K_ST = 2, K_TSP = 3, K_MIS = 4, K_DS = 5, K_MC = 6, K_DC = 7, K_PPD = 8 };
#define PPD_MAXP 5   /* perishable_product_delivery.py:38-41: HAS_P = cols 1..5, NEEDS_P = 6..10, TIME_LEFT = 11..15 */
#define NSTRUCT 5

typedef struct {
    int kind, N, M, F, Fe, parenting;
    int n_dests, n_choices, n_targets;
    int src, dest, start, head;
    int done;
    float *nodes;      /* N*F  */
    float *edges;      /* M*Fe */
    int32_t *links;    /* M*2  */
    double *w64;       /* M    */
    int32_t *row_ptr;  /* N+1, derived (links are source-sorted) */
    uint8_t *in_range; /* n_targets*N (DistributionCenter) */
    int32_t *targets;  /* n_targets */
    uint8_t *alt;      /* N: membership in alt_G (LongestPath p>=2, TSP p>=2) */
    uint8_t *taken_set;/* N: DensestSubgraph nodes_taken */
    double cost;       /* solution_cost / total_solution_cost (kinds that accumulate in fp64) */
    float cost32;      /* kinds that accumulate in float32 */
    double heuristic;
    double max_distance;
    long edge_cnt;
    int k_taken, steps_taken, constraints;
    uint8_t *scratch;  /* 4*N + M bytes */
    int32_t *queue;    /* N */
} oenv;

typedef struct {
    double reward;
    double solution_cost; /* NaN = key absent from info */
    double heuristic;     /* NaN = key absent */
    int done, solved /* -1 absent */, has_mask, status /* 0 ok, 1 AssertionError */;
} ostep;

static int dyn_cols(int kind) {
    switch (kind) {
    case K_SP: case K_LP: case K_ST: case K_MIS: return 2;
    case K_TSP: case K_MC: return 4;
    case K_DS: return 1;
    case K_DC: return 5;
    case K_PPD: return 1 + 3 * PPD_MAXP;
    }
    return 0;
}
static int edge_cols(int kind) { return (kind == K_ST || kind == K_MC) ? 2 : 1; }

#define ND(e, v, c) ((e)->nodes[(size_t)(v) * (e)->F + (c)])
#define ED(e, i, c) ((e)->edges[(size_t)(i) * (e)->Fe + (c)])
#define SRC(e, i) ((e)->links[2 * (size_t)(i)])
#define DST(e, i) ((e)->links[2 * (size_t)(i) + 1])

void oenv_free(oenv *e) {
    if (!e) return;
    free(e->nodes); free(e->edges); free(e->links); free(e->w64); free(e->row_ptr);
    free(e->in_range); free(e->targets); free(e->alt); free(e->taken_set); free(e->scratch); free(e->queue);
    free(e);
}

/* adj[head, a] of the dense float64 matrix (shortest_path.py:82): 0 when there is no edge. */
static double adj_lookup(const oenv *e, int u, int v) {
    for (int i = e->row_ptr[u]; i < e->row_ptr[u + 1]; ++i)
        if (DST(e, i) == v) return e->w64[i];
    return 0.0;
}
static int is_neighbor(const oenv *e, int u, int v) {
    /* _get_neighbors: edge_links[edge_links[:,0]==node, 1]  (shortest_path.py:101-103) */
    for (int i = 0; i < e->M; ++i)
        if (SRC(e, i) == u && DST(e, i) == v) return 1;
    return 0;
}

/* ---- cutoff Dijkstra value semantics (nx:algorithms/shortest_paths/weighted.py:853-881) ----
 * O(N^2) selection; dist = fp64 left-fold along the path, relaxations with dist+w > cutoff skipped. */
static void dijkstra64(const oenv *e, int s, double cutoff, int use_cutoff, double *dist) {
    int N = e->N;
    uint8_t *fin = e->scratch;
    for (int v = 0; v < N; ++v) { dist[v] = INFINITY; fin[v] = 0; }
    dist[s] = 0.0;
    for (;;) {
        int u = -1; double best = INFINITY;
        for (int v = 0; v < N; ++v) if (!fin[v] && dist[v] < best) { best = dist[v]; u = v; }
        if (u < 0) break;
        fin[u] = 1;
        for (int i = e->row_ptr[u]; i < e->row_ptr[u + 1]; ++i) {
            int v = DST(e, i);
            double nd = dist[u] + e->w64[i];
            if (use_cutoff && nd > cutoff) continue;
            if (!fin[v] && nd < dist[v]) dist[v] = nd;
        }
    }
}

/* Kruskal total weight, summed in Kruskal's yield order (ascending weight, stable) -- steiner_tree.py:81 */
typedef struct { double w; int u, v, idx; } kedge;
static int kedge_cmp(const void *a, const void *b) {
    const kedge *x = a, *y = b;
    if (x->w < y->w) return -1;
    if (x->w > y->w) return 1;
    return x->idx - y->idx;
}
static int uf_find(int *p, int x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
double oenv_mst_weight(const oenv *e) {
    int cnt = 0;
    kedge *ke = malloc(sizeof(kedge) * (size_t)e->M);
    for (int i = 0; i < e->M; ++i)
        if (SRC(e, i) < DST(e, i)) { ke[cnt].w = e->w64[i]; ke[cnt].u = SRC(e, i); ke[cnt].v = DST(e, i); ke[cnt].idx = cnt; cnt++; }
    qsort(ke, (size_t)cnt, sizeof(kedge), kedge_cmp);
    int *p = malloc(sizeof(int) * (size_t)e->N);
    for (int v = 0; v < e->N; ++v) p[v] = v;
    double total = 0.0; /* python sum(): int 0 start, left fold */
    for (int i = 0; i < cnt; ++i) {
        int a = uf_find(p, ke[i].u), b = uf_find(p, ke[i].v);
        if (a != b) { p[a] = b; total += ke[i].w; }
    }
    free(p); free(ke);
    return total;
}

double oenv_sssp_pair(oenv *e, int s, int t) {
    double *dist = malloc(sizeof(double) * (size_t)e->N);
    dijkstra64(e, s, 0, 0, dist);
    double d = dist[t];
    free(dist);
    return d;
}
void oenv_sssp_all(oenv *e, int s, double cutoff, int use_cutoff, double *dist_out) {
    dijkstra64(e, s, cutoff, use_cutoff, dist_out);
}

/* =============================== masks =============================== */

static int bfs_reach(const oenv *e, const uint8_t *present, int s, uint8_t *seen) {
    /* BFS over the subgraph induced on `present`; returns #reached */
    int N = e->N, qh = 0, qt = 0, cnt = 0;
    memset(seen, 0, (size_t)N);
    seen[s] = 1; e->queue[qt++] = s;
    while (qh < qt) {
        int u = e->queue[qh++]; cnt++;
        for (int i = e->row_ptr[u]; i < e->row_ptr[u + 1]; ++i) {
            int v = DST(e, i);
            if (present[v] && !seen[v]) { seen[v] = 1; e->queue[qt++] = v; }
        }
    }
    return cnt;
}

static void mask_sp(const oenv *e, uint8_t *mask) { /* shortest_path.py:105-109 */
    memset(mask, 0, (size_t)e->N);
    for (int i = 0; i < e->M; ++i) if (SRC(e, i) == e->head) mask[DST(e, i)] = 1;
    for (int v = 0; v < e->N; ++v) if (ND(e, v, 0) == 1.0f) mask[v] = 0;
}

static void mask_lp(const oenv *e, uint8_t *mask) { /* longest_path.py:125-145 */
    int N = e->N;
    if (e->parenting == 0) { memset(mask, 1, (size_t)N); return; }
    mask_sp(e, mask);
    if (e->parenting >= 2) {
        if (!e->alt[e->dest]) return;
        uint8_t *seen = e->scratch + N;
        for (int k = 0; k < N; ++k)
            if (mask[k]) { /* nx.has_path(alt_G, k, dest): k is in alt_G because mask excludes visited */
                bfs_reach(e, e->alt, k, seen);
                if (!seen[e->dest]) mask[k] = 0;
            }
        if (e->parenting == 3) {
            int n_alt = 0;
            for (int v = 0; v < N; ++v) n_alt += e->alt[v];
            if (n_alt <= N / 3)
                for (int v = 0; v < N; ++v) if (e->alt[v]) mask[v] = 1;
        }
    }
}

static void mask_st(const oenv *e, uint8_t *mask) { /* steiner_tree.py:116-120 */
    for (int i = 0; i < e->M; ++i) {
        mask[i] = 1;
        if (ND(e, SRC(e, i), 0) < 0.5f) mask[i] = 0;
        if (ND(e, DST(e, i), 0) > 0.5f) mask[i] = 0;
    }
}

static void mask_tsp(const oenv *e, uint8_t *mask) { /* tsp.py:174-199 */
    int N = e->N;
    memset(mask, 0, (size_t)N);
    for (int i = 0; i < e->M; ++i) if (SRC(e, i) == e->head) mask[DST(e, i)] = 1;
    float taken_sum = 0.f;
    for (int v = 0; v < N; ++v) { if (ND(e, v, 0) == 1.0f) mask[v] = 0; taken_sum += ND(e, v, 0); }
    if (taken_sum < (float)(N - 1)) mask[e->start] = 0;
    if (e->parenting >= 2) {
        uint8_t *copy = e->scratch + N, *seen = e->scratch + 2 * N;
        for (int v = 0; v < N; ++v) {
            if (!mask[v]) continue;       /* valid_nodes computed before the loop; entries only get cleared at v itself */
            if (v == e->start) continue;
            memcpy(copy, e->alt, (size_t)N);
            copy[v] = 0;                  /* G_copy.remove_node(v) */
            int n = 0, first = -1;
            for (int u = 0; u < N; ++u) if (copy[u]) { n++; if (first < 0) first = u; }
            if (n == 0) break;
            if (bfs_reach(e, copy, first, seen) != n) mask[v] = 0;
        }
    }
}

static void mask_mis(const oenv *e, uint8_t *mask) { /* max_independent_set.py:92-100 */
    for (int v = 0; v < e->N; ++v) mask[v] = (ND(e, v, 1) == 0.0f);
}

static void mask_ds(const oenv *e, uint8_t *mask) { /* densest_subgraph.py:105-129 */
    int N = e->N;
    float s = 0.f;
    for (int v = 0; v < N; ++v) s += ND(e, v, 0);
    if (s == 0.f) { memset(mask, 1, (size_t)N); return; }
    if (e->parenting == 0) {
        for (int v = 0; v < N; ++v) mask[v] = !(ND(e, v, 0) == 1.0f);
    } else {
        memset(mask, 0, (size_t)N);
        for (int i = 0; i < e->M; ++i) if (ND(e, SRC(e, i), 0) == 1.0f) mask[DST(e, i)] = 1;
        for (int v = 0; v < N; ++v) if (ND(e, v, 0) == 1.0f) mask[v] = 0;
    }
}

static void mask_mc(const oenv *e, uint8_t *mask) { /* multicast_routing.py:155-188 */
    int M = e->M, N = e->N;
    for (int i = 0; i < M; ++i) {
        mask[i] = 1;
        if (ED(e, i, 1) > 0.5f) mask[i] = 0;
        if (e->parenting >= 2) {
            if (ND(e, SRC(e, i), 0) < 0.5f) mask[i] = 0;
            if (ND(e, DST(e, i), 0) > 0.5f) mask[i] = 0;
        }
    }
    if (e->parenting >= 3) {
        uint8_t *isV = e->scratch + N; /* np.unique(edge_links[mask,1]) -> ascending */
        memset(isV, 0, (size_t)N);
        for (int i = 0; i < M; ++i) if (mask[i]) isV[DST(e, i)] = 1;
        memset(mask, 0, (size_t)M);
        for (int v = 0; v < N; ++v) {
            if (!isV[v]) continue;
            float best = INFINITY; int best_e = 0; /* np.argmin: first minimum; all-inf -> 0 (cannot happen) */
            for (int i = 0; i < M; ++i) {
                float d = INFINITY;
                if (DST(e, i) == v && ND(e, SRC(e, i), 0) > 0.5f)
                    d = ND(e, SRC(e, i), 3) + ED(e, i, 0); /* float32 + float32 */
                if (d < best) { best = d; best_e = i; }
            }
            mask[best_e] = 1;
        }
    }
}

static void mask_dc(const oenv *e, uint8_t *mask) { /* distribution_center.py:129-141 */
    int N = e->N;
    memset(mask, 0, (size_t)N);
    if (e->parenting == 2) {
        for (int t = 0; t < e->n_targets; ++t) {
            int tv = e->targets[t];
            if (ND(e, tv, 2) == 1.0f && ND(e, tv, 3) == 0.0f)
                for (int v = 0; v < N; ++v) if (e->in_range[(size_t)t * N + v]) mask[v] = 1;
        }
    } else memset(mask, 1, (size_t)N);
    for (int v = 0; v < N; ++v) if (ND(e, v, 1) == 1.0f) mask[v] = 0;
}

static void mask_ppd(const oenv *e, uint8_t *mask) { /* perishable_product_delivery.py:175-196 (parenting must be 1) */
    memset(mask, 0, (size_t)e->N);
    for (int i = 0; i < e->n_dests; ++i)
        if (ND(e, e->head, 1 + i) == 1.0f) mask[e->head] = 1;               /* a product waits at the head: "pick up" = head */
    for (int i = e->row_ptr[e->head]; i < e->row_ptr[e->head + 1]; ++i) mask[DST(e, i)] = 1;
}

int oenv_mask_len(const oenv *e) { return (e->kind == K_ST || e->kind == K_MC) ? e->M : e->N; }

void oenv_mask(const oenv *e, uint8_t *mask) {
    switch (e->kind) {
    case K_SP: mask_sp(e, mask); break;
    case K_LP: mask_lp(e, mask); break;
    case K_ST: mask_st(e, mask); break;
    case K_TSP: mask_tsp(e, mask); break;
    case K_MIS: mask_mis(e, mask); break;
    case K_DS: mask_ds(e, mask); break;
    case K_MC: mask_mc(e, mask); break;
    case K_DC: mask_dc(e, mask); break;
    case K_PPD: mask_ppd(e, mask); break;
    }
}

static int mask_count(const oenv *e, uint8_t *mask) {
    int c = 0, n = oenv_mask_len(e);
    for (int i = 0; i < n; ++i) c += mask[i];
    return c;
}

/* =============================== construction =============================== */

/* Builds the env exactly as the tail of each reference reset() leaves it, from an exported
 * instance: links (reference order), w64, per-kind terminals/params, features float32[N,5]. */
oenv *oenv_create(int kind, int N, int M, int parenting, const int32_t *links, const double *w64,
                  const float *features /* N*5 or NULL */, const int32_t *iparams, const double *dparams,
                  const int32_t *dests /* n_dests or n_targets */, const double *node_cost /* N or NULL */,
                  const double *node_xy /* 2N or NULL */) {
    /* iparams: [src, dest, n_dests, n_choices, n_targets, eval]   dparams: [max_distance, heuristic] */
    oenv *e = calloc(1, sizeof(oenv));
    e->kind = kind; e->N = N; e->M = M; e->parenting = parenting;
    e->F = dyn_cols(kind) + NSTRUCT; e->Fe = edge_cols(kind);
    e->nodes = calloc((size_t)N * e->F, sizeof(float));
    e->edges = calloc((size_t)M * e->Fe, sizeof(float));
    e->links = malloc(sizeof(int32_t) * 2 * (size_t)M);
    e->w64 = malloc(sizeof(double) * (size_t)M);
    e->row_ptr = calloc((size_t)N + 1, sizeof(int32_t));
    e->alt = calloc((size_t)N, 1);
    e->taken_set = calloc((size_t)N, 1);
    e->scratch = calloc(4 * (size_t)N + (size_t)M + 16, 1);
    e->queue = malloc(sizeof(int32_t) * (size_t)N);
    memcpy(e->links, links, sizeof(int32_t) * 2 * (size_t)M);
    memcpy(e->w64, w64, sizeof(double) * (size_t)M);
    for (int i = 0; i < M; ++i) e->row_ptr[SRC(e, i) + 1]++;
    for (int v = 0; v < N; ++v) e->row_ptr[v + 1] += e->row_ptr[v];
    e->src = iparams[0]; e->dest = iparams[1]; e->n_dests = iparams[2]; e->n_choices = iparams[3];
    e->n_targets = iparams[4];
    e->max_distance = dparams[0]; e->heuristic = dparams[1];
    int dc = dyn_cols(kind);
    if (features)
        for (int v = 0; v < N; ++v)
            for (int c = 0; c < NSTRUCT; ++c) ND(e, v, dc + c) = features[(size_t)v * NSTRUCT + c];
    for (int i = 0; i < M; ++i) ED(e, i, 0) = (kind == K_MIS || kind == K_DS) ? 1.0f : (float)w64[i];
    switch (kind) {
    case K_SP: case K_LP:
        ND(e, e->src, 0) = 1; ND(e, e->dest, 1) = 1;
        if (kind == K_LP && parenting == 0) ND(e, e->src, 1) = 2;       /* longest_path.py:85-86 */
        if (kind == K_LP && parenting >= 2) { memset(e->alt, 1, (size_t)N); e->alt[e->src] = 0; }
        e->head = e->src;
        break;
    case K_ST:
        ND(e, e->src, 0) = 1;
        for (int i = 0; i < e->n_dests; ++i) ND(e, dests[i], 1) = 1;
        break;
    case K_TSP:
        e->start = 0; e->head = 0;
        if (node_xy) for (int v = 0; v < N; ++v) { ND(e, v, 2) = (float)node_xy[2 * v]; ND(e, v, 3) = (float)node_xy[2 * v + 1]; }
        ND(e, 0, 1) = 1;
        if (parenting >= 2) { memset(e->alt, 1, (size_t)N); e->alt[0] = 0; }
        break;
    case K_MIS:
        for (int v = 0; v < N; ++v) ND(e, v, 0) = (float)node_cost[v];
        break;
    case K_DS: break;
    case K_MC:
        e->src = 0;
        ND(e, 0, 0) = 1;
        for (int i = 0; i < e->n_dests; ++i) ND(e, dests[i], 1) = 1;
        for (int v = 0; v < N; ++v) { ND(e, v, 2) = (float)e->max_distance; ND(e, v, 3) = -1; }
        ND(e, 0, 3) = 0;
        break;
    case K_PPD:  /* perishable_product_delivery.py:120-133; dests = pickups[n] then dropoffs[n]; max_distance = delivery_time */
        e->targets = malloc(sizeof(int32_t) * 2 * PPD_MAXP);
        for (int i = 0; i < e->n_dests; ++i) {
            e->targets[i] = dests[i]; e->targets[PPD_MAXP + i] = dests[e->n_dests + i];
            ND(e, dests[i], 1 + i) = 1;
            ND(e, dests[e->n_dests + i], 1 + PPD_MAXP + i) = 1;
            for (int v = 0; v < N; ++v) ND(e, v, 1 + 2 * PPD_MAXP + i) = (float)e->max_distance;
        }
        e->head = 0;
        ND(e, 0, 0) = 1;
        break;
    case K_DC: {
        e->targets = malloc(sizeof(int32_t) * (size_t)(e->n_targets > 0 ? e->n_targets : 1));
        e->in_range = calloc((size_t)(e->n_targets > 0 ? e->n_targets : 1) * N, 1);
        double *dist = malloc(sizeof(double) * (size_t)N);
        for (int v = 0; v < N; ++v) { ND(e, v, 0) = (float)node_cost[v]; ND(e, v, 4) = (float)e->max_distance; }
        for (int t = 0; t < e->n_targets; ++t) {
            e->targets[t] = dests[t];
            ND(e, dests[t], 2) = 1;
            dijkstra64(e, dests[t], e->max_distance, 1, dist); /* distribution_center.py:113-116 */
            for (int v = 0; v < N; ++v) e->in_range[(size_t)t * N + v] = dist[v] <= e->max_distance;
        }
        free(dist);
        break; }
    }
    return e;
}

int oenv_obs_len(const oenv *e) { return e->N * e->F + e->M * e->Fe + 2 * e->M; }
void oenv_obs(const oenv *e, float *out) { /* utils.py:87-88 */
    size_t p = 0;
    memcpy(out, e->nodes, sizeof(float) * (size_t)e->N * e->F); p += (size_t)e->N * e->F;
    memcpy(out + p, e->edges, sizeof(float) * (size_t)e->M * e->Fe); p += (size_t)e->M * e->Fe;
    for (int i = 0; i < 2 * e->M; ++i) out[p + i] = (float)e->links[i];
}
const float *oenv_nodes(const oenv *e) { return e->nodes; }
const float *oenv_edges(const oenv *e) { return e->edges; }
int oenv_F(const oenv *e) { return e->F; }
int oenv_Fe(const oenv *e) { return e->Fe; }
int oenv_done(const oenv *e) { return e->done; }
double oenv_in_range(const oenv *e, int t, int v) { return e->in_range[(size_t)t * e->N + v]; }

/* =============================== steps =============================== */
#define FAIL(r) do { (r)->status = 1; return; } while (0)

static void step_sp(oenv *e, int a, ostep *r, uint8_t *mask) { /* shortest_path.py:111-141 */
    if (!(a >= 0 && a < e->N)) FAIL(r);
    mask_sp(e, mask);
    if (!mask[a]) FAIL(r);
    double reward = -adj_lookup(e, e->head, a);
    e->cost -= reward;
    if (ND(e, a, 1) == 1.0f) { r->done = 1; r->solved = 1; }
    ND(e, a, 0) = 1; e->head = a;
    mask_sp(e, mask); r->has_mask = 1;
    if (!r->done && mask_count(e, mask) == 0) { r->done = 1; reward = -(double)e->N; r->solved = 0; }
    if (r->done) { r->heuristic = e->heuristic; r->solution_cost = e->cost; }
    r->reward = reward;
}

static void step_lp(oenv *e, int a, ostep *r, uint8_t *mask) { /* longest_path.py:147-196 */
    if (!(a >= 0 && a < e->N)) FAIL(r);
    mask_lp(e, mask);
    if (!mask[a]) FAIL(r);
    int nb = is_neighbor(e, e->head, a), vis = (ND(e, a, 0) == 1.0f);
    if (e->parenting >= 1 && (!nb || vis)) FAIL(r);
    double reward = adj_lookup(e, e->head, a);
    e->cost -= reward;
    r->heuristic = e->heuristic; r->solution_cost = e->cost;
    if (!nb || vis) { r->done = 1; r->solved = 0; r->reward = -2.0 * e->N; r->has_mask = 0; return; }
    e->head = a; ND(e, a, 0) = 1;
    if (ND(e, a, 1) == 1.0f) { r->done = 1; r->solved = 1; }
    if (e->parenting >= 2) e->alt[a] = 0;
    mask_lp(e, mask); r->has_mask = 1;
    if (!r->done && mask_count(e, mask) == 0) { r->done = 1; reward = -2.0 * e->N; r->solved = 0; }
    r->reward = reward;
}

static void step_st(oenv *e, int a, ostep *r, uint8_t *mask) { /* steiner_tree.py:123-157 */
    if (!(a >= 0 && a < e->M)) FAIL(r);
    int v = DST(e, a);
    mask_st(e, mask);
    if (!mask[a]) FAIL(r);
    float reward = -ED(e, a, 0);
    e->cost32 -= reward;
    ND(e, v, 0) = 1;
    int left = 0;
    for (int x = 0; x < e->N; ++x) left += (ND(e, x, 0) == 0.0f && ND(e, x, 1) == 1.0f);
    if (left == 0) r->done = 1;
    mask_st(e, mask); r->has_mask = 1;
    if (r->done) { r->heuristic = e->heuristic; r->solved = 1; r->solution_cost = e->cost32; }
    r->reward = reward;
}

static void step_tsp(oenv *e, int a, ostep *r, uint8_t *mask) { /* tsp.py:201-258 */
    if (a == e->start && e->head == e->start) {
        r->done = 1; r->reward = -(double)e->N; r->solved = 0; r->heuristic = e->heuristic; r->solution_cost = -1;
        mask_tsp(e, mask); r->has_mask = 1;
        return;
    }
    if (!(a >= 0 && a < e->N)) FAIL(r);
    mask_tsp(e, mask);
    if (!mask[a]) FAIL(r);
    e->steps_taken++;
    double w = adj_lookup(e, e->head, a);
    double reward = 0 - w;
    e->cost += w;
    ND(e, a, 0) = 1;
    if (e->parenting >= 2 && a != e->start) e->alt[a] = 0;
    e->head = a;
    int any_untaken = 0;
    for (int v = 0; v < e->N; ++v) any_untaken |= (fabsf(ND(e, v, 0)) <= 1e-8f); /* np.isclose(x, 0) */
    if (!any_untaken && a == e->start) { r->done = 1; r->solved = 1; }
    mask_tsp(e, mask); r->has_mask = 1;
    if (!r->done && mask_count(e, mask) == 0) { r->done = 1; reward -= e->N * 2; r->solved = 0; }
    if (r->done) { r->heuristic = e->heuristic; r->solution_cost = e->cost; }
    r->reward = reward;
}

static void step_mis(oenv *e, int a, ostep *r, uint8_t *mask) { /* max_independent_set.py:102-124 */
    if (!(a >= 0 && a < e->N)) FAIL(r);
    mask_mis(e, mask);
    if (!mask[a]) FAIL(r);
    float reward = -ND(e, a, 0);
    e->cost32 -= reward;
    ND(e, a, 1) = 1;
    mask_mis(e, mask); r->has_mask = 1;
    if (mask_count(e, mask) == 0) { r->done = 1; r->solved = 1; }
    if (r->done) { r->heuristic = e->heuristic; r->solution_cost = e->cost32; }
    r->reward = reward;
}

static void step_ds(oenv *e, int a, ostep *r, uint8_t *mask) { /* densest_subgraph.py:135-196 */
    if (!(a >= 0 && a < e->N)) FAIL(r);
    mask_ds(e, mask);
    if (!mask[a]) FAIL(r);
    r->heuristic = e->heuristic; r->solved = 1;
    if (a == e->N - 1) {
        r->reward = 0; r->done = 1; r->solution_cost = e->cost; mask_ds(e, mask); r->has_mask = 1;
        return;
    }
    long new_edges = 0;
    for (int i = 0; i < e->M; ++i) if (SRC(e, i) == a && e->taken_set[DST(e, i)]) new_edges++;
    double reward;
    if (e->k_taken == 0) reward = 0;
    else reward = ((double)(e->edge_cnt + new_edges) / (double)(e->k_taken + 1)) - ((double)e->edge_cnt / (double)e->k_taken);
    e->edge_cnt += new_edges;
    e->taken_set[a] = 1; e->k_taken++;
    ND(e, a, 0) = 1;
    e->cost = (double)e->edge_cnt / (double)e->k_taken;
    mask_ds(e, mask); r->has_mask = 1;
    if (e->k_taken == e->n_choices) r->done = 1;
    if (r->done) r->solution_cost = e->cost;
    r->reward = reward;
}

static void step_mc(oenv *e, int a, ostep *r, uint8_t *mask) { /* multicast_routing.py:191-266 */
    if (!(a >= 0 && a < e->M)) FAIL(r);
    int u = SRC(e, a), v = DST(e, a);
    mask_mc(e, mask);
    if (!mask[a]) FAIL(r);
    float penalty = (float)(-2 * e->N * e->n_dests);
    float reward = -ED(e, a, 0);
    r->heuristic = e->heuristic; r->solution_cost = -1;
    e->cost32 -= reward;
    if (ND(e, u, 0) == 0.0f || ND(e, v, 0) == 1.0f) { /* only reachable with parenting<=1 */
        r->reward = penalty; r->done = 1; r->solved = 0; mask_mc(e, mask); r->has_mask = 1;
        return;
    }
    ND(e, v, 0) = 1; ED(e, a, 1) = 1;
    ND(e, v, 3) = ND(e, u, 3) + ED(e, a, 0); /* float32 add */
    if (ND(e, v, 1) == 1.0f) {
        volatile float lim = ND(e, v, 2) + 1e-4f; /* float32 under numpy 2 (NEP 50), multicast_routing.py:232 */
        if (ND(e, v, 3) > lim) {
            r->reward = penalty; r->done = 1; r->solved = 0; mask_mc(e, mask); r->has_mask = 1;
            return;
        }
        reward += 1;
        e->constraints++;
    }
    int left = 0;
    for (int x = 0; x < e->N; ++x) left += (ND(e, x, 0) < 1e-5f && ND(e, x, 1) > (1 - 1e-5f));
    mask_mc(e, mask); r->has_mask = 1;
    if (left == 0) { r->done = 1; r->solved = 1; }
    else if (mask_count(e, mask) == 0) { r->reward = penalty; r->done = 1; r->solved = 0; return; }
    if (r->done) { r->heuristic = e->heuristic; r->solution_cost = e->cost32; }
    r->reward = reward;
}

static void step_dc(oenv *e, int a, ostep *r, uint8_t *mask) { /* distribution_center.py:144-174 */
    if (!(a >= 0 && a < e->N)) FAIL(r);
    mask_dc(e, mask);
    if (!mask[a]) FAIL(r);
    float reward = -ND(e, a, 0);
    e->cost32 -= reward;
    ND(e, a, 1) = 1;
    double *dist = malloc(sizeof(double) * (size_t)e->N);
    dijkstra64(e, a, e->max_distance, 1, dist); /* find_nodes_in_range, distribution_center.py:25-26 */
    for (int nd = 0; nd < e->N; ++nd) {
        if (!(dist[nd] <= e->max_distance)) continue;
        if (ND(e, nd, 3) == 1.0f) continue;
        ND(e, nd, 3) = 1;
        if (ND(e, nd, 2) == 1.0f) reward += 1;
    }
    free(dist);
    mask_dc(e, mask); r->has_mask = 1;
    int left = 0;
    for (int v = 0; v < e->N; ++v) left += (ND(e, v, 2) == 1.0f && ND(e, v, 3) == 0.0f);
    if (left == 0) { r->done = 1; r->solved = 1; }
    if (r->done) { r->heuristic = e->heuristic; r->solution_cost = e->cost32; }
    r->reward = reward;
}

static void step_ppd(oenv *e, int a, ostep *r, uint8_t *mask) { /* perishable_product_delivery.py:198-271 */
    const int N = e->N, P = e->n_dests;
    if (!(a >= 0 && a < N)) FAIL(r);
    mask_ppd(e, mask);
    if (!mask[a]) FAIL(r);
    double reward = 0;
    r->heuristic = e->heuristic;            /* info every step (:210-212); solution_cost is the value BEFORE this move */
    r->solution_cost = e->cost;
    if (a == e->head) {                     /* pick up (:214-221): the one product waiting here goes into transit */
        int prod = -1;
        for (int i = 0; i < PPD_MAXP; ++i) if (ND(e, e->head, 1 + i) == 1.0f) { prod = i; break; }
        for (int v = 0; v < N; ++v) ND(e, v, 1 + prod) = -1;
        reward += 2;
        e->steps_taken++;
    } else {                                /* move (:223-249) */
        reward = -adj_lookup(e, e->head, a);
        e->cost -= reward;
        e->steps_taken++;
        ND(e, e->head, 0) = 0; ND(e, a, 0) = 1;
        e->head = a;
        for (int i = 0; i < P; ++i) {
            if (ND(e, e->head, 1 + i) == -1.0f) {
                const float dt = (float)adj_lookup(e, e->head, a);   /* self.adj[self.head, action] AFTER head = action: adj[a, a] = 0 (:234) */
                float sum = 0;                                       /* numpy float32 pairwise sum; only its sign matters */
                for (int v = 0; v < N; ++v) { ND(e, v, 1 + 2 * PPD_MAXP + i) -= dt; sum += ND(e, v, 1 + 2 * PPD_MAXP + i); }
                if (sum < 0 - 1e-6) {                                /* :236-241 early return, no mask */
                    r->done = 1; r->reward = -2.0 * N * P; r->solved = 0; r->has_mask = 0;
                    return;
                }
                if (ND(e, e->head, 1 + PPD_MAXP + i) == 1.0f) {      /* delivered (:243-249) */
                    reward += 2;
                    for (int v = 0; v < N; ++v) { ND(e, v, 1 + i) = 0; ND(e, v, 1 + PPD_MAXP + i) = 0; ND(e, v, 1 + 2 * PPD_MAXP + i) = 0; }
                }
            }
        }
    }
    double has = 0;
    for (int v = 0; v < N; ++v) for (int i = 0; i < PPD_MAXP; ++i) has += ND(e, v, 1 + i);
    if (has == 0) { r->done = 1; r->solved = 1; reward += 2.0 * N; }                                   /* :252-255 */
    else if (e->steps_taken >= N * P * 50) { r->done = 1; r->solved = 0; reward = -2.0 * N * P; }      /* :256-259 */
    mask_ppd(e, mask); r->has_mask = 1;
    r->reward = reward;
}

/* mask_out receives info['mask'] when has_mask (length oenv_mask_len). */
void oenv_step(oenv *e, int action, ostep *r, uint8_t *mask_out) {
    r->reward = 0; r->solution_cost = NAN; r->heuristic = NAN;
    r->done = 0; r->solved = -1; r->has_mask = 0; r->status = 0;
    switch (e->kind) {
    case K_SP: step_sp(e, action, r, mask_out); break;
    case K_LP: step_lp(e, action, r, mask_out); break;
    case K_ST: step_st(e, action, r, mask_out); break;
    case K_TSP: step_tsp(e, action, r, mask_out); break;
    case K_MIS: step_mis(e, action, r, mask_out); break;
    case K_DS: step_ds(e, action, r, mask_out); break;
    case K_MC: step_mc(e, action, r, mask_out); break;
    case K_DC: step_dc(e, action, r, mask_out); break;
    case K_PPD: step_ppd(e, action, r, mask_out); break;
    }
    if (r->status == 0 && r->done) e->done = 1;
}

/* Re-run the state part of reset() on the same instance (used by rollouts with auto-reset). */
void oenv_reset_state(oenv *e) {
    int N = e->N, dc = dyn_cols(e->kind);
    e->done = 0; e->cost = 0; e->cost32 = 0; e->edge_cnt = 0; e->k_taken = 0; e->steps_taken = 0; e->constraints = 0;
    memset(e->taken_set, 0, (size_t)N);
    switch (e->kind) {
    case K_SP: case K_LP:
        for (int v = 0; v < N; ++v) ND(e, v, 0) = 0;
        ND(e, e->src, 0) = 1; e->head = e->src;
        if (e->kind == K_LP && e->parenting >= 2) { memset(e->alt, 1, (size_t)N); e->alt[e->src] = 0; }
        break;
    case K_ST:
        for (int v = 0; v < N; ++v) ND(e, v, 0) = 0;
        ND(e, e->src, 0) = 1;
        break;
    case K_TSP:
        for (int v = 0; v < N; ++v) ND(e, v, 0) = 0;
        e->head = e->start;
        if (e->parenting >= 2) { memset(e->alt, 1, (size_t)N); e->alt[e->start] = 0; }
        break;
    case K_MIS: for (int v = 0; v < N; ++v) ND(e, v, 1) = 0; break;
    case K_DS: for (int v = 0; v < N; ++v) ND(e, v, 0) = 0; break;
    case K_MC:
        for (int v = 0; v < N; ++v) { ND(e, v, 0) = 0; ND(e, v, 3) = -1; }
        ND(e, 0, 0) = 1; ND(e, 0, 3) = 0;
        for (int i = 0; i < e->M; ++i) ED(e, i, 1) = 0;
        break;
    case K_DC: for (int v = 0; v < N; ++v) { ND(e, v, 1) = 0; ND(e, v, 3) = 0; } break;
    case K_PPD:
        for (int v = 0; v < N; ++v) for (int c = 0; c < 1 + 3 * PPD_MAXP; ++c) ND(e, v, c) = 0;
        for (int i = 0; i < e->n_dests; ++i) {
            ND(e, e->targets[i], 1 + i) = 1;
            ND(e, e->targets[PPD_MAXP + i], 1 + PPD_MAXP + i) = 1;
            for (int v = 0; v < N; ++v) ND(e, v, 1 + 2 * PPD_MAXP + i) = (float)e->max_distance;
        }
        e->head = 0; ND(e, 0, 0) = 1;
        break;
    }
    (void)dc;
}

/* =============================== structural features =============================== */
/* feature_extraction.py:6-37 on the DIRECTED symmetric graph; values in float64, caller rounds
 * to float32 like torch.tensor(sf) does.  out[v*5 + {0..4}] = degree, betweenness, closeness,
 * pagerank, clustering.  `weighted_pr`: TSP stores its weights under the attribute 'weight',
 * which nx.pagerank picks up (tsp.py:90); every other env -> unweighted. */
int oenv_features(const oenv *e, int weighted_pr, double *out) {
    int N = e->N, M = e->M;
    const int32_t *rp = e->row_ptr;
    int *Q = malloc(sizeof(int) * (size_t)N), *D = malloc(sizeof(int) * (size_t)N);
    double *sigma = malloc(sizeof(double) * (size_t)N), *delta = malloc(sizeof(double) * (size_t)N);
    double *bt = calloc((size_t)N, sizeof(double));
    for (int v = 0; v < N; ++v) out[5 * v + 0] = 2.0 * (rp[v + 1] - rp[v]); /* G.degree = in + out */
    for (int s = 0; s < N; ++s) {
        /* nx:centrality/betweenness.py _single_source_shortest_path_basic + _accumulate_basic */
        for (int v = 0; v < N; ++v) { D[v] = -1; sigma[v] = 0.0; delta[v] = 0.0; }
        int qh = 0, qt = 0;
        sigma[s] = 1.0; D[s] = 0; Q[qt++] = s;
        while (qh < qt) {
            int v = Q[qh++];
            for (int i = rp[v]; i < rp[v + 1]; ++i) {
                int w = DST(e, i);
                if (D[w] < 0) { Q[qt++] = w; D[w] = D[v] + 1; }
                if (D[w] == D[v] + 1) sigma[w] += sigma[v];
            }
        }
        /* closeness (nx:centrality/closeness.py) shares the BFS: reversed graph == graph (symmetric) */
        double totsp = 0; for (int i = 0; i < qt; ++i) totsp += D[Q[i]];
        double cc = 0.0;
        if (totsp > 0.0 && N > 1) { cc = (qt - 1.0) / totsp; cc *= (qt - 1.0) / (N - 1); }
        out[5 * s + 2] = cc;
        for (int i = qt - 1; i >= 0; --i) {
            int w = Q[i];
            double coeff = (1 + delta[w]) / sigma[w];
            /* P[w] = neighbours one BFS level up (graph is symmetric); each v in P[w] updates its
             * own delta[v], so only the (reverse-BFS) order over w matters for the sums. */
            for (int k = rp[w]; k < rp[w + 1]; ++k) {
                int v = DST(e, k);
                if (D[v] + 1 == D[w]) delta[v] += sigma[v] * coeff;
            }
            if (w != s) bt[w] += delta[w];
        }
    }
    if (N - 1 >= 2) { double scale = 1.0 / ((double)(N - 1) * (N - 2)); for (int v = 0; v < N; ++v) bt[v] *= scale; }
    for (int v = 0; v < N; ++v) out[5 * v + 1] = bt[v];
    /* pagerank: nx:link_analysis/pagerank_alg.py _pagerank_scipy, alpha .85, tol 1e-6, <=100 it */
    {
        double *S = calloc((size_t)N, sizeof(double)), *A = malloc(sizeof(double) * (size_t)M);
        double *x = malloc(sizeof(double) * (size_t)N), *y = malloc(sizeof(double) * (size_t)N);
        for (int i = 0; i < M; ++i) S[SRC(e, i)] += weighted_pr ? e->w64[i] : 1.0;
        for (int v = 0; v < N; ++v) if (S[v] != 0) S[v] = 1.0 / S[v];
        for (int i = 0; i < M; ++i) A[i] = S[SRC(e, i)] * (weighted_pr ? e->w64[i] : 1.0);
        double p = 1.0 / N, alpha = 0.85;
        for (int v = 0; v < N; ++v) x[v] = p;
        int it, ok = 0;
        for (it = 0; it < 100; ++it) {
            double dsum = 0; for (int v = 0; v < N; ++v) if (S[v] == 0) dsum += x[v];
            for (int v = 0; v < N; ++v) y[v] = 0;
            for (int i = 0; i < M; ++i) y[DST(e, i)] += A[i] * x[SRC(e, i)]; /* csc order: ascending source row */
            double err = 0;
            for (int v = 0; v < N; ++v) {
                double nx_ = alpha * (y[v] + dsum * p) + (1 - alpha) * p;
                err += fabs(nx_ - x[v]); y[v] = nx_;
            }
            double *t = x; x = y; y = t;
            if (err < N * 1.0e-6) { ok = 1; break; }
        }
        for (int v = 0; v < N; ++v) out[5 * v + 3] = x[v];
        free(S); free(A); free(x); free(y);
        if (!ok) { free(Q); free(D); free(sigma); free(delta); free(bt); return -1; }
    }
    /* clustering: nx:algorithms/cluster.py directed Fagiolo form; on a symmetric digraph
     * t = 8*S, S = sum_{j in N(i)} |N(i) & N(j)|, denom = 2*(dt(dt-1) - 2*db), dt = 2d, db = d. */
    {
        uint8_t *mark = calloc((size_t)N, 1);
        for (int i = 0; i < N; ++i) {
            long Ssum = 0; long d = rp[i + 1] - rp[i];
            for (int k = rp[i]; k < rp[i + 1]; ++k) mark[DST(e, k)] = 1;
            for (int k = rp[i]; k < rp[i + 1]; ++k) {
                int j = DST(e, k);
                for (int q = rp[j]; q < rp[j + 1]; ++q) Ssum += mark[DST(e, q)];
            }
            for (int k = rp[i]; k < rp[i + 1]; ++k) mark[DST(e, k)] = 0;
            long t = 8 * Ssum, dt = 2 * d, db = d;
            out[5 * i + 4] = (t == 0) ? 0.0 : (double)t / (double)((dt * (dt - 1) - 2 * db) * 2);
        }
        free(mark);
    }
    free(Q); free(D); free(sigma); free(delta); free(bt);
    return 0;
}

/* =============================== batched rollouts (CPU baseline) =============================== */
/* Same counter-based action sampler as the CUDA engine (graphenvs_b200/csrc): uniform over the
 * valid mask bits, r = mix(seed, env, t).  Kept bit-identical so full-size trajectories can be
 * compared through checksums. */
static inline uint32_t ge_mix(uint64_t seed, uint32_t env, uint32_t t) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * ((uint64_t)env * 0x100000001ull + (((uint64_t)t) << 32 | 0x5bd1e995u));
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}
int oenv_sample_action(const uint8_t *mask, int n, uint64_t seed, uint32_t env, uint32_t t) {
    int cnt = 0;
    for (int i = 0; i < n; ++i) cnt += mask[i];
    if (cnt == 0) return -1;
    uint32_t r = (uint32_t)(((uint64_t)ge_mix(seed, env, t) * (uint64_t)cnt) >> 32);
    for (int i = 0; i < n; ++i) if (mask[i]) { if (r == 0) return i; r--; }
    return -1;
}

#ifdef _OPENMP
#include <omp.h>
#endif
/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline sets its team size explicitly. */
int oenv_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* Steps every env `n_steps` times with the sampler above and auto-reset-on-done.
 * Outputs (per env): steps counted, episodes finished, sum of rewards, xor-rotate checksum of
 * (action, done, solved) to compare with the device path.  Uses OpenMP over envs. */
void oenv_rollout(oenv **envs, int n_envs, int env_id0, int n_steps, uint64_t seed, int t0,
                  double *sum_reward, int64_t *episodes, uint64_t *checksum) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int b = 0; b < n_envs; ++b) {
        oenv *e = envs[b];
        int n = oenv_mask_len(e);
        uint8_t *mask = malloc((size_t)n);
        oenv_mask(e, mask);
        if (e->kind == K_TSP && mask_count(e, mask) == 0) mask[e->start] = 1;
        double sr = 0; int64_t ep = 0; uint64_t cs = 0;
        for (int t = 0; t < n_steps; ++t) {
            int a = oenv_sample_action(mask, n, seed, (uint32_t)(env_id0 + b), (uint32_t)(t0 + t));
            ostep r;
            oenv_step(e, a, &r, mask);
            sr += r.reward;
            cs = ((cs << 7) | (cs >> 57)) ^ (uint64_t)(uint32_t)a ^ ((uint64_t)r.done << 40) ^ ((uint64_t)(r.solved & 3) << 44) ^ ((uint64_t)r.status << 48);
            if (r.done) {
                ep++;
                oenv_reset_state(e);
                oenv_mask(e, mask);
                if (e->kind == K_TSP && mask_count(e, mask) == 0) mask[e->start] = 1;
            }
        }
        sum_reward[b] = sr; episodes[b] = ep; checksum[b] = cs;
        free(mask);
    }
}
