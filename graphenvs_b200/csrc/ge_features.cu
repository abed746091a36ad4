// ge_features.cu -- feature_extraction.generate_features (feature_extraction.py:6-37) on the GPU.
// Per node, in the reference's column order: G.degree (in+out of the directed symmetric graph),
// betweenness (Brandes, nx:centrality/betweenness.py), closeness (Wasserman-Faust,
// nx:centrality/closeness.py), pagerank (scipy power iteration, nx:link_analysis/pagerank_alg.py,
// alpha .85, tol 1e-6, <= 100 it., same stopping rule), clustering (directed Fagiolo form,
// nx:algorithms/cluster.py).  All arithmetic in fp64, rounded to float32 at the end exactly like
// `torch.tensor(sf)` (feature_extraction.py:36).
//
// One warp per environment.  Betweenness + closeness share one BFS per source; BFS frontiers are
// bitsets, edge work goes through the load-balanced expand_set primitive; sigma / delta use
// shared-memory fp64 atomics (sigma sums are integers < 2^53 => order-independent; delta sums
// differ from networkx only by fp64 reassociation, ~1e-16 relative).
#include "ge_common.cuh"

using namespace ge;

extern "C" int ge_set_error(int code, const char *fmt, ...);

namespace {

struct FScr {
    double *sigma, *delta, *bt;  // [N] each
    int *D;                      // [N]
    uint32_t *f0, *f1, *lvl;     // [NW] each
};

__host__ __device__ inline int feat_words(int N, int NW) { return ((6 * N + N + 3 * NW) + 3) & ~3; }

__device__ inline FScr fcarve(uint32_t *base, int N, int NW) {
    FScr s;
    s.sigma = reinterpret_cast<double *>(base);
    s.delta = s.sigma + N;
    s.bt = s.delta + N;
    s.D = reinterpret_cast<int *>(s.bt + N);
    s.f0 = reinterpret_cast<uint32_t *>(s.D + N);
    s.f1 = s.f0 + NW;
    s.lvl = s.f1 + NW;
    return s;
}

__global__ void __launch_bounds__(GE_WPB * 32) features_kernel(ge_batch d, int words_per_warp, int wpb) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * wpb + warp;
    if (b >= d.B) return;
    const int N = d.N, NW = d.NW;
    FScr s = fcarve(smem + (size_t)warp * words_per_warp, N, NW);
    const int32_t *rp = d.row_ptr + (size_t)b * d.RP;
    const int32_t *col = d.col + (size_t)b * d.MP;
    float *out = d.features + (size_t)b * N * 5;

    for (int v = lane; v < N; v += 32) s.bt[v] = 0.0;
    __syncwarp();

    // ---------------- betweenness + closeness: one BFS per source
    for (int src = 0; src < N; ++src) {
        for (int v = lane; v < N; v += 32) { s.D[v] = -1; s.sigma[v] = 0.0; s.delta[v] = 0.0; }
        for (int w = lane; w < NW; w += 32) { s.f0[w] = 0; s.f1[w] = 0; }
        __syncwarp();
        if (lane == 0) { s.D[src] = 0; s.sigma[src] = 1.0; s.f0[src >> 5] = 1u << (src & 31); }
        __syncwarp();
        int level = 0;
        long long totsp = 0;
        int reached = 1;
        for (;;) {  // forward sweep (_single_source_shortest_path_basic)
            double sg = 0.0;
            expand_set(
                rp, s.f0, NW, lane, [&](int u) { sg = u >= 0 ? s.sigma[u] : 0.0; },
                [&](int owner, int e, bool active) {
                    double su = __shfl_sync(GE_FULL, sg, owner);
                    if (active) {
                        int w = col[e];
                        int old = atomicCAS(&s.D[w], -1, level + 1);
                        if (old == -1 || old == level + 1) {
                            atomicAdd(&s.sigma[w], su);
                            if (old == -1) atomicOr(&s.f1[w >> 5], 1u << (w & 31));
                        }
                    }
                });
            __syncwarp();
            int cnt = 0;
            uint32_t any = 0;
            for (int w = lane; w < NW; w += 32) { uint32_t n = s.f1[w]; s.f0[w] = n; s.f1[w] = 0; any |= n; cnt += __popc(n); }
            cnt = __reduce_add_sync(GE_FULL, cnt);
            __syncwarp();
            if (!__any_sync(GE_FULL, any != 0)) break;
            ++level;
            reached += cnt;
            totsp += (long long)cnt * level;
        }
        if (lane == 0) {  // closeness: ((r-1)/totsp) * ((r-1)/(N-1)), 0 when totsp == 0
            double cc = 0.0;
            if (totsp > 0 && N > 1) {
                cc = ((double)reached - 1.0) / (double)totsp;
                cc *= ((double)reached - 1.0) / (double)(N - 1);
            }
            out[src * 5 + 2] = (float)cc;
        }
        // backward sweep (_accumulate_basic): levels from the deepest up
        for (int L = level; L >= 1; --L) {
            for (int w = lane; w < NW; w += 32) {
                uint32_t bits = 0;
                for (int j = 0; j < 32; ++j) {
                    int v = (w << 5) + j;
                    if (v < N && s.D[v] == L) bits |= 1u << j;
                }
                s.lvl[w] = bits;
            }
            __syncwarp();
            double coeff = 0.0;
            expand_set(
                rp, s.lvl, NW, lane,
                [&](int w) {
                    if (w >= 0) {
                        coeff = (1.0 + s.delta[w]) / s.sigma[w];
                        s.bt[w] += s.delta[w];  // w != src at level >= 1; delta[w] is final here
                    } else coeff = 0.0;
                },
                [&](int owner, int e, bool active) {
                    double c = __shfl_sync(GE_FULL, coeff, owner);
                    if (active) {
                        int v = col[e];
                        if (s.D[v] == L - 1) atomicAdd(&s.delta[v], s.sigma[v] * c);
                    }
                });
            __syncwarp();
        }
    }
    {
        double scale = (N - 1 >= 2) ? 1.0 / ((double)(N - 1) * (double)(N - 2)) : 1.0;  // _rescale, directed, normalized
        for (int v = lane; v < N; v += 32) {
            out[v * 5 + 1] = (float)(s.bt[v] * scale);
            out[v * 5 + 0] = (float)(2 * (rp[v + 1] - rp[v]));
        }
    }
    __syncwarp();

    // ---------------- pagerank (power iteration, scipy formulation)
    {
        double *x = s.sigma, *y = s.delta, *invS = s.bt;
        const bool weighted = (d.flags & GE_FLAG_WEIGHTED_PR) && d.w64;
        const double *w64 = d.w64 ? d.w64 + (size_t)b * d.MP : nullptr;
        const double p = 1.0 / (double)N, alpha = 0.85;
        for (int v = lane; v < N; v += 32) {
            double S = 0.0;
            for (int e = rp[v]; e < rp[v + 1]; ++e) S += weighted ? w64[e] : 1.0;
            invS[v] = S != 0.0 ? 1.0 / S : 0.0;
            x[v] = p;
        }
        __syncwarp();
        for (int it = 0; it < 100; ++it) {
            double dsum = 0.0;
            for (int v = lane; v < N; v += 32) if (invS[v] == 0.0) dsum += x[v];
            for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(GE_FULL, dsum, o);
            double err = 0.0;
            for (int v = lane; v < N; v += 32) {
                double acc = 0.0;  // (x @ A)[v] = sum_u x[u] * (invS[u] * w(u,v)); symmetric => in-neighbours = row v
                for (int e = rp[v]; e < rp[v + 1]; ++e) {
                    int u = col[e];
                    double a = invS[u] * (weighted ? w64[e] : 1.0);
                    acc += a * x[u];
                }
                double nx = alpha * (acc + dsum * p) + (1.0 - alpha) * p;
                err += fabs(nx - x[v]);
                y[v] = nx;
            }
            for (int o = 16; o > 0; o >>= 1) err += __shfl_xor_sync(GE_FULL, err, o);
            __syncwarp();
            double *t = x; x = y; y = t;
            if (err < (double)N * 1.0e-6) break;
        }
        for (int v = lane; v < N; v += 32) out[v * 5 + 3] = (float)x[v];
        __syncwarp();
    }

    // ---------------- clustering: S_i = sum_{j in N(i)} |N(i) & N(j)|, c = S / (d (d-1))
    for (int i = 0; i < N; ++i) {
        int lo = rp[i], hi = rp[i + 1];
        for (int w = lane; w < NW; w += 32) s.lvl[w] = 0;
        __syncwarp();
        for (int e = lo + lane; e < hi; e += 32) { int j = col[e]; atomicOr(&s.lvl[j >> 5], 1u << (j & 31)); }
        __syncwarp();
        long long cnt = 0;
        expand_set(
            rp, s.lvl, NW, lane, [](int) {},
            [&](int, int e, bool active) {
                if (active) cnt += tbit(s.lvl, col[e]);
            });
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(GE_FULL, cnt, o);
        if (lane == 0) {
            long long dg = hi - lo;
            long long t = 8 * cnt, dt = 2 * dg, db = dg;
            out[i * 5 + 4] = (t == 0) ? 0.f : (float)((double)t / (double)((dt * (dt - 1) - 2 * db) * 2));
        }
        __syncwarp();
    }
}

}  // namespace

extern "C" int ge_features(const ge_batch *d, void *stream) {
    if (!d || !d->features) return ge_set_error(GE_ERR_ARG, "ge_features: features buffer is null");
    int wpw = feat_words(d->N, d->NW);
    size_t per_warp = (size_t)wpw * sizeof(uint32_t);
    int wpb = (int)((200 * 1024) / per_warp);
    if (wpb < 1) return ge_set_error(GE_ERR_UNSUPPORTED, "ge_features: N=%d too large for shared scratch", d->N);
    if (wpb > GE_WPB) wpb = GE_WPB;
    size_t smem = per_warp * wpb;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return ge_set_error(GE_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    features_kernel<<<(d->B + wpb - 1) / wpb, wpb * 32, smem, (cudaStream_t)stream>>>(*d, wpw, wpb);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return ge_set_error(GE_ERR_CUDA, "features_kernel launch: %s", cudaGetErrorString(e));
    return GE_OK;
}
