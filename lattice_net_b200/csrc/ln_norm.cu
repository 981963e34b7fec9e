// GroupNorm (+ optional ReLU) over lattice values stored vertex-major [nv x C] -- SURVEY.md section 8f
// rank 2: the normalisation that sits between every two lattice convolutions
// (/root/reference/latticenet_py/lattice/lattice_modules.py:585-614, 935-960 do it with
// unsqueeze/transpose + torch.nn.GroupNorm + ReLU, i.e. two layout copies and four kernels forward).
// Statistics of group g run over (channels of g) x (all vertices), biased variance, like
// torch.nn.GroupNorm on a [1, C, nv] tensor.
//
// Three families, chosen by size (ln_group_norm_fwd / _bwd):
//   * <= 2048 rows, 1..8 channels per group: one CTA per group with the rows cached in registers, one launch each way
//     (the ShapeNet-sized lattices, where launch count and latency are the cost);
//   * scene-sized lattices: row-tiled kernels over all SMs (stats partial -> finalize -> apply), see gn_tiled_*;
//   * anything else (no workspace given, C % 4 != 0, C > 1024): the generic one-CTA-per-group kernels below --
//     mean, variance and the normalised output in a single launch, backward without atomics.
#include "ln_common.cuh"

namespace ln {

constexpr int kGnThreads = 512;

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                       // protects `red` from the previous call
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0f;
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) red[0] = t;
    }
    __syncthreads();
    return red[0];
}

__global__ void __launch_bounds__(kGnThreads)
group_norm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                      int nv_rows, const int* __restrict__ nv_dev, int c, int cpg, float eps, int relu,
                      float* __restrict__ y, float* __restrict__ stats) {
    LN_PDL_ENTRY();
    __shared__ float red[32];
    const int g = blockIdx.x;
    const int c0 = g * cpg;
    // static-shape mode: the tensors have nv_rows rows, only the first *nv_dev are vertices; padding rows
    // take no part in the statistics and come out as zeros
    const int nv = nv_dev ? min(nv_rows, __ldg(nv_dev)) : nv_rows;
    for (long long i = (long long)nv * cpg + threadIdx.x; i < (long long)nv_rows * cpg; i += blockDim.x) {
        const long long v = i / cpg;
        y[v * c + c0 + (int)(i - v * cpg)] = 0.0f;
    }
    const long long m = (long long)nv * cpg;
    float s = 0.0f;
    for (long long i = threadIdx.x; i < m; i += blockDim.x) {
        const long long v = i / cpg;
        const int j = (int)(i - v * cpg);
        s += __ldg(x + v * c + c0 + j);
    }
    const float mean = block_sum(s, red) / (float)m;
    float sq = 0.0f;
    for (long long i = threadIdx.x; i < m; i += blockDim.x) {
        const long long v = i / cpg;
        const int j = (int)(i - v * cpg);
        const float d = __ldg(x + v * c + c0 + j) - mean;
        sq = fmaf(d, d, sq);
    }
    const float var = block_sum(sq, red) / (float)m;
    const float rstd = rsqrtf(var + eps);
    if (threadIdx.x == 0) {
        stats[2 * g] = mean;
        stats[2 * g + 1] = rstd;
    }
    for (long long i = threadIdx.x; i < m; i += blockDim.x) {
        const long long v = i / cpg;
        const int j = (int)(i - v * cpg);
        const int ch = c0 + j;
        float o = fmaf((__ldg(x + v * c + ch) - mean) * rstd, __ldg(gamma + ch), __ldg(beta + ch));
        if (relu) o = fmaxf(o, 0.0f);
        y[v * c + ch] = o;
    }
}

// dy' = dy * [y > 0] (when relu);  xhat = (x - mean) * rstd
// dgamma_c = sum_v dy' xhat ; dbeta_c = sum_v dy'
// dx = rstd * ( dy' gamma - (xhat * ds + db) / m ),  ds = sum dy' gamma xhat, db = sum dy' gamma  (over the group)
__global__ void __launch_bounds__(kGnThreads)
group_norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                      const float* __restrict__ gamma, const float* __restrict__ stats, const float* __restrict__ dx_add,
                      int nv_rows, const int* __restrict__ nv_dev, int c, int cpg, int relu, float* __restrict__ dx,
                      float* __restrict__ dgamma, float* __restrict__ dbeta) {
    LN_PDL_ENTRY();
    __shared__ float red[32];
    const int g = blockIdx.x;
    const int c0 = g * cpg;
    const int nv = nv_dev ? min(nv_rows, __ldg(nv_dev)) : nv_rows;
    for (long long i = (long long)nv * cpg + threadIdx.x; i < (long long)nv_rows * cpg; i += blockDim.x) {
        const long long v = i / cpg;
        const long long o = v * c + c0 + (int)(i - v * cpg);
        dx[o] = dx_add ? __ldg(dx_add + o) : 0.0f;
    }
    const float mean = stats[2 * g], rstd = stats[2 * g + 1];
    const long long m = (long long)nv * cpg;
    float ds = 0.0f, db = 0.0f;
    for (int j = 0; j < cpg; j++) {          // per channel: rows strided over the block
        const int ch = c0 + j;
        float a = 0.0f, b = 0.0f;
        for (int v = threadIdx.x; v < nv; v += blockDim.x) {
            const size_t o = (size_t)v * c + ch;
            float d = __ldg(dy + o);
            if (relu && !(__ldg(y + o) > 0.0f)) d = 0.0f;
            a = fmaf(d, (__ldg(x + o) - mean) * rstd, a);
            b += d;
        }
        a = block_sum(a, red);
        b = block_sum(b, red);
        if (threadIdx.x == 0) {
            dgamma[ch] = a;
            dbeta[ch] = b;
        }
        const float gm = __ldg(gamma + ch);
        ds = fmaf(a, gm, ds);
        db = fmaf(b, gm, db);
    }
    const float inv_m = 1.0f / (float)m;
    for (long long i = threadIdx.x; i < m; i += blockDim.x) {
        const long long v = i / cpg;
        const int j = (int)(i - v * cpg);
        const size_t o = (size_t)v * c + c0 + j;
        float d = __ldg(dy + o);
        if (relu && !(__ldg(y + o) > 0.0f)) d = 0.0f;
        const float xhat = (__ldg(x + o) - mean) * rstd;
        dx[o] = rstd * (d * __ldg(gamma + c0 + j) - (xhat * ds + db) * inv_m) + (dx_add ? __ldg(dx_add + o) : 0.0f);
    }
}


// ---------------------------------------------------------------------------------------------
// Low-latency variants for the lattice sizes where a group's slice (nv x CPG floats) fits the registers of one
// CTA: every row is read from global memory ONCE, the passes (mean, centred variance, normalise / the backward
// sums and dx) run on the register copy, and all per-channel sums of a pass share one reduction round.
constexpr int kGnRows = 4;     // rows cached per thread: nv <= kGnRows * kGnThreads = 2048

template <int N>
__device__ __forceinline__ void block_sum_n(float (&v)[N], float* red /* [32][N] */) {
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    __syncthreads();                       // protects `red` from the previous round
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < N; i++) red[warp * N + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; i++) {
        float t = 0.0f;
        for (int w = 0; w < nwarps; w++) t += red[w * N + i];   // same order in every thread: identical results
        v[i] = t;
    }
}

template <int CPG>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&r)[CPG]) {
    if (CPG % 4 == 0) {
#pragma unroll
        for (int k = 0; k < CPG; k += 4) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(p + k));
            r[k] = q.x; r[k + 1] = q.y; r[k + 2] = q.z; r[k + 3] = q.w;
        }
    } else if (CPG % 2 == 0) {
#pragma unroll
        for (int k = 0; k < CPG; k += 2) {
            const float2 q = __ldg(reinterpret_cast<const float2*>(p + k));
            r[k] = q.x; r[k + 1] = q.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < CPG; k++) r[k] = __ldg(p + k);
    }
}
template <int CPG>
__device__ __forceinline__ void store_row(float* __restrict__ p, const float (&r)[CPG]) {
    if (CPG % 4 == 0) {
#pragma unroll
        for (int k = 0; k < CPG; k += 4) *reinterpret_cast<float4*>(p + k) = make_float4(r[k], r[k + 1], r[k + 2], r[k + 3]);
    } else if (CPG % 2 == 0) {
#pragma unroll
        for (int k = 0; k < CPG; k += 2) *reinterpret_cast<float2*>(p + k) = make_float2(r[k], r[k + 1]);
    } else {
#pragma unroll
        for (int k = 0; k < CPG; k++) p[k] = r[k];
    }
}

template <int CPG>
__global__ void __launch_bounds__(kGnThreads)
group_norm_fwd_small_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                            int nv_rows, const int* __restrict__ nv_dev, int c, float eps, int relu,
                            float* __restrict__ y, float* __restrict__ stats) {
    LN_PDL_ENTRY();
    __shared__ float red[32 * 1];
    const int g = blockIdx.x;
    const int c0 = g * CPG;
    const int nv = nv_dev ? min(nv_rows, __ldg(nv_dev)) : nv_rows;
    float xr[kGnRows][CPG];
    float s[1] = {0.0f};
#pragma unroll
    for (int i = 0; i < kGnRows; i++) {
        const int v = threadIdx.x + i * kGnThreads;
        if (v < nv) {
            load_row<CPG>(x + (size_t)v * c + c0, xr[i]);
#pragma unroll
            for (int k = 0; k < CPG; k++) s[0] += xr[i][k];
        }
    }
    const float m = (float)nv * (float)CPG;
    block_sum_n<1>(s, red);
    const float mean = s[0] / m;
    float sq[1] = {0.0f};
#pragma unroll
    for (int i = 0; i < kGnRows; i++) {
        const int v = threadIdx.x + i * kGnThreads;
        if (v < nv) {
#pragma unroll
            for (int k = 0; k < CPG; k++) {
                const float d = xr[i][k] - mean;
                sq[0] = fmaf(d, d, sq[0]);
            }
        }
    }
    block_sum_n<1>(sq, red);
    const float rstd = rsqrtf(sq[0] / m + eps);
    if (threadIdx.x == 0) {
        stats[2 * g] = mean;
        stats[2 * g + 1] = rstd;
    }
    float gm[CPG], bt[CPG];
#pragma unroll
    for (int k = 0; k < CPG; k++) {
        gm[k] = __ldg(gamma + c0 + k);
        bt[k] = __ldg(beta + c0 + k);
    }
#pragma unroll
    for (int i = 0; i < kGnRows; i++) {
        const int v = threadIdx.x + i * kGnThreads;
        if (v < nv_rows) {
            float o[CPG];
#pragma unroll
            for (int k = 0; k < CPG; k++) {
                o[k] = 0.0f;                 // padding rows (static-shape mode) come out as zeros
                if (v < nv) {
                    o[k] = fmaf((xr[i][k] - mean) * rstd, gm[k], bt[k]);
                    if (relu) o[k] = fmaxf(o[k], 0.0f);
                }
            }
            store_row<CPG>(y + (size_t)v * c + c0, o);
        }
    }
}

template <int CPG>
__global__ void __launch_bounds__(kGnThreads)
group_norm_bwd_small_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                            const float* __restrict__ gamma, const float* __restrict__ stats, const float* __restrict__ dx_add,
                            int nv_rows, const int* __restrict__ nv_dev, int c, int relu, float* __restrict__ dx,
                            float* __restrict__ dgamma, float* __restrict__ dbeta) {
    LN_PDL_ENTRY();
    __shared__ float red[32 * 2 * CPG];
    const int g = blockIdx.x;
    const int c0 = g * CPG;
    const int nv = nv_dev ? min(nv_rows, __ldg(nv_dev)) : nv_rows;
    const float mean = stats[2 * g], rstd = stats[2 * g + 1];
    float dr[kGnRows][CPG], xh[kGnRows][CPG];
    float ab[2 * CPG];                      // [0..CPG): sum dy' xhat (dgamma), [CPG..2CPG): sum dy' (dbeta)
#pragma unroll
    for (int k = 0; k < 2 * CPG; k++) ab[k] = 0.0f;
#pragma unroll
    for (int i = 0; i < kGnRows; i++) {
        const int v = threadIdx.x + i * kGnThreads;
        if (v < nv) {
            const size_t o = (size_t)v * c + c0;
            load_row<CPG>(dy + o, dr[i]);
            load_row<CPG>(x + o, xh[i]);
            if (relu) {
                float yr[CPG];
                load_row<CPG>(y + o, yr);
#pragma unroll
                for (int k = 0; k < CPG; k++)
                    if (!(yr[k] > 0.0f)) dr[i][k] = 0.0f;
            }
#pragma unroll
            for (int k = 0; k < CPG; k++) {
                xh[i][k] = (xh[i][k] - mean) * rstd;
                ab[k] = fmaf(dr[i][k], xh[i][k], ab[k]);
                ab[CPG + k] += dr[i][k];
            }
        }
    }
    block_sum_n<2 * CPG>(ab, red);
    float gm[CPG];
    float ds = 0.0f, db = 0.0f;
#pragma unroll
    for (int k = 0; k < CPG; k++) {
        gm[k] = __ldg(gamma + c0 + k);
        ds = fmaf(ab[k], gm[k], ds);
        db = fmaf(ab[CPG + k], gm[k], db);
    }
#pragma unroll
    for (int k = 0; k < CPG; k++) {        // unrolled selects: ab[] stays in registers
        if (threadIdx.x == k) {
            dgamma[c0 + k] = ab[k];
            dbeta[c0 + k] = ab[CPG + k];
        }
    }
    const float inv_m = 1.0f / ((float)nv * (float)CPG);
#pragma unroll
    for (int i = 0; i < kGnRows; i++) {
        const int v = threadIdx.x + i * kGnThreads;
        if (v < nv_rows) {
            float o[CPG];
#pragma unroll
            for (int k = 0; k < CPG; k++) o[k] = 0.0f;
            if (dx_add != nullptr) load_row<CPG>(dx_add + (size_t)v * c + c0, o);   // gradient of the skip connection that forked off x
#pragma unroll
            for (int k = 0; k < CPG; k++)
                if (v < nv) o[k] += rstd * (dr[i][k] * gm[k] - (xh[i][k] * ds + db) * inv_m);
            store_row<CPG>(dx + (size_t)v * c + c0, o);
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Row-tiled variants for scene-sized lattices (SemanticKITTI / ScanNet: 10^4..10^5 vertices per level).  The
// one-CTA-per-group kernels above put 32 CTAs on a 148-SM machine and walk a group's channels at stride C; on a
// KITTI-sized pass they were 31 % of all kernel time (profiles/r01i_launches_kitti_pass.md).  Here every CTA owns a
// block of consecutive ROWS and all C channels: loads are whole 16-byte-vectorised rows (coalesced), the grid
// covers the machine, and the statistics are combined from per-CTA partials in a fixed order (no atomics:
// results are reproducible run to run).
//   forward : stats partial (per CTA and group: local mean, centred M2 -- rows cached in registers between the
//             two passes)  ->  finalize (Chan's parallel combination, one CTA per group)  ->  apply
//   backward: per-channel partial sums of dy' xhat and dy'  ->  finalize (dgamma, dbeta, per-group ds / db)  ->  apply
// Thread layout: tpr = C/4 threads per row (one float4 each), rpp = 256 / tpr rows per pass, kGtIter passes.
constexpr int kGtIter = 8;
constexpr int kGtRedRows = 148;      // rows the backward partials are pre-reduced to on very large lattices
constexpr int kGtMaxThreads = 256;

struct GtPlan {
    int tpr, rpp, threads, rows_per_cta, ctas;
};
static bool gt_plan(int nv_rows, int c, int cpg, GtPlan* p) {
    if (c % 4 != 0 || c / 4 > kGtMaxThreads || cpg > 128) return false;   // cpg <= 128: one finalize CTA holds whole groups
    p->tpr = c / 4;
    p->rpp = kGtMaxThreads / p->tpr;
    p->threads = p->tpr * p->rpp;
    p->rows_per_cta = p->rpp * kGtIter;
    p->ctas = (nv_rows + p->rows_per_cta - 1) / p->rows_per_cta;
    return true;
}
static size_t gt_workspace_floats(const GtPlan& p, int c, int groups) {   // partials, pre-reduced partials, per-group (ds, db)
    return (size_t)p.ctas * 2 * c + (size_t)kGtRedRows * 2 * c + 2 * (size_t)groups;
}

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// partial [ctas][G][2] = (local mean, local centred sum of squares) of the CTA's rows, per group
__global__ void __launch_bounds__(kGtMaxThreads)
gn_tiled_stats_kernel(const float* __restrict__ x, int nv_rows, const int* __restrict__ nv_dev, int c, int cpg, int tpr,
                      int rows_per_cta, float* __restrict__ partial) {
    LN_PDL_ENTRY();
    extern __shared__ __align__(16) float sh[];   // [rpp][c] per-row-slot channel sums, [c] channel sums, [G] group means
    const int rpp = blockDim.x / tpr;
    float* part = sh;
    float* ch_sum = sh + (size_t)rpp * c;
    float* g_mean = ch_sum + c;
    const int G = c / cpg;
    const int nv = nv_dev ? min(nv_rows, __ldg(nv_dev)) : nv_rows;
    const int col = threadIdx.x % tpr, rsub = threadIdx.x / tpr;
    const int r0 = blockIdx.x * rows_per_cta;
    const int rows_here = max(0, min(rows_per_cta, nv - r0));
    const int ch = col * 4;
    float4 xr[kGtIter];
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kGtIter; i++) {
        const int r = rsub + i * rpp;
        xr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows_here) {
            xr[i] = ld4(x + (size_t)(r0 + r) * c + ch);
            s.x += xr[i].x; s.y += xr[i].y; s.z += xr[i].z; s.w += xr[i].w;
        }
    }
    *reinterpret_cast<float4*>(part + (size_t)rsub * c + ch) = s;
    __syncthreads();
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        float t = 0.0f;
        for (int rs = 0; rs < rpp; rs++) t += part[(size_t)rs * c + i];
        ch_sum[i] = t;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        float t = 0.0f;
        for (int j = 0; j < cpg; j++) t += ch_sum[g * cpg + j];
        g_mean[g] = rows_here > 0 ? t / ((float)rows_here * (float)cpg) : 0.0f;
    }
    __syncthreads();
    const float m0 = g_mean[ch / cpg], m1 = g_mean[(ch + 1) / cpg], m2 = g_mean[(ch + 2) / cpg], m3 = g_mean[(ch + 3) / cpg];
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kGtIter; i++) {
        const int r = rsub + i * rpp;
        if (r < rows_here) {
            float d = xr[i].x - m0; q.x = fmaf(d, d, q.x);
            d = xr[i].y - m1; q.y = fmaf(d, d, q.y);
            d = xr[i].z - m2; q.z = fmaf(d, d, q.z);
            d = xr[i].w - m3; q.w = fmaf(d, d, q.w);
        }
    }
    *reinterpret_cast<float4*>(part + (size_t)rsub * c + ch) = q;     // every thread rewrites only its own slot
    __syncthreads();
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        float t = 0.0f;
        for (int rs = 0; rs < rpp; rs++) t += part[(size_t)rs * c + i];
        ch_sum[i] = t;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        float t = 0.0f;
        for (int j = 0; j < cpg; j++) t += ch_sum[g * cpg + j];
        partial[((size_t)blockIdx.x * G + g) * 2] = g_mean[g];
        partial[((size_t)blockIdx.x * G + g) * 2 + 1] = t;
    }
}

// Chan et al.: mean = sum n_i mean_i / N,  M2 = sum (M2_i + n_i (mean_i - mean)^2); one CTA per group
__global__ void __launch_bounds__(128)
gn_tiled_finalize_kernel(const float* __restrict__ partial, int ctas, int G, int rows_per_cta, int nv_rows,
                         const int* __restrict__ nv_dev, int cpg, float eps, float* __restrict__ stats) {
    LN_PDL_ENTRY();
    __shared__ float red[32];
    const int g = blockIdx.x;
    const int nv = nv_dev ? min(nv_rows, __ldg(nv_dev)) : nv_rows;
    float a = 0.0f;
    for (int i = threadIdx.x; i < ctas; i += blockDim.x) {
        const float n_i = (float)max(0, min(rows_per_cta, nv - i * rows_per_cta)) * (float)cpg;
        a = fmaf(n_i, partial[((size_t)i * G + g) * 2], a);
    }
    const float m = (float)nv * (float)cpg;
    const float mean = block_sum(a, red) / m;
    float b = 0.0f;
    for (int i = threadIdx.x; i < ctas; i += blockDim.x) {
        const float n_i = (float)max(0, min(rows_per_cta, nv - i * rows_per_cta)) * (float)cpg;
        const float d = partial[((size_t)i * G + g) * 2] - mean;
        b += partial[((size_t)i * G + g) * 2 + 1] + n_i * d * d;
    }
    const float var = block_sum(b, red) / m;
    if (threadIdx.x == 0) {
        stats[2 * g] = mean;
        stats[2 * g + 1] = rsqrtf(var + eps);
    }
}

__global__ void __launch_bounds__(kGtMaxThreads)
gn_tiled_apply_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                      const float* __restrict__ stats, int nv_rows, const int* __restrict__ nv_dev, int c, int cpg, int tpr,
                      int rows_per_cta, int relu, float* __restrict__ y) {
    LN_PDL_ENTRY();
    const int rpp = blockDim.x / tpr;
    const int nv = nv_dev ? min(nv_rows, __ldg(nv_dev)) : nv_rows;
    const int col = threadIdx.x % tpr, rsub = threadIdx.x / tpr;
    const int ch = col * 4;
    const float4 gm = ld4(gamma + ch), bt = ld4(beta + ch);
    float mean[4], rstd[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int g = (ch + k) / cpg;
        mean[k] = __ldg(stats + 2 * g);
        rstd[k] = __ldg(stats + 2 * g + 1);
    }
    const int r0 = blockIdx.x * rows_per_cta;
#pragma unroll
    for (int i = 0; i < kGtIter; i++) {
        const int r = r0 + rsub + i * rpp;
        if (r >= nv_rows) break;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);          // padding rows (static-shape mode) come out as zeros
        if (r < nv) {
            const float4 v = ld4(x + (size_t)r * c + ch);
            o.x = fmaf((v.x - mean[0]) * rstd[0], gm.x, bt.x);
            o.y = fmaf((v.y - mean[1]) * rstd[1], gm.y, bt.y);
            o.z = fmaf((v.z - mean[2]) * rstd[2], gm.z, bt.z);
            o.w = fmaf((v.w - mean[3]) * rstd[3], gm.w, bt.w);
            if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        }
        *reinterpret_cast<float4*>(y + (size_t)r * c + ch) = o;
    }
}

// partial_ab [ctas][2c]: per channel sum dy' xhat (first c) and sum dy' (second c) over the CTA's rows
__global__ void __launch_bounds__(kGtMaxThreads)
gn_tiled_bwd_partial_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                            const float* __restrict__ stats, int nv_rows, const int* __restrict__ nv_dev, int c, int cpg,
                            int tpr, int rows_per_cta, int relu, float* __restrict__ partial_ab) {
    LN_PDL_ENTRY();
    extern __shared__ __align__(16) float sh[];   // [rpp][2c]
    const int rpp = blockDim.x / tpr;
    const int nv = nv_dev ? min(nv_rows, __ldg(nv_dev)) : nv_rows;
    const int col = threadIdx.x % tpr, rsub = threadIdx.x / tpr;
    const int ch = col * 4;
    float mean[4], rstd[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int g = (ch + k) / cpg;
        mean[k] = __ldg(stats + 2 * g);
        rstd[k] = __ldg(stats + 2 * g + 1);
    }
    const int r0 = blockIdx.x * rows_per_cta;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kGtIter; i++) {
        const int r = r0 + rsub + i * rpp;
        if (r < nv) {
            const size_t o = (size_t)r * c + ch;
            float4 d = ld4(dy + o);
            const float4 v = ld4(x + o);
            if (relu) {
                const float4 yy = ld4(y + o);
                if (!(yy.x > 0.f)) d.x = 0.f;
                if (!(yy.y > 0.f)) d.y = 0.f;
                if (!(yy.z > 0.f)) d.z = 0.f;
                if (!(yy.w > 0.f)) d.w = 0.f;
            }
            a.x = fmaf(d.x, (v.x - mean[0]) * rstd[0], a.x); b.x += d.x;
            a.y = fmaf(d.y, (v.y - mean[1]) * rstd[1], a.y); b.y += d.y;
            a.z = fmaf(d.z, (v.z - mean[2]) * rstd[2], a.z); b.z += d.z;
            a.w = fmaf(d.w, (v.w - mean[3]) * rstd[3], a.w); b.w += d.w;
        }
    }
    *reinterpret_cast<float4*>(sh + (size_t)rsub * 2 * c + ch) = a;
    *reinterpret_cast<float4*>(sh + (size_t)rsub * 2 * c + c + ch) = b;
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) {
        float t = 0.0f;
        for (int rs = 0; rs < rpp; rs++) t += sh[(size_t)rs * 2 * c + i];
        partial_ab[(size_t)blockIdx.x * 2 * c + i] = t;
    }
}

// Pre-reduction of the per-CTA partial rows when there are many of them (lattices beyond ~10^5 vertices): out[s][col] =
// sum over rows r = s, s + S, s + 2S, ... of in[r][col], lane = column (coalesced), fixed order (deterministic).  The
// finalize kernel then walks S rows instead of thousands with a single CTA per 128 channels.
__global__ void __launch_bounds__(128)
gn_tiled_reduce_rows_kernel(const float* __restrict__ in, int rows, int width, float* __restrict__ out) {
    LN_PDL_ENTRY();
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= width) return;
    float acc = 0.0f;
    for (int r = blockIdx.y; r < rows; r += gridDim.y) acc += __ldg(in + (size_t)r * width + col);
    out[(size_t)blockIdx.y * width + col] = acc;
}

// dgamma / dbeta per channel and gstat[g] = (ds, db) = sum_j (a_j, b_j) gamma_j per group.  A CTA owns `cpc` consecutive
// channels (a whole number of groups, <= 128): lane = channel, so the loads of one partial row are coalesced; four
// slices of CTAs are summed side by side and combined in a fixed order (deterministic).
constexpr int kGtFinLanes = 128;
constexpr int kGtFinSlices = 4;
__global__ void __launch_bounds__(kGtFinLanes * kGtFinSlices)
gn_tiled_bwd_finalize_kernel(const float* __restrict__ partial_ab, int ctas, int c, int cpg, int cpc, const float* __restrict__ gamma,
                             float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ gstat) {
    LN_PDL_ENTRY();
    __shared__ float red[kGtFinSlices][2][kGtFinLanes];
    __shared__ float ch_a[kGtFinLanes], ch_b[kGtFinLanes];
    const int lane = threadIdx.x % kGtFinLanes, slice = threadIdx.x / kGtFinLanes;
    const int ch0 = blockIdx.x * cpc;
    const int nch = min(cpc, c - ch0);                 // channels of this CTA (whole groups: c and cpc are multiples of cpg)
    float a = 0.0f, b = 0.0f;
    if (lane < nch) {
        const float* p = partial_ab + ch0 + lane;
        for (int i = slice; i < ctas; i += kGtFinSlices) {
            a += __ldg(p + (size_t)i * 2 * c);
            b += __ldg(p + (size_t)i * 2 * c + c);
        }
    }
    red[slice][0][lane] = a;
    red[slice][1][lane] = b;
    __syncthreads();
    if (slice == 0 && lane < nch) {
        float ta = 0.0f, tb = 0.0f;
#pragma unroll
        for (int k = 0; k < kGtFinSlices; k++) {
            ta += red[k][0][lane];
            tb += red[k][1][lane];
        }
        dgamma[ch0 + lane] = ta;
        dbeta[ch0 + lane] = tb;
        const float gm = __ldg(gamma + ch0 + lane);
        ch_a[lane] = ta * gm;
        ch_b[lane] = tb * gm;
    }
    __syncthreads();
    if (threadIdx.x < nch / cpg) {
        float ds = 0.0f, db = 0.0f;
        for (int j = 0; j < cpg; j++) {
            ds += ch_a[threadIdx.x * cpg + j];
            db += ch_b[threadIdx.x * cpg + j];
        }
        const int g = ch0 / cpg + threadIdx.x;
        gstat[2 * g] = ds;
        gstat[2 * g + 1] = db;
    }
}

__global__ void __launch_bounds__(kGtMaxThreads)
gn_tiled_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                          const float* __restrict__ gamma, const float* __restrict__ stats, const float* __restrict__ gstat,
                          const float* __restrict__ dx_add, int nv_rows, const int* __restrict__ nv_dev, int c, int cpg, int tpr,
                          int rows_per_cta, int relu, float* __restrict__ dx) {
    LN_PDL_ENTRY();
    const int rpp = blockDim.x / tpr;
    const int nv = nv_dev ? min(nv_rows, __ldg(nv_dev)) : nv_rows;
    const int col = threadIdx.x % tpr, rsub = threadIdx.x / tpr;
    const int ch = col * 4;
    const float4 gm4 = ld4(gamma + ch);
    const float gm[4] = {gm4.x, gm4.y, gm4.z, gm4.w};
    float mean[4], rstd[4], ds[4], db[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int g = (ch + k) / cpg;
        mean[k] = __ldg(stats + 2 * g);
        rstd[k] = __ldg(stats + 2 * g + 1);
        ds[k] = __ldg(gstat + 2 * g);
        db[k] = __ldg(gstat + 2 * g + 1);
    }
    const float inv_m = 1.0f / ((float)nv * (float)cpg);
    const int r0 = blockIdx.x * rows_per_cta;
#pragma unroll
    for (int i = 0; i < kGtIter; i++) {
        const int r = r0 + rsub + i * rpp;
        if (r >= nv_rows) break;
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        if (dx_add != nullptr) {
            const float4 a4 = ld4(dx_add + (size_t)r * c + ch);
            o[0] = a4.x; o[1] = a4.y; o[2] = a4.z; o[3] = a4.w;
        }
        if (r < nv) {
            const size_t off = (size_t)r * c + ch;
            const float4 d4 = ld4(dy + off), v4 = ld4(x + off);
            float d[4] = {d4.x, d4.y, d4.z, d4.w};
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
            if (relu) {
                const float4 y4 = ld4(y + off);
                const float yy[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (!(yy[k] > 0.f)) d[k] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float xhat = (v[k] - mean[k]) * rstd[k];
                o[k] += rstd[k] * (d[k] * gm[k] - (xhat * ds[k] + db[k]) * inv_m);
            }
        }
        *reinterpret_cast<float4*>(dx + (size_t)r * c + ch) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

static bool gn_small_ok(int nv, int cpg) {
    return nv <= kGnRows * kGnThreads && (cpg == 1 || cpg == 2 || cpg == 3 || cpg == 4 || cpg == 6 || cpg == 8);
}

}  // namespace ln

using namespace ln;

extern "C" {

long long ln_group_norm_workspace_bytes(int nv, int c, int groups) {
    if (nv < 1 || c < 1 || groups < 1 || c % groups != 0) return 0;
    GtPlan p;
    if (gn_small_ok(nv, c / groups) || !gt_plan(nv, c, c / groups, &p)) return 0;
    return (long long)(gt_workspace_floats(p, c, groups) * sizeof(float));
}

int ln_group_norm_fwd(const float* x, const float* gamma, const float* beta, int nv, const int* nv_dev, int c, int groups,
                      float eps, int relu, float* y, float* stats, float* workspace, void* stream) {
    LN_REQUIRE(x && gamma && beta && y && stats, "ln_group_norm_fwd: null pointer");
    LN_REQUIRE(nv >= 1 && c >= 1 && groups >= 1 && c % groups == 0, "ln_group_norm_fwd: bad size nv=%d c=%d groups=%d", nv, c, groups);
    const int cpg = c / groups;
    cudaStream_t s = (cudaStream_t)stream;
    if (gn_small_ok(nv, cpg)) {
#define LN_GN_FWD(CPG) launch_k(group_norm_fwd_small_kernel<CPG>, dim3(groups), dim3(kGnThreads), 0, s, x, gamma, beta, nv, nv_dev, c, eps, relu, y, stats)
        switch (cpg) {
            case 1: LN_GN_FWD(1); break;
            case 2: LN_GN_FWD(2); break;
            case 3: LN_GN_FWD(3); break;
            case 4: LN_GN_FWD(4); break;
            case 6: LN_GN_FWD(6); break;
            default: LN_GN_FWD(8); break;
        }
#undef LN_GN_FWD
    } else if (GtPlan p; workspace != nullptr && gt_plan(nv, c, cpg, &p)) {
        const size_t smem = ((size_t)p.rpp * c + c + groups) * sizeof(float);
        launch_k(gn_tiled_stats_kernel, dim3(p.ctas), dim3(p.threads), smem, s, x, nv, nv_dev, c, cpg, p.tpr, p.rows_per_cta, workspace);
        launch_k(gn_tiled_finalize_kernel, dim3(groups), dim3(128), 0, s, workspace, p.ctas, groups, p.rows_per_cta, nv, nv_dev, cpg, eps, stats);
        launch_k(gn_tiled_apply_kernel, dim3(p.ctas), dim3(p.threads), 0, s, x, gamma, beta, stats, nv, nv_dev, c, cpg, p.tpr, p.rows_per_cta, relu, y);
        count_launch();
        count_launch();
    } else {
        launch_k(group_norm_fwd_kernel, dim3(groups), dim3(kGnThreads), 0, s, x, gamma, beta, nv, nv_dev, c, cpg, eps, relu, y, stats);
    }
    count_launch();
    return check_launch("group_norm_fwd");
}

int ln_group_norm_bwd(const float* dy, const float* x, const float* y, const float* gamma, const float* stats,
                      const float* dx_add, int nv, const int* nv_dev, int c, int groups, int relu, float* dx, float* dgamma,
                      float* dbeta, float* workspace, void* stream) {
    LN_REQUIRE(dy && x && gamma && stats && dx && dgamma && dbeta, "ln_group_norm_bwd: null pointer");
    LN_REQUIRE(!relu || y, "ln_group_norm_bwd: the forward output is needed for the ReLU mask");
    LN_REQUIRE(nv >= 1 && c >= 1 && groups >= 1 && c % groups == 0, "ln_group_norm_bwd: bad size");
    const int cpg = c / groups;
    cudaStream_t s = (cudaStream_t)stream;
    if (gn_small_ok(nv, cpg)) {
#define LN_GN_BWD(CPG) launch_k(group_norm_bwd_small_kernel<CPG>, dim3(groups), dim3(kGnThreads), 0, s, dy, x, y, gamma, stats, dx_add, nv, nv_dev, c, relu, dx, dgamma, dbeta)
        switch (cpg) {
            case 1: LN_GN_BWD(1); break;
            case 2: LN_GN_BWD(2); break;
            case 3: LN_GN_BWD(3); break;
            case 4: LN_GN_BWD(4); break;
            case 6: LN_GN_BWD(6); break;
            default: LN_GN_BWD(8); break;
        }
#undef LN_GN_BWD
    } else if (GtPlan p; workspace != nullptr && gt_plan(nv, c, cpg, &p)) {
        float* partial_ab = workspace;
        float* reduced_ab = workspace + (size_t)p.ctas * 2 * c;
        float* gstat = reduced_ab + (size_t)kGtRedRows * 2 * c;
        const size_t smem = (size_t)p.rpp * 2 * c * sizeof(float);
        launch_k(gn_tiled_bwd_partial_kernel, dim3(p.ctas), dim3(p.threads), smem, s, dy, x, y, stats, nv, nv_dev, c, cpg, p.tpr, p.rows_per_cta, relu, partial_ab);
        const float* fin_in = partial_ab;
        int fin_rows = p.ctas;
        if (p.ctas > 4 * kGtRedRows) {
            launch_k(gn_tiled_reduce_rows_kernel, dim3(dim3((2 * c + 127) / 128, kGtRedRows)), dim3(128), 0, s, partial_ab, p.ctas, 2 * c, reduced_ab);
            count_launch();
            fin_in = reduced_ab;
            fin_rows = kGtRedRows;
        }
        const int cpc = cpg >= kGtFinLanes ? cpg : (kGtFinLanes / cpg) * cpg;       // whole groups per CTA
        launch_k(gn_tiled_bwd_finalize_kernel, dim3((c + cpc - 1) / cpc), dim3(kGtFinLanes * kGtFinSlices), 0, s, fin_in, fin_rows, c, cpg, cpc, gamma, dgamma, dbeta, gstat);
        launch_k(gn_tiled_bwd_apply_kernel, dim3(p.ctas), dim3(p.threads), 0, s, dy, x, y, gamma, stats, gstat, dx_add, nv, nv_dev, c, cpg, p.tpr, p.rows_per_cta, relu, dx);
        count_launch();
        count_launch();
    } else {
        launch_k(group_norm_bwd_kernel, dim3(groups), dim3(kGnThreads), 0, s, dy, x, y, gamma, stats, dx_add, nv, nv_dev, c, cpg, relu, dx, dgamma, dbeta);
    }
    count_launch();
    return check_launch("group_norm_bwd");
}

}  // extern "C"
