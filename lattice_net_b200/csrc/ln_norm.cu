// GroupNorm (+ optional ReLU) over lattice values stored vertex-major [nv x C] -- SURVEY.md section 8f
// rank 2: the normalisation that sits between every two lattice convolutions
// (/root/reference/latticenet_py/lattice/lattice_modules.py:585-614, 935-960 do it with
// unsqueeze/transpose + torch.nn.GroupNorm + ReLU, i.e. two layout copies and four kernels forward).
// Statistics of group g run over (channels of g) x (all vertices), biased variance, like
// torch.nn.GroupNorm on a [1, C, nv] tensor.
//
// One CTA per group: mean, variance and the normalised output in a single launch (the group's slice is
// nv * C/G floats and stays in L1/L2 between the passes); the backward pass likewise produces dx,
// dgamma and dbeta in one launch without atomics (every channel belongs to exactly one group).
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
                      const float* __restrict__ gamma, const float* __restrict__ stats, int nv_rows,
                      const int* __restrict__ nv_dev, int c, int cpg, int relu, float* __restrict__ dx,
                      float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ float red[32];
    const int g = blockIdx.x;
    const int c0 = g * cpg;
    const int nv = nv_dev ? min(nv_rows, __ldg(nv_dev)) : nv_rows;
    for (long long i = (long long)nv * cpg + threadIdx.x; i < (long long)nv_rows * cpg; i += blockDim.x) {
        const long long v = i / cpg;
        dx[v * c + c0 + (int)(i - v * cpg)] = 0.0f;
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
        dx[o] = rstd * (d * __ldg(gamma + c0 + j) - (xhat * ds + db) * inv_m);
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
                            const float* __restrict__ gamma, const float* __restrict__ stats, int nv_rows,
                            const int* __restrict__ nv_dev, int c, int relu, float* __restrict__ dx,
                            float* __restrict__ dgamma, float* __restrict__ dbeta) {
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
            for (int k = 0; k < CPG; k++) {
                o[k] = 0.0f;
                if (v < nv) o[k] = rstd * (dr[i][k] * gm[k] - (xh[i][k] * ds + db) * inv_m);
            }
            store_row<CPG>(dx + (size_t)v * c + c0, o);
        }
    }
}

static bool gn_small_ok(int nv, int cpg) {
    return nv <= kGnRows * kGnThreads && (cpg == 1 || cpg == 2 || cpg == 3 || cpg == 4 || cpg == 6 || cpg == 8);
}

}  // namespace ln

using namespace ln;

extern "C" {

int ln_group_norm_fwd(const float* x, const float* gamma, const float* beta, int nv, const int* nv_dev, int c, int groups,
                      float eps, int relu, float* y, float* stats, void* stream) {
    LN_REQUIRE(x && gamma && beta && y && stats, "ln_group_norm_fwd: null pointer");
    LN_REQUIRE(nv >= 1 && c >= 1 && groups >= 1 && c % groups == 0, "ln_group_norm_fwd: bad size nv=%d c=%d groups=%d", nv, c, groups);
    const int cpg = c / groups;
    cudaStream_t s = (cudaStream_t)stream;
    if (gn_small_ok(nv, cpg)) {
#define LN_GN_FWD(CPG) group_norm_fwd_small_kernel<CPG><<<groups, kGnThreads, 0, s>>>(x, gamma, beta, nv, nv_dev, c, eps, relu, y, stats)
        switch (cpg) {
            case 1: LN_GN_FWD(1); break;
            case 2: LN_GN_FWD(2); break;
            case 3: LN_GN_FWD(3); break;
            case 4: LN_GN_FWD(4); break;
            case 6: LN_GN_FWD(6); break;
            default: LN_GN_FWD(8); break;
        }
#undef LN_GN_FWD
    } else {
        group_norm_fwd_kernel<<<groups, kGnThreads, 0, s>>>(x, gamma, beta, nv, nv_dev, c, cpg, eps, relu, y, stats);
    }
    count_launch();
    return check_launch("group_norm_fwd");
}

int ln_group_norm_bwd(const float* dy, const float* x, const float* y, const float* gamma, const float* stats, int nv,
                      const int* nv_dev, int c, int groups, int relu, float* dx, float* dgamma, float* dbeta, void* stream) {
    LN_REQUIRE(dy && x && gamma && stats && dx && dgamma && dbeta, "ln_group_norm_bwd: null pointer");
    LN_REQUIRE(!relu || y, "ln_group_norm_bwd: the forward output is needed for the ReLU mask");
    LN_REQUIRE(nv >= 1 && c >= 1 && groups >= 1 && c % groups == 0, "ln_group_norm_bwd: bad size");
    const int cpg = c / groups;
    cudaStream_t s = (cudaStream_t)stream;
    if (gn_small_ok(nv, cpg)) {
#define LN_GN_BWD(CPG) group_norm_bwd_small_kernel<CPG><<<groups, kGnThreads, 0, s>>>(dy, x, y, gamma, stats, nv, nv_dev, c, relu, dx, dgamma, dbeta)
        switch (cpg) {
            case 1: LN_GN_BWD(1); break;
            case 2: LN_GN_BWD(2); break;
            case 3: LN_GN_BWD(3); break;
            case 4: LN_GN_BWD(4); break;
            case 6: LN_GN_BWD(6); break;
            default: LN_GN_BWD(8); break;
        }
#undef LN_GN_BWD
    } else {
        group_norm_bwd_kernel<<<groups, kGnThreads, 0, s>>>(dy, x, y, gamma, stats, nv, nv_dev, c, cpg, relu, dx, dgamma, dbeta);
    }
    count_launch();
    return check_launch("group_norm_bwd");
}

}  // extern "C"
