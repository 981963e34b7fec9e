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

}  // namespace ln

using namespace ln;

extern "C" {

int ln_group_norm_fwd(const float* x, const float* gamma, const float* beta, int nv, const int* nv_dev, int c, int groups,
                      float eps, int relu, float* y, float* stats, void* stream) {
    LN_REQUIRE(x && gamma && beta && y && stats, "ln_group_norm_fwd: null pointer");
    LN_REQUIRE(nv >= 1 && c >= 1 && groups >= 1 && c % groups == 0, "ln_group_norm_fwd: bad size nv=%d c=%d groups=%d", nv, c, groups);
    group_norm_fwd_kernel<<<groups, kGnThreads, 0, (cudaStream_t)stream>>>(x, gamma, beta, nv, nv_dev, c, c / groups, eps, relu, y, stats);
    count_launch();
    return check_launch("group_norm_fwd");
}

int ln_group_norm_bwd(const float* dy, const float* x, const float* y, const float* gamma, const float* stats, int nv,
                      const int* nv_dev, int c, int groups, int relu, float* dx, float* dgamma, float* dbeta, void* stream) {
    LN_REQUIRE(dy && x && gamma && stats && dx && dgamma && dbeta, "ln_group_norm_bwd: null pointer");
    LN_REQUIRE(!relu || y, "ln_group_norm_bwd: the forward output is needed for the ReLU mask");
    LN_REQUIRE(nv >= 1 && c >= 1 && groups >= 1 && c % groups == 0, "ln_group_norm_bwd: bad size");
    group_norm_bwd_kernel<<<groups, kGnThreads, 0, (cudaStream_t)stream>>>(dy, x, y, gamma, stats, nv, nv_dev, c, c / groups, relu, dx, dgamma, dbeta);
    count_launch();
    return check_launch("group_norm_bwd");
}

}  // extern "C"
