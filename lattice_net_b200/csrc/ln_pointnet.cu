// Fused "distribute + PointNet" front end of LatticeNet.
//
// Reference (/root/reference/latticenet_py/lattice/lattice_modules.py:52-96, 620-733): `distribute` materialises one row
// per (point, simplex vertex) -- [N(d+1) x (d+V+1)] --, torch_scatter computes the per-vertex mean position, the rows are
// centred, pushed through three weight-normalised Linear + LeakyReLU(0.2) layers ([N(d+1) x 16/32/64] intermediates),
// max-pooled per vertex with torch_scatter.scatter_max, the barycentric weight of each winning row is gathered, vertices
// with fewer than 4 points and vertex 0 are zeroed.  ~40 launches forward, ~40 backward, and at ScanNet size (600 k rows)
// ~0.5 GB of intermediates.
//
// Here: nothing per-row is ever stored.
//   forward   memset -> pn_mean_kernel (per-vertex position sums / counts)
//                    -> pn_mlp_max_kernel (thread per row: centre, 3-layer MLP in registers, weights normalised in the
//                       CTA prologue, packed 64-bit atomicMax per (vertex, channel))
//                    -> pn_finish_kernel (unpack max / argmax, gather barycentric weights, masks)
//   backward  memset -> pn_mlp_bwd_kernel (thread per row: recompute the activations, route the pooled gradient to the
//                       winning rows, back-propagate; weight gradients as per-CTA register-tiled G^T.H products over row
//                       tiles staged in shared memory, one vector of atomics per CTA)
//                    -> pn_wn_bwd_kernel (weight-norm backward: grads of weight_v, weight_g, bias of the three layers)
// Rows whose vertex id is < 0 (not inserted: table overflow / row bound exceeded) take no part; with `vertex0_quirk` the
// rows of vertex 0 are masked as in the reference (lattice_modules.py:72-94, 712).
#include "ln_common.cuh"

namespace ln {

constexpr int kPnThreads = 128;
constexpr float kLeaky = 0.2f;

struct PnLayers {            // device pointers of the three LinearWN layers (weight_v [out x in], weight_g [out x 1], bias [out])
    const float* v[3];
    const float* g[3];
    const float* b[3];
};
struct PnGrads {
    float* v[3];
    float* g[3];
    float* b[3];
};

__device__ __forceinline__ float lrelu(float x) { return x > 0.0f ? x : kLeaky * x; }
__device__ __forceinline__ float lrelu_grad(float x) { return x > 0.0f ? 1.0f : kLeaky; }
__device__ __forceinline__ unsigned int pn_ordered(float f) {
    const unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float pn_unordered(unsigned int o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

__device__ __forceinline__ float pn_block_sum(float v, float* red) {     // blockDim = kPnThreads
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < kPnThreads / 32; w++) t += red[w];
    return t;
}

// normalised weights of all three layers into shared memory: w = v * (g[out] / ||v||_F)  (utils.py:72-158, g_dim = 0)
template <int IN, int H1, int H2, int H3>
__device__ __forceinline__ void pn_load_weights(const PnLayers& L, float* w1, float* w2, float* w3, float* b1, float* b2, float* b3,
                                                float* red) {
    const int sizes[3] = {H1 * IN, H2 * H1, H3 * H2};
    const int outs[3] = {H1, H2, H3};
    const int ins[3] = {IN, H1, H2};
    float* dst[3] = {w1, w2, w3};
    float* bd[3] = {b1, b2, b3};
#pragma unroll
    for (int l = 0; l < 3; l++) {
        float s = 0.0f;
        for (int i = threadIdx.x; i < sizes[l]; i += kPnThreads) {
            const float x = __ldg(L.v[l] + i);
            s = fmaf(x, x, s);
        }
        const float norm = sqrtf(pn_block_sum(s, red));
        for (int i = threadIdx.x; i < sizes[l]; i += kPnThreads) dst[l][i] = __ldg(L.v[l] + i) * (__ldg(L.g[l] + i / ins[l]) / norm);
        for (int i = threadIdx.x; i < outs[l]; i += kPnThreads) bd[l][i] = __ldg(L.b[l] + i);
    }
    __syncthreads();
}

// sums [nv x D] (+ count in column D): acc[v*(D+1) + j]
template <int D>
__global__ void __launch_bounds__(256)
pn_mean_kernel(const float* __restrict__ positions_raw, const float* __restrict__ sigmas, const int* __restrict__ indices, int n,
               float* __restrict__ acc) {
    LN_PDL_ENTRY();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * (D + 1)) return;
    const int p = (int)(t / (D + 1));
    const int id = __ldg(indices + t);
    if (id < 0) return;
    float* a = acc + (size_t)id * (D + 1);
#pragma unroll
    for (int i = 0; i < D; i++) atomicAdd(a + i, __fdiv_rn(__ldg(positions_raw + (size_t)p * D + i), __ldg(sigmas + i)));
    atomicAdd(a + D, 1.0f);
}

// input features of one row: [ p/sigma - mean(vertex) | values ]
template <int D, int IN>
__device__ __forceinline__ void pn_row_input(const float* __restrict__ positions_raw, const float* __restrict__ sigmas,
                                             const float* __restrict__ values, const float* __restrict__ acc, int p, int id,
                                             int vertex0_quirk, float* x0) {
    const float* a = acc + (size_t)id * (D + 1);
    const float cnt = fmaxf(__ldg(a + D), 1.0f);
#pragma unroll
    for (int i = 0; i < D; i++) {
        float mean = __ldg(a + i) / cnt;
        if (vertex0_quirk && id == 0) mean = 0.0f;
        x0[i] = __fdiv_rn(__ldg(positions_raw + (size_t)p * D + i), __ldg(sigmas + i)) - mean;
    }
#pragma unroll
    for (int i = D; i < IN; i++) x0[i] = __ldg(values + (size_t)p * (IN - D) + (i - D));
}

template <int D, int IN, int H1, int H2, int H3>
__global__ void __launch_bounds__(kPnThreads)
pn_mlp_max_kernel(const float* __restrict__ positions_raw, const float* __restrict__ sigmas, const float* __restrict__ values,
                  const int* __restrict__ indices, int n, PnLayers L, const float* __restrict__ acc, int vertex0_quirk,
                  unsigned long long* __restrict__ packed) {
    __shared__ __align__(16) float w1[H1 * IN], w2[H2 * H1], w3[H3 * H2], b1[H1], b2[H2], b3[H3], red[4];
    LN_PDL_ENTRY();
    pn_load_weights<IN, H1, H2, H3>(L, w1, w2, w3, b1, b2, b3, red);
    const long long rows = (long long)n * (D + 1);
    for (long long row = (long long)blockIdx.x * kPnThreads + threadIdx.x; row < rows; row += (long long)gridDim.x * kPnThreads) {
        const int id = __ldg(indices + row);
        if (id < 0 || (vertex0_quirk && id == 0)) continue;      // masked rows only ever reach vertex 0, which is zeroed afterwards
        const int p = (int)(row / (D + 1));
        float x0[IN];
        pn_row_input<D, IN>(positions_raw, sigmas, values, acc, p, id, vertex0_quirk, x0);
        float h1[H1], h2[H2];
#pragma unroll
        for (int o = 0; o < H1; o++) {
            float s = b1[o];
#pragma unroll
            for (int i = 0; i < IN; i++) s = fmaf(w1[o * IN + i], x0[i], s);
            h1[o] = lrelu(s);
        }
#pragma unroll
        for (int o = 0; o < H2; o++) {
            float s = b2[o];
#pragma unroll
            for (int i = 0; i < H1; i++) s = fmaf(w2[o * H1 + i], h1[i], s);
            h2[o] = lrelu(s);
        }
        unsigned long long* dst = packed + (size_t)id * H3;
        const unsigned long long low = (unsigned long long)(0xffffffffu - (unsigned)row);     // ties: the smallest row wins
#pragma unroll 4
        for (int o = 0; o < H3; o++) {
            float s = b3[o];
#pragma unroll
            for (int i = 0; i < H2; i++) s = fmaf(w3[o * H2 + i], h2[i], s);
            atomicMax(dst + o, ((unsigned long long)pn_ordered(lrelu(s)) << 32) | low);
        }
    }
}

// out [nv_rows x 2*H3] = [ max | barycentric weight of the winning row ], arg [nv_rows x H3] (rows_total = "none / masked")
template <int D>
__global__ void __launch_bounds__(256)
pn_finish_kernel(const unsigned long long* __restrict__ packed, const float* __restrict__ weights, const float* __restrict__ acc,
                 int nv_rows, int h3, long long rows_total, int vertex0_quirk, int min_points, float* __restrict__ out,
                 int* __restrict__ arg) {
    LN_PDL_ENTRY();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nv_rows * h3) return;
    const int v = (int)(t / h3), c = (int)(t - (long long)v * h3);
    const unsigned long long key = packed[t];
    float mx = 0.0f, bary = 0.0f;
    int a = (int)rows_total;
    const bool masked = __ldg(acc + (size_t)v * (D + 1) + D) < (float)min_points || (vertex0_quirk && v == 0);
    if (key != 0ull && !masked) {
        a = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
        mx = pn_unordered((unsigned)(key >> 32));
        bary = __ldg(weights + a);
    }
    out[(size_t)v * 2 * h3 + c] = mx;
    out[(size_t)v * 2 * h3 + h3 + c] = bary;
    arg[t] = a;
}

// ---- backward ------------------------------------------------------------------------------------------------------
// grad accumulators (one flat zeroed buffer): dW1 [H1 x IN] | dW2 [H2 x H1] | dW3 [H3 x H2] | db1 | db2 | db3
template <int IN, int H1, int H2, int H3>
struct PnAccLayout {
    static constexpr int w1 = 0, w2 = w1 + H1 * IN, w3 = w2 + H2 * H1, b1 = w3 + H3 * H2, b2 = b1 + H1, b3 = b2 + H2, total = b3 + H3;
};

template <int D, int IN, int H1, int H2, int H3>
__global__ void __launch_bounds__(kPnThreads)
pn_mlp_bwd_kernel(const float* __restrict__ positions_raw, const float* __restrict__ sigmas, const float* __restrict__ values,
                  const int* __restrict__ indices, int n, PnLayers L, const float* __restrict__ acc, int vertex0_quirk,
                  const float* __restrict__ grad_reduced /* [nv x 2*H3] */, const int* __restrict__ arg, float* __restrict__ gacc) {
    using A = PnAccLayout<IN, H1, H2, H3>;
    LN_PDL_ENTRY();
    constexpr int kStride = H3 + 2 * H2 + 2 * H1 + IN + 1;       // odd for the reference widths: conflict-free row-parallel stores
    extern __shared__ __align__(16) float pn_smem[];
    float* w1 = pn_smem;
    float* w2 = w1 + H1 * IN;
    float* w3 = w2 + H2 * H1;
    float* b1 = w3 + H3 * H2;
    float* b2 = b1 + H1;
    float* b3 = b2 + H2;
    float* red = b3 + H3;
    float* tile = red + 4;                                        // [kPnThreads][kStride]: g3 | h2 | g2 | h1 | g1 | x0
    pn_load_weights<IN, H1, H2, H3>(L, w1, w2, w3, b1, b2, b3, red);
    constexpr int kE3 = (H3 * H2 + kPnThreads - 1) / kPnThreads, kE2 = (H2 * H1 + kPnThreads - 1) / kPnThreads,
                  kE1 = (H1 * IN + kPnThreads - 1) / kPnThreads;
    float a3[kE3], a2[kE2], a1[kE1], ab = 0.0f;                   // this thread's entries of dW3 / dW2 / dW1 and one bias entry
#pragma unroll
    for (int k = 0; k < kE3; k++) a3[k] = 0.0f;
#pragma unroll
    for (int k = 0; k < kE2; k++) a2[k] = 0.0f;
#pragma unroll
    for (int k = 0; k < kE1; k++) a1[k] = 0.0f;
    const long long rows = (long long)n * (D + 1);
    const long long n_tiles = (rows + kPnThreads - 1) / kPnThreads;
    for (long long tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
        const long long row = tl * kPnThreads + threadIdx.x;
        float* my = tile + (size_t)threadIdx.x * kStride;
        float* g3 = my;
        float* h2s = g3 + H3;
        float* g2s = h2s + H2;
        float* h1s = g2s + H2;
        float* g1s = h1s + H1;
        float* x0s = g1s + H1;
        int id = -1;
        if (row < rows) id = __ldg(indices + row);
        const bool live = id >= 0 && !(vertex0_quirk && id == 0);
        if (!live) {
            for (int k = 0; k < kStride - 1; k++) my[k] = 0.0f;
        } else {
            const int p = (int)(row / (D + 1));
            float x0[IN];
            pn_row_input<D, IN>(positions_raw, sigmas, values, acc, p, id, vertex0_quirk, x0);
            float pre1[H1], h1[H1], pre2[H2], h2[H2];
#pragma unroll
            for (int o = 0; o < H1; o++) {
                float s = b1[o];
#pragma unroll
                for (int i = 0; i < IN; i++) s = fmaf(w1[o * IN + i], x0[i], s);
                pre1[o] = s;
                h1[o] = lrelu(s);
            }
#pragma unroll
            for (int o = 0; o < H2; o++) {
                float s = b2[o];
#pragma unroll
                for (int i = 0; i < H1; i++) s = fmaf(w2[o * H1 + i], h1[i], s);
                pre2[o] = s;
                h2[o] = lrelu(s);
            }
            float g2[H2];
#pragma unroll
            for (int i = 0; i < H2; i++) g2[i] = 0.0f;
            const int* arow = arg + (size_t)id * H3;
            const float* grow = grad_reduced + (size_t)id * 2 * H3;
#pragma unroll 4
            for (int o = 0; o < H3; o++) {
                float g = 0.0f;
                if (__ldg(arow + o) == (int)row) {                 // this row won channel o of its vertex
                    float s = b3[o];
#pragma unroll
                    for (int i = 0; i < H2; i++) s = fmaf(w3[o * H2 + i], h2[i], s);
                    g = __ldg(grow + o) * lrelu_grad(s);
#pragma unroll
                    for (int i = 0; i < H2; i++) g2[i] = fmaf(g, w3[o * H2 + i], g2[i]);
                }
                g3[o] = g;
            }
            float g1[H1];
#pragma unroll
            for (int i = 0; i < H1; i++) g1[i] = 0.0f;
#pragma unroll
            for (int o = 0; o < H2; o++) {
                const float g = g2[o] * lrelu_grad(pre2[o]);
                g2s[o] = g;
                h2s[o] = h2[o];
#pragma unroll
                for (int i = 0; i < H1; i++) g1[i] = fmaf(g, w2[o * H1 + i], g1[i]);
            }
#pragma unroll
            for (int o = 0; o < H1; o++) {
                g1s[o] = g1[o] * lrelu_grad(pre1[o]);
                h1s[o] = h1[o];
            }
#pragma unroll
            for (int i = 0; i < IN; i++) x0s[i] = x0[i];
        }
        __syncthreads();
        // weight gradients of the tile: dW[o][i] += sum_rows G[row][o] * H[row][i]  (i fastest across a warp: the H reads
        // hit consecutive banks, the G read is a broadcast)
#pragma unroll
        for (int k = 0; k < kE3; k++) {
            const int e = threadIdx.x + k * kPnThreads;
            if (e < H3 * H2) {
                const int o = e / H2, i = e - o * H2;
                float s = a3[k];
#pragma unroll 8
                for (int r = 0; r < kPnThreads; r++) s = fmaf(tile[r * kStride + o], tile[r * kStride + H3 + i], s);
                a3[k] = s;
            }
        }
#pragma unroll
        for (int k = 0; k < kE2; k++) {
            const int e = threadIdx.x + k * kPnThreads;
            if (e < H2 * H1) {
                const int o = e / H1, i = e - o * H1;
                float s = a2[k];
#pragma unroll 8
                for (int r = 0; r < kPnThreads; r++) s = fmaf(tile[r * kStride + H3 + H2 + o], tile[r * kStride + H3 + 2 * H2 + i], s);
                a2[k] = s;
            }
        }
#pragma unroll
        for (int k = 0; k < kE1; k++) {
            const int e = threadIdx.x + k * kPnThreads;
            if (e < H1 * IN) {
                const int o = e / IN, i = e - o * IN;
                float s = a1[k];
#pragma unroll 8
                for (int r = 0; r < kPnThreads; r++) s = fmaf(tile[r * kStride + H3 + 2 * H2 + H1 + o], tile[r * kStride + H3 + 2 * H2 + 2 * H1 + i], s);
                a1[k] = s;
            }
        }
        if (threadIdx.x < H1 + H2 + H3) {      // bias gradients: column sums of g1 | g2 | g3
            const int j = threadIdx.x;
            const int col = j < H1 ? H3 + 2 * H2 + H1 + j : j < H1 + H2 ? H3 + H2 + (j - H1) : (j - H1 - H2);
            float s = ab;
#pragma unroll 8
            for (int r = 0; r < kPnThreads; r++) s += tile[r * kStride + col];
            ab = s;
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < kE3; k++) {
        const int e = threadIdx.x + k * kPnThreads;
        if (e < H3 * H2) atomicAdd(gacc + A::w3 + e, a3[k]);
    }
#pragma unroll
    for (int k = 0; k < kE2; k++) {
        const int e = threadIdx.x + k * kPnThreads;
        if (e < H2 * H1) atomicAdd(gacc + A::w2 + e, a2[k]);
    }
#pragma unroll
    for (int k = 0; k < kE1; k++) {
        const int e = threadIdx.x + k * kPnThreads;
        if (e < H1 * IN) atomicAdd(gacc + A::w1 + e, a1[k]);
    }
    if (threadIdx.x < H1 + H2 + H3) atomicAdd(gacc + A::b1 + threadIdx.x, ab);
}

// weight-norm backward, one CTA per layer: w = v * g[o] / n, n = ||v||_F
//   dg[o] = sum_i dW[o,i] v[o,i] / n;   dv = dW * g[o] / n  -  v * (sum_{o,i} dW[o,i] v[o,i] g[o]) / n^3
__global__ void __launch_bounds__(kPnThreads)
pn_wn_bwd_kernel(PnLayers L, PnGrads G, const float* __restrict__ gacc, int in0, int h1, int h2, int h3) {
    LN_PDL_ENTRY();
    __shared__ float red[4];
    const int l = blockIdx.x;
    const int outs[3] = {h1, h2, h3}, ins[3] = {in0, h1, h2};
    const int woff[3] = {0, h1 * in0, h1 * in0 + h2 * h1};
    const int boff[3] = {woff[2] + h3 * h2, woff[2] + h3 * h2 + h1, woff[2] + h3 * h2 + h1 + h2};
    const int out = outs[l], in = ins[l], size = out * in;
    const float* v = L.v[l];
    const float* g = L.g[l];
    const float* dw = gacc + woff[l];
    float s = 0.0f, dot = 0.0f;
    for (int i = threadIdx.x; i < size; i += kPnThreads) {
        const float x = __ldg(v + i);
        s = fmaf(x, x, s);
        dot = fmaf(dw[i] * x, __ldg(g + i / in), dot);
    }
    const float n2 = pn_block_sum(s, red);
    const float total = pn_block_sum(dot, red);
    const float n = sqrtf(n2);
    for (int i = threadIdx.x; i < size; i += kPnThreads)
        G.v[l][i] = dw[i] * (__ldg(g + i / in) / n) - __ldg(v + i) * (total / (n2 * n));
    for (int o = threadIdx.x; o < out; o += kPnThreads) {
        float d = 0.0f;
        for (int i = 0; i < in; i++) d = fmaf(dw[o * in + i], __ldg(v + o * in + i), d);
        G.g[l][o] = d / n;
        G.b[l][o] = gacc[boff[l] + o];
    }
}

template <int D, int IN>
static int pn_forward(const float* positions_raw, const float* sigmas, const float* values, const int* indices, const float* weights, int n,
                      const PnLayers& L, int nv_rows, int vertex0_quirk, int min_points, float* acc, unsigned long long* packed, float* out,
                      int* arg, cudaStream_t s) {
    const long long rows = (long long)n * (D + 1);
    launch_k(pn_mean_kernel<D>, dim3(cdiv(rows, 256)), dim3(256), 0, s, positions_raw, sigmas, indices, n, acc);
    const int grid = (int)min((long long)148 * 8, (rows + kPnThreads - 1) / kPnThreads);
    launch_k((pn_mlp_max_kernel<D, IN, 16, 32, 64>), dim3(grid), dim3(kPnThreads), 0, s, positions_raw, sigmas, values, indices, n, L, acc, vertex0_quirk, packed);
    launch_k(pn_finish_kernel<D>, dim3(cdiv((long long)nv_rows * 64, 256)), dim3(256), 0, s, packed, weights, acc, nv_rows, 64, rows, vertex0_quirk, min_points, out, arg);
    count_launch(3);
    return check_launch("pointnet_fwd");
}

template <int D, int IN>
static int pn_backward(const float* positions_raw, const float* sigmas, const float* values, const int* indices, int n, const PnLayers& L,
                       const PnGrads& G, int vertex0_quirk, const float* acc, const float* grad_reduced, const int* arg, float* gacc,
                       cudaStream_t s) {
    constexpr int H1 = 16, H2 = 32, H3 = 64;
    constexpr int kStride = H3 + 2 * H2 + 2 * H1 + IN + 1;
    const size_t smem = (size_t)(H1 * IN + H2 * H1 + H3 * H2 + H1 + H2 + H3 + 4 + kPnThreads * kStride) * sizeof(float);
    const void* kern = (const void*)pn_mlp_bwd_kernel<D, IN, H1, H2, H3>;
    const cudaError_t err = allow_max_smem(kern);
    if (err != cudaSuccess) {
        set_error("pointnet_bwd: %s", cudaGetErrorString(err));
        return LN_ERR_CUDA;
    }
    const long long rows = (long long)n * (D + 1);
    const int grid = (int)min((long long)148 * 2, (rows + kPnThreads - 1) / kPnThreads);
    launch_k((pn_mlp_bwd_kernel<D, IN, H1, H2, H3>), dim3(grid), dim3(kPnThreads), smem, s, positions_raw, sigmas, values, indices, n, L, acc, vertex0_quirk,
                                                                         grad_reduced, arg, gacc);
    launch_k(pn_wn_bwd_kernel, dim3(3), dim3(kPnThreads), 0, s, L, G, gacc, IN, H1, H2, H3);
    count_launch(2);
    return check_launch("pointnet_bwd");
}

}  // namespace ln

using namespace ln;

extern "C" {

// 1 if the (pos_dim, val_dim, layer widths) combination is built
int ln_pointnet_supported(int pos_dim, int val_dim, int h1, int h2, int h3) {
    const int in = pos_dim + val_dim;
    return ((pos_dim == 3 && in >= 4 && in <= 8) || (pos_dim == 5 && in >= 6 && in <= 9)) && h1 == 16 && h2 == 32 && h3 == 64;
}
// floats of zeroed scratch the forward needs: per-vertex sums + counts, then the packed maxima (2 floats each)
long long ln_pointnet_scratch_floats(int pos_dim, int nv_rows, int h3) { return (long long)nv_rows * (pos_dim + 1) + 2ll * nv_rows * h3 + 4; }
long long ln_pointnet_grad_scratch_floats(int pos_dim, int val_dim, int h1, int h2, int h3) {
    const int in = pos_dim + val_dim;
    return (long long)h1 * in + h2 * h1 + h3 * h2 + h1 + h2 + h3;
}

#define LN_PN_DISPATCH(CALL3, CALL5)                                           \
    switch (pos_dim * 16 + pos_dim + val_dim) {                                 \
        case 3 * 16 + 4: return CALL3(4);                                       \
        case 3 * 16 + 5: return CALL3(5);                                       \
        case 3 * 16 + 6: return CALL3(6);                                       \
        case 3 * 16 + 7: return CALL3(7);                                       \
        case 3 * 16 + 8: return CALL3(8);                                       \
        case 5 * 16 + 6: return CALL5(6);                                       \
        case 5 * 16 + 7: return CALL5(7);                                       \
        case 5 * 16 + 8: return CALL5(8);                                       \
        case 5 * 16 + 9: return CALL5(9);                                       \
    }

int ln_pointnet_fwd(const float* positions_raw, const float* sigmas, const float* values, const int* indices, const float* weights,
                    int n, int pos_dim, int val_dim, const float* const* layer_ptrs /* v0,g0,b0,v1,g1,b1,v2,g2,b2 (host array) */,
                    int h1, int h2, int h3, int nv_rows, int vertex0_quirk, int min_points, float* scratch_zeroed, float* out, int* arg,
                    void* stream) {
    LN_REQUIRE(positions_raw && sigmas && values && indices && weights && layer_ptrs && scratch_zeroed && out && arg, "ln_pointnet_fwd: null pointer");
    LN_REQUIRE(n >= 0 && nv_rows >= 1, "ln_pointnet_fwd: bad size");
    LN_REQUIRE(ln_pointnet_supported(pos_dim, val_dim, h1, h2, h3), "ln_pointnet_fwd: pos_dim %d / val_dim %d / widths %d,%d,%d not built", pos_dim,
               val_dim, h1, h2, h3);
    PnLayers L;
    for (int l = 0; l < 3; l++) {
        L.v[l] = layer_ptrs[3 * l];
        L.g[l] = layer_ptrs[3 * l + 1];
        L.b[l] = layer_ptrs[3 * l + 2];
    }
    float* acc = scratch_zeroed;
    // 8-byte aligned start of the packed maxima
    unsigned long long* packed = reinterpret_cast<unsigned long long*>(scratch_zeroed + (((size_t)nv_rows * (pos_dim + 1) + 1) & ~(size_t)1));
    cudaStream_t s = (cudaStream_t)stream;
#define LN_PN_F3(IN) pn_forward<3, IN>(positions_raw, sigmas, values, indices, weights, n, L, nv_rows, vertex0_quirk, min_points, acc, packed, out, arg, s)
#define LN_PN_F5(IN) pn_forward<5, IN>(positions_raw, sigmas, values, indices, weights, n, L, nv_rows, vertex0_quirk, min_points, acc, packed, out, arg, s)
    LN_PN_DISPATCH(LN_PN_F3, LN_PN_F5)
#undef LN_PN_F3
#undef LN_PN_F5
    return LN_ERR_UNSUPPORTED;
}

int ln_pointnet_bwd(const float* positions_raw, const float* sigmas, const float* values, const int* indices, int n, int pos_dim,
                    int val_dim, const float* const* layer_ptrs, float* const* grad_ptrs /* dv0,dg0,db0,... (host array) */, int h1, int h2,
                    int h3, int vertex0_quirk, const float* fwd_scratch, const float* grad_reduced, const int* arg, float* grad_scratch_zeroed,
                    void* stream) {
    LN_REQUIRE(positions_raw && sigmas && values && indices && layer_ptrs && grad_ptrs && fwd_scratch && grad_reduced && arg && grad_scratch_zeroed,
               "ln_pointnet_bwd: null pointer");
    LN_REQUIRE(ln_pointnet_supported(pos_dim, val_dim, h1, h2, h3), "ln_pointnet_bwd: shape not built");
    PnLayers L;
    PnGrads G;
    for (int l = 0; l < 3; l++) {
        L.v[l] = layer_ptrs[3 * l];
        L.g[l] = layer_ptrs[3 * l + 1];
        L.b[l] = layer_ptrs[3 * l + 2];
        G.v[l] = grad_ptrs[3 * l];
        G.g[l] = grad_ptrs[3 * l + 1];
        G.b[l] = grad_ptrs[3 * l + 2];
    }
    cudaStream_t s = (cudaStream_t)stream;
#define LN_PN_B3(IN) pn_backward<3, IN>(positions_raw, sigmas, values, indices, n, L, G, vertex0_quirk, fwd_scratch, grad_reduced, arg, grad_scratch_zeroed, s)
#define LN_PN_B5(IN) pn_backward<5, IN>(positions_raw, sigmas, values, indices, n, L, G, vertex0_quirk, fwd_scratch, grad_reduced, arg, grad_scratch_zeroed, s)
    LN_PN_DISPATCH(LN_PN_B3, LN_PN_B5)
#undef LN_PN_B3
#undef LN_PN_B5
    return LN_ERR_UNSUPPORTED;
}

}  // extern "C"
