// Lattice convolution as an implicit GEMM over the neighbour table -- exact-fp32 SIMT path,
// weight gradient, and the filter re-layout for the data gradient.  The tcgen05 tensor-core path
// lives in ln_conv_tc.cu and is selected with precision = 1 / 2.
//
// Reference: Lattice::convolve_im2row_standalone (/root/reference/src/Lattice.cu:424-474) =
// im2row kernel (LatticeGPU.cuh:1464-1688) + cuBLAS SGEMM; backward algebra in
// /root/reference/latticenet_py/lattice/lattice_funcs.py:294-313.
#include <mutex>
#include "ln_common.cuh"

namespace ln {

// ln_conv_tc.cu
int conv_fwd_tc(const float* nbr_values, const int* neighbours, const float* slabs, const float* bias, const float* residual, int nv_query,
                int F, int c_in, int c_out, int flip, int precision, float* out, int cta_budget, cudaStream_t s);
int filter_prepare(const float* filter, int F, int c_in, int c_out, int transposed, int precision, float* slabs, cudaStream_t s);
int filter_prepare_batch(const void* jobs_device, int n_jobs, long long total_threads, cudaStream_t s);
size_t conv_tc_workspace_bytes(int F, int c_in, int c_out);
bool conv_tc_supported(int F, int c_in, int c_out);
bool conv_tc_needs_zero(int nv_query, int F, int c_in);
bool conv_wgrad_tc_supported(int F, int c_in, int c_out);
bool conv_wgrad_tc_needs_zero(int nv_query, int F, int c_in);
int conv_wgrad_tc(const float* nbr_values, const int* neighbours, const float* grad_out, int nv_query, int F, int c_in,
                  int c_out, int precision, float* grad_filter, int cta_budget, cudaStream_t s);

constexpr int kThreads = 256;
constexpr int BM = 64, BN = 64, BK = 16;

// out[q0:q0+64, n0:n0+64] tile per block, 4x4 outputs per thread, K walked slot by slot.
__global__ void __launch_bounds__(kThreads)
conv_fwd_simt_kernel(const float* __restrict__ values, const int* __restrict__ neighbours,
                     const float* __restrict__ filter, const float* __restrict__ bias, const float* __restrict__ residual,
                     int nv_query, int F, int c_in, int c_out, int flip, int transposed, float* __restrict__ out) {
    LN_PDL_ENTRY();
    __shared__ float a_sh[BK][BM + 4];
    __shared__ float b_sh[BK][BN + 4];
    __shared__ int nbr_sh[BM];
    const int q0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;

    const int a_row = tid >> 2;          // 0..63
    const int a_col = (tid & 3) * 4;     // 0,4,8,12
    const int b_row = tid >> 4;          // 0..15
    const int b_col = (tid & 15) * 4;    // 0..60

    for (int slot = 0; slot < F; slot++) {
        const int src_slot = (flip && slot < F - 1) ? (slot ^ 1) : slot;
        __syncthreads();
        if (tid < BM) {
            const int q = q0 + tid;
            nbr_sh[tid] = (q < nv_query) ? __ldg(neighbours + (size_t)q * F + src_slot) : -1;
        }
        __syncthreads();
        for (int c0 = 0; c0 < c_in; c0 += BK) {
            // stage A: gathered neighbour rows (zeros where the neighbour is absent)
            {
                const int id = nbr_sh[a_row];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int c = c0 + a_col + k;
                    a_sh[a_col + k][a_row] = (id >= 0 && c < c_in) ? __ldg(values + (size_t)id * c_in + c) : 0.0f;
                }
            }
            // stage B: filter rows slot*c_in + c0 .. +BK
            {
                const int c = c0 + b_row;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int n = n0 + b_col + k;
                    float wv = 0.0f;
                    if (c < c_in && n < c_out)
                        wv = transposed ? __ldg(filter + ((size_t)slot * c_out + n) * c_in + c)      // forward bank read as its own transpose
                                        : __ldg(filter + ((size_t)slot * c_in + c) * c_out + n);
                    b_sh[b_row][b_col + k] = wv;
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; k++) {
                float a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; i++) a[i] = a_sh[k][ty * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; j++) b[j] = b_sh[k][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int q = q0 + ty * 4 + i;
        if (q >= nv_query) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int n = n0 + tx * 4 + j;
            if (n < c_out)
                out[(size_t)q * c_out + n] = acc[i][j] + (bias ? __ldg(bias + n) : 0.0f) + (residual ? __ldg(residual + (size_t)q * c_out + n) : 0.0f);
        }
    }
}

// grad_filter[slot*c_in + ci, co] += sum_{q in chunk} values[nbr[q,slot], ci] * grad_out[q, co]
// grid: x = q-chunk, y = (ci tile, co tile), z = slot
__global__ void __launch_bounds__(kThreads)
conv_wgrad_simt_kernel(const float* __restrict__ values, const int* __restrict__ neighbours,
                       const float* __restrict__ grad_out, int nv_query, int F, int c_in, int c_out, int q_chunk,
                       int co_tiles, float* __restrict__ grad_filter) {
    LN_PDL_ENTRY();
    __shared__ float a_sh[BK][BM + 4];   // [q][ci]
    __shared__ float g_sh[BK][BN + 4];   // [q][co]
    const int slot = blockIdx.z;
    const int ci0 = (blockIdx.y / co_tiles) * BM;
    const int co0 = (blockIdx.y % co_tiles) * BN;
    const int q_begin = blockIdx.x * q_chunk;
    const int q_end = min(q_begin + q_chunk, nv_query);
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;
    const int l_row = tid >> 4;          // 0..15  (q within the step)
    const int l_col = (tid & 15) * 4;    // 0..60
    for (int qs = q_begin; qs < q_end; qs += BK) {
        const int q = qs + l_row;
        const int id = (q < q_end) ? __ldg(neighbours + (size_t)q * F + slot) : -1;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int ci = ci0 + l_col + k;
            a_sh[l_row][l_col + k] = (id >= 0 && ci < c_in) ? __ldg(values + (size_t)id * c_in + ci) : 0.0f;
            const int co = co0 + l_col + k;
            g_sh[l_row][l_col + k] = (q < q_end && co < c_out) ? __ldg(grad_out + (size_t)q * c_out + co) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; k++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = a_sh[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = g_sh[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int ci = ci0 + ty * 4 + i;
        if (ci >= c_in) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int co = co0 + tx * 4 + j;
            if (co < c_out) atomicAdd(grad_filter + ((size_t)slot * c_in + ci) * c_out + co, acc[i][j]);
        }
    }
}

// filter_bw[(slot*c_out + co), ci] = filter[(slot*c_in + ci), co]
__global__ void __launch_bounds__(kThreads)
filter_for_dgrad_kernel(const float* __restrict__ filter, int c_in, int c_out, float* __restrict__ filter_bw) {
    __shared__ float tile[32][33];
    const int slot = blockIdx.z;
    const int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const float* src = filter + (size_t)slot * c_in * c_out;
    float* dst = filter_bw + (size_t)slot * c_in * c_out;
    for (int r = ty; r < 32; r += 8) {
        const int ci = ci0 + r, co = co0 + tx;
        tile[r][tx] = (ci < c_in && co < c_out) ? __ldg(src + (size_t)ci * c_out + co) : 0.0f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int co = co0 + r, ci = ci0 + tx;
        if (co < c_out && ci < c_in) dst[(size_t)co * c_in + ci] = tile[tx][r];
    }
}

}  // namespace ln

using namespace ln;

// Second stream for the weight gradient: it depends on grad_out and the forward inputs only, never on the data
// gradient, so ln_conv_bwd runs the two kernels side by side (fork at entry, join before returning).  Inside a
// CUDA-graph capture the fork / join become graph edges.
namespace {
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
SideStream* side_stream() {
    static SideStream per_device[64];
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    SideStream& ss = per_device[dev];
    if (ss.stream == nullptr) {
        if (cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) != cudaSuccess) {
            ss.stream = nullptr;
            cudaGetLastError();
            return nullptr;
        }
    }
    return &ss;
}

int conv_wgrad_launch(const float* nbr_values, const int* neighbours, const float* grad_out, int nv_query,
                      int filter_extent, int c_in, int c_out, int precision, float* grad_filter, bool already_zero,
                      int cta_budget, cudaStream_t s) {
    const size_t bytes = (size_t)filter_extent * c_in * c_out * sizeof(float);
    const bool tc = precision != 0 && conv_wgrad_tc_supported(filter_extent, c_in, c_out);
    const bool needs_zero = nv_query == 0 || !tc || conv_wgrad_tc_needs_zero(nv_query, filter_extent, c_in);
    if (needs_zero && !already_zero && cudaMemsetAsync(grad_filter, 0, bytes, s) != cudaSuccess) return check_launch("conv_wgrad memset");
    if (nv_query == 0) return LN_OK;
    if (tc) return conv_wgrad_tc(nbr_values, neighbours, grad_out, nv_query, filter_extent, c_in, c_out, precision, grad_filter, cta_budget, s);
    const int ci_tiles = cdiv(c_in, BM), co_tiles = cdiv(c_out, BN);
    // enough q-chunks to fill the machine (~4 waves of 148 SMs), at least 256 rows each
    const int tiles = ci_tiles * co_tiles * filter_extent;
    int chunks = max(1, min(cdiv(nv_query, 256), cdiv(148 * 4, tiles)));
    int q_chunk = cdiv(cdiv(nv_query, chunks), BK) * BK;
    chunks = cdiv(nv_query, q_chunk);
    dim3 grid(chunks, ci_tiles * co_tiles, filter_extent);
    launch_k(conv_wgrad_simt_kernel, dim3(grid), dim3(kThreads), 0, s, nbr_values, neighbours, grad_out, nv_query, filter_extent, c_in, c_out, q_chunk, co_tiles, grad_filter);
    count_launch();
    return check_launch("conv_wgrad_simt");
}

// out = conv(values through `neighbours`) [+ bias] [+ residual]; tensor cores when the shape allows and precision != 0
int conv_launch(const float* nbr_values, const int* neighbours, const float* filter, const float* bias, const float* residual, int nv_query,
                int filter_extent, int c_in, int c_out, int flip, int transposed_filter, int precision, float* slabs, int slabs_prepared,
                int out_is_zero, float* out, int cta_budget, cudaStream_t s, const char* what) {
    if (precision != 0 && conv_tc_supported(filter_extent, c_in, c_out)) {
        if (slabs == nullptr) {
            set_error("%s: precision %d needs a slab buffer of ln_conv_workspace_bytes() bytes", what, precision);
            return LN_ERR_BAD_ARG;
        }
        if (!slabs_prepared) {
            const int rc = filter_prepare(filter, filter_extent, c_in, c_out, transposed_filter, precision, slabs, s);
            if (rc != LN_OK) return rc;
        }
        if (!out_is_zero && conv_tc_needs_zero(nv_query, filter_extent, c_in) &&
            cudaMemsetAsync(out, 0, (size_t)nv_query * c_out * sizeof(float), s) != cudaSuccess)
            return check_launch("conv memset");
        return conv_fwd_tc(nbr_values, neighbours, slabs, bias, residual, nv_query, filter_extent, c_in, c_out, flip, precision, out, cta_budget, s);
    }
    dim3 grid(cdiv(nv_query, BM), cdiv(c_out, BN));
    launch_k(conv_fwd_simt_kernel, dim3(grid), dim3(kThreads), 0, s, nbr_values, neighbours, filter, bias, residual, nv_query, filter_extent, c_in, c_out, flip,
                                                   transposed_filter, out);
    count_launch();
    return check_launch(what);
}
}  // namespace

extern "C" {

long long ln_conv_workspace_bytes(int filter_extent, int c_in, int c_out, int precision) {
    if (precision == 0 || !conv_tc_supported(filter_extent, c_in, c_out)) return 0;
    return (long long)conv_tc_workspace_bytes(filter_extent, c_in, c_out);
}

int ln_conv_needs_zero(int nv_query, int filter_extent, int c_in, int c_out, int precision) {
    return (precision != 0 && conv_tc_supported(filter_extent, c_in, c_out) && conv_tc_needs_zero(nv_query, filter_extent, c_in)) ? 1 : 0;
}

int ln_filter_prepare(const float* filter, int filter_extent, int c_in, int c_out, int transposed_filter, int precision, float* slabs,
                      void* stream) {
    LN_REQUIRE(filter && slabs, "ln_filter_prepare: null pointer");
    LN_REQUIRE(precision == 1 || precision == 2, "ln_filter_prepare: precision must be 1 (3xTF32) or 2 (TF32)");
    LN_REQUIRE(conv_tc_supported(filter_extent, c_in, c_out), "ln_filter_prepare: shape F=%d c_in=%d c_out=%d does not run on the tensor cores",
               filter_extent, c_in, c_out);
    return filter_prepare(filter, filter_extent, c_in, c_out, transposed_filter, precision, slabs, (cudaStream_t)stream);
}

int ln_filter_prepare_batch(const void* jobs_device, int n_jobs, long long total_threads, void* stream) {
    LN_REQUIRE(jobs_device != nullptr || n_jobs == 0, "ln_filter_prepare_batch: null pointer");
    return filter_prepare_batch(jobs_device, n_jobs, total_threads, (cudaStream_t)stream);
}

int ln_conv_fwd(const float* nbr_values, const int* neighbours, const float* filter, const float* bias, const float* residual,
                int nv_query, int filter_extent, int c_in, int c_out, int flip, int transposed_filter, int precision, float* slabs,
                int slabs_prepared, int out_is_zero, float* out, void* stream) {
    LN_REQUIRE(nbr_values && neighbours && filter && out, "ln_conv_fwd: null pointer");
    LN_REQUIRE(nv_query >= 0 && filter_extent >= 1 && (filter_extent & 1) && c_in >= 1 && c_out >= 1, "ln_conv_fwd: bad size");
    LN_REQUIRE(precision >= 0 && precision <= 2, "ln_conv_fwd: precision must be 0 (fp32), 1 (3xTF32) or 2 (TF32)");
    if (nv_query == 0) return LN_OK;
    return conv_launch(nbr_values, neighbours, filter, bias, residual, nv_query, filter_extent, c_in, c_out, flip, transposed_filter, precision,
                       slabs, slabs_prepared, out_is_zero, out, 0, (cudaStream_t)stream, "conv_fwd_simt");
}

int ln_conv_wgrad(const float* nbr_values, const int* neighbours, const float* grad_out, int nv_query,
                  int filter_extent, int c_in, int c_out, int precision, int grad_is_zero, float* grad_filter, void* stream) {
    LN_REQUIRE(nbr_values && neighbours && grad_out && grad_filter, "ln_conv_wgrad: null pointer");
    LN_REQUIRE(nv_query >= 0 && filter_extent >= 1 && c_in >= 1 && c_out >= 1, "ln_conv_wgrad: bad size");
    LN_REQUIRE(precision >= 0 && precision <= 2, "ln_conv_wgrad: precision must be 0 (fp32), 1 (3xTF32) or 2 (TF32)");
    return conv_wgrad_launch(nbr_values, neighbours, grad_out, nv_query, filter_extent, c_in, c_out, precision, grad_filter, grad_is_zero != 0,
                             0, (cudaStream_t)stream);
}

int ln_conv_bwd(const float* nbr_values, const int* neighbours_fwd, const float* grad_out, const int* neighbours_bwd,
                const float* filter, int nv_query, int nv_nbr, int filter_extent, int c_in, int c_out, int precision,
                float* slabs_bwd, int slabs_prepared, float* grad_nbr_values, int grad_nbr_is_zero, float* grad_filter,
                int grad_filter_is_zero, int linear_weight, int defer_join, void* stream) {
    LN_REQUIRE(nbr_values && neighbours_fwd && grad_out && filter, "ln_conv_bwd: null pointer");
    LN_REQUIRE(nv_query >= 0 && nv_nbr >= 0 && filter_extent >= 1 && c_in >= 1 && c_out >= 1, "ln_conv_bwd: bad size");
    LN_REQUIRE(grad_nbr_values == nullptr || neighbours_bwd != nullptr, "ln_conv_bwd: the data gradient needs the reverse neighbour table");
    LN_REQUIRE(!linear_weight || filter_extent == 1, "ln_conv_bwd: a Linear weight is a filter bank of extent 1");
    cudaStream_t s = (cudaStream_t)stream;
    const bool want_dgrad = grad_nbr_values != nullptr && nv_nbr > 0;
    // weight gradient on the side stream while the data gradient runs on the caller's
    SideStream* ss = (want_dgrad && grad_filter != nullptr && nv_query > 0) ? side_stream() : nullptr;
    if (ss != nullptr && (cudaEventRecord(ss->fork, s) != cudaSuccess || cudaStreamWaitEvent(ss->stream, ss->fork, 0) != cudaSuccess))
        return check_launch("conv_bwd fork");
    // both kernels take an SM per CTA: when they run side by side each gets half the machine, so neither waits for the other's SMs
    int half = 0;
    if (ss != nullptr) {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        // (only in the latency-bound regime of small lattices; scene-sized levels keep persistent full-machine grids)
        if (cdiv(nv_nbr, 128) < sms && cdiv(nv_query, 128) < sms) half = sms / 2;
    }
    int rw = LN_OK;
    if (grad_filter != nullptr) {
        if (linear_weight)   // dW [c_out x c_in] = G^T X: the weight gradient of the transposed problem (roles of X and G swapped)
            rw = conv_wgrad_launch(grad_out, neighbours_fwd, nbr_values, nv_query, 1, c_out, c_in, precision, grad_filter, grad_filter_is_zero != 0,
                                   half, ss ? ss->stream : s);
        else
            rw = conv_wgrad_launch(nbr_values, neighbours_fwd, grad_out, nv_query, filter_extent, c_in, c_out, precision, grad_filter,
                                   grad_filter_is_zero != 0, half, ss ? ss->stream : s);
    }
    int rd = LN_OK;
    if (want_dgrad)   // flipped convolution of grad_out at the neighbour lattice's vertices, forward bank read transposed (c_in <-> c_out);
                      // a Linear weight [c_out x c_in] is stored transposed already: dx = dy W is its plain reading
        rd = conv_launch(grad_out, neighbours_bwd, filter, nullptr, nullptr, nv_nbr, filter_extent, c_out, c_in, linear_weight ? 0 : 1,
                         linear_weight ? 0 : 1, precision, slabs_bwd, slabs_prepared, grad_nbr_is_zero, grad_nbr_values, half, s,
                         "conv_dgrad_simt");
    // defer_join: the caller promises to call ln_conv_bwd_join() before anything reads grad_filter and to keep nbr_values /
    // grad_out alive until then -- the weight gradients of a whole backward pass then trail the data-gradient chain on
    // the side stream instead of holding it up layer by layer
    if (ss != nullptr && !defer_join && (cudaEventRecord(ss->join, ss->stream) != cudaSuccess || cudaStreamWaitEvent(s, ss->join, 0) != cudaSuccess))
        return check_launch("conv_bwd join");
    return rw != LN_OK ? rw : rd;
}

int ln_conv_bwd_join(void* stream) {
    SideStream* ss = side_stream();
    if (ss == nullptr) return LN_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaEventRecord(ss->join, ss->stream) != cudaSuccess || cudaStreamWaitEvent(s, ss->join, 0) != cudaSuccess) return check_launch("conv_bwd join");
    return LN_OK;
}

int ln_filter_for_dgrad(const float* filter, int filter_extent, int c_in, int c_out, float* filter_bw, void* stream) {
    LN_REQUIRE(filter && filter_bw, "ln_filter_for_dgrad: null pointer");
    LN_REQUIRE(filter_extent >= 1 && c_in >= 1 && c_out >= 1, "ln_filter_for_dgrad: bad size");
    dim3 grid(cdiv(c_out, 32), cdiv(c_in, 32), filter_extent);
    filter_for_dgrad_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(filter, c_in, c_out, filter_bw);
    count_launch();
    return check_launch("filter_for_dgrad");
}

}  // extern "C"
