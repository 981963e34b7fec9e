// The rest of the reference's training step around the lattice operators:
//   * the loss of ln_train.py:156-158 -- 0.5 * Lovasz-softmax (lovasz_loss.py:41-72) + 0.5 * NLL -- forward AND the
//     gradient w.r.t. the log-probabilities in one launch (the torch formulation is ~55 launches: one_hot, sort, gather,
//     two cumsums and two dozen elementwise kernels, twice over for autograd);
//   * AdamW with amsgrad (ln_train.py:163-165: torch.optim.AdamW(lr, weight_decay, amsgrad=True)) as ONE kernel over the
//     flat parameter / gradient buffers (torch's fused implementation is six multi-tensor launches over the 154 tensors).
#include "ln_common.cuh"

namespace ln {

// ---------------------------------------------------------------------------------------------
// Lovasz-softmax + NLL.  One CTA per class: errors e_i = |fg_i - p_ic| of all points are sorted in shared memory
// (bitonic, 64-bit keys = error bits | fg | point), the Jaccard gradient follows from an integer prefix sum of the sorted
// foreground flags, and G[i,c] = d loss_c / d logp_ic is scattered back through the point index.  Classes that are absent
// (or ignored) contribute nothing, exactly like the reference's `continue` (lovasz_loss.py:44-50).
// acc (zeroed by the caller): [0] sum of class losses, [1] classes present, [2] sum of -logp[label], [3] valid points,
// [4] CTAs done (as float).  result: [0] loss, [1] classes present, [2] valid points (read by the backward kernel).
constexpr int kLossThreads = 1024;

__device__ __forceinline__ float block_sum_1024(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = (threadIdx.x < 32) ? red[threadIdx.x] : 0.0f;
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

__global__ void __launch_bounds__(kLossThreads)
seg_loss_kernel(const float* __restrict__ logp, const long long* __restrict__ labels, int n, int nr_classes, int n_pad,
                int ignore_index, float* __restrict__ grad_lov, float* __restrict__ acc, float* __restrict__ result) {
    LN_PDL_ENTRY();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);          // [n_pad]
    int* cum = reinterpret_cast<int*>(keys + n_pad);                                      // [n_pad] inclusive prefix of sorted fg
    __shared__ float red[33];
    __shared__ int warp_tot[32];
    const int c = blockIdx.x;
    const int tid = threadIdx.x;
    const bool ignored = c == ignore_index;

    float gts_f = 0.0f, nll = 0.0f;
    for (int i = tid; i < n_pad; i += kLossThreads) {
        unsigned long long key = 0ull;
        if (i < n) {
            const long long lbl = labels[i];
            const bool fg = lbl == c;
            const float lp = __ldg(logp + (size_t)i * nr_classes + c);
            const float p = expf(lp);
            const float e = fabsf((fg ? 1.0f : 0.0f) - p);
            key = ((unsigned long long)__float_as_uint(e) << 32) | ((unsigned long long)(fg ? 1u : 0u) << 31) | (unsigned long long)(i + 1);   // > 0: sorts before every padding key
            if (fg) {
                gts_f += 1.0f;
                nll -= lp;
            }
        }
        keys[i] = key;
    }
    const float gts = block_sum_1024(gts_f, red);
    const float nll_c = block_sum_1024(nll, red);
    const bool present = gts > 0.0f && !ignored;      // block-uniform
    float loss_c = 0.0f;
    if (present) {
        // ---- bitonic sort, descending ----
        for (int k = 2; k <= n_pad; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                __syncthreads();
                for (int i = tid; i < n_pad; i += kLossThreads) {
                    const int l = i ^ j;
                    if (l > i) {
                        const unsigned long long a = keys[i], b = keys[l];
                        const bool desc = (i & k) == 0;
                        if (desc ? (a < b) : (a > b)) {
                            keys[i] = b;
                            keys[l] = a;
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- inclusive prefix sum of the sorted foreground flags: thread t owns elements [t*per, (t+1)*per) ----
        const int per = n_pad / kLossThreads > 0 ? n_pad / kLossThreads : 1;
        const int begin = tid * per;
        int local = 0;
        if (begin < n_pad)
            for (int q = 0; q < per; q++) local += (int)((keys[begin + q] >> 31) & 1ull);
        int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += up;
        }
        if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            int w = warp_tot[tid];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, w, o);
                if (tid >= o) w += up;
            }
            warp_tot[tid] = w;
        }
        __syncthreads();
        int run = incl - local + ((tid >> 5) > 0 ? warp_tot[(tid >> 5) - 1] : 0);      // exclusive prefix of this thread's range
        if (begin < n_pad)
            for (int q = 0; q < per; q++) {
                run += (int)((keys[begin + q] >> 31) & 1ull);
                cum[begin + q] = run;
            }
        __syncthreads();
        // ---- Jaccard gradient, class loss, scatter of d loss_c / d logp ----
        for (int j = tid; j < n_pad; j += kLossThreads) {
            const unsigned long long key = keys[j];
            const int i = (int)(key & 0x7fffffffull) - 1;
            const bool fg = ((key >> 31) & 1ull) != 0;
            const float e = __uint_as_float((unsigned)(key >> 32));
            const int cf = cum[j];
            // intersection = gts - cumsum(fg); union = gts + cumsum(1 - fg)   (lovasz_loss.py:8-20)
            const float jac = 1.0f - (gts - (float)cf) / (gts + (float)(j + 1 - cf));
            float g = jac;
            if (j > 0) {
                const int cfp = cf - (fg ? 1 : 0);
                g -= 1.0f - (gts - (float)cfp) / (gts + (float)(j - cfp));
            }
            if (j < n) {      // the n real keys are all > 0 and the padding keys 0: positions [0, n) hold exactly the points
                loss_c = fmaf(e, g, loss_c);
                // d e / d p = -1 (fg) / +1 (bg);  d p / d logp = p
                const float p = fg ? 1.0f - e : e;
                grad_lov[(size_t)i * nr_classes + c] = fg ? -g * p : g * p;
            }
        }
    } else {
        for (int i = tid; i < n; i += kLossThreads) grad_lov[(size_t)i * nr_classes + c] = 0.0f;
    }
    loss_c = block_sum_1024(loss_c, red);
    if (tid == 0) {
        if (present) {
            atomicAdd(acc + 0, loss_c);
            atomicAdd(acc + 1, 1.0f);
        }
        if (!ignored) {
            atomicAdd(acc + 2, nll_c);
            atomicAdd(acc + 3, gts);
        }
        __threadfence();
        const float done = atomicAdd(acc + 4, 1.0f);
        if (done == (float)(gridDim.x - 1)) {            // last class: combine
            __threadfence();
            const float lov = atomicAdd(acc + 0, 0.0f), np = atomicAdd(acc + 1, 0.0f);
            const float ns = atomicAdd(acc + 2, 0.0f), nvld = atomicAdd(acc + 3, 0.0f);
            result[0] = 0.5f * (lov / fmaxf(np, 1.0f)) + 0.5f * (nvld > 0.0f ? ns / nvld : 0.0f);
            result[1] = np;
            result[2] = nvld;
        }
    }
}

// grad_logp[i,c] = grad_loss * ( 0.5 * G[i,c] / classes_present  -  [c == label_i, c != ignore] * 0.5 / valid_points )
__global__ void __launch_bounds__(256)
seg_loss_bwd_kernel(const float* __restrict__ grad_lov, const long long* __restrict__ labels, const float* __restrict__ result,
                    const float* __restrict__ grad_loss, int n, int nr_classes, int ignore_index, float* __restrict__ grad_logp) {
    LN_PDL_ENTRY();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * nr_classes) return;
    const int i = (int)(t / nr_classes), c = (int)(t - (long long)i * nr_classes);
    const float go = __ldg(grad_loss);
    const float np = fmaxf(__ldg(result + 1), 1.0f), nvld = __ldg(result + 2);
    float g = 0.5f * __ldg(grad_lov + t) / np;
    if (labels[i] == c && c != ignore_index && nvld > 0.0f) g -= 0.5f / nvld;
    grad_logp[t] = go * g;
}

// ---------------------------------------------------------------------------------------------
// AdamW + amsgrad over flat buffers, torch.optim.AdamW's update order (torch/optim/adamw.py, single-tensor path):
//   p *= 1 - lr*wd;  m = lerp(m, g, 1-b1);  v = b2*v + (1-b2) g^2;  vmax = max(vmax, v);
//   p -= (lr / (1 - b1^t)) * m / (sqrt(vmax) / sqrt(1 - b2^t) + eps)
// state: [0] step count t (float, incremented here), [1] CTAs done.  skip (may be NULL): device float, != 0 = leave
// everything untouched (a cloud exceeded its vertex bound, graphed.py).  grad_scale: multiplied into g (1/world_size).
__global__ void __launch_bounds__(256)
adamw_amsgrad_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                     float* __restrict__ vmax, long long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                     float grad_scale, float* __restrict__ state, const float* __restrict__ skip) {
    LN_PDL_ENTRY();
    const bool skipped = skip != nullptr && *skip != 0.0f;
    const float step = state[0] + 1.0f;
    if (!skipped) {
        const float bc1 = 1.0f - powf(beta1, step);
        const float bc2_sqrt = sqrtf(1.0f - powf(beta2, step));
        const float step_size = lr / bc1;
        const float decay = 1.0f - lr * weight_decay;
        const long long n4 = n >> 2;
        const long long stride = (long long)gridDim.x * blockDim.x;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 pp = reinterpret_cast<float4*>(p)[i];
            float4 gg = __ldcs(reinterpret_cast<const float4*>(g) + i);
            float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i], xx = reinterpret_cast<float4*>(vmax)[i];
            float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x; float* xa = &xx.x;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float gk = ga[k] * grad_scale;
                pa[k] *= decay;
                ma[k] = ma[k] + (gk - ma[k]) * (1.0f - beta1);
                va[k] = va[k] * beta2 + (1.0f - beta2) * gk * gk;
                xa[k] = fmaxf(xa[k], va[k]);
                pa[k] -= step_size * (ma[k] / (sqrtf(xa[k]) / bc2_sqrt + eps));
            }
            reinterpret_cast<float4*>(p)[i] = pp;
            reinterpret_cast<float4*>(m)[i] = mm;
            reinterpret_cast<float4*>(v)[i] = vv;
            reinterpret_cast<float4*>(vmax)[i] = xx;
        }
        for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const float gk = g[i] * grad_scale;
            float pk = p[i] * decay;
            const float mk = m[i] + (gk - m[i]) * (1.0f - beta1);
            const float vk = v[i] * beta2 + (1.0f - beta2) * gk * gk;
            const float xk = fmaxf(vmax[i], vk);
            pk -= step_size * (mk / (sqrtf(xk) / bc2_sqrt + eps));
            p[i] = pk; m[i] = mk; v[i] = vk; vmax[i] = xk;
        }
    }
    // every CTA read state[0] above; the LAST one to get here advances the step count and resets the arrival counter
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const float done = atomicAdd(state + 1, 1.0f);
        if (done == (float)(gridDim.x - 1)) {
            if (!skipped) state[0] = step;
            state[1] = 0.0f;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Weight normalisation of a filter bank, reference utils.weight_norm_wrapper(cls, g_dim, v_dim=None)
// (/root/reference/latticenet_py/utils/utils.py:72-158): w = v * (g / ||v||_F), one gain per slice of dimension g_dim.
// v is [rows x cols]; gain_per_col != 0: g has `cols` entries (g_dim = 1, the lattice filter banks), else `rows`
// (g_dim = 0, Linear).  One CTA: the bank of the one weight-normalised convolution of LatticeNet is 37 k floats.
__global__ void __launch_bounds__(kLossThreads)
weight_norm_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g, int rows, int cols, int gain_per_col,
                       float* __restrict__ w) {
    LN_PDL_ENTRY();
    __shared__ float red[33];
    const int n = rows * cols;
    float s = 0.0f;
    for (int i = threadIdx.x; i < n; i += kLossThreads) {
        const float x = __ldg(v + i);
        s = fmaf(x, x, s);
    }
    const float norm = sqrtf(block_sum_1024(s, red));
    for (int i = threadIdx.x; i < n; i += kLossThreads) w[i] = __ldg(v + i) * (__ldg(g + (gain_per_col ? i % cols : i / cols)) / norm);
}

//   dg[j] = sum_{i in slice j} dw_i v_i / n;   dv_i = dw_i g[j(i)] / n  -  v_i (sum_i dw_i v_i g[j(i)]) / n^3
__global__ void __launch_bounds__(kLossThreads)
weight_norm_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ dw, int rows, int cols,
                       int gain_per_col, float* __restrict__ dv, float* __restrict__ dg) {
    LN_PDL_ENTRY();
    __shared__ float red[33];
    extern __shared__ float gain_acc[];          // [n_gain]: sum over each gain's slice of dw * v
    const int n = rows * cols;
    const int n_gain = gain_per_col ? cols : rows;
    for (int j = threadIdx.x; j < n_gain; j += kLossThreads) gain_acc[j] = 0.0f;
    __syncthreads();
    float s = 0.0f, dot = 0.0f;
    for (int i = threadIdx.x; i < n; i += kLossThreads) {
        const float x = __ldg(v + i);
        const int j = gain_per_col ? i % cols : i / cols;
        const float p = __ldg(dw + i) * x;
        s = fmaf(x, x, s);
        dot = fmaf(p, __ldg(g + j), dot);
        atomicAdd(gain_acc + j, p);              // shared-memory reduction: lanes of a warp hit consecutive (or one) gain
    }
    const float n2 = block_sum_1024(s, red);
    const float total = block_sum_1024(dot, red);
    const float norm = sqrtf(n2);
    for (int i = threadIdx.x; i < n; i += kLossThreads)
        dv[i] = __ldg(dw + i) * (__ldg(g + (gain_per_col ? i % cols : i / cols)) / norm) - __ldg(v + i) * (total / (n2 * norm));
    for (int j = threadIdx.x; j < n_gain; j += kLossThreads) dg[j] = gain_acc[j] / norm;
}

}  // namespace ln

using namespace ln;

extern "C" {

int ln_seg_loss_max_points(void) { return 8192; }

int ln_seg_loss_fwd(const float* logp, const long long* labels, int n, int nr_classes, int ignore_index, float* grad_lov,
                    float* acc_zeroed, float* result, void* stream) {
    LN_REQUIRE(logp && labels && grad_lov && acc_zeroed && result, "ln_seg_loss_fwd: null pointer");
    LN_REQUIRE(n >= 1 && n <= ln_seg_loss_max_points() && nr_classes >= 1, "ln_seg_loss_fwd: bad size n=%d (1..%d) nr_classes=%d", n,
               ln_seg_loss_max_points(), nr_classes);
    int n_pad = kLossThreads;
    while (n_pad < n) n_pad <<= 1;
    const size_t smem = (size_t)n_pad * (sizeof(unsigned long long) + sizeof(int));
    if (smem > 48 * 1024) {
        const cudaError_t err = allow_max_smem((const void*)seg_loss_kernel);
        if (err != cudaSuccess) {
            set_error("ln_seg_loss_fwd: %s", cudaGetErrorString(err));
            return LN_ERR_CUDA;
        }
    }
    launch_k(seg_loss_kernel, dim3(nr_classes), dim3(kLossThreads), smem, (cudaStream_t)stream, logp, labels, n, nr_classes, n_pad, ignore_index, grad_lov,
                                                                            acc_zeroed, result);
    count_launch();
    return check_launch("seg_loss");
}

int ln_seg_loss_bwd(const float* grad_lov, const long long* labels, const float* result, const float* grad_loss, int n,
                    int nr_classes, int ignore_index, float* grad_logp, void* stream) {
    LN_REQUIRE(grad_lov && labels && result && grad_loss && grad_logp, "ln_seg_loss_bwd: null pointer");
    LN_REQUIRE(n >= 1 && nr_classes >= 1, "ln_seg_loss_bwd: bad size");
    launch_k(seg_loss_bwd_kernel, dim3(cdiv((long long)n * nr_classes, 256)), dim3(256), 0, (cudaStream_t)stream, grad_lov, labels, result, grad_loss, n,
                                                                                                nr_classes, ignore_index, grad_logp);
    count_launch();
    return check_launch("seg_loss_bwd");
}

int ln_adamw_amsgrad(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, long long n,
                     float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale, float* state,
                     const float* skip, void* stream) {
    LN_REQUIRE(params && grads && exp_avg && exp_avg_sq && max_exp_avg_sq && state, "ln_adamw_amsgrad: null pointer");
    LN_REQUIRE(n >= 0, "ln_adamw_amsgrad: bad size");
    LN_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq | (uintptr_t)max_exp_avg_sq) & 15) == 0,
               "ln_adamw_amsgrad: buffers must be 16-byte aligned");
    if (n == 0) return LN_OK;
    const int grid = (int)min((long long)148 * 8, (n / 4 + 255) / 256 + 1);
    launch_k(adamw_amsgrad_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq, max_exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                  weight_decay, grad_scale, state, skip);
    count_launch();
    return check_launch("adamw_amsgrad");
}

int ln_weight_norm_fwd(const float* v, const float* g, int rows, int cols, int gain_per_col, float* w, void* stream) {
    LN_REQUIRE(v && g && w && rows >= 1 && cols >= 1, "ln_weight_norm_fwd: bad argument");
    launch_k(weight_norm_fwd_kernel, dim3(1), dim3(kLossThreads), 0, (cudaStream_t)stream, v, g, rows, cols, gain_per_col, w);
    count_launch();
    return check_launch("weight_norm_fwd");
}

int ln_weight_norm_bwd(const float* v, const float* g, const float* dw, int rows, int cols, int gain_per_col, float* dv, float* dg,
                       void* stream) {
    LN_REQUIRE(v && g && dw && dv && dg && rows >= 1 && cols >= 1, "ln_weight_norm_bwd: bad argument");
    const int n_gain = gain_per_col ? cols : rows;
    LN_REQUIRE(n_gain <= 8192, "ln_weight_norm_bwd: at most 8192 gains");
    launch_k(weight_norm_bwd_kernel, dim3(1), dim3(kLossThreads), (size_t)n_gain * sizeof(float), (cudaStream_t)stream, v, g, dw, rows, cols, gain_per_col, dv, dg);
    count_launch();
    return check_launch("weight_norm_bwd");
}

}  // extern "C"
