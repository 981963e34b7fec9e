// Barycentric gather / scatter kernels: slice, gather, fused slice+classify, and their backward
// passes.  Reference semantics: LatticeGPU.cuh:2552-2595 (slice), 2886-2929 (gather), 3387-3464
// (slice_classify), 3540-3623 / 3761-3817 / 3628-3756 (backwards).
//
// Layout rule used throughout: lanes run along the channel dimension of one vertex row, so every
// global access is a contiguous (vectorised where val_dim % 4 == 0) run -- the reference maps one
// thread to one point and walks channels serially (stride-V across the warp).
#include "ln_common.cuh"

namespace ln {

constexpr int kBlock = 256;
constexpr int kMaxSpv = 8;   // pos_dim + 1 <= 8

static inline int lanes_per_point(int vectors_per_row) {
    int lpp = 1;
    while (lpp < vectors_per_row && lpp < 32) lpp <<= 1;
    return lpp;
}
static inline int ilog2(int x) {
    int l = 0;
    while ((1 << l) < x) l++;
    return l;
}

// ---------------------------------------------------------------------------------------------
// L2 residency: the per-vertex table (values / gradient rows, re-read or re-accumulated (D+1)x per point in random
// order) is the only tensor with reuse; the per-point streams (indices, weights, sliced rows) are touched once.
// ncu on the 1M-point sweep showed 2x the algorithmic DRAM traffic when they compete for L2 on equal terms, so
// the table is accessed with an evict_last policy and the streams with evict-first (.cs) accesses.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ldg_v4_hint(const float* a, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ float ldg_f32_hint(const float* a, uint64_t pol) {
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ void red_v4_hint(float* a, float4 v, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_f32_hint(float* a, float v, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(a), "f"(v), "l"(pol) : "memory");
}

// ids / weights of one point: one 16-byte streaming load each when the simplex has four vertices
template <int SPV>
__device__ __forceinline__ void load_simplex(const int* __restrict__ indices, const float* __restrict__ weights, long long p,
                                             int* id, float* w) {
    if (SPV == 4) {
        const int4 i4 = __ldcs(reinterpret_cast<const int4*>(indices) + p);
        const float4 w4 = __ldcs(reinterpret_cast<const float4*>(weights) + p);
        id[0] = i4.x; id[1] = i4.y; id[2] = i4.z; id[3] = i4.w;
        w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
    } else {
#pragma unroll
        for (int r = 0; r < SPV; r++) {
            id[r] = __ldcs(indices + p * SPV + r);
            w[r] = __ldcs(weights + p * SPV + r);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Work decomposition shared by slice forward and the row scatter (ncu --set full on the 1M-point sweep,
// profiles/r01i_ops_ncu.md holds the current capture, is what shaped it):
//   * CHANNEL SLABS.  The vertex table is hit (D+1)x per point in random order; once nv*V*4 bytes outgrow the L2 it
//     is re-fetched from HBM several times (V=64, nv=463k: 118 MB table, 31 % L2 hit rate, 1.8x the compulsory DRAM
//     traffic).  The channels are therefore processed in slabs of `slab_ch` (a multiple of 32 floats = one 128-byte
//     line per row) sized so that one slab of the table stays L2-resident; all blocks walk the slabs in the same
//     order.  The price is re-reading the 8(D+1) index/weight bytes per point once per slab.
//   * PERSISTENT BLOCKS + REGISTER DOUBLE BUFFERING.  The old one-shot blocks (16 points each, 62 500 of them) lived
//     for two dependent memory latencies and spent the second one with only the row gathers in flight; the kernels
//     were latency bound (stall_long_scoreboard > 85 % of all stalls) at 60 % achieved occupancy.  Now a fixed grid
//     strides over the points and every thread requests the ids / weights (and, for the scatter, the source chunk)
//     of its NEXT point before it gathers the rows of the current one.
// Lanes run along the channels of one point (lpp = lanes per point, a power of two covering slab_ch / VEC chunks).
struct SlabPlan {
    int slab_ch;     // channels per slab (== val_dim when one slab suffices)
    int n_slabs;
    int lpp_log2;
    int grid;
};
static int blocks_per_sm(const void* kernel) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kBlock, 0) != cudaSuccess || nb <= 0) {
        cudaGetLastError();
        nb = 4;
    }
    return nb;
}
static int device_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}
static SlabPlan plan_slabs(int n, int nr_vertices, int val_dim, int vec, int resident_blocks) {
    constexpr long long kL2Budget = 48ll << 20;   // of the 126 MB L2 (two 63 MB halves), leaving room for the streams
    const long long rows = nr_vertices > 0 ? nr_vertices : n;
    SlabPlan pl;
    pl.slab_ch = val_dim;
    if (vec == 4 && val_dim > 32) {
        long long fit = kL2Budget / (4 * rows) / 32 * 32;
        fit = fit < 32 ? 32 : fit > 128 ? 128 : fit;       // 128 channels = 32 lanes x float4
        if (fit < val_dim) pl.slab_ch = (int)fit;
        if (val_dim > 128 && pl.slab_ch > 128) pl.slab_ch = 128;
    }
    pl.n_slabs = (val_dim + pl.slab_ch - 1) / pl.slab_ch;
    const int lpp = lanes_per_point(min(pl.slab_ch / vec, 32));
    pl.lpp_log2 = ilog2(lpp);
    const long long want = ((long long)n * lpp + kBlock - 1) / kBlock;
    const long long cap = (long long)device_sms() * resident_blocks;
    pl.grid = (int)(want < cap ? want : cap);
    return pl;
}

// SPV = simplex vertices per point (pos_dim + 1) as a compile-time constant: ids / weights stay in registers and
// the SPV row gathers of one thread are independent loads in flight together.
template <int VEC, int SPV>
__global__ void __launch_bounds__(kBlock)
slice_fwd_kernel(const float* __restrict__ lattice_values, const int* __restrict__ indices,
                 const float* __restrict__ weights, int n, int val_dim, int slab_ch, int n_slabs, int lpp_log2,
                 float* __restrict__ out) {
    LN_PDL_ENTRY();
    const int lpp = 1 << lpp_log2;
    const int g = threadIdx.x & (lpp - 1);
    const int pts_per_block = kBlock >> lpp_log2;
    const long long p_first = (long long)blockIdx.x * pts_per_block + (threadIdx.x >> lpp_log2);
    const long long p_stride = (long long)gridDim.x * pts_per_block;
    const uint64_t keep = l2_policy_evict_last();
    for (int slab = 0; slab < n_slabs; slab++) {
        const int ch_begin = slab * slab_ch;
        const int ch_end = min(val_dim, ch_begin + slab_ch);
        long long p = p_first;
        int id[SPV], idn[SPV];
        float w[SPV], wn[SPV];
        if (p < n) load_simplex<SPV>(indices, weights, p, id, w);
        while (p < n) {
            const long long pn = p + p_stride;
            if (pn < n) load_simplex<SPV>(indices, weights, pn, idn, wn);
            for (int ch = ch_begin + g * VEC; ch < ch_end; ch += lpp * VEC) {
                float4 x[SPV];
#pragma unroll
                for (int r = 0; r < SPV; r++) {
                    x[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (id[r] >= 0) {
                        const float* src = lattice_values + (size_t)id[r] * val_dim + ch;
                        if (VEC == 4)
                            x[r] = ldg_v4_hint(src, keep);
                        else
                            x[r].x = ldg_f32_hint(src, keep);
                    }
                }
                float acc[VEC];
#pragma unroll
                for (int k = 0; k < VEC; k++) acc[k] = 0.0f;
#pragma unroll
                for (int r = 0; r < SPV; r++) {
                    if (id[r] >= 0) {   // same FMA chain, in the same order, as LatticeGPU.cuh:2575-2585
                        acc[0] = fmaf(x[r].x, w[r], acc[0]);
                        if (VEC == 4) {
                            acc[1] = fmaf(x[r].y, w[r], acc[1]);
                            acc[2] = fmaf(x[r].z, w[r], acc[2]);
                            acc[3] = fmaf(x[r].w, w[r], acc[3]);
                        }
                    }
                }
                float* dst = out + (size_t)p * val_dim + ch;
                if (VEC == 4)
                    __stcs(reinterpret_cast<float4*>(dst), make_float4(acc[0], acc[1], acc[2], acc[3]));
                else
                    __stcs(dst, acc[0]);
            }
#pragma unroll
            for (int r = 0; r < SPV; r++) {
                id[r] = idn[r];
                w[r] = wn[r];
            }
            p = pn;
        }
    }
}

// rows[idx[p,r], :] += src[p, :] * w[p,r]: the backward of slice AND the value accumulation of splat
// (splatCacheNaive, LatticeGPU.cuh:926-973) are this one scatter.  Lanes run along the channels of one point, so
// every reduction instruction is a run of coalesced 16-byte vector REDs (red.global.add.v4.f32).  The L2's reduction
// units bound it (1 GB of RED payload per 1M points x 64 channels; ncu: lts throughput 60 %), the slabs keep the
// read-modify-write lines of the table from bouncing to HBM.
template <int VEC, int SPV>
__global__ void __launch_bounds__(kBlock)
scatter_rows_kernel(const float* __restrict__ src, const int* __restrict__ indices,
                    const float* __restrict__ weights, int n, int val_dim, int slab_ch, int n_slabs, int lpp_log2,
                    float* __restrict__ rows) {
    LN_PDL_ENTRY();
    const int lpp = 1 << lpp_log2;
    const int g = threadIdx.x & (lpp - 1);
    const int pts_per_block = kBlock >> lpp_log2;
    const long long p_first = (long long)blockIdx.x * pts_per_block + (threadIdx.x >> lpp_log2);
    const long long p_stride = (long long)gridDim.x * pts_per_block;
    const uint64_t keep = l2_policy_evict_last();
    for (int slab = 0; slab < n_slabs; slab++) {
        const int ch_begin = slab * slab_ch;
        const int ch_end = min(val_dim, ch_begin + slab_ch);
        const int ch0 = ch_begin + g * VEC;            // first chunk of this lane (the common case has exactly one)
        auto load_src = [&](long long pp, int ch) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* s = src + (size_t)pp * val_dim + ch;
            if (VEC == 4)
                v = __ldcs(reinterpret_cast<const float4*>(s));
            else
                v.x = __ldcs(s);
            return v;
        };
        long long p = p_first;
        int id[SPV], idn[SPV];
        float w[SPV], wn[SPV];
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f), xn = x;
        if (p < n) {
            load_simplex<SPV>(indices, weights, p, id, w);
            if (ch0 < ch_end) x = load_src(p, ch0);
        }
        while (p < n) {
            const long long pn = p + p_stride;
            if (pn < n) {
                load_simplex<SPV>(indices, weights, pn, idn, wn);
                if (ch0 < ch_end) xn = load_src(pn, ch0);
            }
            for (int ch = ch0; ch < ch_end; ch += lpp * VEC) {
                if (ch != ch0) x = load_src(p, ch);
#pragma unroll
                for (int r = 0; r < SPV; r++) {
                    if (id[r] < 0) continue;
                    float* dst = rows + (size_t)id[r] * val_dim + ch;
                    if (VEC == 4)
                        red_v4_hint(dst, make_float4(x.x * w[r], x.y * w[r], x.z * w[r], x.w * w[r]), keep);
                    else
                        red_f32_hint(dst, x.x * w[r], keep);
                }
            }
#pragma unroll
            for (int r = 0; r < SPV; r++) {
                id[r] = idn[r];
                w[r] = wn[r];
            }
            x = xn;
            p = pn;
        }
    }
}

int launch_scatter_rows(const float* src, const int* indices, const float* weights, int n, int pos_dim, int val_dim,
                        int nr_vertices, float* rows, cudaStream_t s, const char* what) {
    const int vec = (val_dim % 4 == 0) ? 4 : 1;
    const int spv = pos_dim + 1;
    if (spv != 4 && spv != 6) {
        set_error("%s: pos_dim %d not built (3 and 5 are)", what, pos_dim);
        return LN_ERR_UNSUPPORTED;
    }
#define LN_LAUNCH_SCATTER(VEC, SPV)                                                                                  \
    do {                                                                                                             \
        static const int resident = blocks_per_sm((const void*)scatter_rows_kernel<VEC, SPV>);                       \
        const SlabPlan pl = plan_slabs(n, nr_vertices, val_dim, VEC, resident);                                      \
        launch_k((scatter_rows_kernel<VEC, SPV>), dim3(pl.grid), dim3(kBlock), 0, s, src, indices, weights, n, val_dim, pl.slab_ch, pl.n_slabs, pl.lpp_log2, rows); \
    } while (0)
    if (vec == 4) {
        if (spv == 4) LN_LAUNCH_SCATTER(4, 4); else LN_LAUNCH_SCATTER(4, 6);
    } else {
        if (spv == 4) LN_LAUNCH_SCATTER(1, 4); else LN_LAUNCH_SCATTER(1, 6);
    }
#undef LN_LAUNCH_SCATTER
    count_launch();
    return check_launch(what);
}

// ---------------------------------------------------------------------------------------------
// gather: one thread per output element; row p = (D+1) chunks of [w*values[idx] (V) | w].
__global__ void __launch_bounds__(kBlock)
gather_fwd_kernel(const float* __restrict__ lattice_values, const int* __restrict__ indices,
                  const float* __restrict__ weights, int n, int spv, int val_dim, float* __restrict__ out) {
    LN_PDL_ENTRY();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = val_dim + 1;
    if (t >= (long long)n * spv * chunk) return;
    const int j = (int)(t % chunk);
    const long long pr = t / chunk;
    const int id = __ldg(indices + pr);
    float r = 0.0f;
    if (id >= 0) {
        const float w = __ldg(weights + pr);
        r = (j < val_dim) ? __ldg(lattice_values + (size_t)id * val_dim + j) * w : w;
    }
    out[t] = r;
}

__global__ void __launch_bounds__(kBlock)
gather_bwd_kernel(const float* __restrict__ grad_out, const int* __restrict__ indices,
                  const float* __restrict__ weights, int n, int spv, int val_dim, float* __restrict__ grad_values) {
    LN_PDL_ENTRY();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * spv * val_dim) return;
    const int j = (int)(t % val_dim);
    const long long pr = t / val_dim;
    const int id = __ldg(indices + pr);
    if (id < 0) return;
    const float w = __ldg(weights + pr);
    // the gradient of the trailing weight column is dropped, as in LatticeGPU.cuh:3796-3803
    atomicAdd(grad_values + (size_t)id * val_dim + j, __ldg(grad_out + pr * (val_dim + 1) + j) * w);
}

// ---------------------------------------------------------------------------------------------
// slice_classify forward: one warp per point, lanes along channels (KV = ceil(V/32) channels per
// lane), classifier weights staged cooperatively in shared memory (the reference lets thread 0
// copy them serially and accumulates logits through global read-modify-writes).
template <int KV>
__global__ void __launch_bounds__(kBlock)
slice_classify_fwd_kernel(const float* __restrict__ lattice_values, const int* __restrict__ indices,
                          const float* __restrict__ weights, const float* __restrict__ delta_weights,
                          const float* __restrict__ cls_weight, const float* __restrict__ cls_bias, int n, int spv,
                          int val_dim, int nr_classes, float* __restrict__ logits) {
    LN_PDL_ENTRY();
    extern __shared__ float smem[];
    float* w_sh = smem;   // [nr_classes][val_dim]
    for (int i = threadIdx.x; i < nr_classes * val_dim; i += blockDim.x) w_sh[i] = __ldg(cls_weight + i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (long long p = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); p < n;
         p += (long long)gridDim.x * warps_per_block) {
        float s[KV];
#pragma unroll
        for (int k = 0; k < KV; k++) s[k] = 0.0f;
        for (int r = 0; r < spv; r++) {
            const int id = __ldg(indices + p * spv + r);
            if (id < 0) continue;
            const float w = __ldg(weights + p * spv + r) + __ldg(delta_weights + p * spv + r);
#pragma unroll
            for (int k = 0; k < KV; k++) {
                const int v = lane + 32 * k;
                if (v < val_dim) s[k] = fmaf(__ldg(lattice_values + (size_t)id * val_dim + v), w, s[k]);
            }
        }
        float mine = 0.0f;   // lane c%32 keeps logit c (+32: second round)
        for (int c = 0; c < nr_classes; c++) {
            float part = 0.0f;
#pragma unroll
            for (int k = 0; k < KV; k++) {
                const int v = lane + 32 * k;
                if (v < val_dim) part = fmaf(w_sh[c * val_dim + v], s[k], part);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if ((c & 31) == lane) mine = part + __ldg(cls_bias + c);
            if ((c & 31) == 31 || c == nr_classes - 1) {
                const int cc = (c & ~31) + lane;
                if (cc <= c) logits[p * nr_classes + cc] = mine;
            }
        }
    }
}

// Tiled forward (val_dim % 4 == 0).  ncu on the 10^6-point sweep showed the warp-per-point kernel above issue-bound
// (74 % issue slots, one 5-step butterfly reduction per class and point).  Here a CTA walks tiles of kScfPoints points:
//   phase A  thread = (point, float4 column): the SPV ids / weights are read first, then all SPV row loads are issued
//            together and folded into s_p = sum_r (w_r + dw_r) values[idx_r]; the tile of s rows is parked in smem;
//   phase B  thread = (point, group of 8 classes): logits[p, c] = <s_p, W_c> + b_c as a register-tiled mini GEMM --
//            one float4 of s_p and one broadcast float4 of W_c per four FMAs, no cross-lane reduction at all.
// Row stride of the s tile is V + 4 floats: the float4 reads of 8 consecutive points fall into 32 distinct banks.
constexpr int kScfPoints = 64;
template <int SPV>
__global__ void __launch_bounds__(kBlock)
slice_classify_fwd_tiled_kernel(const float* __restrict__ lattice_values, const int* __restrict__ indices,
                                const float* __restrict__ weights, const float* __restrict__ delta_weights,
                                const float* __restrict__ cls_weight, const float* __restrict__ cls_bias, int n,
                                int val_dim, int nr_classes, float* __restrict__ logits) {
    LN_PDL_ENTRY();
    extern __shared__ __align__(16) float smem_t[];
    const int lds = val_dim + 4;
    float* w_sh = smem_t;                            // [nc][V]
    float* s_sh = w_sh + nr_classes * val_dim;       // [kScfPoints][V + 4]
    for (int i = threadIdx.x; i < nr_classes * val_dim; i += blockDim.x) w_sh[i] = __ldg(cls_weight + i);
    const int cv = val_dim >> 2;
    const int n_tiles = (n + kScfPoints - 1) / kScfPoints;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();                             // w_sh ready / phase B of the previous tile done
        const long long p0 = (long long)tile * kScfPoints;
        for (int e = threadIdx.x; e < kScfPoints * cv; e += blockDim.x) {
            const int pl = e / cv, j = e - pl * cv;
            const long long p = p0 + pl;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < n) {
                int id[SPV];
                float w[SPV];
#pragma unroll
                for (int r = 0; r < SPV; r++) {
                    id[r] = __ldg(indices + p * SPV + r);
                    w[r] = __ldg(weights + p * SPV + r) + __ldg(delta_weights + p * SPV + r);
                }
                float4 x[SPV];
#pragma unroll
                for (int r = 0; r < SPV; r++)
                    x[r] = id[r] >= 0 ? __ldg(reinterpret_cast<const float4*>(lattice_values + (size_t)id[r] * val_dim) + j)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int r = 0; r < SPV; r++) {
                    acc.x = fmaf(x[r].x, w[r], acc.x);
                    acc.y = fmaf(x[r].y, w[r], acc.y);
                    acc.z = fmaf(x[r].z, w[r], acc.z);
                    acc.w = fmaf(x[r].w, w[r], acc.w);
                }
            }
            *reinterpret_cast<float4*>(s_sh + pl * lds + 4 * j) = acc;
        }
        __syncthreads();
        const int pl = threadIdx.x & (kScfPoints - 1);
        const long long p = p0 + pl;
        const float* srow = s_sh + pl * lds;
        for (int c0 = (threadIdx.x / kScfPoints) * 8; c0 < nr_classes; c0 += (kBlock / kScfPoints) * 8) {
            float acc[8];
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = 0.0f;
            for (int v = 0; v < val_dim; v += 4) {
                const float4 s4 = *reinterpret_cast<const float4*>(srow + v);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (c0 + k < nr_classes) {       // uniform across the warp: no divergence
                        const float4 w4 = *reinterpret_cast<const float4*>(w_sh + (c0 + k) * val_dim + v);
                        acc[k] = fmaf(s4.x, w4.x, acc[k]);
                        acc[k] = fmaf(s4.y, w4.y, acc[k]);
                        acc[k] = fmaf(s4.z, w4.z, acc[k]);
                        acc[k] = fmaf(s4.w, w4.w, acc[k]);
                    }
                }
            }
            if (p < n) {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (c0 + k < nr_classes) logits[p * nr_classes + c0 + k] = acc[k] + __ldg(cls_bias + c0 + k);
            }
        }
    }
}

// slice_classify backward.  Persistent blocks walk tiles of kTile points:
//   phase A (warp per point): s_p = sum_r (w+dw) values[idx_r], t_p = g_p * W, scatter
//            (w+dw)*t_p into the lattice gradient, grad_dw[p,r] = <values[idx_r], t_p>;
//            s_p and g_p are parked in shared memory;
//   phase B (all threads): grad_W += G^T S for the tile as a register-accumulated mini GEMM.
// grad_W / grad_b reach global memory once per block (the reference issues N*nc*V global atomics
// onto nc*V addresses, LatticeGPU.cuh:3716-3722).
constexpr int kTile = 32;
template <int KV>
__global__ void __launch_bounds__(kBlock)
slice_classify_bwd_kernel(const float* __restrict__ grad_logits, const float* __restrict__ lattice_values,
                          const int* __restrict__ indices, const float* __restrict__ weights,
                          const float* __restrict__ delta_weights, const float* __restrict__ cls_weight, int n,
                          int spv, int val_dim, int nr_classes, float* __restrict__ grad_lattice_values,
                          float* __restrict__ grad_delta_weights, float* __restrict__ grad_cls_weight,
                          float* __restrict__ grad_cls_bias) {
    LN_PDL_ENTRY();
    extern __shared__ float smem[];
    float* w_sh = smem;                                 // [nc][V]
    float* s_sh = w_sh + nr_classes * val_dim;          // [kTile][V]
    float* g_sh = s_sh + kTile * val_dim;               // [kTile][nc]
    for (int i = threadIdx.x; i < nr_classes * val_dim; i += blockDim.x) w_sh[i] = __ldg(cls_weight + i);

    constexpr int kMaxAcc = 24;                         // nc*V <= kMaxAcc*kBlock  (checked on the host)
    float gw_acc[kMaxAcc];
#pragma unroll
    for (int i = 0; i < kMaxAcc; i++) gw_acc[i] = 0.0f;
    float gb_acc = 0.0f;
    const int n_out = nr_classes * val_dim;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int n_tiles = (n + kTile - 1) / kTile;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();   // w_sh ready / previous tile's phase B done
        const int p0 = tile * kTile;
        const int tile_n = min(kTile, n - p0);
        for (int pl = warp; pl < kTile; pl += warps_per_block) {
            float* s_row = s_sh + pl * val_dim;
            float* g_row = g_sh + pl * nr_classes;
            if (pl >= tile_n) {   // pad the tile with zeros so phase B needs no bounds
                for (int v = lane; v < val_dim; v += 32) s_row[v] = 0.0f;
                for (int c = lane; c < nr_classes; c += 32) g_row[c] = 0.0f;
                continue;
            }
            const long long p = p0 + pl;
            for (int c = lane; c < nr_classes; c += 32) g_row[c] = __ldg(grad_logits + p * nr_classes + c);
            __syncwarp();
            float t[KV], s[KV];
#pragma unroll
            for (int k = 0; k < KV; k++) {
                t[k] = 0.0f;
                s[k] = 0.0f;
            }
            for (int c = 0; c < nr_classes; c++) {
                const float g = g_row[c];
#pragma unroll
                for (int k = 0; k < KV; k++) {
                    const int v = lane + 32 * k;
                    if (v < val_dim) t[k] = fmaf(g, w_sh[c * val_dim + v], t[k]);
                }
            }
            for (int r = 0; r < spv; r++) {
                const int id = __ldg(indices + p * spv + r);
                float dot = 0.0f;
                if (id >= 0) {
                    const float w = __ldg(weights + p * spv + r) + __ldg(delta_weights + p * spv + r);
#pragma unroll
                    for (int k = 0; k < KV; k++) {
                        const int v = lane + 32 * k;
                        if (v < val_dim) {
                            const float x = __ldg(lattice_values + (size_t)id * val_dim + v);
                            s[k] = fmaf(x, w, s[k]);
                            dot = fmaf(x, t[k], dot);
                            atomicAdd(grad_lattice_values + (size_t)id * val_dim + v, t[k] * w);
                        }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                if (lane == 0) grad_delta_weights[p * spv + r] += dot;   // (p,r) is owned by this warp
            }
#pragma unroll
            for (int k = 0; k < KV; k++) {
                const int v = lane + 32 * k;
                if (v < val_dim) s_row[v] = s[k];
            }
        }
        __syncthreads();
        // phase B: grad_W[c][v] += sum_p G[p][c] * S[p][v]
#pragma unroll
        for (int i = 0; i < kMaxAcc; i++) {
            const int o = threadIdx.x + i * kBlock;
            if (o < n_out) {
                const int c = o / val_dim, v = o - c * val_dim;
                float acc = gw_acc[i];
#pragma unroll 8
                for (int pl = 0; pl < kTile; pl++) acc = fmaf(g_sh[pl * nr_classes + c], s_sh[pl * val_dim + v], acc);
                gw_acc[i] = acc;
            }
        }
        if (threadIdx.x < nr_classes) {
            for (int pl = 0; pl < kTile; pl++) gb_acc += g_sh[pl * nr_classes + threadIdx.x];
        }
    }
#pragma unroll
    for (int i = 0; i < kMaxAcc; i++) {
        const int o = threadIdx.x + i * kBlock;
        if (o < n_out) atomicAdd(grad_cls_weight + o, gw_acc[i]);
    }
    if (threadIdx.x < nr_classes) atomicAdd(grad_cls_bias + threadIdx.x, gb_acc);
}

// Vectorised backward (val_dim % 4 == 0, SPV known at compile time).  The kernel above is latency-bound (ncu, 10^6
// points: 25 % occupancy, 5.5 % L2 throughput, one dependent index -> row -> atomic chain per simplex vertex, 4-byte
// reductions).  Same tiling and phase B, but in phase A a lane owns float4 column(s) of the row: the SPV ids /
// weights are fetched by SPV lanes and broadcast, ALL row loads of the simplex are issued before anything consumes
// them (t_p = g_p W is computed while they are in flight), and the lattice gradient receives 16-byte reductions
// (red.global.add.v4.f32, evict_last like the row scatter) -- a quarter of the L2 atomic operations.
template <int KV4, int SPV>   // KV4 float4 columns per lane: val_dim <= 128 * KV4
__global__ void __launch_bounds__(kBlock)
slice_classify_bwd_vec_kernel(const float* __restrict__ grad_logits, const float* __restrict__ lattice_values,
                              const int* __restrict__ indices, const float* __restrict__ weights,
                              const float* __restrict__ delta_weights, const float* __restrict__ cls_weight, int n,
                              int val_dim, int nr_classes, float* __restrict__ grad_lattice_values,
                              float* __restrict__ grad_delta_weights, float* __restrict__ grad_cls_weight,
                              float* __restrict__ grad_cls_bias) {
    LN_PDL_ENTRY();
    extern __shared__ __align__(16) float smem_v[];
    float* w_sh = smem_v;                               // [nc][V]
    float* s_sh = w_sh + nr_classes * val_dim;          // [kTile][V]
    float* g_sh = s_sh + kTile * val_dim;               // [kTile][nc]
    for (int i = threadIdx.x; i < nr_classes * val_dim; i += blockDim.x) w_sh[i] = __ldg(cls_weight + i);

    constexpr int kMaxAcc = 24;                         // nc*V <= kMaxAcc*kBlock  (checked on the host)
    float gw_acc[kMaxAcc];
#pragma unroll
    for (int i = 0; i < kMaxAcc; i++) gw_acc[i] = 0.0f;
    float gb_acc = 0.0f;
    const int n_out = nr_classes * val_dim;
    const int cv = val_dim >> 2;
    const uint64_t keep = l2_policy_evict_last();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int n_tiles = (n + kTile - 1) / kTile;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();   // w_sh ready / previous tile's phase B done
        const int p0 = tile * kTile;
        const int tile_n = min(kTile, n - p0);
        for (int pl = warp; pl < kTile; pl += warps_per_block) {
            float* s_row = s_sh + pl * val_dim;
            float* g_row = g_sh + pl * nr_classes;
            if (pl >= tile_n) {   // pad the tile with zeros so phase B needs no bounds
                for (int v = lane; v < val_dim; v += 32) s_row[v] = 0.0f;
                for (int c = lane; c < nr_classes; c += 32) g_row[c] = 0.0f;
                continue;
            }
            const long long p = p0 + pl;
            int my_id = -1;
            float my_w = 0.0f;
            if (lane < SPV) {
                my_id = __ldg(indices + p * SPV + lane);
                my_w = __ldg(weights + p * SPV + lane) + __ldg(delta_weights + p * SPV + lane);
            }
            for (int c = lane; c < nr_classes; c += 32) g_row[c] = __ldg(grad_logits + p * nr_classes + c);
            int id[SPV];
            float w[SPV];
#pragma unroll
            for (int r = 0; r < SPV; r++) {
                id[r] = __shfl_sync(0xffffffffu, my_id, r);
                w[r] = __shfl_sync(0xffffffffu, my_w, r);
            }
            float4 x[SPV][KV4];
#pragma unroll
            for (int r = 0; r < SPV; r++)
#pragma unroll
                for (int k = 0; k < KV4; k++) {
                    const int j = lane + 32 * k;
                    x[r][k] = (id[r] >= 0 && j < cv) ? __ldg(reinterpret_cast<const float4*>(lattice_values + (size_t)id[r] * val_dim) + j)
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            __syncwarp();   // g_row written by all lanes
            float4 t[KV4], sacc[KV4];
#pragma unroll
            for (int k = 0; k < KV4; k++) {
                t[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                sacc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            for (int c = 0; c < nr_classes; c++) {
                const float g = g_row[c];
#pragma unroll
                for (int k = 0; k < KV4; k++) {
                    const int j = lane + 32 * k;
                    if (j < cv) {
                        const float4 w4 = *reinterpret_cast<const float4*>(w_sh + c * val_dim + 4 * j);
                        t[k].x = fmaf(g, w4.x, t[k].x);
                        t[k].y = fmaf(g, w4.y, t[k].y);
                        t[k].z = fmaf(g, w4.z, t[k].z);
                        t[k].w = fmaf(g, w4.w, t[k].w);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < SPV; r++) {
                float dot = 0.0f;
#pragma unroll
                for (int k = 0; k < KV4; k++) {
                    const int j = lane + 32 * k;
                    const float4 xv = x[r][k];
                    sacc[k].x = fmaf(xv.x, w[r], sacc[k].x);
                    sacc[k].y = fmaf(xv.y, w[r], sacc[k].y);
                    sacc[k].z = fmaf(xv.z, w[r], sacc[k].z);
                    sacc[k].w = fmaf(xv.w, w[r], sacc[k].w);
                    dot = fmaf(xv.x, t[k].x, dot);
                    dot = fmaf(xv.y, t[k].y, dot);
                    dot = fmaf(xv.z, t[k].z, dot);
                    dot = fmaf(xv.w, t[k].w, dot);
                    if (id[r] >= 0 && j < cv)
                        red_v4_hint(grad_lattice_values + (size_t)id[r] * val_dim + 4 * j,
                                    make_float4(t[k].x * w[r], t[k].y * w[r], t[k].z * w[r], t[k].w * w[r]), keep);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                if (lane == 0) grad_delta_weights[p * SPV + r] += dot;   // (p,r) is owned by this warp
            }
#pragma unroll
            for (int k = 0; k < KV4; k++) {
                const int j = lane + 32 * k;
                if (j < cv) *reinterpret_cast<float4*>(s_row + 4 * j) = sacc[k];
            }
        }
        __syncthreads();
        // phase B: grad_W[c][v] += sum_p G[p][c] * S[p][v]
#pragma unroll
        for (int i = 0; i < kMaxAcc; i++) {
            const int o = threadIdx.x + i * kBlock;
            if (o < n_out) {
                const int c = o / val_dim, v = o - c * val_dim;
                float acc = gw_acc[i];
#pragma unroll 8
                for (int pl = 0; pl < kTile; pl++) acc = fmaf(g_sh[pl * nr_classes + c], s_sh[pl * val_dim + v], acc);
                gw_acc[i] = acc;
            }
        }
        if (threadIdx.x < nr_classes) {
            for (int pl = 0; pl < kTile; pl++) gb_acc += g_sh[pl * nr_classes + threadIdx.x];
        }
    }
#pragma unroll
    for (int i = 0; i < kMaxAcc; i++) {
        const int o = threadIdx.x + i * kBlock;
        if (o < n_out) atomicAdd(grad_cls_weight + o, gw_acc[i]);
    }
    if (threadIdx.x < nr_classes) atomicAdd(grad_cls_bias + threadIdx.x, gb_acc);
}

// ---------------------------------------------------------------------------------------------
// PointNet glue: segmented max (+argmax) and sum/count over the rows that share a vertex.
__device__ __forceinline__ unsigned int float_to_ordered(float f) {
    const unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned int o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void __launch_bounds__(kBlock)
scatter_max_pack_kernel(const float* __restrict__ src, const int* __restrict__ index, int m, int c,
                        unsigned long long* __restrict__ packed) {
    LN_PDL_ENTRY();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)m * c) return;
    const int row = (int)(t / c);
    const int col = (int)(t - (long long)row * c);
    const int v = __ldg(index + row);
    // max over (value, -row): the largest value wins, ties go to the smallest row (deterministic)
    const unsigned long long key =
        ((unsigned long long)float_to_ordered(__ldg(src + t)) << 32) | (unsigned long long)(0xffffffffu - (unsigned)row);
    atomicMax(packed + (size_t)v * c + col, key);
}

__global__ void __launch_bounds__(kBlock)
scatter_max_unpack_kernel(const unsigned long long* __restrict__ packed, long long total, int m,
                          float* __restrict__ out_max, int* __restrict__ out_arg) {
    LN_PDL_ENTRY();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const unsigned long long key = packed[t];
    if (key == 0ull) {   // no row maps to this vertex: torch_scatter yields 0 and arg == m
        out_max[t] = 0.0f;
        out_arg[t] = m;
    } else {
        out_max[t] = ordered_to_float((unsigned int)(key >> 32));
        out_arg[t] = (int)(0xffffffffu - (unsigned int)(key & 0xffffffffu));
    }
}

__global__ void __launch_bounds__(kBlock)
scatter_sum_count_kernel(const float* __restrict__ src, const int* __restrict__ index, int m, int c,
                         float* __restrict__ out_sum, float* __restrict__ out_count) {
    LN_PDL_ENTRY();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)m * c) return;
    const int row = (int)(t / c);
    const int col = (int)(t - (long long)row * c);
    const int v = __ldg(index + row);
    atomicAdd(out_sum + (size_t)v * c + col, __ldg(src + t));
    if (col == 0 && out_count != nullptr) atomicAdd(out_count + v, 1.0f);
}


// ---------------------------------------------------------------------------------------------
// Learned barycentric offsets of the DeformSlice head (lattice_modules.py:465-567), one kernel each way instead of
// gather -> view -> max -> affine -> subtract -> Linear(9 -> 1) -> reshape (and their ~20 autograd kernels):
//   g[r] = [ w_r * values[idx_r, 0:8] | w_r ]   (zeros where idx_r < 0: gather_with_precomputation, LatticeGPU.cuh:2886-2929)
//   m = max_r g[r];   t[r] = g[r] - (gamma * m + beta);   delta_w[p, r] = <W, t[r]> + bias
constexpr int kDwC = 8;            // bottleneck width of the slice head
constexpr int kDwF = kDwC + 1;     // features per simplex vertex

template <int SPV>
__device__ __forceinline__ void deltaw_features(const float* __restrict__ values, const int* __restrict__ indices,
                                                const float* __restrict__ weights, long long p, int* id, float* w, float (*g)[kDwF]) {
    load_simplex<SPV>(indices, weights, p, id, w);
#pragma unroll
    for (int r = 0; r < SPV; r++) {
        if (id[r] >= 0) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(values + (size_t)id[r] * kDwC));
            const float4 b = __ldg(reinterpret_cast<const float4*>(values + (size_t)id[r] * kDwC) + 1);
            g[r][0] = a.x * w[r]; g[r][1] = a.y * w[r]; g[r][2] = a.z * w[r]; g[r][3] = a.w * w[r];
            g[r][4] = b.x * w[r]; g[r][5] = b.y * w[r]; g[r][6] = b.z * w[r]; g[r][7] = b.w * w[r];
            g[r][8] = w[r];
        } else {
#pragma unroll
            for (int j = 0; j < kDwF; j++) g[r][j] = 0.0f;
        }
    }
}

template <int SPV>
__global__ void __launch_bounds__(kBlock)
deltaw_fwd_kernel(const float* __restrict__ values, const int* __restrict__ indices, const float* __restrict__ weights,
                  const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ lin_w,
                  const float* __restrict__ lin_b, int n, float* __restrict__ delta_w) {
    LN_PDL_ENTRY();
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int id[SPV];
    float w[SPV], g[SPV][kDwF];
    deltaw_features<SPV>(values, indices, weights, p, id, w, g);
    float out[SPV];
    const float b0 = __ldg(lin_b);
#pragma unroll
    for (int r = 0; r < SPV; r++) out[r] = b0;
#pragma unroll
    for (int j = 0; j < kDwF; j++) {
        float m = g[0][j];
#pragma unroll
        for (int r = 1; r < SPV; r++) m = fmaxf(m, g[r][j]);
        const float shift = __ldg(gamma + j) * m + __ldg(beta + j);
        const float wj = __ldg(lin_w + j);
#pragma unroll
        for (int r = 0; r < SPV; r++) out[r] = fmaf(wj, g[r][j] - shift, out[r]);
    }
#pragma unroll
    for (int r = 0; r < SPV; r++) delta_w[p * SPV + r] = out[r];
}

// the four small gradients (zeroed by the caller) are accumulated with one atomic per CTA and element
template <int SPV>
__global__ void __launch_bounds__(kBlock)
deltaw_bwd_kernel(const float* __restrict__ values, const int* __restrict__ indices, const float* __restrict__ weights,
                  const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ lin_w,
                  const float* __restrict__ grad_delta_w, int n, float* __restrict__ grad_values, float* __restrict__ g_lin_w,
                  float* __restrict__ g_lin_b, float* __restrict__ g_gamma, float* __restrict__ g_beta) {
    LN_PDL_ENTRY();
    __shared__ float red[3 * kDwF + 1][kBlock / 32];
    float acc[3 * kDwF + 1];
#pragma unroll
    for (int k = 0; k < 3 * kDwF + 1; k++) acc[k] = 0.0f;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        int id[SPV];
        float w[SPV], g[SPV][kDwF], d[SPV];
        deltaw_features<SPV>(values, indices, weights, p, id, w, g);
        float dsum = 0.0f;
#pragma unroll
        for (int r = 0; r < SPV; r++) {
            d[r] = __ldg(grad_delta_w + p * SPV + r);
            dsum += d[r];
        }
        acc[kDwF] += dsum;                                   // d bias
        float dv[SPV][kDwC];
#pragma unroll
        for (int j = 0; j < kDwF; j++) {
            float m = g[0][j];
            int arg = 0;
#pragma unroll
            for (int r = 1; r < SPV; r++)
                if (g[r][j] > m) {                           // first maximum, as torch.max(dim) reports it
                    m = g[r][j];
                    arg = r;
                }
            const float gm = __ldg(gamma + j), wj = __ldg(lin_w + j);
            const float shift = gm * m + __ldg(beta + j);
            float dwj = 0.0f;
#pragma unroll
            for (int r = 0; r < SPV; r++) dwj = fmaf(d[r], g[r][j] - shift, dwj);
            acc[j] += dwj;                                   // d lin_w[j]
            const float dt_sum = wj * dsum;                  // sum_r d t[r][j]
            acc[kDwF + 1 + j] -= m * dt_sum;                 // d gamma[j]
            acc[2 * kDwF + 1 + j] -= dt_sum;                 // d beta[j]
            if (j < kDwC) {
#pragma unroll
                for (int r = 0; r < SPV; r++) dv[r][j] = (d[r] * wj - (r == arg ? gm * dt_sum : 0.0f)) * w[r];
            }
        }
#pragma unroll
        for (int r = 0; r < SPV; r++) {
            if (id[r] < 0) continue;
            float* dst = grad_values + (size_t)id[r] * kDwC;
            atomicAdd(reinterpret_cast<float4*>(dst), make_float4(dv[r][0], dv[r][1], dv[r][2], dv[r][3]));
            atomicAdd(reinterpret_cast<float4*>(dst) + 1, make_float4(dv[r][4], dv[r][5], dv[r][6], dv[r][7]));
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3 * kDwF + 1; k++) {
        float v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3 * kDwF + 1) {
        float t = 0.0f;
#pragma unroll
        for (int wv = 0; wv < kBlock / 32; wv++) t += red[threadIdx.x][wv];
        const int k = threadIdx.x;       // acc layout: [0:9) d lin_w, [9] d lin_b, [10:19) d gamma, [19:28) d beta
        float* dst = k < kDwF ? g_lin_w + k : k == kDwF ? g_lin_b : k < 2 * kDwF + 1 ? g_gamma + (k - kDwF - 1) : g_beta + (k - 2 * kDwF - 1);
        atomicAdd(dst, t);
    }
}

}  // namespace ln

using namespace ln;

#define LN_SLICE_ARGS_OK(name)                                                                              \
    LN_REQUIRE(n >= 0 && pos_dim >= 1 && pos_dim + 1 <= kMaxSpv && val_dim >= 1, name ": bad size n=%d pos_dim=%d val_dim=%d", n, pos_dim, val_dim)

extern "C" {

int ln_slice_fwd(const float* lattice_values, const int* indices, const float* weights, int n, int pos_dim,
                 int val_dim, int nr_vertices, float* out, void* stream) {
    LN_REQUIRE(lattice_values && indices && weights && out, "ln_slice_fwd: null pointer");
    LN_SLICE_ARGS_OK("ln_slice_fwd");
    if (n == 0) return LN_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int vec = (val_dim % 4 == 0) ? 4 : 1;
#define LN_LAUNCH_SLICE(VEC, SPV)                                                                                    \
    do {                                                                                                             \
        static const int resident = blocks_per_sm((const void*)slice_fwd_kernel<VEC, SPV>);                          \
        const SlabPlan pl = plan_slabs(n, nr_vertices, val_dim, VEC, resident);                                      \
        launch_k((slice_fwd_kernel<VEC, SPV>), dim3(pl.grid), dim3(kBlock), 0, s, lattice_values, indices, weights, n, val_dim, pl.slab_ch, pl.n_slabs, pl.lpp_log2, out); \
    } while (0)
    const int spv = pos_dim + 1;
    if (spv != 4 && spv != 6) {
        set_error("ln_slice_fwd: pos_dim %d not built (3 and 5 are)", pos_dim);
        return LN_ERR_UNSUPPORTED;
    }
    if (vec == 4) {
        if (spv == 4) LN_LAUNCH_SLICE(4, 4); else LN_LAUNCH_SLICE(4, 6);
    } else {
        if (spv == 4) LN_LAUNCH_SLICE(1, 4); else LN_LAUNCH_SLICE(1, 6);
    }
#undef LN_LAUNCH_SLICE
    count_launch();
    return check_launch("slice_fwd");
}

int ln_slice_bwd(const float* grad_out, const int* indices, const float* weights, int n, int pos_dim, int val_dim,
                 int nr_vertices, float* grad_values, void* stream) {
    LN_REQUIRE(grad_out && indices && weights && grad_values, "ln_slice_bwd: null pointer");
    LN_SLICE_ARGS_OK("ln_slice_bwd");
    if (n == 0) return LN_OK;
    return launch_scatter_rows(grad_out, indices, weights, n, pos_dim, val_dim, nr_vertices, grad_values, (cudaStream_t)stream, "slice_bwd");
}

int ln_gather_fwd(const float* lattice_values, const int* indices, const float* weights, int n, int pos_dim,
                  int val_dim, float* out, void* stream) {
    LN_REQUIRE(lattice_values && indices && weights && out, "ln_gather_fwd: null pointer");
    LN_SLICE_ARGS_OK("ln_gather_fwd");
    if (n == 0) return LN_OK;
    const long long total = (long long)n * (pos_dim + 1) * (val_dim + 1);
    launch_k(gather_fwd_kernel, dim3(cdiv(total, kBlock)), dim3(kBlock), 0, (cudaStream_t)stream, lattice_values, indices, weights, n, pos_dim + 1, val_dim, out);
    count_launch();
    return check_launch("gather_fwd");
}

int ln_gather_bwd(const float* grad_out, const int* indices, const float* weights, int n, int pos_dim, int val_dim,
                  float* grad_values, void* stream) {
    LN_REQUIRE(grad_out && indices && weights && grad_values, "ln_gather_bwd: null pointer");
    LN_SLICE_ARGS_OK("ln_gather_bwd");
    if (n == 0) return LN_OK;
    const long long total = (long long)n * (pos_dim + 1) * val_dim;
    launch_k(gather_bwd_kernel, dim3(cdiv(total, kBlock)), dim3(kBlock), 0, (cudaStream_t)stream, grad_out, indices, weights, n, pos_dim + 1, val_dim, grad_values);
    count_launch();
    return check_launch("gather_bwd");
}

int ln_slice_classify_fwd(const float* lattice_values, const int* indices, const float* weights,
                          const float* delta_weights, const float* cls_weight, const float* cls_bias, int n,
                          int pos_dim, int val_dim, int nr_classes, float* logits, void* stream) {
    LN_REQUIRE(lattice_values && indices && weights && delta_weights && cls_weight && cls_bias && logits,
               "ln_slice_classify_fwd: null pointer");
    LN_SLICE_ARGS_OK("ln_slice_classify_fwd");
    LN_REQUIRE(nr_classes >= 1, "ln_slice_classify_fwd: nr_classes must be positive");
    if (n == 0) return LN_OK;
    if (val_dim > 256) {
        set_error("ln_slice_classify_fwd: val_dim %d > 256 not built", val_dim);
        return LN_ERR_UNSUPPORTED;
    }
    cudaStream_t s = (cudaStream_t)stream;
    {
        const size_t smem_t = ((size_t)nr_classes * val_dim + (size_t)kScfPoints * (val_dim + 4)) * sizeof(float);
        if (val_dim % 4 == 0 && (pos_dim == 3 || pos_dim == 5) && smem_t <= 200 * 1024) {
            // as many CTAs as stay resident (ncu r01s: with 2 per SM the kernel sat at 25 % occupancy, latency-bound);
            // the tile loop is grid-stride, so the grid size only affects scheduling
            const void* kern = pos_dim == 3 ? (const void*)slice_classify_fwd_tiled_kernel<4> : (const void*)slice_classify_fwd_tiled_kernel<6>;
            cudaError_t err = smem_t > 48 * 1024 ? allow_max_smem(kern) : cudaSuccess;
            static int cached_resident[2] = {0, 0};          // per kernel variant; queried once per shared-memory size
            static size_t cached_smem[2] = {0, 0};
            const int slot = pos_dim == 3 ? 0 : 1;
            int resident = cached_smem[slot] == smem_t ? cached_resident[slot] : 0;
            if (resident < 1) {
                if (err != cudaSuccess || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, kBlock, smem_t) != cudaSuccess || resident < 1) {
                    cudaGetLastError();
                    resident = 2;
                }
                cached_resident[slot] = resident;
                cached_smem[slot] = smem_t;
            }
            const int grid_t = min(cdiv(n, kScfPoints), device_sms() * min(resident, 8));
            if (pos_dim == 3) {
                if (smem_t > 48 * 1024) err = allow_max_smem((const void*)slice_classify_fwd_tiled_kernel<4>);
                if (err == cudaSuccess)
                    launch_k(slice_classify_fwd_tiled_kernel<4>, dim3(grid_t), dim3(kBlock), smem_t, s, lattice_values, indices, weights, delta_weights, cls_weight, cls_bias, n, val_dim, nr_classes, logits);
            } else {
                if (smem_t > 48 * 1024) err = allow_max_smem((const void*)slice_classify_fwd_tiled_kernel<6>);
                if (err == cudaSuccess)
                    launch_k(slice_classify_fwd_tiled_kernel<6>, dim3(grid_t), dim3(kBlock), smem_t, s, lattice_values, indices, weights, delta_weights, cls_weight, cls_bias, n, val_dim, nr_classes, logits);
            }
            if (err != cudaSuccess) {
                set_error("ln_slice_classify_fwd: %s", cudaGetErrorString(err));
                return LN_ERR_CUDA;
            }
            count_launch();
            return check_launch("slice_classify_fwd_tiled");
        }
    }
    const size_t smem = (size_t)nr_classes * val_dim * sizeof(float);
    const int grid = min(cdiv(n, kBlock / 32), 148 * 8);
    const int kv = cdiv(val_dim, 32);
#define LN_LAUNCH_SCF(KV)                                                                                          \
    do {                                                                                                           \
        if (smem > 48 * 1024) allow_max_smem((const void*)slice_classify_fwd_kernel<KV>); \
        launch_k(slice_classify_fwd_kernel<KV>, dim3(grid), dim3(kBlock), smem, s, lattice_values, indices, weights, delta_weights, cls_weight, cls_bias, n, pos_dim + 1, val_dim, nr_classes, logits); \
    } while (0)
    if (kv <= 1) LN_LAUNCH_SCF(1);
    else if (kv <= 2) LN_LAUNCH_SCF(2);
    else if (kv <= 4) LN_LAUNCH_SCF(4);
    else LN_LAUNCH_SCF(8);
#undef LN_LAUNCH_SCF
    count_launch();
    return check_launch("slice_classify_fwd");
}

int ln_slice_classify_bwd(const float* grad_logits, const float* lattice_values, const int* indices,
                          const float* weights, const float* delta_weights, const float* cls_weight, int n,
                          int pos_dim, int val_dim, int nr_classes, float* grad_lattice_values,
                          float* grad_delta_weights, float* grad_cls_weight, float* grad_cls_bias, void* stream) {
    LN_REQUIRE(grad_logits && lattice_values && indices && weights && delta_weights && cls_weight && grad_lattice_values &&
                   grad_delta_weights && grad_cls_weight && grad_cls_bias,
               "ln_slice_classify_bwd: null pointer");
    LN_SLICE_ARGS_OK("ln_slice_classify_bwd");
    LN_REQUIRE(nr_classes >= 1, "ln_slice_classify_bwd: nr_classes must be positive");
    if (n == 0) return LN_OK;
    if (val_dim > 256 || (long long)nr_classes * val_dim > 24LL * kBlock || nr_classes > kBlock) {
        set_error("ln_slice_classify_bwd: val_dim %d x nr_classes %d outside the built range (val_dim<=256, product<=6144)", val_dim, nr_classes);
        return LN_ERR_UNSUPPORTED;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = ((size_t)nr_classes * val_dim + (size_t)kTile * val_dim + (size_t)kTile * nr_classes) * sizeof(float);
    const int grid = min(cdiv(n, kTile), 148 * 2);
    if (val_dim % 4 == 0 && (pos_dim == 3 || pos_dim == 5)) {
#define LN_LAUNCH_SCBV(KV4, SPV)                                                                                   \
    do {                                                                                                           \
        cudaError_t err = smem > 48 * 1024 ? allow_max_smem((const void*)slice_classify_bwd_vec_kernel<KV4, SPV>) : cudaSuccess; \
        if (err != cudaSuccess) {                                                                                  \
            set_error("ln_slice_classify_bwd: %s", cudaGetErrorString(err));                                       \
            return LN_ERR_CUDA;                                                                                    \
        }                                                                                                          \
        launch_k((slice_classify_bwd_vec_kernel<KV4, SPV>), dim3(grid), dim3(kBlock), smem, s, grad_logits, lattice_values, indices, weights, delta_weights, cls_weight, n, val_dim, nr_classes, grad_lattice_values, grad_delta_weights, grad_cls_weight, grad_cls_bias); \
    } while (0)
        if (val_dim <= 128) {
            if (pos_dim == 3) LN_LAUNCH_SCBV(1, 4); else LN_LAUNCH_SCBV(1, 6);
        } else {
            if (pos_dim == 3) LN_LAUNCH_SCBV(2, 4); else LN_LAUNCH_SCBV(2, 6);
        }
#undef LN_LAUNCH_SCBV
        count_launch();
        return check_launch("slice_classify_bwd_vec");
    }
    const int kv = cdiv(val_dim, 32);
#define LN_LAUNCH_SCB(KV)                                                                                          \
    do {                                                                                                           \
        if (smem > 48 * 1024) allow_max_smem((const void*)slice_classify_bwd_kernel<KV>); \
        launch_k(slice_classify_bwd_kernel<KV>, dim3(grid), dim3(kBlock), smem, s, grad_logits, lattice_values, indices, weights, delta_weights, cls_weight, n, pos_dim + 1, val_dim, nr_classes, grad_lattice_values, grad_delta_weights, grad_cls_weight, grad_cls_bias); \
    } while (0)
    if (kv <= 1) LN_LAUNCH_SCB(1);
    else if (kv <= 2) LN_LAUNCH_SCB(2);
    else if (kv <= 4) LN_LAUNCH_SCB(4);
    else LN_LAUNCH_SCB(8);
#undef LN_LAUNCH_SCB
    count_launch();
    return check_launch("slice_classify_bwd");
}

int ln_scatter_max(const float* src, const int* index, int m, int c, int nv, float* out_max, int* out_arg,
                   unsigned long long* workspace, void* stream) {
    LN_REQUIRE(src && index && out_max && out_arg && workspace, "ln_scatter_max: null pointer");
    LN_REQUIRE(m >= 0 && c >= 1 && nv >= 0, "ln_scatter_max: bad size");
    if (nv == 0) return LN_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const long long total = (long long)nv * c;
    if (cudaMemsetAsync(workspace, 0, total * sizeof(unsigned long long), s) != cudaSuccess) return check_launch("scatter_max memset");
    if (m > 0) {
        launch_k(scatter_max_pack_kernel, dim3(cdiv((long long)m * c, kBlock)), dim3(kBlock), 0, s, src, index, m, c, workspace);
        count_launch();
    }
    launch_k(scatter_max_unpack_kernel, dim3(cdiv(total, kBlock)), dim3(kBlock), 0, s, workspace, total, m, out_max, out_arg);
    count_launch();
    return check_launch("scatter_max");
}

int ln_scatter_sum_count(const float* src, const int* index, int m, int c, int nv, float* out_sum, float* out_count,
                         void* stream) {
    LN_REQUIRE(src && index && out_sum, "ln_scatter_sum_count: null pointer");
    LN_REQUIRE(m >= 0 && c >= 1 && nv >= 0, "ln_scatter_sum_count: bad size");
    if (nv == 0) return LN_OK;
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(out_sum, 0, (size_t)nv * c * sizeof(float), s);
    if (out_count) cudaMemsetAsync(out_count, 0, (size_t)nv * sizeof(float), s);
    if (m > 0) {
        launch_k(scatter_sum_count_kernel, dim3(cdiv((long long)m * c, kBlock)), dim3(kBlock), 0, s, src, index, m, c, out_sum, out_count);
        count_launch();
    }
    return check_launch("scatter_sum_count");
}

int ln_deltaw_fwd(const float* values, const int* indices, const float* weights, const float* gamma, const float* beta,
                  const float* lin_w, const float* lin_b, int n, int pos_dim, int val_dim, float* delta_w, void* stream) {
    LN_REQUIRE(values && indices && weights && gamma && beta && lin_w && lin_b && delta_w, "ln_deltaw_fwd: null pointer");
    LN_REQUIRE(n >= 0 && val_dim == kDwC && (pos_dim == 3 || pos_dim == 5), "ln_deltaw_fwd: built for val_dim 8 and pos_dim 3 / 5 (got %d, %d)", val_dim, pos_dim);
    if (n == 0) return LN_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (pos_dim == 3)
        launch_k(deltaw_fwd_kernel<4>, dim3(cdiv(n, kBlock)), dim3(kBlock), 0, s, values, indices, weights, gamma, beta, lin_w, lin_b, n, delta_w);
    else
        launch_k(deltaw_fwd_kernel<6>, dim3(cdiv(n, kBlock)), dim3(kBlock), 0, s, values, indices, weights, gamma, beta, lin_w, lin_b, n, delta_w);
    count_launch();
    return check_launch("deltaw_fwd");
}

int ln_deltaw_bwd(const float* values, const int* indices, const float* weights, const float* gamma, const float* beta,
                  const float* lin_w, const float* grad_delta_w, int n, int pos_dim, int val_dim, float* grad_values_zeroed,
                  float* grad_lin_w_zeroed, float* grad_lin_b_zeroed, float* grad_gamma_zeroed, float* grad_beta_zeroed, void* stream) {
    LN_REQUIRE(values && indices && weights && gamma && beta && lin_w && grad_delta_w && grad_values_zeroed && grad_lin_w_zeroed &&
                   grad_lin_b_zeroed && grad_gamma_zeroed && grad_beta_zeroed,
               "ln_deltaw_bwd: null pointer");
    LN_REQUIRE(n >= 0 && val_dim == kDwC && (pos_dim == 3 || pos_dim == 5), "ln_deltaw_bwd: built for val_dim 8 and pos_dim 3 / 5");
    if (n == 0) return LN_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = min(cdiv(n, kBlock), device_sms() * 4);
    if (pos_dim == 3)
        launch_k(deltaw_bwd_kernel<4>, dim3(grid), dim3(kBlock), 0, s, values, indices, weights, gamma, beta, lin_w, grad_delta_w, n, grad_values_zeroed, grad_lin_w_zeroed, grad_lin_b_zeroed, grad_gamma_zeroed, grad_beta_zeroed);
    else
        launch_k(deltaw_bwd_kernel<6>, dim3(grid), dim3(kBlock), 0, s, values, indices, weights, gamma, beta, lin_w, grad_delta_w, n, grad_values_zeroed, grad_lin_w_zeroed, grad_lin_b_zeroed, grad_gamma_zeroed, grad_beta_zeroed);
    count_launch();
    return check_launch("deltaw_bwd");
}

}  // extern "C"
