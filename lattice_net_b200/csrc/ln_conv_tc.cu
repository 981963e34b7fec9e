// Lattice convolution on the 5th-generation tensor cores (tcgen05 / UMMA, accumulators in TMEM).
//
//   out[q, :] = sum_slot  values[nbr[q, slot'], :] . W[slot*c_in : (slot+1)*c_in, :]      (+ bias)
//
// is a GEMM whose A rows are GATHERED through the neighbour table: M = query vertices (tile 128),
// N = c_out (<= 256, one tile), K = F * c_in walked in blocks of 32 floats (one 128-byte swizzle row).
// No im2row buffer exists anywhere (the reference writes nv*F*c_in floats and reads them back through
// cuBLAS SGEMM, /root/reference/src/Lattice.cu:454-462).
//
// Arithmetic: kind::tf32 with fp32 accumulation.  precision 1 = 3xTF32 error-compensated split
// (A = Ah + Al, B = Bh + Bl;  D += Ah.Bh + Ah.Bl + Al.Bh), which reproduces fp32 SGEMM to ~1e-6 and
// keeps the parity tolerance of the fp32 reference; precision 2 = single TF32 pass.
//
// CTA layout (160 threads):
//   warps 0-3  producers: gather A rows (ld.global.v4 -> split -> st.shared into the 128B-swizzled
//              K-major UMMA layout); thread 0 also launches the bulk-async copy (UBLKCP) of the
//              B slab, which the prep kernel stored pre-swizzled so one copy lands a whole stage;
//              after the K loop the same warps are the epilogue (tcgen05.ld -> bias -> st.global)
//   warp 4     TMEM allocation, and one elected lane issues tcgen05.mma / tcgen05.commit
// smem ring of kStages, full/empty mbarriers between producers and the MMA lane, one mbarrier for
// "accumulator complete".
#include "ln_common.cuh"

namespace ln {

constexpr int kTcThreads = 160;
constexpr int kTileM = 128;
constexpr int kBlockK = 32;                 // floats per K block = 128 bytes = one swizzle row
constexpr int kRowBytes = kBlockK * 4;
constexpr int kATileBytes = kTileM * kRowBytes;   // 16 KB

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem], kind::tf32, issued by ONE thread for the whole CTA
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {   // 32 lanes x 16 consecutive columns
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// K-major, 128B-swizzled shared-memory matrix descriptor (sm_100 format: version 1, SBO = 8 rows * 128 B)
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
    uint64_t desc = 0;
    desc |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address, 16-byte units
    desc |= (uint64_t)0 << 16;                                // leading byte offset: unused for swizzled K-major
    desc |= (uint64_t)(1024 >> 4) << 32;                      // stride byte offset between 8-row groups
    desc |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
    desc |= (uint64_t)2 << 61;                                // layout type: SWIZZLE_128B
    return desc;
}
// kind::tf32 instruction descriptor: D fp32, A/B tf32, both K-major, M x N
__device__ __forceinline__ uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ float to_tf32(float x) {   // round-to-nearest TF32, returned as fp32 bits
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ---- filter preparation ----------------------------------------------------------------------------
// W [F*c_in x c_out] (row = slot*c_in + ci) -> per K block kb a slab [n_pad rows x 128 B] holding
// B^T (n-major rows, 32 k-values each) already in the 128B-swizzled order the UMMA descriptor expects:
// 16-byte chunk j of row n sits at chunk position j ^ (n % 8).  hi = tf32(W), lo = tf32(W - hi).
__global__ void __launch_bounds__(256)
filter_prep_kernel(const float* __restrict__ filter, int k_total, int c_in, int c_out, int n_pad, int split,
                   int transposed, float* __restrict__ b_hi, float* __restrict__ b_lo,
                   float* __restrict__ zero_a, long long n_a, float* __restrict__ zero_b, long long n_b) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // buffers that later kernels accumulate into with atomics (split-K output, weight gradient) are
    // cleared here instead of by separate memset launches
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = t; i < n_a; i += stride) zero_a[i] = 0.0f;
    for (long long i = t; i < n_b; i += stride) zero_b[i] = 0.0f;
    const long long total = (long long)(k_total / kBlockK) * n_pad * kBlockK;
    if (t >= total) return;
    const int kk = (int)(t % kBlockK);
    const long long rest = t / kBlockK;
    const int n = (int)(rest % n_pad);
    const int kb = (int)(rest / n_pad);
    const int k = kb * kBlockK + kk;
    // transposed: `filter` is the forward bank [F*c_out x c_in] of the convolution whose data gradient this is
    // (element (slot, k, n) lives at [(slot*c_out + n), k]), lattice_funcs.py:304-311 without the copy
    float w = 0.0f;
    if (n < c_out) {
        if (transposed) {
            const int slot = k / c_in, ci = k - slot * c_in;
            w = __ldg(filter + ((size_t)slot * c_out + n) * c_in + ci);
        } else {
            w = __ldg(filter + (size_t)k * c_out + n);
        }
    }
    const int chunk = kk >> 2, within = kk & 3;
    const size_t dst = ((size_t)kb * n_pad + n) * kBlockK + (size_t)((chunk ^ (n & 7)) << 2) + within;
    const float hi = to_tf32(w);
    b_hi[dst] = hi;
    if (split) b_lo[dst] = to_tf32(w - hi);
}

// ---- main kernel ----------------------------------------------------------------------------------
template <int kSplit>   // 1: 3xTF32, 0: single pass
__global__ void __launch_bounds__(kTcThreads, 1)
conv_fwd_tc_kernel(const float* __restrict__ values, const int* __restrict__ neighbours,
                   const float* __restrict__ b_hi, const float* __restrict__ b_lo, const float* __restrict__ bias,
                   int nv_query, int F, int c_in, int c_out, int n_pad, int flip, int stages, int kb_per_split,
                   float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages] x { A_hi, (A_lo), B_hi, (B_lo) } tiles (all multiples of 1024 B), then indices, barriers
    const uint32_t b_tile_bytes = (uint32_t)n_pad * kRowBytes;
    const uint32_t stage_bytes = (kSplit ? 2 : 1) * (kATileBytes + b_tile_bytes);
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    int* nbr_sh = (int*)(base + (size_t)stages * stage_bytes);                  // [kTileM][F]
    uint64_t* bars = (uint64_t*)(((uintptr_t)(nbr_sh + kTileM * F) + 15) & ~(uintptr_t)15);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * stages + 1);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int q0 = blockIdx.x * kTileM;
    const uint32_t base_u32 = smem_u32(base);
    const uint32_t bars_u32 = smem_u32(bars);
    auto full_bar = [&](int s) { return bars_u32 + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bars_u32 + 8u * (uint32_t)(stages + s); };
    const uint32_t accum_bar = bars_u32 + 8u * (uint32_t)(2 * stages);
    auto a_hi = [&](int s) { return base_u32 + (uint32_t)s * stage_bytes; };
    auto a_lo = [&](int s) { return a_hi(s) + kATileBytes; };
    auto b_hi_s = [&](int s) { return a_hi(s) + (kSplit ? 2 : 1) * kATileBytes; };
    auto b_lo_s = [&](int s) { return b_hi_s(s) + b_tile_bytes; };

    // neighbour ids of this tile (all slots), coalesced
    for (int i = tid; i < kTileM * F; i += kTcThreads) {
        const int q = q0 + i / F;
        nbr_sh[i] = (q < nv_query) ? __ldg(neighbours + (size_t)q0 * F + i) : -1;
    }
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < n_pad) tmem_cols <<= 1;
    if (tid == 0) {
        for (int s = 0; s < stages; s++) {
            mbar_init(full_bar(s), 128 + 1);   // 128 producer threads + the expect_tx arrival for B
            mbar_init(empty_bar(s), 1);        // one tcgen05.commit
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int cpb = c_in / kBlockK;          // K blocks per slot
    // split-K: blockIdx.y owns K blocks [kb_begin, kb_end); partial tiles are reduced with fp32 atomics
    // into the pre-zeroed output (small lattices: a 1000-vertex level is only 8 M tiles)
    const int kb_begin = blockIdx.y * kb_per_split;
    const int kb_end = min(F * cpb, kb_begin + kb_per_split);
    const int num_kb = kb_end - kb_begin;
    const bool split_k = gridDim.y > 1;

    if (warp < 4) {
        // ================= producers =================
        const int chunk = tid & 7;
        const int row0 = tid >> 3;           // rows row0 + 16*i
        for (int it = 0; it < num_kb; it++) {
            const int kb = kb_begin + it;
            const int s = it % stages;
            const uint32_t ph = (uint32_t)(it / stages) & 1u;
            mbar_wait(empty_bar(s), ph ^ 1u);
            if (tid == 0) {
                mbar_arrive_expect_tx(full_bar(s), (kSplit ? 2u : 1u) * b_tile_bytes);
                bulk_copy_g2s(b_hi_s(s), b_hi + (size_t)kb * n_pad * kBlockK, b_tile_bytes, full_bar(s));
                if (kSplit) bulk_copy_g2s(b_lo_s(s), b_lo + (size_t)kb * n_pad * kBlockK, b_tile_bytes, full_bar(s));
            }
            const int slot = kb / cpb;
            const int cb = kb - slot * cpb;
            const int src_slot = (flip && slot < F - 1) ? (slot ^ 1) : slot;
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int row = row0 + 16 * i;
                const int id = nbr_sh[row * F + src_slot];
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (id >= 0) v[i] = __ldg(reinterpret_cast<const float4*>(values + (size_t)id * c_in + cb * kBlockK) + chunk);
            }
            uint8_t* a_hi_p = base + (size_t)s * stage_bytes;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int row = row0 + 16 * i;
                const uint32_t off = (uint32_t)row * kRowBytes + (uint32_t)((chunk ^ (row & 7)) << 4);
                float4 h = make_float4(to_tf32(v[i].x), to_tf32(v[i].y), to_tf32(v[i].z), to_tf32(v[i].w));
                *reinterpret_cast<float4*>(a_hi_p + off) = h;
                if (kSplit) {
                    float4 l = make_float4(to_tf32(v[i].x - h.x), to_tf32(v[i].y - h.y), to_tf32(v[i].z - h.z), to_tf32(v[i].w - h.w));
                    *reinterpret_cast<float4*>(a_hi_p + kATileBytes + off) = l;
                }
            }
            fence_proxy_async();             // generic-proxy stores -> visible to the tensor-core (async) proxy
            mbar_arrive(full_bar(s));
        }
        // ================= epilogue =================
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int q = q0 + warp * 32 + (tid & 31);
        float* orow = out + (size_t)q * c_out;
        for (int n0 = 0; n0 < n_pad; n0 += 16) {
            float acc[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, acc);
            if (q < nv_query && split_k) {
#pragma unroll
                for (int j = 0; j < 16; j++)
                    if (n0 + j < c_out) atomicAdd(orow + n0 + j, acc[j] + ((bias && blockIdx.y == 0) ? __ldg(bias + n0 + j) : 0.0f));
            } else if (q < nv_query) {
                if (n0 + 16 <= c_out && (c_out & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float4 o = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                        if (bias) {
                            o.x += __ldg(bias + n0 + j); o.y += __ldg(bias + n0 + j + 1);
                            o.z += __ldg(bias + n0 + j + 2); o.w += __ldg(bias + n0 + j + 3);
                        }
                        *reinterpret_cast<float4*>(orow + n0 + j) = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (n0 + j < c_out) orow[n0 + j] = acc[j] + (bias ? __ldg(bias + n0 + j) : 0.0f);
                }
            }
        }
        tc_fence_before();
    } else {
        // ================= MMA issuer (warp 4, one lane) =================
        if ((tid & 31) == 0) {
            const uint32_t idesc = umma_idesc_tf32(kTileM, n_pad);
            for (int kb = 0; kb < num_kb; kb++) {   // kb counts this CTA's K blocks from 0
                const int s = kb % stages;
                const uint32_t ph = (uint32_t)(kb / stages) & 1u;
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint64_t da_hi = umma_desc_kmajor_sw128(a_hi(s));
                const uint64_t db_hi = umma_desc_kmajor_sw128(b_hi_s(s));
                const uint64_t da_lo = umma_desc_kmajor_sw128(a_lo(s));
                const uint64_t db_lo = umma_desc_kmajor_sw128(b_lo_s(s));
#pragma unroll
                for (int ks = 0; ks < kBlockK / 8; ks++) {   // UMMA K = 8 tf32 = 32 bytes = 2 x 16-byte units
                    const uint64_t adv = (uint64_t)(ks * 2);
                    if (kSplit) {                             // small cross terms first, then the main product
                        umma_tf32(tmem_base, da_lo + adv, db_hi + adv, idesc, (kb | ks) != 0 ? 1u : 0u);
                        umma_tf32(tmem_base, da_hi + adv, db_lo + adv, idesc, 1u);
                        umma_tf32(tmem_base, da_hi + adv, db_hi + adv, idesc, 1u);
                    } else {
                        umma_tf32(tmem_base, da_hi + adv, db_hi + adv, idesc, (kb | ks) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(empty_bar(s));   // frees the stage when these MMAs have read it
            }
            umma_commit(accum_bar);          // accumulator complete -> epilogue
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

size_t conv_tc_workspace_bytes(int F, int c_in, int c_out) {
    const int n_pad = (c_out + 15) / 16 * 16;
    return (size_t)2 * F * c_in * n_pad * sizeof(float);
}

bool conv_tc_supported(int F, int c_in, int c_out) { return c_in % kBlockK == 0 && c_out >= 1 && c_out <= 256 && F >= 3; }

int conv_fwd_tc(const float* nbr_values, const int* neighbours, const float* filter, const float* bias, int nv_query,
                int F, int c_in, int c_out, int flip, int precision, int transposed, float* workspace, float* out,
                float* also_zero, long long also_zero_n, cudaStream_t s) {
    const int n_pad = (c_out + 15) / 16 * 16;
    const int k_total = F * c_in;
    const int split = precision == 1 ? 1 : 0;
    float* b_hi = workspace;
    float* b_lo = workspace + (size_t)k_total * n_pad;
    const int num_kb = F * (c_in / kBlockK);
    // enough CTAs for the machine: split K when there are few M tiles (148 SMs, one CTA each)
    const int m_tiles = cdiv(nv_query, kTileM);
    int splits = 1;
    if (m_tiles < 148) splits = max(1, min(num_kb, 148 / m_tiles));
    const int kb_per_split = cdiv(num_kb, splits);
    splits = cdiv(num_kb, kb_per_split);
    {
        const long long total = (long long)k_total * n_pad;
        filter_prep_kernel<<<cdiv(total, 256), 256, 0, s>>>(filter, k_total, c_in, c_out, n_pad, split, transposed, b_hi, b_lo,
                                                            out, splits > 1 ? (long long)nv_query * c_out : 0, also_zero, also_zero_n);
        count_launch();
    }
    const size_t b_tile = (size_t)n_pad * kRowBytes;
    const size_t stage_bytes = (split ? 2 : 1) * (kATileBytes + b_tile);
    const size_t fixed = (size_t)kTileM * F * sizeof(int) + 16 + (2 * 8 + 1) * 8 + 16 + 1024;
    int stages = (int)((227 * 1024 - fixed) / stage_bytes);
    stages = min(stages, 8);
    stages = min(stages, kb_per_split);
    if (stages < 2 && kb_per_split >= 2) {
        set_error("ln_conv_fwd: tensor-core tile does not fit shared memory (c_out=%d)", c_out);
        return LN_ERR_UNSUPPORTED;
    }
    const size_t smem = (size_t)stages * stage_bytes + fixed;
    const dim3 grid(m_tiles, splits);
    cudaError_t err;
    if (split) {
        err = cudaFuncSetAttribute(conv_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err == cudaSuccess)
            conv_fwd_tc_kernel<1><<<grid, kTcThreads, smem, s>>>(nbr_values, neighbours, b_hi, b_lo, bias, nv_query, F, c_in, c_out, n_pad, flip, stages, kb_per_split, out);
    } else {
        err = cudaFuncSetAttribute(conv_fwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err == cudaSuccess)
            conv_fwd_tc_kernel<0><<<grid, kTcThreads, smem, s>>>(nbr_values, neighbours, b_hi, b_lo, bias, nv_query, F, c_in, c_out, n_pad, flip, stages, kb_per_split, out);
    }
    if (err != cudaSuccess) {
        set_error("conv_fwd_tc: %s", cudaGetErrorString(err));
        return LN_ERR_CUDA;
    }
    count_launch();
    return check_launch("conv_fwd_tc");
}

}  // namespace ln
