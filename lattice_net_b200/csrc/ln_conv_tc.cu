// Lattice convolution on the 5th-generation tensor cores (tcgen05 / UMMA, accumulators in TMEM).
//
//   out[q, :] = sum_slot  values[nbr[q, slot'], :] . W[slot*c_in : (slot+1)*c_in, :]      (+ bias)
//
// is a GEMM whose A rows are GATHERED through the neighbour table: M = query vertices (tile 128),
// N = c_out (one tile of up to 256 columns; wider layers run as 256-column chunks), K = F * c_in walked in blocks of
// 32 floats (one 128-byte swizzle row).
// No im2row buffer exists anywhere (the reference writes nv*F*c_in floats and reads them back through
// cuBLAS SGEMM, /root/reference/src/Lattice.cu:454-462).
//
// Arithmetic: kind::tf32 with fp32 accumulation.  precision 1 = 3xTF32 error-compensated split
// (A = Ah + Al, B = Bh + Bl;  D += Ah.Bh + Ah.Bl + Al.Bh), which reproduces fp32 SGEMM to ~1e-6 and
// keeps the parity tolerance of the fp32 reference; precision 2 = single TF32 pass.
//
// Persistent CTAs (one per SM) walk work items = (M tile, K split); the smem stage ring and the two TMEM
// accumulator buffers run straight across item boundaries, so the gathers of item i+1 overlap the MMAs of
// item i and the epilogue of item i-1.  CTA layout (416 threads):
//   warps 0-7  producers: neighbour rows are gathered with cp.async (LDGSTS, 16 B, zero-fill for absent
//              neighbours) DIRECTLY into the 128B-swizzled K-major UMMA layout, `lookahead` K blocks in
//              flight per thread; when a block has landed its owner derives the TF32 high / low parts
//              in place (3xTF32 only), fences the async proxy and arrives on the stage's "full" barrier.
//              Thread 0 also launches the bulk-async copy (UBLKCP) of the pre-swizzled B slab of that stage.
//   warp 8     TMEM allocation; one lane issues tcgen05.mma / tcgen05.commit
//   warps 9-12 epilogue: tcgen05.ld (32 columns = one 128-byte line per thread) -> bias / residual -> st.global
//              (vector fp32 atomics when K is split across CTAs)
//
// The filter bank reaches the kernel as PREPARED SLABS (filter_prep kernels below): per K block a [n_pad x 128 B]
// tile of B^T, pre-swizzled, split into TF32 high / low parts.  A bank is prepared once per optimizer step for the
// forward AND the transposed (data-gradient) reading -- one batched launch for all banks of a model -- instead of
// once per convolution call.
#include <mutex>
#include <set>
#include <utility>
#include "ln_common.cuh"

namespace ln {

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

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {   // 32 lanes x 32 consecutive columns
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
// 16-byte async copy global -> shared; src_bytes = 0 zero-fills the destination (absent neighbour)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void producer_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 producer warps of conv_tc2
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {   // round-to-nearest TF32, returned as fp32 bits
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
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


// ---- filter preparation ----------------------------------------------------------------------------
// W [F*c_in x c_out] (row = slot*c_in + ci) -> prepared slabs.  The output channels are cut into chunks of kMaxTileN
// (one UMMA N tile); chunk j (n_off = j*kMaxTileN, n_pad = its width rounded up to 16) owns the floats
// [2*K*n_off, 2*K*(n_off + n_pad)) of the slab buffer: first the TF32 high parts, K*n_pad floats, then the low parts.
// Inside a chunk, K block kb is a tile [n_pad rows x 128 B] holding B^T (n-major rows, 32 k-values each) already in the
// 128B-swizzled order the UMMA descriptor expects: 16-byte unit j of row n sits at unit position j ^ (n % 8).
//   hi = tf32(W), lo = tf32(W - hi).
// transposed: `filter` is the FORWARD bank of the convolution whose DATA GRADIENT is being computed, read in place
// (lattice_funcs.py:304-311 without the copy): this GEMM's K runs over (slot, forward output channel), its N over the
// forward input channels, so element (slot, k, n) lives at filter[(slot*N + n) * c_in + k] with c_in = this reading's
// channels per slot (= the forward c_out) and N = this reading's c_out (= the forward c_in).
constexpr int kMaxTileN = 256;

struct FilterPrepJob {          // one bank; 48 bytes, built on the host (lattice.py packs the same layout)
    const float* src;
    float* dst;
    int k_total;                // GEMM K = F * c_in (of THIS reading)
    int c_in;                   // channels per filter slot of this reading
    int c_out;                  // GEMM N
    int transposed;
    int split;                  // 1: write the low parts too (3xTF32)
    int pad_;
    long long first_thread;     // prefix sum of TILE counts (32 x 32 slab elements, one CTA each) over the jobs of a batch
};

__host__ __device__ __forceinline__ int prep_n_pad_sum(int c_out) {   // padded columns over all chunks
    const int full = c_out / kMaxTileN, rest = c_out - full * kMaxTileN;
    return full * kMaxTileN + (rest + 15) / 16 * 16;
}

// One CTA of 1024 threads per tile of 32 (k) x 32 (n) slab elements.  The transposed reading is contiguous along k on both
// sides (loads from the forward bank's rows, stores into a slab row): straight through.  The plain reading is contiguous
// along n in the bank and along k in the slab: the tile is transposed through shared memory so that loads and stores are
// both coalesced (the first version loaded 32 different bank rows per warp instruction: 66 us for the 14.7 MB of LatticeNet).
__device__ __forceinline__ void filter_prep_tile(const FilterPrepJob& j, long long tile, float (*sh)[33]) {
    const int n_pad_sum = prep_n_pad_sum(j.c_out);
    const int n_tiles = (n_pad_sum + 31) / 32;                 // n_pad_sum is a multiple of 16
    const int kb = (int)(tile / n_tiles);
    const int nt = (int)(tile - (long long)kb * n_tiles);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    float w = 0.0f;
    if (j.transposed) {
        const int n = nt * 32 + ty, k = kb * kBlockK + tx;
        if (n < j.c_out) {
            const int slot = k / j.c_in, ci = k - slot * j.c_in;
            w = __ldg(j.src + ((size_t)slot * j.c_out + n) * j.c_in + ci);
        }
    } else {
        const int n = nt * 32 + tx, k = kb * kBlockK + ty;
        sh[ty][tx] = n < j.c_out ? __ldg(j.src + (size_t)k * j.c_out + n) : 0.0f;
        __syncthreads();
        w = sh[tx][ty];                                        // element (k = tx, n = ty) of the tile
    }
    const int kk = tx, np = nt * 32 + ty;
    if (np >= n_pad_sum) return;
    const int n_off = np / kMaxTileN * kMaxTileN;
    const int nl = np - n_off;
    const int n_pad = min(kMaxTileN, n_pad_sum - n_off);
    const int unit = kk >> 2, within = kk & 3;
    float* hi_base = j.dst + (size_t)2 * j.k_total * n_off;
    const size_t dst = ((size_t)kb * n_pad + nl) * kBlockK + (size_t)((unit ^ (nl & 7)) << 2) + within;
    const float hi = to_tf32(w);
    hi_base[dst] = hi;
    if (j.split) hi_base[(size_t)j.k_total * n_pad + dst] = to_tf32(w - hi);
}

// tiles of one job
__host__ __device__ __forceinline__ long long prep_tiles(int k_total, int c_out) {
    return (long long)(k_total / kBlockK) * ((prep_n_pad_sum(c_out) + 31) / 32);
}

__global__ void __launch_bounds__(1024) filter_prep_kernel(FilterPrepJob job) {
    LN_PDL_ENTRY();
    __shared__ float sh[32][33];
    filter_prep_tile(job, blockIdx.x, sh);
}

// all banks of a model in ONE launch: jobs[] lives in device memory, sorted by first_tile (the field `first_thread`)
__global__ void __launch_bounds__(1024) filter_prep_batch_kernel(const FilterPrepJob* __restrict__ jobs, int n_jobs, long long total) {
    LN_PDL_ENTRY();
    __shared__ float sh[32][33];
    const long long t = blockIdx.x;
    if (t >= total) return;
    int lo = 0, hi = n_jobs - 1;              // last job whose first tile <= t
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(&jobs[mid].first_thread) <= t) lo = mid; else hi = mid - 1;
    }
    const FilterPrepJob j = jobs[lo];
    filter_prep_tile(j, t - j.first_thread, sh);
}

// ---- persistent, cp.async-fed kernel (v2) -----------------------------------------------------------
constexpr int kTc2ProducerWarps = 8;
constexpr int kTc2Producers = kTc2ProducerWarps * 32;
constexpr int kTc2Threads = kTc2Producers + 32 + 128;   // producers, MMA warp, 4 epilogue warps
constexpr int kMaxStages = 6;

struct Tc2Item {   // one unit of work of a persistent CTA
    int q0, kb_begin, num_kb, split;
};
__device__ __forceinline__ Tc2Item tc2_item(int item, int m_tiles, int total_kb, int kb_per_split) {
    Tc2Item w;
    w.split = item / m_tiles;                    // items of one split are contiguous: neighbouring CTAs share B slabs in L2
    w.q0 = (item - w.split * m_tiles) * kTileM;
    w.kb_begin = w.split * kb_per_split;
    w.num_kb = min(total_kb, w.kb_begin + kb_per_split) - w.kb_begin;
    return w;
}

template <int kSplit>   // 1: 3xTF32, 0: single pass
__global__ void __launch_bounds__(kTc2Threads, 1)
conv_tc2_kernel(const float* __restrict__ values, const int* __restrict__ neighbours,
                const float* __restrict__ b_hi, const float* __restrict__ b_lo, const float* __restrict__ bias,
                const float* __restrict__ residual, int nv_query, int F, int c_in, int c_out, int ld_out, int n_pad, int flip,
                int stages, int lookahead, int m_tiles, int n_items, int kb_per_split, float* __restrict__ out) {
    pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t b_tile_bytes = (uint32_t)n_pad * kRowBytes;
    const uint32_t stage_bytes = (kSplit ? 2 : 1) * (kATileBytes + b_tile_bytes);
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    int* nbr_sh = (int*)(base + (size_t)stages * stage_bytes);                      // [2][kTileM * F]
    uint64_t* bars = (uint64_t*)(((uintptr_t)(nbr_sh + 2 * kTileM * F) + 15) & ~(uintptr_t)15);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kMaxStages + 4);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const uint32_t base_u32 = smem_u32(base);
    const uint32_t bars_u32 = smem_u32(bars);
    auto full_bar = [&](int s) { return bars_u32 + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bars_u32 + 8u * (uint32_t)(kMaxStages + s); };
    auto acc_full_bar = [&](int a) { return bars_u32 + 8u * (uint32_t)(2 * kMaxStages + a); };
    auto acc_empty_bar = [&](int a) { return bars_u32 + 8u * (uint32_t)(2 * kMaxStages + 2 + a); };
    auto a_hi = [&](int s) { return base_u32 + (uint32_t)s * stage_bytes; };
    auto a_lo = [&](int s) { return a_hi(s) + kATileBytes; };
    auto b_hi_s = [&](int s) { return a_hi(s) + (kSplit ? 2 : 1) * kATileBytes; };
    auto b_lo_s = [&](int s) { return b_hi_s(s) + b_tile_bytes; };

    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 2 * n_pad) tmem_cols <<= 1;       // two accumulator buffers
    if (tid == 0) {
        for (int s = 0; s < stages; s++) {
            mbar_init(full_bar(s), kTc2Producers + 1);   // the producer threads + the expect_tx arrival for B
            mbar_init(empty_bar(s), 1);        // one tcgen05.commit
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(acc_full_bar(a), 1);     // one tcgen05.commit
            mbar_init(acc_empty_bar(a), 128);  // the 128 epilogue threads
        }
        fence_barrier_init();
    }
    if (warp == kTc2ProducerWarps) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();          // everything above ran while the previous kernel was still finishing; global memory from here on

    const int cpb = c_in / kBlockK;          // K blocks per slot
    const int total_kb = F * cpb;
    const bool split_k = n_items > m_tiles;

    if (warp < kTc2ProducerWarps) {
        // ================= producers (8 warps: the loop is instruction-bound, two warps per scheduler) =================
        // All per-block state (stage address, barrier address, phase, slot / channel-block position, B source) is
        // advanced incrementally: no division, no 64-bit multiply and no stage arithmetic inside the K loop.
        const int chunk = tid & 7;
        const int row0 = tid >> 3;           // this thread's rows: row0 + 32*i, i < 4; (row & 7) is the same for all of them
        const uint32_t my_off = (uint32_t)row0 * kRowBytes + (uint32_t)((chunk ^ (row0 & 7)) << 4);
        const int tile_ids = kTileM * F;
        const int row_stride = 32 * F;       // neighbour ids between two of this thread's rows
        const int per_thread = (tile_ids + kTc2Producers - 1) / kTc2Producers;      // ids staged per thread and item (<= 7)
        int ibuf = 0;
        {   // neighbour ids of the first item
            const Tc2Item w = tc2_item(blockIdx.x, m_tiles, total_kb, kb_per_split);
            for (int i = tid; i < tile_ids; i += kTc2Producers) {
                const int q = w.q0 + i / F;
                nbr_sh[i] = (q < nv_query) ? __ldg(neighbours + (size_t)w.q0 * F + i) : -1;
            }
        }
        producer_bar_sync();
        const uint32_t ring_end = base_u32 + (uint32_t)stages * stage_bytes;
        uint32_t issue_addr = base_u32, issue_full = full_bar(0), issue_empty = empty_bar(0), issue_phase = 0;   // next block to issue
        uint32_t pub_addr = base_u32, pub_full = full_bar(0);                                                  // next block to publish
        int in_flight = 0;
        const float* col0 = values + chunk * 4;
        const size_t b_block = (size_t)n_pad * kBlockK;          // floats of one B slab
        auto publish = [&]() {               // the oldest in-flight block has landed in this thread's view: finish it, signal the MMA lane
            if (kSplit) {
                const uint32_t hi = pub_addr + my_off, lo = pub_addr + kATileBytes + my_off;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float4 x = lds128(hi + (uint32_t)i * 32u * kRowBytes);
                    // explicit round-to-nearest high part, independent of how the tensor core reads a raw fp32 operand
                    const float4 h = make_float4(to_tf32(x.x), to_tf32(x.y), to_tf32(x.z), to_tf32(x.w));
                    sts128(hi + (uint32_t)i * 32u * kRowBytes, h);
                    sts128(lo + (uint32_t)i * 32u * kRowBytes,
                           make_float4(to_tf32(x.x - h.x), to_tf32(x.y - h.y), to_tf32(x.z - h.z), to_tf32(x.w - h.w)));
                }
            }
            fence_proxy_async();             // generic-proxy writes (cp.async data, lo tile) -> visible to the tensor-core proxy
            mbar_arrive(pub_full);
            pub_addr += stage_bytes;
            pub_full += 8;
            if (pub_addr == ring_end) {
                pub_addr = base_u32;
                pub_full = full_bar(0);
            }
            in_flight--;
        };
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const Tc2Item w = tc2_item(item, m_tiles, total_kb, kb_per_split);
            const int* ids = nbr_sh + ibuf * tile_ids + row0 * F;
            // prefetch the next item's neighbour ids into registers (stored to the other buffer at the end of this item)
            int next_ids[7];
            const int next = item + gridDim.x;
            if (next < n_items) {
                const Tc2Item wn = tc2_item(next, m_tiles, total_kb, kb_per_split);
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    const int i = tid + kTc2Producers * j;
                    next_ids[j] = -1;
                    if (j < per_thread && i < tile_ids && wn.q0 + i / F < nv_query) next_ids[j] = __ldg(neighbours + (size_t)wn.q0 * F + i);
                }
            }
            int slot = w.kb_begin / cpb;
            int cb = w.kb_begin - slot * cpb;
            const float* b_src_hi = b_hi + (size_t)w.kb_begin * b_block;
            const float* b_src_lo = b_lo + (size_t)w.kb_begin * b_block;
            int id[4];
            bool reload = true;
            for (int it = 0; it < w.num_kb; it++) {
                if (reload) {                // a new filter slot: this thread's four neighbour ids change
                    const int src_slot = (flip && slot < F - 1) ? (slot ^ 1) : slot;
#pragma unroll
                    for (int i = 0; i < 4; i++) id[i] = ids[i * row_stride + src_slot];
                    reload = false;
                }
                mbar_wait(issue_empty, issue_phase ^ 1u);
                if (tid == 0) {
                    mbar_arrive_expect_tx(issue_full, (kSplit ? 2u : 1u) * b_tile_bytes);
                    bulk_copy_g2s(issue_addr + (kSplit ? 2 : 1) * kATileBytes, b_src_hi, b_tile_bytes, issue_full);
                    if (kSplit) bulk_copy_g2s(issue_addr + 2 * kATileBytes + b_tile_bytes, b_src_lo, b_tile_bytes, issue_full);
                }
                const uint32_t dst = issue_addr + my_off;
                const float* col = col0 + cb * kBlockK;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const bool have = id[i] >= 0;
                    const float* src = have ? col + (unsigned long long)(unsigned)id[i] * (unsigned)c_in : values;   // mul.wide.u32
                    cp_async16(dst + (uint32_t)i * 32u * kRowBytes, src, have ? 16u : 0u);
                }
                cp_async_commit();
                in_flight++;
                // advance the issue state
                b_src_hi += b_block;
                b_src_lo += b_block;
                if (++cb == cpb) {
                    cb = 0;
                    slot++;
                    reload = true;
                }
                issue_addr += stage_bytes;
                issue_full += 8;
                issue_empty += 8;
                if (issue_addr == ring_end) {
                    issue_addr = base_u32;
                    issue_full = full_bar(0);
                    issue_empty = empty_bar(0);
                    issue_phase ^= 1u;
                }
                if (in_flight > lookahead) {
                    switch (lookahead) {     // cp.async.wait_group takes an immediate
                        case 1: cp_async_wait<1>(); break;
                        case 2: cp_async_wait<2>(); break;
                        case 3: cp_async_wait<3>(); break;
                        case 4: cp_async_wait<4>(); break;
                        default: cp_async_wait<5>(); break;
                    }
                    publish();
                }
            }
            if (next < n_items) {
                int* dst_ids = nbr_sh + (ibuf ^ 1) * tile_ids;
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    const int i = tid + kTc2Producers * j;
                    if (j < per_thread && i < tile_ids) dst_ids[i] = next_ids[j];
                }
            }
            producer_bar_sync();
            ibuf ^= 1;
        }
        cp_async_wait<0>();
        while (in_flight > 0) publish();
    } else if (warp == kTc2ProducerWarps) {
        // ================= MMA issuer (one lane) =================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(kTileM, n_pad);
            int g = 0, j = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, j++) {
                const Tc2Item w = tc2_item(item, m_tiles, total_kb, kb_per_split);
                const int a = j & 1;
                mbar_wait(acc_empty_bar(a), (((uint32_t)(j >> 1)) & 1u) ^ 1u);   // epilogue has drained this buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(a * n_pad);
                for (int it = 0; it < w.num_kb; it++, g++) {
                    const int s = g % stages;
                    mbar_wait(full_bar(s), ((uint32_t)(g / stages)) & 1u);
                    tc_fence_after();
                    const uint64_t da_hi = umma_desc_kmajor_sw128(a_hi(s));
                    const uint64_t db_hi = umma_desc_kmajor_sw128(b_hi_s(s));
                    const uint64_t da_lo = umma_desc_kmajor_sw128(a_lo(s));
                    const uint64_t db_lo = umma_desc_kmajor_sw128(b_lo_s(s));
#pragma unroll
                    for (int ks = 0; ks < kBlockK / 8; ks++) {   // UMMA K = 8 tf32 = 32 bytes = 2 x 16-byte units
                        const uint64_t adv = (uint64_t)(ks * 2);
                        if (kSplit) {                             // small cross terms first, then the main product
                            umma_tf32(d_tmem, da_lo + adv, db_hi + adv, idesc, (it | ks) != 0 ? 1u : 0u);
                            umma_tf32(d_tmem, da_hi + adv, db_lo + adv, idesc, 1u);
                            umma_tf32(d_tmem, da_hi + adv, db_hi + adv, idesc, 1u);
                        } else {
                            umma_tf32(d_tmem, da_hi + adv, db_hi + adv, idesc, (it | ks) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(empty_bar(s));   // frees the stage when these MMAs have read it
                }
                umma_commit(acc_full_bar(a));    // accumulator complete -> epilogue
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue (4 warps; TMEM lane quadrant = warp % 4) =================
        const int quad = warp & 3;
        int j = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, j++) {
            const Tc2Item w = tc2_item(item, m_tiles, total_kb, kb_per_split);
            const int a = j & 1;
            mbar_wait(acc_full_bar(a), ((uint32_t)(j >> 1)) & 1u);
            tc_fence_after();
            const int q = w.q0 + quad * 32 + lane;
            const bool live = q < nv_query;
            float* orow = out + (size_t)q * ld_out;     // `out` / `bias` / `residual` already point at this launch's first channel
            const bool add_bias = bias != nullptr && w.split == 0;
            const float* rrow = (residual != nullptr && w.split == 0) ? residual + (size_t)q * ld_out : nullptr;
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * n_pad);
            for (int n0 = 0; n0 < n_pad; n0 += 32) {
                float acc[32];
                if (n0 + 32 <= n_pad) {
                    tmem_ld32(t_row + (uint32_t)n0, acc);
                } else {                       // n_pad is a multiple of 16: a 16-column tail
                    tmem_ld16(t_row + (uint32_t)n0, acc);
#pragma unroll
                    for (int k = 16; k < 32; k++) acc[k] = 0.0f;
                }
                if (!live) continue;
                if (((c_out | ld_out) & 3) == 0) {      // 16-byte aligned rows and whole float4 groups
#pragma unroll
                    for (int k = 0; k < 32; k += 4) {
                        if (n0 + k < c_out) {
                            float4 o = make_float4(acc[k], acc[k + 1], acc[k + 2], acc[k + 3]);
                            if (add_bias) {
                                o.x += __ldg(bias + n0 + k); o.y += __ldg(bias + n0 + k + 1);
                                o.z += __ldg(bias + n0 + k + 2); o.w += __ldg(bias + n0 + k + 3);
                            }
                            if (rrow != nullptr) {
                                const float4 r4 = __ldg(reinterpret_cast<const float4*>(rrow + n0 + k));
                                o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
                            }
                            if (split_k)
                                atomicAdd(reinterpret_cast<float4*>(orow + n0 + k), o);
                            else
                                *reinterpret_cast<float4*>(orow + n0 + k) = o;
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 32; k++) {
                        if (n0 + k < c_out) {
                            const float o = acc[k] + (add_bias ? __ldg(bias + n0 + k) : 0.0f) + (rrow != nullptr ? __ldg(rrow + n0 + k) : 0.0f);
                            if (split_k)
                                atomicAdd(orow + n0 + k, o);
                            else
                                orow[n0 + k] = o;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty_bar(a));     // this thread has read its rows of the buffer
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kTc2ProducerWarps) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ---- 3xTF32 kernel with decoupled rings (v3) ----------------------------------------------------------------------
// ncu on the scene-sized sweep (10^6 points, 64 -> 64: tensor pipe 22 %, L2 throughput 19 %) and on ShapeNet-sized calls
// (~1 us per extra K block) showed conv_tc2<1> LATENCY-bound: a 3xTF32 stage carries A_hi + A_lo + B_hi + B_lo, so only
// 2..4 stages fit and at most 1..2 gathers are in flight per thread -- ~48 KB per SM against the ~100 KB an L2 latency of
// ~1 us needs.  The low part of A, however, is COMPUTED, not loaded: it needs no shared memory while its gather is in
// flight.  Three rings instead of one:
//   landing ring  `landing` x 16 KB: raw fp32 rows straight from cp.async, up to 8 K blocks in flight.  Every producer
//                 thread re-reads exactly the 16-byte units it gathered itself, so a slot is recycled without any barrier;
//   A ring        `a_stages` (2..4) x (hi 16 KB | lo 16 KB): written by the producers when a block has landed (round-to-nearest TF32 high
//                 part + residual), consumed by the MMA lane;
//   B ring        `b_stages` x (hi | lo) pre-split filter slabs, fed by a dedicated bulk-copy warp that runs ahead of the
//                 MMAs independently of the gathers.
// Warps: 0-7 producers, 8 MMA issue, 9 filter-slab loader, 10-13 epilogue.
constexpr int kTc3Threads = kTc2Producers + 32 + 32 + 128;
constexpr int kTc3MaxA = 4;
constexpr int kTc3MaxB = 4;
constexpr int kTc3MaxLanding = 8;
constexpr int kTc3Bars = 2 * kTc3MaxA + 2 * kTc3MaxB + 4;

__global__ void __launch_bounds__(kTc3Threads, 1)
conv_tc3_kernel(const float* __restrict__ values, const int* __restrict__ neighbours,
                const float* __restrict__ b_hi, const float* __restrict__ b_lo, const float* __restrict__ bias,
                const float* __restrict__ residual, int nv_query, int F, int c_in, int c_out, int ld_out, int n_pad, int flip,
                int landing, int a_stages, int b_stages, int m_tiles, int n_items, int kb_per_split, float* __restrict__ out) {
    pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t b_tile_bytes = (uint32_t)n_pad * kRowBytes;
    const uint32_t a_stage_bytes = 2u * kATileBytes, b_stage_bytes = 2u * b_tile_bytes;
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t base_u32 = smem_u32(base);
    const uint32_t a_base = base_u32;
    const uint32_t b_base = a_base + (uint32_t)a_stages * a_stage_bytes;
    const uint32_t land_base = b_base + (uint32_t)b_stages * b_stage_bytes;
    int* nbr_sh = (int*)(base + (size_t)a_stages * a_stage_bytes + (size_t)b_stages * b_stage_bytes + (size_t)landing * kATileBytes);   // [2][kTileM * F]
    uint64_t* bars = (uint64_t*)(((uintptr_t)(nbr_sh + 2 * kTileM * F) + 15) & ~(uintptr_t)15);
    uint32_t* tmem_slot = (uint32_t*)(bars + kTc3Bars);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const uint32_t bars_u32 = smem_u32(bars);
    auto full_a = [&](int s) { return bars_u32 + 8u * (uint32_t)s; };
    auto empty_a = [&](int s) { return bars_u32 + 8u * (uint32_t)(kTc3MaxA + s); };
    auto full_b = [&](int s) { return bars_u32 + 8u * (uint32_t)(2 * kTc3MaxA + s); };
    auto empty_b = [&](int s) { return bars_u32 + 8u * (uint32_t)(2 * kTc3MaxA + kTc3MaxB + s); };
    auto acc_full_bar = [&](int a) { return bars_u32 + 8u * (uint32_t)(2 * kTc3MaxA + 2 * kTc3MaxB + a); };
    auto acc_empty_bar = [&](int a) { return bars_u32 + 8u * (uint32_t)(2 * kTc3MaxA + 2 * kTc3MaxB + 2 + a); };

    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 2 * n_pad) tmem_cols <<= 1;       // two accumulator buffers
    if (tid == 0) {
        for (int s = 0; s < a_stages; s++) {
            mbar_init(full_a(s), kTc2Producers);   // every producer thread has written its units of the hi / lo tiles
            mbar_init(empty_a(s), 1);              // one tcgen05.commit
        }
        for (int s = 0; s < b_stages; s++) {
            mbar_init(full_b(s), 1);               // the expect_tx arrival of the loader (+ the bytes of the two bulk copies)
            mbar_init(empty_b(s), 1);              // one tcgen05.commit
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(acc_full_bar(a), 1);         // one tcgen05.commit
            mbar_init(acc_empty_bar(a), 128);      // the 128 epilogue threads
        }
        fence_barrier_init();
    }
    if (warp == kTc2ProducerWarps) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();          // everything above ran while the previous kernel was still finishing; global memory from here on

    const int cpb = c_in / kBlockK;          // K blocks per slot
    const int total_kb = F * cpb;
    const bool split_k = n_items > m_tiles;

    if (warp < kTc2ProducerWarps) {
        // ================= producers =================
        const int chunk = tid & 7;
        const int row0 = tid >> 3;           // this thread's rows: row0 + 32*i, i < 4; (row & 7) is the same for all of them
        const uint32_t my_off = (uint32_t)row0 * kRowBytes + (uint32_t)((chunk ^ (row0 & 7)) << 4);
        const int tile_ids = kTileM * F;
        const int row_stride = 32 * F;       // neighbour ids between two of this thread's rows
        const int per_thread = (tile_ids + kTc2Producers - 1) / kTc2Producers;      // ids staged per thread and item (<= 7)
        int ibuf = 0;
        {   // neighbour ids of the first item
            const Tc2Item w = tc2_item(blockIdx.x, m_tiles, total_kb, kb_per_split);
            for (int i = tid; i < tile_ids; i += kTc2Producers) {
                const int q = w.q0 + i / F;
                nbr_sh[i] = (q < nv_query) ? __ldg(neighbours + (size_t)w.q0 * F + i) : -1;
            }
        }
        producer_bar_sync();
        const uint32_t land_end = land_base + (uint32_t)landing * kATileBytes;
        uint32_t issue_land = land_base, pub_land = land_base;
        uint32_t pa_addr = a_base, pa_full = full_a(0), pa_empty = empty_a(0), pa_phase = 0;
        int in_flight = 0;
        const float* col0 = values + chunk * 4;
        auto publish = [&]() {               // the oldest gather of this thread has landed: split it into the A ring, signal the MMA lane
            mbar_wait(pa_empty, pa_phase ^ 1u);                  // the MMAs that read this A stage two blocks ago have retired
            const uint32_t src = pub_land + my_off, hi = pa_addr + my_off, lo = hi + kATileBytes;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float4 x = lds128(src + (uint32_t)i * 32u * kRowBytes);
                const float4 h = make_float4(to_tf32(x.x), to_tf32(x.y), to_tf32(x.z), to_tf32(x.w));
                sts128(hi + (uint32_t)i * 32u * kRowBytes, h);
                sts128(lo + (uint32_t)i * 32u * kRowBytes,
                       make_float4(to_tf32(x.x - h.x), to_tf32(x.y - h.y), to_tf32(x.z - h.z), to_tf32(x.w - h.w)));
            }
            fence_proxy_async();             // generic-proxy writes -> visible to the tensor-core proxy
            mbar_arrive(pa_full);
            pub_land += kATileBytes;
            if (pub_land == land_end) pub_land = land_base;
            pa_addr += a_stage_bytes;
            pa_full += 8;
            pa_empty += 8;
            if (pa_addr == a_base + (uint32_t)a_stages * a_stage_bytes) {
                pa_addr = a_base;
                pa_full = full_a(0);
                pa_empty = empty_a(0);
                pa_phase ^= 1u;
            }
            in_flight--;
        };
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const Tc2Item w = tc2_item(item, m_tiles, total_kb, kb_per_split);
            const int* ids = nbr_sh + ibuf * tile_ids + row0 * F;
            // prefetch the next item's neighbour ids into registers (stored to the other buffer at the end of this item)
            int next_ids[7];
            const int next = item + gridDim.x;
            if (next < n_items) {
                const Tc2Item wn = tc2_item(next, m_tiles, total_kb, kb_per_split);
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    const int i = tid + kTc2Producers * j;
                    next_ids[j] = -1;
                    if (j < per_thread && i < tile_ids && wn.q0 + i / F < nv_query) next_ids[j] = __ldg(neighbours + (size_t)wn.q0 * F + i);
                }
            }
            int slot = w.kb_begin / cpb;
            int cb = w.kb_begin - slot * cpb;
            int id[4];
            bool reload = true;
            for (int it = 0; it < w.num_kb; it++) {
                if (reload) {                // a new filter slot: this thread's four neighbour ids change
                    const int src_slot = (flip && slot < F - 1) ? (slot ^ 1) : slot;
#pragma unroll
                    for (int i = 0; i < 4; i++) id[i] = ids[i * row_stride + src_slot];
                    reload = false;
                }
                const uint32_t dst = issue_land + my_off;
                const float* col = col0 + cb * kBlockK;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const bool have = id[i] >= 0;
                    const float* src = have ? col + (unsigned long long)(unsigned)id[i] * (unsigned)c_in : values;   // mul.wide.u32
                    cp_async16(dst + (uint32_t)i * 32u * kRowBytes, src, have ? 16u : 0u);
                }
                cp_async_commit();
                in_flight++;
                if (++cb == cpb) {
                    cb = 0;
                    slot++;
                    reload = true;
                }
                issue_land += kATileBytes;
                if (issue_land == land_end) issue_land = land_base;
                if (in_flight == landing) {  // the ring is full: the oldest block must land (cp.async.wait_group takes an immediate)
                    switch (landing) {
                        case 1: cp_async_wait<0>(); break;
                        case 2: cp_async_wait<1>(); break;
                        case 3: cp_async_wait<2>(); break;
                        case 4: cp_async_wait<3>(); break;
                        case 5: cp_async_wait<4>(); break;
                        case 6: cp_async_wait<5>(); break;
                        case 7: cp_async_wait<6>(); break;
                        default: cp_async_wait<7>(); break;
                    }
                    publish();
                }
            }
            if (next < n_items) {
                int* dst_ids = nbr_sh + (ibuf ^ 1) * tile_ids;
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    const int i = tid + kTc2Producers * j;
                    if (j < per_thread && i < tile_ids) dst_ids[i] = next_ids[j];
                }
            }
            producer_bar_sync();
            ibuf ^= 1;
        }
        cp_async_wait<0>();
        while (in_flight > 0) publish();
    } else if (warp == kTc2ProducerWarps) {
        // ================= MMA issuer (one lane) =================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(kTileM, n_pad);
            int g = 0, j = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, j++) {
                const Tc2Item w = tc2_item(item, m_tiles, total_kb, kb_per_split);
                const int a = j & 1;
                mbar_wait(acc_empty_bar(a), (((uint32_t)(j >> 1)) & 1u) ^ 1u);   // epilogue has drained this buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(a * n_pad);
                for (int it = 0; it < w.num_kb; it++, g++) {
                    const int sa = g % a_stages, sb = g % b_stages;
                    mbar_wait(full_b(sb), ((uint32_t)(g / b_stages)) & 1u);
                    mbar_wait(full_a(sa), ((uint32_t)(g / a_stages)) & 1u);
                    tc_fence_after();
                    const uint32_t a_addr = a_base + (uint32_t)sa * a_stage_bytes, b_addr = b_base + (uint32_t)sb * b_stage_bytes;
                    const uint64_t da_hi = umma_desc_kmajor_sw128(a_addr);
                    const uint64_t da_lo = umma_desc_kmajor_sw128(a_addr + kATileBytes);
                    const uint64_t db_hi = umma_desc_kmajor_sw128(b_addr);
                    const uint64_t db_lo = umma_desc_kmajor_sw128(b_addr + b_tile_bytes);
#pragma unroll
                    for (int ks = 0; ks < kBlockK / 8; ks++) {   // UMMA K = 8 tf32 = 32 bytes = 2 x 16-byte units
                        const uint64_t adv = (uint64_t)(ks * 2);
                        umma_tf32(d_tmem, da_lo + adv, db_hi + adv, idesc, (it | ks) != 0 ? 1u : 0u);   // small cross terms first
                        umma_tf32(d_tmem, da_hi + adv, db_lo + adv, idesc, 1u);
                        umma_tf32(d_tmem, da_hi + adv, db_hi + adv, idesc, 1u);
                    }
                    umma_commit(empty_a(sa));    // frees the A stage and the B stage when these MMAs have read them
                    umma_commit(empty_b(sb));
                }
                umma_commit(acc_full_bar(a));    // accumulator complete -> epilogue
            }
        }
        __syncwarp();
    } else if (warp == kTc2ProducerWarps + 1) {
        // ================= filter-slab loader (one lane): bulk-async copies, b_stages blocks ahead of the MMAs =================
        if (lane == 0) {
            const size_t b_block = (size_t)n_pad * kBlockK;          // floats of one B slab
            int g = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const Tc2Item w = tc2_item(item, m_tiles, total_kb, kb_per_split);
                const float* src_hi = b_hi + (size_t)w.kb_begin * b_block;
                const float* src_lo = b_lo + (size_t)w.kb_begin * b_block;
                for (int it = 0; it < w.num_kb; it++, g++) {
                    const int sb = g % b_stages;
                    mbar_wait(empty_b(sb), (((uint32_t)(g / b_stages)) & 1u) ^ 1u);
                    const uint32_t dst = b_base + (uint32_t)sb * b_stage_bytes;
                    mbar_arrive_expect_tx(full_b(sb), 2u * b_tile_bytes);
                    bulk_copy_g2s(dst, src_hi, b_tile_bytes, full_b(sb));
                    bulk_copy_g2s(dst + b_tile_bytes, src_lo, b_tile_bytes, full_b(sb));
                    src_hi += b_block;
                    src_lo += b_block;
                }
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue (4 warps; TMEM lane quadrant = warp % 4) =================
        const int quad = warp & 3;
        int j = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, j++) {
            const Tc2Item w = tc2_item(item, m_tiles, total_kb, kb_per_split);
            const int a = j & 1;
            mbar_wait(acc_full_bar(a), ((uint32_t)(j >> 1)) & 1u);
            tc_fence_after();
            const int q = w.q0 + quad * 32 + lane;
            const bool live = q < nv_query;
            float* orow = out + (size_t)q * ld_out;     // `out` / `bias` / `residual` already point at this launch's first channel
            const bool add_bias = bias != nullptr && w.split == 0;
            const float* rrow = (residual != nullptr && w.split == 0) ? residual + (size_t)q * ld_out : nullptr;
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * n_pad);
            for (int n0 = 0; n0 < n_pad; n0 += 32) {
                float acc[32];
                if (n0 + 32 <= n_pad) {
                    tmem_ld32(t_row + (uint32_t)n0, acc);
                } else {                       // n_pad is a multiple of 16: a 16-column tail
                    tmem_ld16(t_row + (uint32_t)n0, acc);
#pragma unroll
                    for (int k = 16; k < 32; k++) acc[k] = 0.0f;
                }
                if (!live) continue;
                if (((c_out | ld_out) & 3) == 0) {      // 16-byte aligned rows and whole float4 groups
#pragma unroll
                    for (int k = 0; k < 32; k += 4) {
                        if (n0 + k < c_out) {
                            float4 o = make_float4(acc[k], acc[k + 1], acc[k + 2], acc[k + 3]);
                            if (add_bias) {
                                o.x += __ldg(bias + n0 + k); o.y += __ldg(bias + n0 + k + 1);
                                o.z += __ldg(bias + n0 + k + 2); o.w += __ldg(bias + n0 + k + 3);
                            }
                            if (rrow != nullptr) {
                                const float4 r4 = __ldg(reinterpret_cast<const float4*>(rrow + n0 + k));
                                o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
                            }
                            if (split_k)
                                atomicAdd(reinterpret_cast<float4*>(orow + n0 + k), o);
                            else
                                *reinterpret_cast<float4*>(orow + n0 + k) = o;
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 32; k++) {
                        if (n0 + k < c_out) {
                            const float o = acc[k] + (add_bias ? __ldg(bias + n0 + k) : 0.0f) + (rrow != nullptr ? __ldg(rrow + n0 + k) : 0.0f);
                            if (split_k)
                                atomicAdd(orow + n0 + k, o);
                            else
                                orow[n0 + k] = o;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty_bar(a));     // this thread has read its rows of the buffer
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kTc2ProducerWarps) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// floats of the prepared slabs of one bank reading (both parts are always reserved, the low half stays unused in
// single-pass TF32 mode, so a buffer survives a precision change)
size_t conv_tc_workspace_bytes(int F, int c_in, int c_out) {
    return (size_t)2 * F * c_in * prep_n_pad_sum(c_out) * sizeof(float);
}

// One launch covers up to kMaxTileN output channels (UMMA N <= 256, two accumulator buffers = all 512 TMEM columns);
// wider layers (the 384- and 512-channel levels of the SemanticKITTI architecture) run as chunks of kMaxTileN columns.
constexpr int kMaxCoutTc = 1024;

// F = 1 with an identity "neighbour" table is a plain row-major GEMM: the 1x1 layers of the bottleneck blocks and of the
// slice head run through the same kernels (lattice_modules.py:806-832 uses torch.nn.Linear there)
bool conv_tc_supported(int F, int c_in, int c_out) {
    return c_in % kBlockK == 0 && c_out >= 1 && c_out <= kMaxCoutTc && F >= 1;
}
// Opt a kernel into the full 227 KB of dynamic shared memory ONCE per (kernel, device), not per launch: the call is
// not free on the host, and an attribute change in the middle of a stream capture trips profilers.
cudaError_t allow_max_smem(const void* kernel) {
    static std::mutex mu;
    static std::set<std::pair<int, const void*>> done;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    std::lock_guard<std::mutex> lock(mu);
    if (done.count({dev, kernel})) return cudaSuccess;
    cudaFuncAttributes attr;
    err = cudaFuncGetAttributes(&attr, kernel);          // static shared memory counts against the same 227 KB
    if (err == cudaSuccess)
        err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)attr.sharedSizeBytes);
    if (err == cudaSuccess)
        done.insert({dev, kernel});
    else
        cudaGetLastError();                               // do not leave a sticky error behind for the next runtime call
    return err;
}
static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

int filter_prepare(const float* filter, int F, int c_in, int c_out, int transposed, int precision, float* slabs, cudaStream_t s) {
    FilterPrepJob job{filter, slabs, F * c_in, c_in, c_out, transposed, precision == 1 ? 1 : 0, 0, 0};
    launch_k(filter_prep_kernel, dim3((unsigned)prep_tiles(job.k_total, c_out)), dim3(1024), 0, s, job);
    count_launch();
    return check_launch("filter_prep");
}

int filter_prepare_batch(const void* jobs_device, int n_jobs, long long total_tiles, cudaStream_t s) {
    if (n_jobs <= 0 || total_tiles <= 0) return LN_OK;
    launch_k(filter_prep_batch_kernel, dim3((unsigned)total_tiles), dim3(1024), 0, s, (const FilterPrepJob*)jobs_device, n_jobs, total_tiles);
    count_launch();
    return check_launch("filter_prep_batch");
}


// ---- weight gradient on the tensor cores -------------------------------------------------------------
//   grad_filter[slot*c_in + ci, co] = sum_q values[nbr[q, slot], ci] * grad_out[q, co]
// per slot a GEMM  D[c_in x c_out] = A^T[c_in x nv] . G[nv x c_out]  whose reduction runs over the VERTICES.
// Both operands sit in shared memory the way they sit in HBM -- one 128-byte row segment (32 channels) per
// vertex -- which is the MN-major UMMA layout (channels contiguous, K = vertices down the rows), so the gather
// needs no transposition.  32-bit MN-major operands must use the SWIZZLE_128B_BASE32B pattern: the four 32-byte
// units of a row are XOR-ed with (row & 3), period 4 rows.  M = 128 channels of c_in (4 groups of 32), N = c_out
// (groups of 32), K = 8 vertices per tcgen05.mma, 32 vertices per pipeline stage.
// Work item = (slot, 128-channel tile of c_in, vertex range); partial sums of different vertex ranges are
// combined with vector fp32 atomics into the pre-zeroed gradient.
constexpr int kWgRows = 32;                          // vertices per stage
constexpr int kWgGroupBytes = kWgRows * kRowBytes;   // one 32-channel group of a stage: 4 KB
constexpr int kWgProducers = 256;                    // 8 lanes per vertex row (one 16-byte chunk each), 32 rows
constexpr int kWgThreads = kWgProducers + 32;

__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128_32b(uint32_t smem_addr) {
    uint64_t desc = 0;
    desc |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address, 16-byte units
    desc |= (uint64_t)(kWgGroupBytes >> 4) << 16;             // leading byte offset: next 32-channel (MN) group
    desc |= (uint64_t)(512 >> 4) << 32;                       // stride byte offset: next 4 vertices (one swizzle period down K)
    desc |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
    desc |= (uint64_t)1 << 61;                                // layout type: SWIZZLE_128B_BASE32B
    return desc;
}
__device__ __forceinline__ uint32_t umma_idesc_tf32_mn(int m, int n) {   // as umma_idesc_tf32, A and B MN-major
    return umma_idesc_tf32(m, n) | (1u << 15) | (1u << 16);
}

template <int kSplit>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_tc_kernel(const float* __restrict__ values, const int* __restrict__ neighbours,
                     const float* __restrict__ grad_out, int nv_query, int F, int c_in, int c_out, int ld_g, int n_pad,
                     int ci_tiles, int q_splits, int chunks_per_split, int stages, int lookahead,
                     float* __restrict__ grad_filter) {
    pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int n_groups = (n_pad + 31) / 32;
    const uint32_t a_bytes = 4u * kWgGroupBytes;                       // 128 channels x 32 vertices = 16 KB
    const uint32_t g_bytes = (uint32_t)n_groups * kWgGroupBytes;
    const uint32_t stage_bytes = (kSplit ? 2 : 1) * (a_bytes + g_bytes);   // [A hi | G hi | A lo | G lo]
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(base + (size_t)stages * stage_bytes);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kMaxStages + 1);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const uint32_t base_u32 = smem_u32(base);
    const uint32_t bars_u32 = smem_u32(bars);
    auto full_bar = [&](int s) { return bars_u32 + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bars_u32 + 8u * (uint32_t)(kMaxStages + s); };
    const uint32_t accum_bar = bars_u32 + 8u * (uint32_t)(2 * kMaxStages);
    const uint32_t lo_off = a_bytes + g_bytes;                          // hi -> lo copy of the same tile

    // work item
    int item = blockIdx.x;
    const int qs = item % q_splits;
    item /= q_splits;
    const int ct = item % ci_tiles;
    const int slot = item / ci_tiles;
    const int ci0 = ct * 128;
    const int ci_n = min(128, c_in - ci0);                             // channels of this tile (multiple of 32)
    const int total_chunks = (nv_query + kWgRows - 1) / kWgRows;
    const int chunk_begin = qs * chunks_per_split;
    const int num_chunks = max(0, min(total_chunks, chunk_begin + chunks_per_split) - chunk_begin);

    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < n_pad) tmem_cols <<= 1;
    if (tid == 0) {
        for (int s = 0; s < stages; s++) {
            mbar_init(full_bar(s), kWgProducers);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp < 8) {
        // ================= producers: one vertex row per thread-octet =================
        const int cc = tid & 7;                   // 16-byte chunk inside a 128-byte group row
        const int r = tid >> 3;                   // row of the stage
        const int a_groups = ci_n / 32;
        // SWIZZLE_128B_BASE32B: 32-byte unit (cc >> 1) goes to unit (cc >> 1) ^ (r & 3); the 16-byte half keeps its place
        const uint32_t my_off = (uint32_t)r * kRowBytes + (uint32_t)((((cc >> 1) ^ (r & 3)) << 5) | ((cc & 1) << 4));
        const uint32_t ring_end = base_u32 + (uint32_t)stages * stage_bytes;
        uint32_t issue_addr = base_u32, issue_full = full_bar(0), issue_empty = empty_bar(0), issue_phase = 0;
        uint32_t pub_addr = base_u32, pub_full = full_bar(0);
        int in_flight = 0;
        auto publish = [&]() {
            if (kSplit) {
                const uint32_t hi0 = pub_addr + my_off;
                for (int grp = 0; grp < 4 + n_groups; grp++) {              // A groups then G groups are contiguous 4 KB blocks
                    if (grp >= a_groups && grp < 4) continue;               // channels past c_in: never gathered, never stored
                    const uint32_t hi = hi0 + (uint32_t)grp * kWgGroupBytes;
                    const float4 x = lds128(hi);
                    const float4 hv = make_float4(to_tf32(x.x), to_tf32(x.y), to_tf32(x.z), to_tf32(x.w));
                    sts128(hi, hv);
                    sts128(hi + lo_off, make_float4(to_tf32(x.x - hv.x), to_tf32(x.y - hv.y), to_tf32(x.z - hv.z), to_tf32(x.w - hv.w)));
                }
            }
            fence_proxy_async();
            mbar_arrive(pub_full);
            pub_addr += stage_bytes;
            pub_full += 8;
            if (pub_addr == ring_end) {
                pub_addr = base_u32;
                pub_full = full_bar(0);
            }
            in_flight--;
        };
        int q = chunk_begin * kWgRows + r;
        const int* nbr_p = neighbours + (size_t)q * F + slot;
        const float* g_p = grad_out + (size_t)q * ld_g + cc * 4;     // ld_g: full width of grad_out / grad_filter rows (N chunking)
        const float* a_col = values + ci0 + cc * 4;
        // the neighbour id of a row is needed before its copies can be issued: it is requested one chunk ahead, so
        // its latency overlaps the wait for a free stage instead of serialising every iteration of this loop
        // (ncu r01g: 8 % L2 throughput, 17 % tensor pipe, one dependent L2 round trip per 32 vertices)
        int id_next = (num_chunks > 0 && q < nv_query) ? __ldg(nbr_p) : -1;
        for (int g = 0; g < num_chunks; g++) {
            const bool in_range = q < nv_query;
            const int id = id_next;
            id_next = (g + 1 < num_chunks && q + kWgRows < nv_query) ? __ldg(nbr_p + (size_t)kWgRows * F) : -1;
            mbar_wait(issue_empty, issue_phase ^ 1u);
            const bool have = id >= 0;
            const float* arow = have ? a_col + (unsigned long long)(unsigned)id * (unsigned)c_in : values;
            const uint32_t dst = issue_addr + my_off;
            for (int grp = 0; grp < a_groups; grp++)
                cp_async16(dst + (uint32_t)grp * kWgGroupBytes, have ? arow + grp * 32 : values, have ? 16u : 0u);
            for (int grp = 0; grp < n_groups; grp++) {
                const bool ok = in_range && grp * 32 + cc * 4 < c_out;      // c_out % 4 == 0: whole chunks only
                cp_async16(dst + a_bytes + (uint32_t)grp * kWgGroupBytes, ok ? g_p + grp * 32 : grad_out, ok ? 16u : 0u);
            }
            cp_async_commit();
            in_flight++;
            q += kWgRows;
            nbr_p += (size_t)kWgRows * F;
            g_p += (size_t)kWgRows * ld_g;
            issue_addr += stage_bytes;
            issue_full += 8;
            issue_empty += 8;
            if (issue_addr == ring_end) {
                issue_addr = base_u32;
                issue_full = full_bar(0);
                issue_empty = empty_bar(0);
                issue_phase ^= 1u;
            }
            if (in_flight > lookahead) {
                switch (lookahead) {
                    case 1: cp_async_wait<1>(); break;
                    case 2: cp_async_wait<2>(); break;
                    case 3: cp_async_wait<3>(); break;
                    case 4: cp_async_wait<4>(); break;
                    default: cp_async_wait<5>(); break;
                }
                publish();
            }
        }
        cp_async_wait<0>();
        while (in_flight > 0) publish();

        // ================= epilogue (warps 0-3): TMEM lane = channel of the tile, columns = c_out =================
        if (warp < 4 && num_chunks > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
            const int m = warp * 32 + lane;
            const bool live = m < ci_n;
            float* orow = grad_filter + ((size_t)slot * c_in + ci0 + m) * ld_g;
            for (int n0 = 0; n0 < n_pad; n0 += 16) {
                float acc[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, acc);
                if (!live) continue;
#pragma unroll
                for (int k = 0; k < 16; k += 4) {
                    if (n0 + k < c_out) {
                        const float4 o = make_float4(acc[k], acc[k + 1], acc[k + 2], acc[k + 3]);
                        if (q_splits > 1)
                            atomicAdd(reinterpret_cast<float4*>(orow + n0 + k), o);
                        else
                            *reinterpret_cast<float4*>(orow + n0 + k) = o;
                    }
                }
            }
            tc_fence_before();
        }
    } else {
        // ================= MMA issuer =================
        if (lane == 0 && num_chunks > 0) {
            const uint32_t idesc = umma_idesc_tf32_mn(128, n_pad);
            uint32_t addr = base_u32, full = full_bar(0), empty = empty_bar(0), phase = 0;
            for (int g = 0; g < num_chunks; g++) {
                mbar_wait(full, phase);
                tc_fence_after();
                const uint64_t da_hi = umma_desc_mnmajor_sw128_32b(addr);
                const uint64_t dg_hi = umma_desc_mnmajor_sw128_32b(addr + a_bytes);
                const uint64_t da_lo = umma_desc_mnmajor_sw128_32b(addr + lo_off);
                const uint64_t dg_lo = umma_desc_mnmajor_sw128_32b(addr + lo_off + a_bytes);
#pragma unroll
                for (int ks = 0; ks < kWgRows / 8; ks++) {       // 8 vertices per MMA = two 512-byte swizzle periods down the rows
                    const uint64_t adv = (uint64_t)(ks * (1024 >> 4));
                    if (kSplit) {
                        umma_tf32(tmem_base, da_lo + adv, dg_hi + adv, idesc, (g | ks) != 0 ? 1u : 0u);
                        umma_tf32(tmem_base, da_hi + adv, dg_lo + adv, idesc, 1u);
                        umma_tf32(tmem_base, da_hi + adv, dg_hi + adv, idesc, 1u);
                    } else {
                        umma_tf32(tmem_base, da_hi + adv, dg_hi + adv, idesc, (g | ks) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(empty);
                addr += stage_bytes;
                full += 8;
                empty += 8;
                if (addr == base_u32 + (uint32_t)stages * stage_bytes) {
                    addr = base_u32;
                    full = full_bar(0);
                    empty = empty_bar(0);
                    phase ^= 1u;
                }
            }
            umma_commit(accum_bar);
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

bool conv_wgrad_tc_supported(int F, int c_in, int c_out) {
    return c_in % 32 == 0 && c_out % 4 == 0 && c_out >= 4 && c_out <= kMaxCoutTc && F >= 1;
}

// K blocks in flight per producer thread = half the ring: a stage published `stages - lookahead` iterations ago has had
// that long for its MMAs to retire before the producer needs it back (lookahead = stages - 1 serialises the two).
// A CTA whose whole K range fits the ring (the split-K regime of small lattices) issues everything up front.
static inline int lookahead_for(int stages, int blocks_per_cta) {
    if (blocks_per_cta <= stages) return max(1, stages - 1);
    return max(1, min(stages - 1, stages / 2));
}

// Vertex-range splits of the weight gradient: enough CTAs for the machine (two waves at most), at least 4 stages of
// vertices per CTA.
static int wgrad_q_splits(int nv_query, int F, int c_in, int cta_budget = 0) {
    const int tiles = F * cdiv(c_in, 128);
    const int total_chunks = cdiv(nv_query, kWgRows);
    const int ctas = cta_budget > 0 ? cta_budget : 2 * sm_count();
    int q_splits = max(1, min(cdiv(total_chunks, 4), ctas / tiles));
    const int chunks_per_split = cdiv(total_chunks, q_splits);
    return cdiv(total_chunks, chunks_per_split);
}
bool conv_wgrad_tc_needs_zero(int nv_query, int F, int c_in) { return wgrad_q_splits(nv_query, F, c_in) > 1; }

// grad_filter must be zero when conv_wgrad_tc_needs_zero() (partial sums of vertex ranges are combined with vector atomics).
static int conv_wgrad_tc_chunk(const float* nbr_values, const int* neighbours, const float* grad_out, int nv_query, int F, int c_in,
                               int c_out, int ld_g, int precision, float* grad_filter, int cta_budget, cudaStream_t s) {
    const int split = precision == 1 ? 1 : 0;
    const int n_pad = (c_out + 15) / 16 * 16;
    const int n_groups = (n_pad + 31) / 32;
    const int ci_tiles = cdiv(c_in, 128);
    const int total_chunks = cdiv(nv_query, kWgRows);
    const int tiles = F * ci_tiles;
    const int q_splits = wgrad_q_splits(nv_query, F, c_in, cta_budget);
    const int chunks_per_split = cdiv(total_chunks, q_splits);
    const size_t stage_bytes = (size_t)(split ? 2 : 1) * (4 + n_groups) * kWgGroupBytes;
    const size_t fixed = (2 * kMaxStages + 1) * 8 + 16 + 1024;
    const int stages = (int)min((size_t)kMaxStages, (227 * 1024 - fixed) / stage_bytes);
    if (stages < 2) {
        set_error("conv_wgrad_tc: tile does not fit shared memory (c_out=%d)", c_out);
        return LN_ERR_UNSUPPORTED;
    }
    const int lookahead = lookahead_for(stages, chunks_per_split);
    // > half of the SM's shared memory: one CTA per SM (TMEM columns, see conv_tc2)
    const size_t smem = max((size_t)stages * stage_bytes + fixed, (size_t)120 * 1024);
    const int grid = tiles * q_splits;
    cudaError_t err;
    if (split) {
        err = allow_max_smem((const void*)conv_wgrad_tc_kernel<1>);
        if (err == cudaSuccess)
            launch_k(conv_wgrad_tc_kernel<1>, dim3(grid), dim3(kWgThreads), smem, s, nbr_values, neighbours, grad_out, nv_query, F, c_in, c_out, ld_g, n_pad, ci_tiles,
                                                                    q_splits, chunks_per_split, stages, lookahead, grad_filter);
    } else {
        err = allow_max_smem((const void*)conv_wgrad_tc_kernel<0>);
        if (err == cudaSuccess)
            launch_k(conv_wgrad_tc_kernel<0>, dim3(grid), dim3(kWgThreads), smem, s, nbr_values, neighbours, grad_out, nv_query, F, c_in, c_out, ld_g, n_pad, ci_tiles,
                                                                    q_splits, chunks_per_split, stages, lookahead, grad_filter);
    }
    if (err != cudaSuccess) {
        set_error("conv_wgrad_tc: %s", cudaGetErrorString(err));
        return LN_ERR_CUDA;
    }
    count_launch();
    return check_launch("conv_wgrad_tc");
}

// cta_budget (0 = the whole machine): upper bound on the CTAs of one launch, so that a data-gradient convolution running on
// another stream at the same time finds free SMs (every CTA of either kernel needs an SM to itself)
int conv_wgrad_tc(const float* nbr_values, const int* neighbours, const float* grad_out, int nv_query, int F, int c_in,
                  int c_out, int precision, float* grad_filter, int cta_budget, cudaStream_t s) {
    // column chunks of grad_out / grad_filter (row stride c_out): each chunk is an independent GEMM over the same gathered A
    for (int n_off = 0; n_off < c_out; n_off += kMaxTileN) {
        const int rc = conv_wgrad_tc_chunk(nbr_values, neighbours, grad_out + n_off, nv_query, F, c_in, min(kMaxTileN, c_out - n_off),
                                           c_out, precision, grad_filter + n_off, cta_budget, s);
        if (rc != LN_OK) return rc;
    }
    return LN_OK;
}

// K splits of the convolution: enough CTAs for the machine when there are few M tiles (148 SMs, one CTA each)
static int conv_k_splits(int nv_query, int F, int c_in, int cta_budget = 0) {
    const int num_kb = F * (c_in / kBlockK);
    const int m_tiles = cdiv(nv_query, kTileM);
    const int ctas = cta_budget > 0 ? min(cta_budget, sm_count()) : sm_count();
    int splits = 1;
    if (m_tiles < ctas) splits = max(1, min(num_kb, ctas / m_tiles));
    const int kb_per_split = cdiv(num_kb, splits);
    return cdiv(num_kb, kb_per_split);
}
// partial tiles of different K ranges are combined with vector atomics: `out` must be zero beforehand
bool conv_tc_needs_zero(int nv_query, int F, int c_in) { return conv_k_splits(nv_query, F, c_in) > 1; }

// One launch for output channels [n_off, n_off + c_out) of a layer that is ld_n wide.  `out` / `bias` / `residual` point at
// the layer's first channel; `slabs` at this chunk's prepared filter.
static int conv_fwd_tc_chunk(const float* nbr_values, const int* neighbours, const float* slabs, const float* bias, const float* residual,
                             int nv_query, int F, int c_in, int c_out, int ld_n, int n_off, int flip, int precision, float* out,
                             int cta_budget, cudaStream_t s) {
    const int n_pad = (c_out + 15) / 16 * 16;
    const int k_total = F * c_in;
    const int split = precision == 1 ? 1 : 0;
    const float* b_hi = slabs;
    const float* b_lo = slabs + (size_t)k_total * n_pad;
    const int num_kb = F * (c_in / kBlockK);
    const int m_tiles = cdiv(nv_query, kTileM);
    const int splits = conv_k_splits(nv_query, F, c_in, cta_budget);
    const int kb_per_split = cdiv(num_kb, splits);
    const size_t b_tile = (size_t)n_pad * kRowBytes;
    const size_t stage_bytes = (split ? 2 : 1) * (kATileBytes + b_tile);
    float* out_chunk = out + n_off;
    const float* bias_chunk = bias != nullptr ? bias + n_off : nullptr;
    const float* res_chunk = residual != nullptr ? residual + n_off : nullptr;
    // persistent kernel: one CTA per SM, items = (M tile, K split)
    const int n_items = m_tiles * splits;
    const int grid = min(n_items, cta_budget > 0 ? min(cta_budget, sm_count()) : sm_count());
    cudaError_t err;
    if (split) {
        // 3xTF32: decoupled landing / A / B rings (conv_tc3)
        const size_t fixed3 = (size_t)2 * kTileM * F * sizeof(int) + 16 + kTc3Bars * 8 + 16 + 1024;
        const size_t b_pair = 2 * b_tile, a_pair = 2 * kATileBytes;
        const size_t budget = 227 * 1024 - fixed3;
        // ring depths: the A ring hides the publish -> MMA hand-over latency (3 deep when it fits), the B ring runs 2..4 slabs
        // ahead, what is left becomes landing slots (gathers in flight)
        int a_stages = 3, b_stages = b_pair <= 8 * 1024 ? 4 : b_pair <= 16 * 1024 ? 3 : 2;
        if (budget < a_stages * a_pair + b_stages * b_pair + 2 * kATileBytes) b_stages = 2;
        if (budget < a_stages * a_pair + b_stages * b_pair + 2 * kATileBytes) a_stages = 2;
        if (budget < a_stages * a_pair + b_stages * b_pair + kATileBytes) {
            set_error("ln_conv_fwd: tensor-core tile does not fit shared memory (c_out=%d)", c_out);
            return LN_ERR_UNSUPPORTED;
        }
        int landing = (int)min((size_t)kTc3MaxLanding, (budget - a_stages * a_pair - b_stages * b_pair) / kATileBytes);
        landing = max(1, min(landing, cdiv(n_items, grid) * kb_per_split));      // never deeper than the work of a CTA
        const size_t smem = max(fixed3 + a_stages * a_pair + b_stages * b_pair + (size_t)landing * kATileBytes, (size_t)120 * 1024);
        err = allow_max_smem((const void*)conv_tc3_kernel);
        if (err == cudaSuccess)
            launch_k(conv_tc3_kernel, dim3(grid), dim3(kTc3Threads), smem, s, nbr_values, neighbours, b_hi, b_lo, bias_chunk, res_chunk, nv_query, F,
                     c_in, c_out, ld_n, n_pad, flip, landing, a_stages, b_stages, m_tiles, n_items, kb_per_split, out_chunk);
        if (err != cudaSuccess) {
            set_error("conv_tc3: %s", cudaGetErrorString(err));
            return LN_ERR_CUDA;
        }
        count_launch();
        return check_launch("conv_tc3");
    }
    const size_t fixed = (size_t)2 * kTileM * F * sizeof(int) + 16 + (2 * kMaxStages + 4) * 8 + 16 + 1024;
    const int stages = min((int)((227 * 1024 - fixed) / stage_bytes), kMaxStages);
    if (stages < 2) {
        set_error("ln_conv_fwd: tensor-core tile does not fit shared memory (c_out=%d)", c_out);
        return LN_ERR_UNSUPPORTED;
    }
    const int lookahead = lookahead_for(stages, cdiv(n_items, grid) * kb_per_split);
    // > half of the SM's shared memory: exactly one CTA per SM, so the 2*n_pad TMEM columns are always available
    const size_t smem = max((size_t)stages * stage_bytes + fixed, (size_t)120 * 1024);
    err = allow_max_smem((const void*)conv_tc2_kernel<0>);
    if (err == cudaSuccess)
        launch_k(conv_tc2_kernel<0>, dim3(grid), dim3(kTc2Threads), smem, s, nbr_values, neighbours, b_hi, b_lo, bias_chunk, res_chunk, nv_query, F, c_in, c_out, ld_n,
                 n_pad, flip, stages, lookahead, m_tiles, n_items, kb_per_split, out_chunk);
    if (err != cudaSuccess) {
        set_error("conv_tc2: %s", cudaGetErrorString(err));
        return LN_ERR_CUDA;
    }
    count_launch();
    return check_launch("conv_tc2");
}

// slabs: prepared filter of this reading (filter_prepare / filter_prepare_batch).  `out` must be zero when
// conv_tc_needs_zero() (the caller clears it or hands over a zeroed buffer).
int conv_fwd_tc(const float* nbr_values, const int* neighbours, const float* slabs, const float* bias, const float* residual, int nv_query,
                int F, int c_in, int c_out, int flip, int precision, float* out, int cta_budget, cudaStream_t s) {
    for (int n_off = 0; n_off < c_out; n_off += kMaxTileN) {
        const int rc = conv_fwd_tc_chunk(nbr_values, neighbours, slabs + (size_t)2 * F * c_in * n_off, bias, residual, nv_query, F, c_in,
                                         min(kMaxTileN, c_out - n_off), c_out, n_off, flip, precision, out, cta_budget, s);
        if (rc != LN_OK) return rc;
    }
    return LN_OK;
}

}  // namespace ln
