// Shared device helpers for the B200 lattice kernels: error plumbing, permutohedral
// geometry (bit-compatible with the reference's fast-math build), and the open-addressing
// vertex table.  Compiled for sm_100a with -ftz=true (NOT --use_fast_math): every
// floating-point operation whose rounding matters for lattice keys is an explicit intrinsic.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include "../../include/lattice_b200.h"

namespace ln {

// ---------------------------------------------------------------------------------------------
// host-side error / launch bookkeeping (defined in ln_api.cu)
void set_error(const char* fmt, ...);
int check_launch(const char* what);
void count_launch(int n = 1);
cudaError_t allow_max_smem(const void* kernel);   // ln_conv_tc.cu: opt a kernel into 227 KB dynamic smem, once per device
// rows[idx[p,r], :] += src[p, :] * w[p,r]  (ln_slice.cu; shared by ln_slice_bwd and ln_splat_accumulate)
int launch_scatter_rows(const float* src, const int* indices, const float* weights, int n, int pos_dim, int val_dim,
                        int nr_vertices, float* rows, cudaStream_t s, const char* what);

#define LN_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            ::ln::set_error(__VA_ARGS__);     \
            return LN_ERR_BAD_ARG;            \
        }                                     \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch.  A ShapeNet-sized training step is a chain of ~450 kernels of a few microseconds each,
// so the launch latency between two dependent kernels is a first-order cost.  Every kernel of this library signals
// `launch_dependents` as its first instruction and waits for its predecessors (`griddepcontrol.wait`: all prerequisite
// grids complete, their memory visible) before its first global-memory access; launched through launch_k() with the
// programmatic-stream-serialization attribute, the NEXT kernel's launch, CTA scheduling and on-chip prologue (barrier
// initialisation, TMEM allocation, weight staging in shared memory that does not depend on earlier kernels) overlap the
// tail of the current one.  Inside a stream capture the attribute becomes a programmatic graph edge.
// Rules kept by every kernel: (1) no global read or write before pdl_wait(); (2) every thread reaches pdl_wait() (no
// early return before it), so a grid can never complete before its predecessor has.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#define LN_PDL_ENTRY()       \
    do {                     \
        ::ln::pdl_trigger(); \
        ::ln::pdl_wait();    \
    } while (0)

bool pdl_enabled();   // ln_api.cu; ln_set_programmatic_launch()

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

constexpr int kEmpty = -1;   // entries[] states, HashTableGPU.cuh:24-26
constexpr int kLocked = -2;

// ---------------------------------------------------------------------------------------------
// memory-ordering helpers: table state lives in L2 (the coherence point); never trust L1 for it.
__device__ __forceinline__ int ld_relaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// entries[] loads that are FOLLOWED by a read of keys[id]: acquire pairs with the writer's st.release, so the key
// words are ordered after the id by the memory model and not merely by the address dependency
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Hash of a lattice key: HashTableGPU::hash (HashTableGPU.cuh:35-50), first D coordinates only.
template <int D>
__device__ __forceinline__ uint32_t key_hash(const int* key) {
    uint32_t k = 0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        k += (uint32_t)key[i];
        k *= 2531011u;
    }
    return k;
}

struct TableView {   // device view of one vertex table (HashTableGPU.cuh:23-28)
    int* keys;
    int* entries;
    int* nr_filled;
    int* status;     // [0] overflow flags (bit 0 table full, bit 1 vertex bound exceeded), [1] max probe length seen by inserts
    int capacity;
    int max_vertices;   // ids >= max_vertices are reported through status[0] bit 1 and handed out as -1 by the callers
};
struct ConstTableView {
    const int* keys;
    const int* entries;
    int capacity;
};

template <int D>
__device__ __forceinline__ bool key_equal_at(const int* keys, int id, const int* key) {
    // vertex keys are written once and published with a release store on entries[]; read them
    // at L2 (ld.cg) so a stale L1 line can never produce a false mismatch (-> duplicate vertex).
    bool same = true;
#pragma unroll
    for (int i = 0; i < D; i++) same &= (__ldcg(keys + (size_t)id * D + i) == key[i]);
    return same;
}

// Insert-or-find.  Same observable contract as HashTableGPU::insert (HashTableGPU.cuh:425-484):
// returns the compact vertex id of `key`, allocating the next id if the key is new.  Differences
// by design: (1) a plain L2 load precedes the CAS, so the common "already present" case costs no
// atomic; (2) probing is bounded by the capacity and overflow is reported through status[0]
// instead of spinning forever; (3) returns the id, not the slot.
template <int D>
__device__ __forceinline__ int table_insert(const TableView& t, const int* key, uint32_t hash) {
    int h = (int)(hash % (uint32_t)t.capacity);
    for (int probe = 0; probe < t.capacity; probe++) {
        int* e = t.entries + h;
        int cur = ld_acquire(e);
        if (cur == kEmpty) {
            cur = atomicCAS(e, kEmpty, kLocked);
            if (cur == kEmpty) {   // we own the slot: allocate the vertex, publish its key
                // the vertex counter is ONE address for the whole grid: lanes of this warp that won a slot in the
                // same probe step share a single atomic (warp-aggregated allocation over the coalesced group)
                const cooperative_groups::coalesced_group winners = cooperative_groups::coalesced_threads();
                int base = 0;
                if (winners.thread_rank() == 0) base = atomicAdd(t.nr_filled, (int)winners.size());
                base = winners.shfl(base, 0);
                const int id = base + (int)winners.thread_rank();
                if (id >= t.max_vertices) atomicOr(t.status, 2);   // caller's row bound exceeded (static-shape mode)
#pragma unroll
                for (int i = 0; i < D; i++) t.keys[(size_t)id * D + i] = key[i];
                st_release(e, id);   // release at gpu scope: the key stores above are visible before the id is
                                     // (a separate __threadfence() here only doubled the membar stall, ncu r01g)
                if (probe > 0) atomicMax(t.status + 1, probe);
                return id;
            }
        }
        while (cur == kLocked) cur = ld_acquire(e);   // another thread is publishing; short wait
        if (key_equal_at<D>(t.keys, cur, key)) return cur;
        h = (h + 1 == t.capacity) ? 0 : h + 1;   // linear probing
    }
    atomicOr(t.status, 1);   // table full
    return -1;
}

// Find only.  HashTableGPU::retrieve (HashTableGPU.cuh:491-519) without the 300-probe cap
// (the cap never triggers below ~0.7 load).
template <int D>
__device__ __forceinline__ int table_find(const ConstTableView& t, const int* key) {
    int h = (int)(key_hash<D>(key) % (uint32_t)t.capacity);
    for (int probe = 0; probe < t.capacity; probe++) {
        const int cur = __ldg(t.entries + h);
        if (cur < 0) return -1;   // empty (tables are read-only here, so never locked)
        bool same = true;
#pragma unroll
        for (int i = 0; i < D; i++) same &= (__ldg(t.keys + (size_t)cur * D + i) == key[i]);
        if (same) return cur;
        h = (h + 1 == t.capacity) ? 0 : h + 1;
    }
    return -1;
}

// ---------------------------------------------------------------------------------------------
// Permutohedral geometry.  The reference compiles its kernels with NVRTC --use_fast_math
// (jitify_helper.cuh:29) and the PTX is then lowered by ptxas / the driver JIT.  What actually runs
// (SASS of kernel_splat / distribute / slice_no_precomputation, identical in all of them; see
// DESIGN.md "bit-exact keys") for LatticeGPU.cuh:718-741 is
//     scale_i  : a CONSTANT -- ptxas folds rsqrt.approx.ftz((i+1)(i+2)) with the correctly rounded
//                1/sqrt and multiplies by fl((D+1)*sqrtf(2/3)); it is not the hardware MUFU.RSQ result
//     cf = fl(p*scale);  e_i = fma(cf, -i, sm)            for i >= 3
//                        e_2 = sm - fma(p, scale, cf)
//     e_1 = fma(p_0, -scale_0, sm), e_0 = fma(p_0, scale_0, sm)   (mul+add contracted by ptxas)
// and everything after it is plain fp32 (fp64 for the two (D+1)-divisions when D != 3).
template <int D>
__device__ __forceinline__ float elevate_scale(int i) {   // scale applied to position coordinate i
    static_assert(D == 3 || D == 5, "pos_dim must be 3 or 5");
    if (D == 3) return __int_as_float(i == 0 ? 0x4013cd3a : i == 1 ? 0x3faaaaab : 0x3f715bef);
    return __int_as_float(i == 0 ? 0x405db3d8 : i == 1 ? 0x40000001 : i == 2 ? 0x3fb504f3 : i == 3 ? 0x3f8c378c : 0x3f64f92e);
}

template <int D>
struct Simplex {
    int rem0[D + 1];
    int rank[D + 1];
    float bary[D + 2];
};

// p: position already divided by sigma
template <int D>
__device__ __forceinline__ void compute_simplex(const float* p, Simplex<D>& s) {
    float e[D + 1];
    float sm = 0.0f;
#pragma unroll
    for (int i = D; i > 0; i--) {
        const float scale = elevate_scale<D>(i - 1);
        if (i == 1) {
            e[1] = __fmaf_rn(p[0], -scale, sm);
            e[0] = __fmaf_rn(p[0], scale, sm);
        } else {
            const float cf = __fmul_rn(p[i - 1], scale);
            e[i] = (i >= 3) ? __fmaf_rn(cf, -(float)i, sm) : __fsub_rn(sm, __fmaf_rn(p[i - 1], scale, cf));
            sm = __fadd_rn(sm, cf);
        }
    }

    // nearest remainder-0 point (LatticeGPU.cuh:746-758)
    int sum = 0;
#pragma unroll
    for (int i = 0; i <= D; i++) {
        const float v = (D == 3) ? __fmul_rn(e[i], 0.25f) : (float)__dmul_rn((double)e[i], 1.0 / (D + 1));
        const float up = __fmul_rn(ceilf(v), (float)(D + 1));
        const float down = __fmul_rn(floorf(v), (float)(D + 1));
        s.rem0[i] = (__fsub_rn(up, e[i]) < __fsub_rn(e[i], down)) ? (int)up : (int)down;
        sum += s.rem0[i];
    }
    sum /= (D + 1);

    // rank of each coordinate's residual (LatticeGPU.cuh:762-772); ties: later index gets +1
    float diff[D + 1];
#pragma unroll
    for (int i = 0; i <= D; i++) {
        diff[i] = __fsub_rn(e[i], (float)s.rem0[i]);
        s.rank[i] = 0;
    }
#pragma unroll
    for (int i = 0; i < D; i++) {
#pragma unroll
        for (int j = i + 1; j <= D; j++) {
            if (diff[i] < diff[j])
                s.rank[i]++;
            else
                s.rank[j]++;
        }
    }
    // bring the point back onto the plane (LatticeGPU.cuh:775-783)
#pragma unroll
    for (int i = 0; i <= D; i++) {
        s.rank[i] += sum;
        if (s.rank[i] < 0) {
            s.rank[i] += D + 1;
            s.rem0[i] += D + 1;
        } else if (s.rank[i] > D) {
            s.rank[i] -= D + 1;
            s.rem0[i] -= D + 1;
        }
    }
    // barycentric coordinates (LatticeGPU.cuh:787-795).  Fully unrolled selects keep bary[] in
    // registers (the reference indexes a local array dynamically); per-element operation order
    // is preserved, so the sums round identically.
#pragma unroll
    for (int k = 0; k <= D + 1; k++) s.bary[k] = 0.0f;
#pragma unroll
    for (int i = 0; i <= D; i++) {
        const float rem0f = (float)s.rem0[i];
        const float d0 = __fsub_rn(e[i], rem0f);
        const float delta = (D == 3) ? __fmul_rn(d0, 0.25f) : (float)__dmul_rn((double)d0, 1.0 / (D + 1));
#pragma unroll
        for (int k = 0; k <= D + 1; k++) {
            if (k == D - s.rank[i]) s.bary[k] = __fadd_rn(s.bary[k], delta);
            if (k == D + 1 - s.rank[i]) s.bary[k] = __fsub_rn(s.bary[k], delta);
        }
    }
    s.bary[0] = (float)__dadd_rn(__dadd_rn((double)s.bary[D + 1], 1.0), (double)s.bary[0]);
}

// key of simplex vertex `r` (first D coordinates; LatticeGPU.cuh:799-806)
template <int D>
__device__ __forceinline__ void simplex_key(const Simplex<D>& s, int r, int* key) {
#pragma unroll
    for (int i = 0; i < D; i++) {
        key[i] = s.rem0[i] + r;
        if (s.rank[i] > D - r) key[i] -= (D + 1);
    }
}

// positions_raw / sigma with IEEE division == torch's `positions_raw / sigmas_tensor`
// (/root/reference/src/Lattice.cu:226), fused into the kernels instead of a separate launch.
template <int D>
__device__ __forceinline__ void load_scaled_position(const float* __restrict__ positions_raw,
                                                     const float* __restrict__ sigmas, int idx, float* p) {
#pragma unroll
    for (int i = 0; i < D; i++) p[i] = __fdiv_rn(__ldg(positions_raw + (size_t)idx * D + i), __ldg(sigmas + i));
}

}  // namespace ln
