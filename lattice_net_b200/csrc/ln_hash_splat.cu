// Lattice construction: table clear, splat-build (simplex + warp-deduplicated insertion),
// distribute, simplex lookup, value accumulation and key coarsening.
//
// Reference semantics: kernel_splat / distribute / splatCacheNaive / coarsen in
// /root/reference/include/lattice_net/kernels/LatticeGPU.cuh:707-842, 534-650, 926-973, 2314-2514.
#include "ln_common.cuh"

namespace ln {

constexpr int kBlock = 256;
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) table_clear_kernel(int4* __restrict__ entries4, int* __restrict__ entries,
                                                             int* nr_filled, int* status, int capacity) {
    LN_PDL_ENTRY();
    const int n4 = capacity >> 2;
    const int stride = gridDim.x * blockDim.x;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = tid; i < n4; i += stride) entries4[i] = make_int4(kEmpty, kEmpty, kEmpty, kEmpty);
    for (int i = (n4 << 2) + tid; i < capacity; i += stride) entries[i] = kEmpty;
    if (tid == 0) {
        *nr_filled = 0;
        status[0] = 0;
        status[1] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// One thread per point, two phases (ncu on the 1M-point sweep: the old kernel, which ran the D+1 insert-or-find chains
// of a thread one after the other, was pure dependent-latency -- stall_long_scoreboard + stall_membar = 88 % of all
// stalls at 19 % issue utilisation):
//   phase 1  LOOK-UP, all D+1 keys of the thread in flight together: load the D+1 home slots, then the D+1 stored
//            keys, compare.  Every vertex is shared by ~9 simplices, so once the table warms up most keys resolve
//            here with two overlapped L2 round trips for the whole simplex and no atomic.
//   phase 2  INSERT for what is left (empty / locked / colliding home slot).  The D+1 simplex vertices of spatially
//            close points coincide most of the time, so each warp first groups equal keys with __match_any_sync and
//            only the group leader probes / inserts; the vertex id is broadcast back with a shuffle.
template <int D, bool kDistribute, bool kInsert>
__global__ void __launch_bounds__(kBlock)
splat_build_kernel(const float* __restrict__ positions_raw, const float* __restrict__ sigmas,
                   const float* __restrict__ values, int n, int val_dim, TableView table,
                   int* __restrict__ indices, float* __restrict__ weights, float* __restrict__ distributed) {
    LN_PDL_ENTRY();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = idx < n;

    Simplex<D> s;
    float p[D];
    if (valid) {
        load_scaled_position<D>(positions_raw, sigmas, idx, p);
        compute_simplex<D>(p, s);
    }

    int ids[D + 1];
    if (kInsert) {
        int key[D + 1][D];
        uint32_t hash[D + 1];
        int cur[D + 1];
        // ---- phase 1 ----
#pragma unroll
        for (int r = 0; r <= D; r++) {
#pragma unroll
            for (int i = 0; i < D; i++) key[r][i] = 0;
            hash[r] = 0;
            cur[r] = kEmpty;
            ids[r] = -1;
            if (valid) {
                simplex_key<D>(s, r, key[r]);
                hash[r] = key_hash<D>(key[r]);
                cur[r] = ld_acquire(table.entries + (int)(hash[r] % (uint32_t)table.capacity));
            }
        }
        int stored[D + 1][D];
#pragma unroll
        for (int r = 0; r <= D; r++) {
#pragma unroll
            for (int i = 0; i < D; i++) stored[r][i] = (cur[r] >= 0) ? __ldcg(table.keys + (size_t)cur[r] * D + i) : 0;
        }
        bool pending[D + 1];
#pragma unroll
        for (int r = 0; r <= D; r++) {
            bool same = cur[r] >= 0;
#pragma unroll
            for (int i = 0; i < D; i++) same &= (stored[r][i] == key[r][i]);
            if (same) ids[r] = cur[r];
            pending[r] = valid && !same;
        }
        // ---- phase 2 ----
#pragma unroll
        for (int r = 0; r <= D; r++) {
            if (!__any_sync(kFull, pending[r])) continue;               // warp-uniform
            const bool need = pending[r];
            // group by hash; lanes with nothing left to insert are masked out of every group
            const unsigned group = __match_any_sync(kFull, hash[r]) & __ballot_sync(kFull, need);
            const int leader = need ? (__ffs(group) - 1) : lane;
            bool same = need;   // same key as the group leader (hash collisions are possible)
#pragma unroll
            for (int i = 0; i < D; i++) same &= (__shfl_sync(kFull, key[r][i], leader) == key[r][i]);
            int id = -1;
            if (need && (lane == leader || !same)) id = table_insert<D>(table, key[r], hash[r]);
            const int leader_id = __shfl_sync(kFull, id, leader);
            if (need) ids[r] = same ? leader_id : id;
        }
    } else {
#pragma unroll
        for (int r = 0; r <= D; r++) {
            int key[D];
            int id = -1;
            if (valid) {
                simplex_key<D>(s, r, key);
                ConstTableView ct{table.keys, table.entries, table.capacity};
                id = table_find<D>(ct, key);
            }
            ids[r] = id;
        }
    }
#pragma unroll
    for (int r = 0; r <= D; r++)
        if (ids[r] >= table.max_vertices) ids[r] = -1;   // beyond the caller's row bound: dropped + flagged
    if (!valid) return;

    if (indices != nullptr) {
        int* irow = indices + (size_t)idx * (D + 1);
        float* wrow = weights + (size_t)idx * (D + 1);
        if (D == 3) {
            *reinterpret_cast<int4*>(irow) = make_int4(ids[0], ids[1], ids[2], ids[3]);
            *reinterpret_cast<float4*>(wrow) =
                make_float4(ids[0] >= 0 ? s.bary[0] : -1.0f, ids[1] >= 0 ? s.bary[1] : -1.0f,
                            ids[2] >= 0 ? s.bary[2] : -1.0f, ids[3] >= 0 ? s.bary[3] : -1.0f);
        } else {
#pragma unroll
            for (int r = 0; r <= D; r++) {
                irow[r] = ids[r];
                wrow[r] = ids[r] >= 0 ? s.bary[r] : -1.0f;
            }
        }
    }
    if (kDistribute) {
        // row p*(D+1)+r = [ scaled position | value | barycentric_r ]  (LatticeGPU.cuh:633-645)
        const int row_len = D + val_dim + 1;
        float* out = distributed + (size_t)idx * (D + 1) * row_len;
        const float* v = values + (size_t)idx * val_dim;
#pragma unroll
        for (int r = 0; r <= D; r++) {
            float* o = out + r * row_len;
#pragma unroll
            for (int i = 0; i < D; i++) o[i] = p[i];
            for (int i = 0; i < val_dim; i++) o[D + i] = __ldg(v + i);
            o[D + val_dim] = s.bary[r];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// values[idx,:] += val[p,:]*w.  Small V (<=4, not a multiple of 4 lanes-wise): one thread per
// (point, simplex vertex); lanes that target the same vertex are summed in-warp first
// (warp-aggregated atomics), one RED per distinct vertex per warp.
template <int V>
__global__ void __launch_bounds__(kBlock)
splat_accumulate_small_kernel(const float* __restrict__ values, const int* __restrict__ indices,
                              const float* __restrict__ weights, int n, int spv /* D+1 */,
                              float* __restrict__ lattice_values) {
    LN_PDL_ENTRY();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n * spv;
    const bool valid = t < total;
    int id = -1;
    float v[V];
#pragma unroll
    for (int j = 0; j < V; j++) v[j] = 0.0f;
    if (valid) {
        id = __ldg(indices + t);
        if (id >= 0) {
            const float w = __ldg(weights + t);
            const int p = (int)(t / spv);
#pragma unroll
            for (int j = 0; j < V; j++) v[j] = __ldg(values + (size_t)p * V + j) * w;
        }
    }
    const unsigned active = __ballot_sync(kFull, id >= 0);
    if (id < 0) return;
    const unsigned peers = __match_any_sync(active, id);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; j++) acc[j] = 0.0f;
    for (unsigned m = peers; m; m &= m - 1) {   // ascending lane order: deterministic per warp
        const int src = __ffs(m) - 1;
#pragma unroll
        for (int j = 0; j < V; j++) acc[j] += __shfl_sync(peers, v[j], src);
    }
    if (lane == leader) {
        float* out = lattice_values + (size_t)id * V;
#pragma unroll
        for (int j = 0; j < V; j++) atomicAdd(out + j, acc[j]);
    }
}

// ---------------------------------------------------------------------------------------------
// coarsen<d>: one thread per (fine vertex, task); task 0 inserts key/2, task 1+2a / 2+2a handle the
// np / nm neighbour of axis a.
template <int D>
__global__ void __launch_bounds__(kBlock)
coarsen_kernel(ConstTableView fine, const int* __restrict__ fine_nr_filled, TableView coarse, int nv_upper) {
    LN_PDL_ENTRY();
    constexpr int kTasks = 1 + 2 * (D + 1);
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int v = (int)(t / kTasks);
    const int task = (int)(t % kTasks);
    if (v >= nv_upper || v >= __ldg(fine_nr_filled)) return;
    int key[D];
    bool all_even = true;
    int sum = 0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        key[i] = __ldg(fine.keys + (size_t)v * D + i);
        all_even &= ((key[i] & 1) == 0);
        sum += key[i];
    }
    all_even &= ((sum & 1) == 0);   // the implied last coordinate is -sum
    if (!all_even) return;
    int half[D];
#pragma unroll
    for (int i = 0; i < D; i++) half[i] = key[i] / 2;   // exact: all coordinates are even
    if (task == 0) {
        table_insert<D>(coarse, half, key_hash<D>(half));
        return;
    }
    const int axis = (task - 1) >> 1;
    const int sgn = ((task - 1) & 1) ? -1 : 1;   // +1: "np", -1: "nm"
    int nk[D], ck[D];
#pragma unroll
    for (int i = 0; i < D; i++) {
        const int step = (i == axis) ? -sgn * D : sgn;
        nk[i] = key[i] + step;
        ck[i] = half[i] + step;
    }
    if (table_find<D>(fine, nk) >= 0) table_insert<D>(coarse, ck, key_hash<D>(ck));
}

// ---------------------------------------------------------------------------------------------
template <int D>
static int launch_splat_build(const float* positions_raw, const float* sigmas, const float* values, int n, int val_dim,
                              TableView t, int* indices, float* weights, float* distributed, bool insert,
                              cudaStream_t s) {
    if (n == 0) return LN_OK;
    const int grid = cdiv(n, kBlock);
    if (distributed != nullptr)
        launch_k((splat_build_kernel<D, true, true>), dim3(grid), dim3(kBlock), 0, s, positions_raw, sigmas, values, n, val_dim, t, indices, weights, distributed);
    else if (insert)
        launch_k((splat_build_kernel<D, false, true>), dim3(grid), dim3(kBlock), 0, s, positions_raw, sigmas, values, n, val_dim, t, indices, weights, nullptr);
    else
        launch_k((splat_build_kernel<D, false, false>), dim3(grid), dim3(kBlock), 0, s, positions_raw, sigmas, values, n, val_dim, t, indices, weights, nullptr);
    count_launch();
    return check_launch("splat_build");
}

}  // namespace ln

using namespace ln;

extern "C" {

int ln_table_clear(int* entries, int* nr_filled, int* status, int capacity, void* stream) {
    LN_REQUIRE(entries && nr_filled && status && capacity > 0, "ln_table_clear: bad argument");
    const int grid = min(cdiv(max(capacity / 4, 1), kBlock), 148 * 8);
    launch_k(table_clear_kernel, dim3(grid), dim3(kBlock), 0, (cudaStream_t)stream, reinterpret_cast<int4*>(entries), entries, nr_filled, status, capacity);
    count_launch();
    return check_launch("table_clear");
}

static inline int vertex_bound(int max_vertices, int capacity) {
    return (max_vertices <= 0 || max_vertices > capacity) ? capacity : max_vertices;
}

int ln_splat_build(const float* positions_raw, const float* sigmas, int n, int pos_dim, int* keys, int* entries,
                   int* nr_filled, int* status, int capacity, int max_vertices, int* indices, float* weights, void* stream) {
    LN_REQUIRE(positions_raw && sigmas && keys && entries && nr_filled && status, "ln_splat_build: null pointer");
    LN_REQUIRE(n >= 0 && capacity > 0, "ln_splat_build: bad size n=%d capacity=%d", n, capacity);
    LN_REQUIRE((indices == nullptr) == (weights == nullptr), "ln_splat_build: indices and weights must both be given or both be NULL");
    TableView t{keys, entries, nr_filled, status, capacity, vertex_bound(max_vertices, capacity)};
    switch (pos_dim) {
        case 3: return launch_splat_build<3>(positions_raw, sigmas, nullptr, n, 0, t, indices, weights, nullptr, true, (cudaStream_t)stream);
        case 5: return launch_splat_build<5>(positions_raw, sigmas, nullptr, n, 0, t, indices, weights, nullptr, true, (cudaStream_t)stream);
    }
    set_error("ln_splat_build: unsupported pos_dim %d (3 and 5 are built)", pos_dim);
    return LN_ERR_UNSUPPORTED;
}

int ln_distribute(const float* positions_raw, const float* sigmas, const float* values, int n, int pos_dim, int val_dim,
                  int* keys, int* entries, int* nr_filled, int* status, int capacity, int max_vertices, int* indices,
                  float* weights, float* distributed, void* stream) {
    LN_REQUIRE(positions_raw && sigmas && values && keys && entries && nr_filled && status && indices && weights && distributed,
               "ln_distribute: null pointer");
    LN_REQUIRE(n >= 0 && capacity > 0 && val_dim >= 1, "ln_distribute: bad size");
    TableView t{keys, entries, nr_filled, status, capacity, vertex_bound(max_vertices, capacity)};
    switch (pos_dim) {
        case 3: return launch_splat_build<3>(positions_raw, sigmas, values, n, val_dim, t, indices, weights, distributed, true, (cudaStream_t)stream);
        case 5: return launch_splat_build<5>(positions_raw, sigmas, values, n, val_dim, t, indices, weights, distributed, true, (cudaStream_t)stream);
    }
    set_error("ln_distribute: unsupported pos_dim %d", pos_dim);
    return LN_ERR_UNSUPPORTED;
}

int ln_lookup_simplex(const float* positions_raw, const float* sigmas, int n, int pos_dim, const int* keys,
                      const int* entries, int capacity, int max_vertices, int* indices, float* weights, void* stream) {
    LN_REQUIRE(positions_raw && sigmas && keys && entries && indices && weights, "ln_lookup_simplex: null pointer");
    LN_REQUIRE(n >= 0 && capacity > 0, "ln_lookup_simplex: bad size");
    // ids at or past the caller's row bound (static-shape mode after an overflow) come back as -1, never as a row to read
    TableView t{const_cast<int*>(keys), const_cast<int*>(entries), nullptr, nullptr, capacity, vertex_bound(max_vertices, capacity)};
    switch (pos_dim) {
        case 3: return launch_splat_build<3>(positions_raw, sigmas, nullptr, n, 0, t, indices, weights, nullptr, false, (cudaStream_t)stream);
        case 5: return launch_splat_build<5>(positions_raw, sigmas, nullptr, n, 0, t, indices, weights, nullptr, false, (cudaStream_t)stream);
    }
    set_error("ln_lookup_simplex: unsupported pos_dim %d", pos_dim);
    return LN_ERR_UNSUPPORTED;
}

int ln_splat_accumulate(const float* values, const int* indices, const float* weights, int n, int pos_dim, int val_dim,
                        int nr_vertices, float* lattice_values, void* stream) {
    LN_REQUIRE(values && indices && weights && lattice_values, "ln_splat_accumulate: null pointer");
    LN_REQUIRE(n >= 0 && pos_dim >= 1 && val_dim >= 1, "ln_splat_accumulate: bad size");
    if (n == 0) return LN_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int spv = pos_dim + 1;
    const long long rows = (long long)n * spv;
    if (val_dim <= 3) {
        const int grid = cdiv(rows, kBlock);
        if (val_dim == 1) launch_k(splat_accumulate_small_kernel<1>, dim3(grid), dim3(kBlock), 0, s, values, indices, weights, n, spv, lattice_values);
        if (val_dim == 2) launch_k(splat_accumulate_small_kernel<2>, dim3(grid), dim3(kBlock), 0, s, values, indices, weights, n, spv, lattice_values);
        if (val_dim == 3) launch_k(splat_accumulate_small_kernel<3>, dim3(grid), dim3(kBlock), 0, s, values, indices, weights, n, spv, lattice_values);
    } else {
        // general V: the same scatter as the backward of slice (ln_slice.cu)
        return launch_scatter_rows(values, indices, weights, n, pos_dim, val_dim, nr_vertices, lattice_values, s, "splat_accumulate");
    }
    count_launch();
    return check_launch("splat_accumulate");
}

int ln_coarsen_keys(const int* fine_keys, const int* fine_entries, const int* fine_nr_filled, int fine_capacity,
                    int* coarse_keys, int* coarse_entries, int* coarse_nr_filled, int* coarse_status, int coarse_capacity,
                    int coarse_max_vertices, int pos_dim, int nv_fine_upper, void* stream) {
    LN_REQUIRE(fine_keys && fine_entries && fine_nr_filled && coarse_keys && coarse_entries && coarse_nr_filled && coarse_status,
               "ln_coarsen_keys: null pointer");
    LN_REQUIRE(fine_capacity > 0 && coarse_capacity > 0 && nv_fine_upper >= 0, "ln_coarsen_keys: bad size");
    if (nv_fine_upper == 0) return LN_OK;
    ConstTableView fine{fine_keys, fine_entries, fine_capacity};
    TableView coarse{coarse_keys, coarse_entries, coarse_nr_filled, coarse_status, coarse_capacity,
                     vertex_bound(coarse_max_vertices, coarse_capacity)};
    cudaStream_t s = (cudaStream_t)stream;
    if (pos_dim == 3)
        launch_k(coarsen_kernel<3>, dim3(cdiv((long long)nv_fine_upper * 9, kBlock)), dim3(kBlock), 0, s, fine, fine_nr_filled, coarse, nv_fine_upper);
    else if (pos_dim == 5)
        launch_k(coarsen_kernel<5>, dim3(cdiv((long long)nv_fine_upper * 13, kBlock)), dim3(kBlock), 0, s, fine, fine_nr_filled, coarse, nv_fine_upper);
    else {
        set_error("ln_coarsen_keys: unsupported pos_dim %d", pos_dim);
        return LN_ERR_UNSUPPORTED;
    }
    count_launch();
    return check_launch("coarsen_keys");
}

}  // extern "C"
