// Neighbourhood of lattice vertices: the hash walk of the reference's im2row / im2rowindices /
// row2im kernels (LatticeGPU.cuh:1464-1688, 1690-1920, 2067-2305) executed once per
// (query lattice, neighbour lattice, dilation) into a compact table, plus the API-parity
// materialisations built from that table.
#include "ln_common.cuh"

namespace ln {

constexpr int kBlock = 256;
// neighbours[] encoding: >= 0 vertex id; -1 looked up and absent; -2 slot not examined by the
// reference traversal (matters only for im2rowindices, whose untouched cells stay 0).
constexpr int kAbsent = -1;
constexpr int kSkipped = -2;

// One thread per (query vertex, slot).
template <int D>
__global__ void __launch_bounds__(kBlock)
neighbour_table_kernel(const int* __restrict__ query_keys, int nv_query, const int* __restrict__ nv_query_dev,
                       ConstTableView nbr, int nbr_max_vertices, int lvl_diff, int dilation, int* __restrict__ neighbours) {
    LN_PDL_ENTRY();
    constexpr int F = 2 * (D + 1) + 1;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nv_query * F) return;
    const int q = (int)(t / F);
    const int slot = (int)(t % F);
    // static-shape mode: the table has nv_query (bound) rows, only the first *nv_query_dev are vertices
    if (nv_query_dev != nullptr && q >= __ldg(nv_query_dev)) {
        neighbours[t] = kSkipped;
        return;
    }

    // full (D+1)-coordinate key, scaled between levels (LatticeGPU.cuh:1479-1495)
    float kf[D + 1];
    float key_sum = 0.0f;
#pragma unroll
    for (int i = 0; i < D; i++) {
        kf[i] = (float)__ldg(query_keys + (size_t)q * D + i);
        key_sum += kf[i];
    }
    kf[D] = -key_sum;
    const float scale = (lvl_diff > 0) ? 2.0f : (lvl_diff < 0 ? 0.5f : 1.0f);
    bool all_integer = true;
#pragma unroll
    for (int i = 0; i <= D; i++) {
        kf[i] *= scale;
        if (scale < 1.0f) all_integer &= (fabsf(kf[i] - truncf(kf[i])) <= 0.0001f);
    }

    int key[D];
    int result;
    if (slot == F - 1) {   // centre (LatticeGPU.cuh:1530-1537)
        result = kSkipped;
        if (all_integer) {
#pragma unroll
            for (int i = 0; i < D; i++) key[i] = (int)roundf(kf[i]);
            const int id = table_find<D>(nbr, key);
            if (id >= 0 && id < nbr_max_vertices) result = id;
        }
    } else {
        // fine query embedded in a coarser lattice: integer keys have no half-step neighbours
        const bool check = !(scale < 1.0f && all_integer);   // LatticeGPU.cuh:1545-1552
        result = kSkipped;
        if (check) {
            const int axis = slot >> 1;
            const float sgn = (slot & 1) ? -1.0f : 1.0f;   // even slot: "np", odd slot: "nm"
            const float mm = (scale < 1.0f) ? scale : 1.0f;
            const float step = mm * (float)dilation;
#pragma unroll
            for (int i = 0; i < D; i++) {
                const float c = (i == axis) ? kf[i] - sgn * (step * (float)D) : kf[i] + sgn * step;
                key[i] = (int)roundf(c);
            }
            result = table_find<D>(nbr, key);   // -1 == kAbsent
            if (result >= nbr_max_vertices) result = kAbsent;
        }
    }
    neighbours[t] = result;
}

// rowified[q, c*V + j] = values[neighbours[q, c'], j]  (zeros where absent); one thread per float4/float
template <int VEC>
__global__ void __launch_bounds__(kBlock)
im2row_kernel(const float* __restrict__ values, const int* __restrict__ neighbours, int nv_query, int F, int val_dim,
              int flip, float* __restrict__ rowified) {
    const int vpr = val_dim / VEC;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)nv_query * F * vpr;
    if (t >= total) return;
    const int j = (int)(t % vpr);
    const long long qc = t / vpr;
    const int c = (int)(qc % F);
    const long long q = qc / F;
    const int src_slot = (flip && c < F - 1) ? (c ^ 1) : c;
    const int id = __ldg(neighbours + q * F + src_slot);
    if (VEC == 4) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (id >= 0) x = __ldg(reinterpret_cast<const float4*>(values + (size_t)id * val_dim) + j);
        reinterpret_cast<float4*>(rowified)[t] = x;
    } else {
        rowified[t] = (id >= 0) ? __ldg(values + (size_t)id * val_dim + j) : 0.0f;
    }
}

__global__ void __launch_bounds__(kBlock)
im2rowindices_kernel(const int* __restrict__ neighbours, int nv_query, int F, int val_dim, int flip,
                     int* __restrict__ rowified) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)nv_query * F * val_dim;
    if (t >= total) return;
    const long long qc = t / val_dim;
    const int c = (int)(qc % F);
    const long long q = qc / F;
    const int src_slot = (flip && c < F - 1) ? (c ^ 1) : c;
    const int id = __ldg(neighbours + q * F + src_slot);
    rowified[t] = (id == kSkipped) ? 0 : id;   // untouched cells of the reference's zeros buffer
}

// out[v, :] = sum_slots rowified[nbr(v, slot), opposite(slot)*V : +V] + rowified[centre(v), (F-1)*V : +V]
// (LatticeGPU.cuh:2196-2284): a vertex pulls the chunk its neighbour stored for it.
template <int VEC>
__global__ void __launch_bounds__(kBlock)
row2im_kernel(const float* __restrict__ rowified, const int* __restrict__ neighbours, int nv, int F, int val_dim,
              float* __restrict__ out) {
    const int vpr = val_dim / VEC;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nv * vpr) return;
    const int j = (int)(t % vpr);
    const long long v = t / vpr;
    const size_t row_len = (size_t)F * val_dim;
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; k++) acc[k] = 0.0f;
    for (int slot = 0; slot < F; slot++) {
        const int id = __ldg(neighbours + v * F + slot);
        if (id < 0) continue;
        const int chunk = (slot < F - 1) ? (slot ^ 1) : slot;
        const float* src = rowified + (size_t)id * row_len + (size_t)chunk * val_dim + (size_t)j * VEC;
        if (VEC == 4) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(src));
            acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
        } else {
            acc[0] += __ldg(src);
        }
    }
    float* dst = out + (size_t)v * val_dim + (size_t)j * VEC;
    if (VEC == 4)
        *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else
        dst[0] = acc[0];
}

}  // namespace ln

using namespace ln;

extern "C" {

int ln_neighbour_table(const int* query_keys, int nv_query, const int* nv_query_dev, int pos_dim, const int* nbr_keys,
                       const int* nbr_entries, int nbr_capacity, int nbr_max_vertices, int lvl_diff, int dilation,
                       int* neighbours, void* stream) {
    LN_REQUIRE(query_keys && nbr_keys && nbr_entries && neighbours, "ln_neighbour_table: null pointer");
    LN_REQUIRE(nv_query >= 0 && nbr_capacity > 0 && dilation >= 1, "ln_neighbour_table: bad size");
    LN_REQUIRE(lvl_diff >= -1 && lvl_diff <= 1, "ln_neighbour_table: query and neighbour lattices may differ by one level at most (got %d)", lvl_diff);
    if (nv_query == 0) return LN_OK;
    ConstTableView nbr{nbr_keys, nbr_entries, nbr_capacity};
    if (nbr_max_vertices <= 0 || nbr_max_vertices > nbr_capacity) nbr_max_vertices = nbr_capacity;
    cudaStream_t s = (cudaStream_t)stream;
    if (pos_dim == 3)
        launch_k(neighbour_table_kernel<3>, dim3(cdiv((long long)nv_query * 9, kBlock)), dim3(kBlock), 0, s, query_keys, nv_query, nv_query_dev, nbr, nbr_max_vertices, lvl_diff, dilation, neighbours);
    else if (pos_dim == 5)
        launch_k(neighbour_table_kernel<5>, dim3(cdiv((long long)nv_query * 13, kBlock)), dim3(kBlock), 0, s, query_keys, nv_query, nv_query_dev, nbr, nbr_max_vertices, lvl_diff, dilation, neighbours);
    else {
        set_error("ln_neighbour_table: unsupported pos_dim %d", pos_dim);
        return LN_ERR_UNSUPPORTED;
    }
    count_launch();
    return check_launch("neighbour_table");
}

int ln_im2row(const float* nbr_values, const int* neighbours, int nv_query, int filter_extent, int val_dim, int flip,
              float* rowified, void* stream) {
    LN_REQUIRE(nbr_values && neighbours && rowified, "ln_im2row: null pointer");
    LN_REQUIRE(nv_query >= 0 && filter_extent >= 3 && val_dim >= 1, "ln_im2row: bad size");
    if (nv_query == 0) return LN_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const long long cells = (long long)nv_query * filter_extent;
    if (val_dim % 4 == 0)
        im2row_kernel<4><<<cdiv(cells * (val_dim / 4), kBlock), kBlock, 0, s>>>(nbr_values, neighbours, nv_query, filter_extent, val_dim, flip, rowified);
    else
        im2row_kernel<1><<<cdiv(cells * val_dim, kBlock), kBlock, 0, s>>>(nbr_values, neighbours, nv_query, filter_extent, val_dim, flip, rowified);
    count_launch();
    return check_launch("im2row");
}

int ln_im2rowindices(const int* neighbours, int nv_query, int filter_extent, int val_dim, int flip, int* rowified,
                     void* stream) {
    LN_REQUIRE(neighbours && rowified, "ln_im2rowindices: null pointer");
    LN_REQUIRE(nv_query >= 0 && filter_extent >= 3 && val_dim >= 1, "ln_im2rowindices: bad size");
    if (nv_query == 0) return LN_OK;
    const long long total = (long long)nv_query * filter_extent * val_dim;
    im2rowindices_kernel<<<cdiv(total, kBlock), kBlock, 0, (cudaStream_t)stream>>>(neighbours, nv_query, filter_extent, val_dim, flip, rowified);
    count_launch();
    return check_launch("im2rowindices");
}

int ln_row2im(const float* rowified, const int* neighbours, int nv, int filter_extent, int val_dim, float* out,
              void* stream) {
    LN_REQUIRE(rowified && neighbours && out, "ln_row2im: null pointer");
    LN_REQUIRE(nv >= 0 && filter_extent >= 3 && val_dim >= 1, "ln_row2im: bad size");
    if (nv == 0) return LN_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (val_dim % 4 == 0)
        row2im_kernel<4><<<cdiv((long long)nv * (val_dim / 4), kBlock), kBlock, 0, s>>>(rowified, neighbours, nv, filter_extent, val_dim, out);
    else
        row2im_kernel<1><<<cdiv((long long)nv * val_dim, kBlock), kBlock, 0, s>>>(rowified, neighbours, nv, filter_extent, val_dim, out);
    count_launch();
    return check_launch("row2im");
}

}  // extern "C"
