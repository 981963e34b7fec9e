// Host-side plumbing of the C ABI: error strings, launch accounting, table status readback.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "ln_common.cuh"

namespace ln {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};   // process-wide: backward launches come from autograd's worker thread

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(err));
        return LN_ERR_CUDA;
    }
    return LN_OK;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::atomic<int> g_pdl{1};
bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; }

struct LevelPtrs {
    const int* nr_filled[8];
    const int* status[8];
};
// per-step bookkeeping of a static-shape lattice pyramid in one launch: vertex count of every level, and 1.0 when any
// level's table overflowed or exceeded its row bound (the "skip this update" flag of the graphed step)
__global__ void levels_status_kernel(LevelPtrs p, int n_levels, int* __restrict__ nv_out, float* __restrict__ overflow_out) {
    LN_PDL_ENTRY();
    if (threadIdx.x == 0) {
        int bad = 0;
        for (int l = 0; l < n_levels; l++) {
            nv_out[l] = *p.nr_filled[l];
            bad |= p.status[l][0];
        }
        *overflow_out = bad != 0 ? 1.0f : 0.0f;
    }
}

}  // namespace ln

extern "C" {

int ln_levels_status(const int* const* nr_filled_ptrs, const int* const* status_ptrs, int n_levels, int* nv_out, float* overflow_out,
                     void* stream) {
    LN_REQUIRE(nr_filled_ptrs && status_ptrs && nv_out && overflow_out && n_levels >= 1 && n_levels <= 8, "ln_levels_status: bad argument");
    ln::LevelPtrs p;
    for (int l = 0; l < 8; l++) {
        p.nr_filled[l] = l < n_levels ? nr_filled_ptrs[l] : nullptr;
        p.status[l] = l < n_levels ? status_ptrs[l] : nullptr;
    }
    ln::launch_k(ln::levels_status_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, p, n_levels, nv_out, overflow_out);
    ln::count_launch();
    return ln::check_launch("levels_status");
}

const char* ln_version(void) { return "lattice_b200 0.1 sm_100a"; }
const char* ln_last_error(void) { return ln::g_error; }
long long ln_launch_count(void) { return ln::g_launches.load(std::memory_order_relaxed); }
void ln_reset_launch_count(void) { ln::g_launches.store(0, std::memory_order_relaxed); }
int ln_set_programmatic_launch(int enabled) {
    const int prev = ln::g_pdl.exchange(enabled ? 1 : 0, std::memory_order_relaxed);
    return prev;
}

int ln_table_status(const int* nr_filled, const int* status, int* nr_filled_host, int* max_probe_host, void* stream) {
    LN_REQUIRE(nr_filled && nr_filled_host, "ln_table_status: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    int st[2] = {0, 0};
    cudaError_t err = cudaMemcpyAsync(nr_filled_host, nr_filled, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess && status) err = cudaMemcpyAsync(st, status, 2 * sizeof(int), cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s);
    if (err != cudaSuccess) {
        ln::set_error("ln_table_status: %s", cudaGetErrorString(err));
        return LN_ERR_CUDA;
    }
    if (max_probe_host) *max_probe_host = st[1];
    if (st[0] & 1) {
        ln::set_error("hash table full: an insert found no free slot (nr_filled=%d); raise hash_table_capacity", *nr_filled_host);
        return LN_ERR_TABLE_FULL;
    }
    if (st[0] & 2) {
        ln::set_error("the lattice has %d vertices, more than the max_vertices bound it was built with", *nr_filled_host);
        return LN_ERR_VERTEX_BOUND;
    }
    return LN_OK;
}

}  // extern "C"
