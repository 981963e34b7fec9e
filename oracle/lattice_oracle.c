/*
 * TEST INFRASTRUCTURE (oracle) -- not product code.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / reference legs may load this.
 *
 * Plain-C restatement of the permutohedral geometry and hash function of AIS-Bonn/lattice_net,
 * following the arithmetic the reference's own build performs.  The reference compiles its kernels
 * with NVRTC `--use_fast_math` (jitify_helper.cuh:29); the operation sequence below is the one NVRTC
 * 12.9 emits for kernel_splat / distribute / slice_no_precomputation
 * (/root/reference/include/lattice_net/kernels/LatticeGPU.cuh:718-806, 544-622, 2608-2680; PTX in
 * oracle/_ref/lattice_ref.ptx):
 *     scale_i = constant (see elevate_scale below)
 *     cf = p*scale ; e_i = fma(cf,-i,sm) (i>=3) ; e_2 = sm - fma(p,scale,cf)
 *     e_1 = fma(p_0,-scale_0,sm) ; e_0 = fma(p_0,scale_0,sm)     <- contracted by ptxas, visible only in SASS
 *     v = e/(d+1) in double (d=5) or e*0.25f (d=3); up/down/rem0/rank/barycentric as in the source.
 *
 * Parity status: pinned against outputs of the reference's own kernels run on a B200
 * (tests/golden/*.npz) -- see DESIGN.md.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC oracle/lattice_oracle.c -o oracle/_build/liboracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MAX_D 8

static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
/* flush denormals to (signed) zero like the .ftz instruction forms */
static inline float ftz(float x) { return (fabsf(x) < 1.17549435e-38f && x != 0.0f) ? copysignf(0.0f, x) : x; }

/* Scale factor applied to position coordinate i.  In the reference's binary this is a constant:
 * ptxas folds `rsqrt.approx.ftz((i+1)(i+2))` with the correctly rounded 1/sqrt (NOT the MUFU.RSQ
 * hardware result, which differs by 1 ulp for 2, 6 and 20 -- tests/golden/rsqrt_approx.json) and
 * multiplies by fl((d+1)*sqrtf(2/3)) = 0f405105EC (d=3) / 0f409CC471 (d=5). */
static const uint32_t k_scale_bits_d3[3] = {0x4013cd3au, 0x3faaaaabu, 0x3f715befu};
static const uint32_t k_scale_bits_d5[5] = {0x405db3d8u, 0x40000001u, 0x3fb504f3u, 0x3f8c378cu, 0x3f64f92eu};

static float elevate_scale(int d, int i) {
    if (d == 3) return u2f(k_scale_bits_d3[i]);
    if (d == 5) return u2f(k_scale_bits_d5[i]);
    return (1.0f / sqrtf((float)((i + 1) * (i + 2)))) * ((float)(d + 1) * sqrtf(2.0f / 3));
}
void oracle_get_scale_table(int d, uint32_t* bits) { for (int i = 0; i < d; i++) bits[i] = f2u(elevate_scale(d, i)); }

/* HashTableGPU::hash, HashTableGPU.cuh:35-50 */
uint32_t oracle_hash(const int* key, int d) {
    uint32_t k = 0;
    for (int i = 0; i < d; i++) { k += (uint32_t)key[i]; k *= 2531011u; }
    return k;
}

/* One point: scaled position p[d] -> rem0[d+1], rank[d+1], bary[d+2]. */
static void simplex_of_point(const float* p, int d, int* rem0, int* rank, float* bary) {
    float e[MAX_D + 1];
    float sm = 0.0f;
    for (int i = d; i > 0; i--) {
        const float scale = elevate_scale(d, i - 1);
        const float pi = ftz(p[i - 1]);
        if (i == 1) {                      /* ptxas contracts  sm -/+ p*scale  into two FMAs */
            e[1] = ftz(fmaf(pi, -scale, sm));
            e[0] = ftz(fmaf(pi, scale, sm));
        } else {
            const float cf = ftz(pi * scale);
            if (i >= 3) e[i] = ftz(fmaf(cf, -(float)i, sm));
            else        e[i] = ftz(sm - ftz(fmaf(pi, scale, cf)));
            sm = ftz(sm + cf);
        }
    }

    int sum = 0;
    for (int i = 0; i <= d; i++) {
        float v;
        if (d == 3) v = ftz(e[i] * 0.25f);
        else        v = ftz((float)((double)e[i] * (1.0 / (d + 1))));
        const float up = ftz(ceilf(v) * (float)(d + 1));
        const float down = ftz(floorf(v) * (float)(d + 1));
        rem0[i] = (ftz(up - e[i]) < ftz(e[i] - down)) ? (int)up : (int)down;
        sum += rem0[i];
    }
    sum /= (d + 1);   /* C integer division truncates toward zero, as on the device */

    float diff[MAX_D + 1];
    for (int i = 0; i <= d; i++) { diff[i] = ftz(e[i] - (float)rem0[i]); rank[i] = 0; }
    for (int i = 0; i < d; i++)
        for (int j = i + 1; j <= d; j++) {
            if (diff[i] < diff[j]) rank[i]++; else rank[j]++;
        }
    for (int i = 0; i <= d; i++) {
        rank[i] += sum;
        if (rank[i] < 0)      { rank[i] += d + 1; rem0[i] += d + 1; }
        else if (rank[i] > d) { rank[i] -= d + 1; rem0[i] -= d + 1; }
    }
    for (int k = 0; k <= d + 1; k++) bary[k] = 0.0f;
    for (int i = 0; i <= d; i++) {
        const float d0 = ftz(e[i] - (float)rem0[i]);
        float delta;
        if (d == 3) delta = ftz(d0 * 0.25f);
        else        delta = ftz((float)((double)d0 * (1.0 / (d + 1))));
        bary[d - rank[i]] = ftz(bary[d - rank[i]] + delta);
        bary[d + 1 - rank[i]] = ftz(bary[d + 1 - rank[i]] - delta);
    }
    bary[0] = ftz((float)(((double)bary[d + 1] + 1.0) + (double)bary[0]));
}

/* positions_raw[n x d] / sigmas[d] -> keys[n x (d+1) x d], bary[n x (d+1)], scaled[n x d] (may be NULL).
 * Division is IEEE fp32 like torch's `positions_raw / sigmas_tensor` (/root/reference/src/Lattice.cu:226). */
int oracle_simplex(const float* positions_raw, const float* sigmas, int n, int d,
                   int* keys, float* bary_out, float* scaled_out) {
    if (d < 1 || d > MAX_D) return -1;
    for (int p = 0; p < n; p++) {
        float ps[MAX_D];
        int rem0[MAX_D + 1], rank[MAX_D + 1];
        float bary[MAX_D + 2];
        for (int i = 0; i < d; i++) ps[i] = positions_raw[(size_t)p * d + i] / sigmas[i];
        if (scaled_out) for (int i = 0; i < d; i++) scaled_out[(size_t)p * d + i] = ps[i];
        simplex_of_point(ps, d, rem0, rank, bary);
        for (int r = 0; r <= d; r++) {
            int* key = keys + ((size_t)p * (d + 1) + r) * d;
            for (int i = 0; i < d; i++) {
                key[i] = rem0[i] + r;
                if (rank[i] > d - r) key[i] -= (d + 1);
            }
            bary_out[(size_t)p * (d + 1) + r] = bary[r];
        }
    }
    return 0;
}

/* Reference hash table, sequential: linear probing over `capacity` slots (HashTableGPU.cuh:425-519).
 * entries[capacity] must be pre-filled with -1.  Returns vertex id, or -1 when the table is full. */
int oracle_table_insert(int* keys, int* entries, int* nr_filled, int capacity, int d, const int* key) {
    int h = (int)(oracle_hash(key, d) % (uint32_t)capacity);
    for (int probe = 0; probe < capacity; probe++) {
        const int e = entries[h];
        if (e == -1) {
            const int id = (*nr_filled)++;
            memcpy(keys + (size_t)id * d, key, sizeof(int) * d);
            entries[h] = id;
            return id;
        }
        if (memcmp(keys + (size_t)e * d, key, sizeof(int) * d) == 0) return e;
        h = (h + 1 == capacity) ? 0 : h + 1;
    }
    return -1;
}
int oracle_table_find(const int* keys, const int* entries, int capacity, int d, const int* key, int max_probes) {
    int h = (int)(oracle_hash(key, d) % (uint32_t)capacity);
    for (int probe = 0; probe < max_probes; probe++) {
        const int e = entries[h];
        if (e == -1) return -1;
        if (memcmp(keys + (size_t)e * d, key, sizeof(int) * d) == 0) return e;
        h = (h + 1 == capacity) ? 0 : h + 1;
    }
    return -1;
}

/* Sequential kernel_splat: insert all simplex keys in point order; returns nr_filled or -1 on overflow.
 * Also reports the longest probe chain (the reference's retrieve gives up after 300, HashTableGPU.cuh:494). */
int oracle_build_table(const int* simplex_keys, long long n_keys, int d, int capacity,
                       int* keys, int* entries, int* indices, int* max_chain_out) {
    int nr_filled = 0, max_chain = 0;
    for (int i = 0; i < capacity; i++) entries[i] = -1;
    for (long long k = 0; k < n_keys; k++) {
        const int* key = simplex_keys + k * d;
        const int id = oracle_table_insert(keys, entries, &nr_filled, capacity, d, key);
        if (id < 0) return -1;
        if (indices) indices[k] = id;
    }
    for (int v = 0; v < nr_filled; v++) {   /* chain length of every stored key */
        int h = (int)(oracle_hash(keys + (size_t)v * d, d) % (uint32_t)capacity), chain = 0;
        while (entries[h] != v) { h = (h + 1 == capacity) ? 0 : h + 1; chain++; }
        if (chain > max_chain) max_chain = chain;
    }
    if (max_chain_out) *max_chain_out = max_chain;
    return nr_filled;
}
