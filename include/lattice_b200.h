/*
 * lattice_b200.h -- C ABI of the B200-native permutohedral-lattice backend.
 *
 * Drop-in boundary for the lattice hot path of AIS-Bonn/lattice_net.  The reference exposes
 * this path as a pybind11 C++ class (`Lattice`, /root/reference/src/PyBridge.cxx:41-113) whose
 * methods allocate torch tensors and then call one `LatticeGPU::<op>()` launcher each
 * (/root/reference/include/lattice_net/kernels/LatticeGPU.cuh:42-412).  This header declares
 * one `extern "C"` entry point per launcher: raw DEVICE pointers, sizes, a `cudaStream_t`
 * passed as `void*`, and an `int` status.  No torch / C++ types cross the boundary; the host
 * side (lattice_net_b200/lattice.py, or a maintainer's own pybind stub)
 * owns all memory.
 *
 * Conventions
 *   - all floating data is fp32, all indices / keys are int32 (as in the reference);
 *   - `stream` is a cudaStream_t (0 = legacy default stream);
 *   - every function returns LN_OK (0) or a negative LN_ERR_* code; ln_last_error() returns a
 *     thread-local human-readable message for the last failure;
 *   - functions are asynchronous with respect to the host unless stated otherwise;
 *   - a hash table is described by four device arrays exactly like the reference's
 *     HashTableGPU (/root/reference/include/lattice_net/kernels/HashTableGPU.cuh:23-28):
 *        keys      int32 [capacity x pos_dim]  compact, row id == vertex id
 *        entries   int32 [capacity]            slot -> vertex id, -1 empty, -2 being written
 *        nr_filled int32 [1]                   number of vertices
 *     plus `capacity` and `pos_dim` passed by value.  `values` [nv x val_dim] are separate.
 *   - index / weight tables are [n_points * (pos_dim+1)], point-major; -1 / -1.0f mean
 *     "not inserted / not found" (/root/reference/src/Lattice.cpp: fill_(-1) at Lattice.cu:214-215).
 *   - supported pos_dim: 3 and 5 (the reference's shipped configs use 3; its sweep uses 5).
 */
#ifndef LATTICE_B200_H_
#define LATTICE_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define LN_OK 0
#define LN_ERR_BAD_ARG (-1)        /* null pointer, negative size, unsupported pos_dim ...   */
#define LN_ERR_CUDA (-2)           /* a CUDA runtime call / launch failed                     */
#define LN_ERR_UNSUPPORTED (-3)    /* shape outside what the kernels implement                */
#define LN_ERR_TABLE_FULL (-4)     /* reported by ln_table_status(): an insert found no slot  */
#define LN_ERR_VERTEX_BOUND (-5)   /* reported by ln_table_status(): more vertices than max_vertices */

/* Version / build info: "lattice_b200 <ver> sm_100a". */
const char* ln_version(void);
/* Message describing the last error returned on this host thread. */
const char* ln_last_error(void);
/* Number of kernels launched by this library (all host threads) since the last reset
 * (bench.py reports it as `gpu_launches`). */
long long ln_launch_count(void);
void ln_reset_launch_count(void);
/* Bookkeeping of a static-shape lattice pyramid without a host round trip: nv_out[l] = vertex count of level l,
 * overflow_out[0] = 1.0 if any level's table filled up or exceeded its row bound, else 0.0.  nr_filled_ptrs / status_ptrs:
 * HOST arrays of n_levels (<= 8) device pointers (the nr_filled / status words of ln_splat_build). */
int ln_levels_status(const int* const* nr_filled_ptrs, const int* const* status_ptrs, int n_levels, int* nv_out, float* overflow_out,
                     void* stream);
/* Programmatic dependent launch of this library's kernels (on by default; see ln_common.cuh): 0 launches them with plain
 * stream ordering.  Returns the previous setting.  Results are identical either way. */
int ln_set_programmatic_launch(int enabled);

/* ---- hash table ---------------------------------------------------------------------------
 * Replaces HashTable::clear() (/root/reference/src/HashTable.cu:49-57) for the structural
 * part: entries <- -1, nr_filled <- 0, status <- 0.  keys need no clearing (rows >= nr_filled
 * are never read).  `status` is a device int32[2]: [0] overflow flag, [1] max probe length. */
int ln_table_clear(int* entries, int* nr_filled, int* status, int capacity, void* stream);

/* Synchronous: copies nr_filled and the status words to the host (one D2H of 12 bytes).
 * Replaces Lattice::nr_lattice_vertices() (/root/reference/src/Lattice.cu:1326-1346).
 * Returns LN_ERR_TABLE_FULL if an insert overflowed (the reference would spin forever,
 * HashTableGPU.cuh:443-484 has no probe limit). */
int ln_table_status(const int* nr_filled, const int* status, int* nr_filled_host, int* max_probe_host, void* stream);

/* ---- splat: build structure -----------------------------------------------------------------
 * Replaces kernel_splat<d,V> (LatticeGPU.cuh:707-842) + the host prep of
 * Lattice::splat_standalone / just_create_verts (/root/reference/src/Lattice.cu:196-290):
 * positions_raw / sigmas, simplex + barycentric computation, insertion of the pos_dim+1 simplex
 * vertices, and (if indices/weights != NULL) the splatting tables.
 *   positions_raw [n x pos_dim], sigmas [pos_dim] (device), indices/weights [n*(pos_dim+1)] or NULL
 * max_vertices (static-shape mode, see DESIGN.md "CUDA-graph step"): the caller's per-vertex tensors have
 * only max_vertices rows.  Vertices numbered >= max_vertices are still inserted (nr_filled keeps counting)
 * but are handed out as index -1 / weight -1 and status[0] gets bit 1 set.  <= 0 or > capacity: no bound. */
int ln_splat_build(const float* positions_raw, const float* sigmas, int n, int pos_dim,
                   int* keys, int* entries, int* nr_filled, int* status, int capacity, int max_vertices,
                   int* indices, float* weights, void* stream);

/* Replaces splatCacheNaive<d,V> (LatticeGPU.cuh:926-973):
 *   lattice_values[indices[p,r], :] += values[p, :] * weights[p,r]      (rows with index < 0 skipped)
 * lattice_values must be zero-initialised by the caller (HashTable::clear does it in the reference).
 * nr_vertices: rows of lattice_values (<= 0: unknown, n is assumed).  Only a tuning input: the kernels walk the
 * channels in slabs sized so that one slab of the vertex table stays resident in the L2. */
int ln_splat_accumulate(const float* values, const int* indices, const float* weights,
                        int n, int pos_dim, int val_dim, int nr_vertices, float* lattice_values, void* stream);

/* Replaces distribute<d,V> (LatticeGPU.cuh:534-650) + host prep of Lattice::distribute
 * (/root/reference/src/Lattice.cu:351-410).  As ln_splat_build, plus
 *   distributed [n*(pos_dim+1) x (pos_dim+val_dim+1)], row p*(pos_dim+1)+r =
 *        [ positions_raw[p]/sigmas | values[p] | barycentric_r ]                                */
int ln_distribute(const float* positions_raw, const float* sigmas, const float* values,
                  int n, int pos_dim, int val_dim,
                  int* keys, int* entries, int* nr_filled, int* status, int capacity, int max_vertices,
                  int* indices, float* weights, float* distributed, void* stream);

/* Structure part of slice_no_precomputation<d,V> (LatticeGPU.cuh:2598-2750): recompute the simplex
 * of every position and *retrieve* (never insert) its vertices; not-found -> index -1, weight -1. */
/* max_vertices: row bound of the caller's per-vertex tensors (0 = capacity); ids at or past it are returned as -1. */
int ln_lookup_simplex(const float* positions_raw, const float* sigmas, int n, int pos_dim,
                      const int* keys, const int* entries, int capacity, int max_vertices,
                      int* indices, float* weights, void* stream);

/* Replaces coarsen<d> (LatticeGPU.cuh:2314-2514) used by Lattice::create_coarse_verts
 * (/root/reference/src/Lattice.cu:670-703): every fine vertex whose key is all-even inserts key/2
 * into the coarse table, plus the coarse image of each of its existing fine 1-hop neighbours.
 * nv_fine is read from fine_nr_filled on the device.  coarse_max_vertices: as max_vertices of ln_splat_build
 * (only the status flag; this call hands out no indices). */
int ln_coarsen_keys(const int* fine_keys, const int* fine_entries, const int* fine_nr_filled, int fine_capacity,
                    int* coarse_keys, int* coarse_entries, int* coarse_nr_filled, int* coarse_status, int coarse_capacity,
                    int coarse_max_vertices, int pos_dim, int nv_fine_upper, void* stream);

/* ---- neighbourhood ---------------------------------------------------------------------------
 * The traversal of im2row / im2rowindices / row2im (LatticeGPU.cuh:1464-1688, 1690-1920, 2067-2305)
 * done ONCE per (query lattice, neighbour lattice, dilation): neighbours[q, slot] = vertex id in the
 * neighbour lattice or -1.  Slot order (filter_extent F = 2(pos_dim+1)+1): slot 2a = "np" of axis a,
 * slot 2a+1 = "nm" of axis a, slot F-1 = centre.  lvl_diff = query_lvl - neighbour_lvl in {-1,0,1}.
 * Static-shape mode: nv_query is the ROW BOUND of the table and nv_query_dev (device int32, may be NULL) the
 * actual vertex count (the query table's nr_filled): rows >= *nv_query_dev get no neighbours, so every
 * convolution over the table yields zeros there.  Neighbour ids >= nbr_max_vertices (<= 0: no bound) are
 * reported as absent. */
int ln_neighbour_table(const int* query_keys, int nv_query, const int* nv_query_dev, int pos_dim,
                       const int* nbr_keys, const int* nbr_entries, int nbr_capacity, int nbr_max_vertices,
                       int lvl_diff, int dilation, int* neighbours, void* stream);

/* API-parity materialisations (the conv kernels below never need them).
 * rowified [nv_query x F*val_dim]; `flip` swaps the np/nm slots (LatticeGPU.cuh:1626,1649). */
int ln_im2row(const float* nbr_values, const int* neighbours, int nv_query, int filter_extent,
              int val_dim, int flip, float* rowified, void* stream);
/* int32 variant: neighbour vertex id replicated val_dim times; slots without a neighbour stay 0
 * exactly like the reference's zeros-initialised buffer (/root/reference/src/Lattice.cu:600). */
int ln_im2rowindices(const int* neighbours, int nv_query, int filter_extent, int val_dim, int flip,
                     int* rowified, void* stream);
/* out[v, :] = sum over slots of rowified[neighbour(v, slot), opposite(slot) chunk] + centre chunk
 * (LatticeGPU.cuh:2067-2305); `neighbours` is the table of the lattice itself. out [nv x val_dim]. */
int ln_row2im(const float* rowified, const int* neighbours, int nv, int filter_extent, int val_dim,
              float* out, void* stream);

/* ---- lattice convolution (implicit GEMM, no im2row buffer) -------------------------------------
 * Replaces im2row + `lattice_rowified.mm(filter_bank)` in Lattice::convolve_im2row_standalone
 * (/root/reference/src/Lattice.cu:424-474):
 *   out[q, co] = sum_slot sum_ci nbr_values[neighbours[q, slot'], ci] * filter[slot*c_in + ci, co]
 *                (+ bias[co]) (+ residual[q, co])
 * slot' = slot^1 for slot < F-1 when flip != 0 (the data-gradient convolution of
 * /root/reference/latticenet_py/lattice/lattice_funcs.py:307-313), else slot.
 * filter_extent = 1 with neighbours[q, 0] = q is a plain row-major GEMM: the 1x1 layers of the bottleneck blocks and of
 * the slice head (torch.nn.Linear in lattice_modules.py:806-832) run through the same kernels.
 * transposed_filter != 0: `filter` is the FORWARD bank [F*c_out x c_in] of the convolution whose data
 * gradient is being computed, read as filter_bw[(slot*c_in + ci), co] = filter[(slot*c_out + co), ci]
 * -- the transpose/view/contiguous re-layout of lattice_funcs.py:304-311 without materialising it.
 * residual (may be NULL): [nv_query x c_out] added in the epilogue (the skip connection of a residual block,
 * lattice_modules.py:1255-1358, without its own kernel).  bias may be NULL.
 * precision: 0 = exact fp32 FMA on the CUDA cores; 1 = tcgen05 tensor cores, 3xTF32 split (fp32-equivalent,
 * ~1e-6 relative); 2 = tcgen05 single-pass TF32 (~1e-3 relative).  The tensor-core path needs
 * c_in % 32 == 0 and c_out <= 1024 (layers wider than 256 run as 256-column chunks); other shapes run the fp32 kernel
 * whatever `precision` says.
 * slabs: for precision 1/2 a device buffer of ln_conv_workspace_bytes() bytes holding the PREPARED filter of this
 * reading (pre-swizzled B tiles, TF32 high / low parts).  slabs_prepared != 0: it already does (ln_filter_prepare /
 * ln_filter_prepare_batch ran after the last change of `filter`); 0: this call prepares it first.  NULL for precision 0.
 * out_is_zero != 0: the caller hands over a zeroed `out`; otherwise the call clears it itself when
 * ln_conv_needs_zero() says the kernel accumulates (K split across CTAs on small lattices). */
int ln_conv_fwd(const float* nbr_values, const int* neighbours, const float* filter, const float* bias, const float* residual,
                int nv_query, int filter_extent, int c_in, int c_out, int flip, int transposed_filter, int precision,
                float* slabs, int slabs_prepared, int out_is_zero, float* out, void* stream);
long long ln_conv_workspace_bytes(int filter_extent, int c_in, int c_out, int precision);
int ln_conv_needs_zero(int nv_query, int filter_extent, int c_in, int c_out, int precision);

/* Filter preparation for the tensor-core convolution, hoisted out of the per-call path: weights change once per
 * optimizer step, so every bank is prepared once per step for both of its readings (forward; transposed for the data
 * gradient) instead of once per convolution call.  ln_filter_prepare: one bank, one launch.
 * ln_filter_prepare_batch: n_jobs banks in ONE launch; jobs_device = device array of 48-byte records
 *   { const float* src; float* dst; int k_total (= F*c_in of the reading); int c_in; int c_out; int transposed;
 *     int split (1 = write the low parts, precision 1); int pad; long long first_thread; }
 * with first_thread the running sum of (k_total / 32) * ceil(n_pad_sum(c_out) / 32) -- the job's number of 32 x 32 tiles, one
 * CTA each -- over the preceding jobs (n_pad_sum = c_out rounded up to 16 within every 256-column chunk) and total_threads
 * that sum over all jobs. */
int ln_filter_prepare(const float* filter, int filter_extent, int c_in, int c_out, int transposed_filter, int precision,
                      float* slabs, void* stream);
int ln_filter_prepare_batch(const void* jobs_device, int n_jobs, long long total_threads, void* stream);

/* Weight gradient, replaces `lattice_rowified.transpose(0,1).mm(grad)` (lattice_funcs.py:302,378,443):
 *   grad_filter[slot*c_in + ci, co] = sum_q nbr_values[neighbours[q, slot], ci] * grad_out[q, co]
 * grad_filter [F*c_in x c_out]: grad_is_zero != 0 = the caller hands over a zeroed buffer (e.g. a slice of a gradient
 * bucket cleared once per step); 0 = the call clears it when its kernel accumulates with reductions.
 * precision as in ln_conv_fwd: 1 / 2 run the gathered-A^T . G product on tcgen05 (MN-major operands, the
 * reduction runs over the vertices) when c_in % 32 == 0 and c_out % 4 == 0, c_out <= 1024 (256-column chunks); else fp32 FMA. */
int ln_conv_wgrad(const float* nbr_values, const int* neighbours, const float* grad_out,
                  int nv_query, int filter_extent, int c_in, int c_out, int precision, int grad_is_zero,
                  float* grad_filter, void* stream);

/* Whole backward pass of one lattice convolution out = conv(query <- neighbours) in one call
 * (/root/reference/latticenet_py/lattice/lattice_funcs.py:294-313, 373-388, 438-454):
 *   grad_nbr_values [nv_nbr x c_in] = flipped convolution of grad_out [nv_query x c_out] at the neighbour
 *                                     lattice's vertices (neighbours_bwd [nv_nbr x F] = table neighbour -> query),
 *                                     forward filter bank read transposed;            NULL = not wanted
 *   grad_filter [F*c_in x c_out]    = im2row(nbr_values)^T . grad_out                 NULL = not wanted
 * The two run side by side (the weight gradient on an internal second stream, joined before returning).
 * slabs_bwd / slabs_prepared: prepared TRANSPOSED reading of `filter` (ln_conv_workspace_bytes(F, c_out, c_in, precision)
 * bytes), as in ln_conv_fwd.  *_is_zero: as in ln_conv_fwd / ln_conv_wgrad.
 * linear_weight != 0 (filter_extent 1): `filter` is a torch.nn.Linear weight [c_out x c_in], i.e. the bank stored
 * transposed: grad_nbr_values = grad_out . filter (its plain reading) and grad_filter [c_out x c_in] = grad_out^T . nbr_values.
 * defer_join != 0: do not make `stream` wait for the weight gradient; the caller calls ln_conv_bwd_join(stream) before
 * anything reads a grad_filter and keeps nbr_values / grad_out alive until then (the weight gradients of a whole backward
 * pass then trail the data-gradient chain instead of holding it up layer by layer). */
int ln_conv_bwd(const float* nbr_values, const int* neighbours_fwd, const float* grad_out, const int* neighbours_bwd,
                const float* filter, int nv_query, int nv_nbr, int filter_extent, int c_in, int c_out, int precision,
                float* slabs_bwd, int slabs_prepared, float* grad_nbr_values, int grad_nbr_is_zero, float* grad_filter,
                int grad_filter_is_zero, int linear_weight, int defer_join, void* stream);
int ln_conv_bwd_join(void* stream);

/* filter_bw[(slot*c_out + co), ci] = filter[(slot*c_in + ci), co]: the re-layout done with
 * transpose/view/contiguous in lattice_funcs.py:304-311. */
int ln_filter_for_dgrad(const float* filter, int filter_extent, int c_in, int c_out, float* filter_bw, void* stream);

/* ---- slice family -------------------------------------------------------------------------------
 * slice_with_precomputation<d,V> (LatticeGPU.cuh:2552-2595): out[p,:] = sum_r w[p,r]*values[idx[p,r],:]
 * nr_vertices: rows of the vertex table, a tuning input as in ln_splat_accumulate (<= 0: unknown). */
int ln_slice_fwd(const float* lattice_values, const int* indices, const float* weights,
                 int n, int pos_dim, int val_dim, int nr_vertices, float* out, void* stream);
/* slice_backwards_with_precomputation_no_homogeneous<d,V> (LatticeGPU.cuh:3540-3623):
 *   grad_values[idx[p,r], :] += grad_out[p,:] * w[p,r]      (grad_values pre-zeroed by caller) */
int ln_slice_bwd(const float* grad_out, const int* indices, const float* weights,
                 int n, int pos_dim, int val_dim, int nr_vertices, float* grad_values, void* stream);
/* gather_with_precomputation<d,V> (LatticeGPU.cuh:2886-2929): out [n x (pos_dim+1)(val_dim+1)],
 * chunk r = [ w_r * values[idx_r,:] | w_r ], zeros where idx_r < 0. */
int ln_gather_fwd(const float* lattice_values, const int* indices, const float* weights,
                  int n, int pos_dim, int val_dim, float* out, void* stream);
/* gather_backwards_with_precomputation<d,V> (LatticeGPU.cuh:3761-3817). grad_values pre-zeroed. */
int ln_gather_bwd(const float* grad_out, const int* indices, const float* weights,
                  int n, int pos_dim, int val_dim, float* grad_values, void* stream);
/* slice_classify_with_precomputation<d,V,nc> (LatticeGPU.cuh:3387-3464):
 *   logits[p,c] = bias[c] + sum_v W[c,v] * sum_r (w[p,r]+dw[p,r]) * values[idx[p,r], v]            */
int ln_slice_classify_fwd(const float* lattice_values, const int* indices, const float* weights,
                          const float* delta_weights, const float* cls_weight, const float* cls_bias,
                          int n, int pos_dim, int val_dim, int nr_classes, float* logits, void* stream);
/* slice_classify_backwards_with_precomputation<d,V,nc> (LatticeGPU.cuh:3628-3756): accumulates into
 * the four caller-zeroed gradients (/root/reference/latticenet_py/lattice/lattice_funcs.py:553-556). */
int ln_slice_classify_bwd(const float* grad_logits, const float* lattice_values, const int* indices,
                          const float* weights, const float* delta_weights, const float* cls_weight,
                          int n, int pos_dim, int val_dim, int nr_classes,
                          float* grad_lattice_values, float* grad_delta_weights,
                          float* grad_cls_weight, float* grad_cls_bias, void* stream);

/* Learned barycentric offsets of the DeformSlice head (lattice_modules.py:465-567: gather, max over the simplex, affine,
 * Linear(9 -> 1)) in one kernel each way.  values [nv x 8] (the head's bottleneck features), gamma / beta / lin_w [9],
 * lin_b [1]; delta_w [n x (pos_dim+1)].  Backward: grad_values_zeroed [nv x 8] and the four zeroed parameter gradients
 * (lin_w [9], lin_b [1], gamma [9], beta [9]) are accumulated into. */
int ln_deltaw_fwd(const float* values, const int* indices, const float* weights, const float* gamma, const float* beta,
                  const float* lin_w, const float* lin_b, int n, int pos_dim, int val_dim, float* delta_w, void* stream);
int ln_deltaw_bwd(const float* values, const int* indices, const float* weights, const float* gamma, const float* beta,
                  const float* lin_w, const float* grad_delta_w, int n, int pos_dim, int val_dim, float* grad_values_zeroed,
                  float* grad_lin_w_zeroed, float* grad_lin_b_zeroed, float* grad_gamma_zeroed, float* grad_beta_zeroed, void* stream);

/* ---- PointNet glue on lattice vertices ---------------------------
 * Segmented reductions over the points that splat onto each vertex, replacing the torch_scatter
 * calls of /root/reference/latticenet_py/lattice/lattice_modules.py:78,688,692.
 * index [m] (already clamped to >= 0), src [m x c]; out_* are [nv x c].
 * scatter_max: out_max[v,c] = max over rows, out_arg[v,c] = row attaining it (m if none, value 0 if none)
 *              -- torch_scatter.scatter_max semantics. */
int ln_scatter_max(const float* src, const int* index, int m, int c, int nv,
                   float* out_max, int* out_arg, unsigned long long* workspace /* [nv x c] */, void* stream);
/* out_sum[v,c] = sum, out_count[v] = number of rows (float) */
int ln_scatter_sum_count(const float* src, const int* index, int m, int c, int nv,
                         float* out_sum, float* out_count, void* stream);

/* Fused distribute + PointNet front end (lattice_modules.py:52-96 + 620-733) without any per-row tensor: per-vertex mean
 * position, centring, three weight-normalised Linear + LeakyReLU(0.2) layers evaluated per (point, simplex vertex) in
 * registers, per-vertex max pooling, barycentric weight of each winning row, "< min_points" and vertex-0 masks.
 *   out [nv_rows x 2*h3] = [ max | barycentric of the argmax row ],  arg [nv_rows x h3] = winning row (n*(pos_dim+1) = none)
 * layer_ptrs: HOST array of 9 device pointers (weight_v [out x in], weight_g [out], bias [out]) x 3 layers.
 * scratch_zeroed: ln_pointnet_scratch_floats() zeroed floats (kept for the backward: its head holds the per-vertex sums).
 * Built for pos_dim 3 (4..8 input features) and 5 (6..9), widths 16/32/64 (every reference config); query with
 * ln_pointnet_supported().  Backward: grad_ptrs = HOST array of the 9 gradient destinations (overwritten);
 * grad_scratch_zeroed: ln_pointnet_grad_scratch_floats() zeroed floats. */
int ln_pointnet_supported(int pos_dim, int val_dim, int h1, int h2, int h3);
long long ln_pointnet_scratch_floats(int pos_dim, int nv_rows, int h3);
long long ln_pointnet_grad_scratch_floats(int pos_dim, int val_dim, int h1, int h2, int h3);
int ln_pointnet_fwd(const float* positions_raw, const float* sigmas, const float* values, const int* indices, const float* weights,
                    int n, int pos_dim, int val_dim, const float* const* layer_ptrs, int h1, int h2, int h3, int nv_rows,
                    int vertex0_quirk, int min_points, float* scratch_zeroed, float* out, int* arg, void* stream);
int ln_pointnet_bwd(const float* positions_raw, const float* sigmas, const float* values, const int* indices, int n, int pos_dim,
                    int val_dim, const float* const* layer_ptrs, float* const* grad_ptrs, int h1, int h2, int h3, int vertex0_quirk,
                    const float* fwd_scratch, const float* grad_reduced, const int* arg, float* grad_scratch_zeroed, void* stream);

/* ---- normalisation between lattice convolutions -----------------------
 * GroupNorm (+ optional fused ReLU) on vertex-major lattice values x [nv x c]: statistics per group over
 * (c/groups channels) x (all nv vertices), biased variance -- torch.nn.GroupNorm(groups, c) applied to the
 * [1, c, nv] view the reference builds with unsqueeze/transpose
 * (/root/reference/latticenet_py/lattice/lattice_modules.py:585-614).  stats [groups x 2] = (mean, rstd).
 * nv_dev (device int32, may be NULL): static-shape mode, only the first min(nv, *nv_dev) rows are vertices;
 * the others are excluded from the statistics and written as zeros.
 * workspace: device scratch of ln_group_norm_workspace_bytes(nv, c, groups) bytes (0 for lattices small enough for
 * the one-CTA-per-group kernels; then it may be NULL).  With it, scene-sized lattices run the row-tiled kernels
 * (stats partial -> finalize -> apply on all SMs, coalesced, no atomics); without it they fall back to one CTA
 * per group.  The same size serves forward and backward. */
long long ln_group_norm_workspace_bytes(int nv, int c, int groups);
int ln_group_norm_fwd(const float* x, const float* gamma, const float* beta, int nv, const int* nv_dev, int c, int groups,
                      float eps, int relu, float* y, float* stats, float* workspace, void* stream);
/* y = forward output (needed for the ReLU mask when relu != 0).  dgamma/dbeta [c] are overwritten.
 * dx_add (may be NULL): [nv x c] added to dx in the same pass -- the gradient of the skip connection that forked off x
 * in a residual block (lattice_modules.py:1255-1358), which autograd would otherwise add with a kernel of its own. */
int ln_group_norm_bwd(const float* dy, const float* x, const float* y, const float* gamma, const float* stats,
                      const float* dx_add, int nv, const int* nv_dev, int c, int groups, int relu, float* dx, float* dgamma,
                      float* dbeta, float* workspace, void* stream);

/* ---- loss and optimizer of the training step -------------------------------------
 * 0.5 * Lovasz-softmax + 0.5 * NLL on log-probabilities (/root/reference/latticenet_py/ln_train.py:156-158,
 * lattice/lovasz_loss.py:41-72), value and d/d logp in one launch: one CTA per class sorts the errors |fg - p| in shared
 * memory.  n <= ln_seg_loss_max_points() (the ShapeNet-sized clouds this matters for; larger scans use the host-side
 * batched formulation).  logp [n x nr_classes], labels int64 [n]; ignore_index: that class is left out of both terms
 * (its points stay in the other classes' error vectors, as in the reference's class loop).
 * grad_lov [n x nr_classes] (out): unnormalised Lovasz gradient; acc_zeroed: 8 floats of zeroed scratch;
 * result (out, 4 floats): [0] loss, [1] classes present, [2] valid points. */
int ln_seg_loss_max_points(void);
int ln_seg_loss_fwd(const float* logp, const long long* labels, int n, int nr_classes, int ignore_index, float* grad_lov,
                    float* acc_zeroed, float* result, void* stream);
/* grad_logp = grad_loss[0] * d loss / d logp from the quantities ln_seg_loss_fwd left behind. */
int ln_seg_loss_bwd(const float* grad_lov, const long long* labels, const float* result, const float* grad_loss, int n,
                    int nr_classes, int ignore_index, float* grad_logp, void* stream);

/* Weight normalisation w = v * (g / ||v||_F) of a [rows x cols] tensor with one gain per column (gain_per_col != 0: the
 * lattice filter banks, g_dim = 1) or per row (Linear, g_dim = 0) -- /root/reference/latticenet_py/utils/utils.py:72-158 --
 * and its backward (dv, dg from dw), one launch each. */
int ln_weight_norm_fwd(const float* v, const float* g, int rows, int cols, int gain_per_col, float* w, void* stream);
int ln_weight_norm_bwd(const float* v, const float* g, const float* dw, int rows, int cols, int gain_per_col, float* dv, float* dg,
                       void* stream);

/* AdamW with amsgrad (ln_train.py:163-165) as ONE kernel over flat fp32 buffers of n elements (16-byte aligned),
 * torch.optim.AdamW's update order.  state: 2 floats, [0] step count (advanced here), [1] scratch (zero).
 * skip (may be NULL): device float, non-zero = leave parameters, moments and step count untouched.
 * grad_scale is multiplied into the gradients first (1 / world_size after a sum all-reduce). */
int ln_adamw_amsgrad(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, long long n,
                     float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale, float* state,
                     const float* skip, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LATTICE_B200_H_ */
