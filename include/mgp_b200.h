/*
 * mgp_b200.h -- C ABI of the B200-native IMGP hot path (libmgp_b200.so).
 *
 * One entry point per operation the reference reaches through faiss / torch_sparse / torch_scatter /
 * linear_operator on this path.  Reference citations are file:line under the
 * reference repository (nash169/manifold-gp).
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer on the current CUDA device unless its name ends in `_host`.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is enqueued on it;
 *     no call synchronises unless its documentation says so.
 *   - No entry point allocates device memory.  Calls that need scratch take `ws` / `ws_bytes`; the matching
 *     `*_ws_bytes` query returns the size to allocate (the Python host allocates it with torch).
 *   - Dense operands are row-major with an explicit leading dimension (`ld*`, in elements).
 *   - Hyper-parameters that live in torch Parameters (graph bandwidth eps, Matern shift 2nu/kappa^2, CG scalars)
 *     are passed as device scalars so that no call forces a device->host read.
 *   - Return value: 0 on success, a negative MGP_E* code otherwise; mgp_last_error() describes the last failure
 *     of the calling thread.  There is NO CPU fallback anywhere behind this ABI.
 *   - `_f32` / `_f64` suffix = arithmetic type of values and vectors.  Indices are int32 on the device
 *     (the reference's int64 COO is accepted / produced at the boundary calls only).
 */
#ifndef MGP_B200_H_
#define MGP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGP_OK 0
#define MGP_EINVAL (-1)    /* bad argument (shape, alignment, null pointer) */
#define MGP_ECUDA (-2)     /* a CUDA runtime call or kernel launch failed  */
#define MGP_EWORKSPACE (-3) /* workspace too small                          */
#define MGP_EUNSUPPORTED (-4)

const char* mgp_last_error(void);
/* Library / build identification: "mgp_b200 <version> sm_100a". */
const char* mgp_version(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches claim). */
int64_t mgp_launch_count(void);
void mgp_reset_launch_count(void);
/* Account for launches replayed from a captured CUDA graph (the host calls above only run at capture time). */
void mgp_add_launch_count(int64_t n);

/* ----------------------------------------------------------------------------------------------------------
 * (a2) Exact brute-force kNN, squared L2, ascending.  Replaces faiss Index{Flat,IVFFlat(nlist=1)}.search
 *      as called by NearestNeighbors.search -- manifold_gp/utils/nearest_neighbors.py:35-37 (index built at :17-33).
 *      db[n,d], q[nq,d] row-major fp32; dist2[nq,k] fp32 ascending; idx[nq,k] int64 (faiss' index type).
 *      Distances are sum_d (q_d - x_d)^2 evaluated in fp32 in ascending d (ties broken by ascending index);
 *      entries beyond n are (+inf, -1).
 * ---------------------------------------------------------------------------------------------------------- */
size_t mgp_knn_search_ws_bytes(int64_t n, int64_t nq, int32_t d, int32_t k);
int mgp_knn_search_f32(const float* db, int64_t n, const float* q, int64_t nq, int32_t d, int32_t k,
                       float* dist2, int64_t* idx, void* ws, size_t ws_bytes, void* stream);

/* Same contract and bit-identical results, for any d and k <= 48 (n >= 256): the q.x distance tiles are TF32 tcgen05.mma
 * contractions (TMA-fed, TMEM accumulators, 3xTF32 split) with a fused top-(k+margin) selection in the epilogue warps;
 * candidates are re-ranked with the exact fp32 form above and every query is certified (all discarded points provably
 * farther than its k-th neighbour) or re-searched exhaustively on the CUDA cores -- no host synchronisation.
 * Replaces the same faiss call (nearest_neighbors.py:37; faiss' own large-batch path is the |q|^2+|x|^2-2q.x BLAS form).
 * stats (device uint32[4]): [0] queries that needed the exhaustive re-search, [1] float bits of the largest
 * |approximate - exact| candidate distance seen, [2] queries processed.  `same` is a sizing hint only (the query-role copy
 * of the points differs from the database-role copy in the norm column and is always prepared).  Returns MGP_EUNSUPPORTED (nothing launched; ws_bytes query returns 0) outside that range. */
/* Minimum size K' of the exact re-rank window of the tensor-core search (0 = default k + 16 rounded to 32; 32; 64).  Process-wide;
 * affects mgp_knn_search_tc_ws_bytes and mgp_knn_search_tc_f32 alike.  NearestNeighbors.search raises it to 64 when a pilot of
 * 1024 queries shows the certificate failing (clouds whose neighbour distances sit below the 3xTF32 error band). */
int mgp_knn_tc_config(int32_t min_kp);
size_t mgp_knn_search_tc_ws_bytes(int64_t n, int64_t nq, int32_t d, int32_t k, int32_t same);
int mgp_knn_search_tc_f32(const float* db, int64_t n, const float* q, int64_t nq, int32_t d, int32_t k,
                          float* dist2, int64_t* idx, void* ws, size_t ws_bytes, uint32_t* stats, void* stream);

/* ----------------------------------------------------------------------------------------------------------
 * (a3) Directed kNN lists -> upper-triangular, lexicographically sorted, mean-coalesced COO.
 *      Replaces NearestNeighbors.graph -- nearest_neighbors.py:39-55 (torch_sparse.coalesce(op='mean') at :51).
 *      dist2/idx are [n,k] as returned by mgp_knn_search_f32; drop_first != 0 drops column 0 (:42-43).
 *      eidx is [2, cap] int64 with cap = n*(k - drop_first) (row 0 = eidx, row 1 = eidx + cap), eval[cap] fp32.
 *      The number of undirected edges M is written to *m_out (device int64); the first M columns are valid.
 * ---------------------------------------------------------------------------------------------------------- */
size_t mgp_graph_symmetrize_ws_bytes(int64_t n, int32_t k);
int mgp_graph_symmetrize_f32(const float* dist2, const int64_t* idx, int64_t n, int32_t k, int32_t drop_first,
                             int64_t* eidx, float* eval, int64_t cap, int64_t* m_out,
                             void* ws, size_t ws_bytes, void* stream);

/* ----------------------------------------------------------------------------------------------------------
 * Row-major directed structure of the symmetric graph, built once per graph (hyper-parameter independent).
 *      From the reference's COO (`idx[2,M]`, row<col as produced at nearest_neighbors.py:48-51; ld = distance in
 *      elements between the two rows of idx) build CSR over BOTH directions of every edge:
 *      rowptr[n+1], col[2M], eid[2M] (undirected edge id of each directed entry).  A diagonal COO entry (i,i)
 *      (nearest_neighbors.py duplicates quirk) contributes two entries to row i, as the two scatter/spmm passes
 *      of graph_laplacian_operator.py:63-69,118-119 do.  Column indices inside a row are ascending.
 * ---------------------------------------------------------------------------------------------------------- */
size_t mgp_csr_build_ws_bytes(int64_t n, int64_t m);
int mgp_csr_build(const int64_t* eidx, int64_t ld, int64_t m, int64_t n,
                  int32_t* rowptr, int32_t* col, int32_t* eid, void* ws, size_t ws_bytes, void* stream);
/* out[p] = val[eid[p]] : per-directed-entry copy of a per-edge array (squared distances). */
int mgp_gather_edge_f32(const float* val, const int32_t* eid, int64_t nnz, float* out, void* stream);
int mgp_gather_edge_f64(const double* val, const int32_t* eid, int64_t nnz, double* out, void* stream);

/* ----------------------------------------------------------------------------------------------------------
 * (a4) Laplacian values for one graph bandwidth eps.  Replaces the cached properties of GraphLaplacianOperator
 *      -- manifold_gp/operators/graph_laplacian_operator.py:52-106:
 *        W = exp(-d2/(4 eps^2)) (:56);  Dt_i = [1] + sum_j W_ij (:60-69);  At = W/(Dt_i Dt_j) (:75);
 *        D_i = [Dt_i^-2] + sum_j At_ij (:79-88);  diag_i = (1 - [Dt_i^-2]/D_i)/eps^2 | 1/eps^2 (:92-97);
 *        a_ij = At_ij/(sqrt(D_i) sqrt(D_j))/eps^2 (:106)            ([.] only with self_loops)
 *      Deterministic row sums over the CSR (no atomics, unlike scatter_add_).  d2csr is the per-directed-entry
 *      squared distance (mgp_gather_edge).  Outputs: deg_unnorm[n], deg[n], diag[n], a[nnz].
 *      `eps` is a device scalar.
 * ---------------------------------------------------------------------------------------------------------- */
int mgp_lap_values_f32(const int32_t* rowptr, const int32_t* col, const float* d2csr, int64_t n,
                       const float* eps, int32_t self_loops,
                       float* deg_unnorm, float* deg, float* diag, float* a, void* stream);
int mgp_lap_values_f64(const int32_t* rowptr, const int32_t* col, const double* d2csr, int64_t n,
                       const double* eps, int32_t self_loops,
                       double* deg_unnorm, double* deg, double* diag, double* a, void* stream);
/* One pass (1, 2 or 3) of the same build on the rows of a row-partitioned structure (SURVEY.md 8e "value build"): n = rows this
 * rank owns, `col` in the rank's extended numbering [own rows | halo rows], deg_unnorm / deg with n_ext entries.  The caller
 * fills the halo part of deg_unnorm (halo exchange) between pass 1 and 2 and the halo part of deg between pass 2 and 3 --
 * the two halo gathers that replace the reference's global scatter_add_ passes (graph_laplacian_operator.py:52-106). */
int mgp_lap_values_pass_f32(int32_t pass, const int32_t* rowptr, const int32_t* col, const float* d2csr, int64_t n, const float* eps,
                            int32_t self_loops, float* deg_unnorm, float* deg, float* diag, float* a, void* stream);
int mgp_lap_values_pass_f64(int32_t pass, const int32_t* rowptr, const int32_t* col, const double* d2csr, int64_t n, const double* eps,
                            int32_t self_loops, double* deg_unnorm, double* deg, double* diag, double* a, void* stream);

/* Backward of mgp_lap_values w.r.t. eps (what autograd through graph_laplacian_operator.py:52-106 yields for
 * raw_graphbandwidth; exercised by the reference's test_grad / test_ml, test/_test_functions.py:59-104):
 *   *g_eps = sum_p g_a[p] da_p/deps + sum_i ( g_diag_i ddiag_i/deps + g_deg_unnorm_i dDt_i/deps + g_deg_i dD_i/deps )
 * computed with forward-mode tangents (three streaming passes, deterministic reduction).  g_deg_unnorm / g_deg may be
 * NULL.  ws: mgp_lap_values_grad_ws_bytes(n) bytes, first 256 bytes zero on first use. */
size_t mgp_lap_values_grad_ws_bytes(int64_t n);
int mgp_lap_values_grad_f32(const int32_t* rowptr, const int32_t* col, const float* d2csr, int64_t n, const float* eps,
                            int32_t self_loops, const float* deg_unnorm, const float* deg, const float* diag,
                            const float* a, const float* g_deg_unnorm, const float* g_deg, const float* g_diag,
                            const float* g_a, float* g_eps, void* ws, void* stream);
int mgp_lap_values_grad_f64(const int32_t* rowptr, const int32_t* col, const double* d2csr, int64_t n, const double* eps,
                            int32_t self_loops, const double* deg_unnorm, const double* deg, const double* diag,
                            const double* a, const double* g_deg_unnorm, const double* g_deg, const double* g_diag,
                            const double* g_a, double* g_eps, void* ws, void* stream);

/* ----------------------------------------------------------------------------------------------------------
 * (a5,a8) Fused Laplacian / Matern-step SpMM:   Y = post .* ( (diag + shift) .* (pre .* X)  -  A (pre .* X) )
 *      Replaces GraphLaplacianOperator._matmul -- graph_laplacian_operator.py:108-124 (two torch_sparse.spmm + diagonal
 *      + D^{+-1/2} scalings) and one step of PrecisionMaternOperator._matmul -- precision_matern_operator.py:26-37
 *      (out <- (out + c L out)/c  ==  (1/c + L) out, i.e. shift = 2 nu / kappa^2).
 *      A = (rowptr, col, a) from mgp_lap_values; `shift` device scalar or NULL (0); `pre`, `post` [n] or NULL (1).
 *      X [n, ncols] ld ldx; Y [n, ncols] ld ldy; X and Y must not alias.
 *      Optional fused reduction (CG's p^T A p, Lanczos' alpha): if dot_out != NULL, dot_out[c] = sum_i dot_with[i,c]*Y[i,c]
 *      (dot_with ld = ldx), accumulated deterministically; needs `dot_ws` of mgp_lap_spmm_dot_ws_bytes(n,ncols).
 * ---------------------------------------------------------------------------------------------------------- */
size_t mgp_lap_spmm_dot_ws_bytes(int64_t n, int32_t ncols);
int mgp_lap_spmm_f32(const int32_t* rowptr, const int32_t* col, const float* a, const float* diag,
                     const float* shift, const float* pre, const float* post,
                     const float* x, int64_t ldx, float* y, int64_t ldy, int64_t n, int32_t ncols,
                     const float* dot_with, float* dot_out, void* dot_ws, void* stream);
int mgp_lap_spmm_f64(const int32_t* rowptr, const int32_t* col, const double* a, const double* diag,
                     const double* shift, const double* pre, const double* post,
                     const double* x, int64_t ldx, double* y, int64_t ldy, int64_t n, int32_t ncols,
                     const double* dot_with, double* dot_out, void* dot_ws, void* stream);

/* Same product, v2 "tile-compacted" kernel: rows are processed in tiles of `tile_rows` (= 128) consecutive rows whose
 * distinct X rows (own rows + a halo list) are staged in shared memory once per tile; per nonzero the kernel streams a
 * 16-bit tile-local column index (lcol: own rows 0..tile_rows-1, halo rows tile_rows + position in the tile's halo list)
 * and the value with 128-bit loads.  Built for graphs whose rows were reordered along a space-filling curve
 * (manifold_gp_b200/graph.py); lmax = max over tiles of tile_rows + halo length, nzmax = max nonzeros of a tile.
 * `a` and `lcol` must be readable 8 entries past nnz.  xmap / ymap (int32[n], optional) give the caller's row of
 * structure row i for X (and dot_with) / Y, so vectors in the caller's order need no separate permutation pass.
 * Returns MGP_EUNSUPPORTED if a tile does not fit in shared memory (use mgp_lap_spmm). */
int mgp_lap_spmm_tiled_f32(const int32_t* rowptr, const uint16_t* lcol, const float* a, const float* diag,
                           const int32_t* halo_ptr, const int32_t* halo_col, int32_t tile_rows, int32_t lmax, int32_t nzmax,
                           const float* shift, const float* pre, const float* post, const int32_t* xmap,
                           const int32_t* ymap, const float* x, int64_t ldx, float* y, int64_t ldy, int64_t n,
                           int32_t ncols, const float* dot_with, float* dot_out, void* dot_ws, void* stream);
int mgp_lap_spmm_tiled_f64(const int32_t* rowptr, const uint16_t* lcol, const double* a, const double* diag,
                           const int32_t* halo_ptr, const int32_t* halo_col, int32_t tile_rows, int32_t lmax, int32_t nzmax,
                           const double* shift, const double* pre, const double* post, const int32_t* xmap,
                           const int32_t* ymap, const double* x, int64_t ldx, double* y, int64_t ldy, int64_t n,
                           int32_t ncols, const double* dot_with, double* dot_out, void* dot_ws, void* stream);

/* Same product, v4 "pipelined" kernel (lap_spmm_pipe.cu): the tile structure of mgp_lap_spmm_tiled with PADDED entry
 * streams -- prowptr[n+1] addresses a stream in which every row holds a multiple of 4 entries (padding entries: plcol =
 * the row's own tile-local index, value 0) and every tile of 128 rows starts at a multiple of 8 entries; plcol and ap
 * must be readable 8 entries past prowptr[n].  pnzmax = max padded entries of a tile (multiple of 8).  No `pre` scaling.
 * mgp_lap_pad_values copies a CSR-ordered value array `a` (per bandwidth) into the padded layout.
 * Replaces the two torch_sparse.spmm calls + diagonal of graph_laplacian_operator.py:117-119 and one step of
 * precision_matern_operator.py:28-32.  Returns MGP_EUNSUPPORTED (nothing launched) when ncols / alignment / shared
 * memory do not qualify: the caller then uses mgp_lap_spmm_tiled or mgp_lap_spmm. */
int mgp_lap_pad_values_f32(const int32_t* rowptr, const int32_t* prowptr, const float* a, int64_t n, float* ap, void* stream);
int mgp_lap_pad_values_f64(const int32_t* rowptr, const int32_t* prowptr, const double* a, int64_t n, double* ap, void* stream);
int mgp_lap_spmm_pipe_f32(const int32_t* prowptr, const uint16_t* plcol, const float* ap, const float* diag,
                          const int32_t* halo_ptr, const int32_t* halo_col, int32_t tile_rows, int32_t lmax, int32_t pnzmax,
                          const float* shift, const float* post, const int32_t* xmap, const int32_t* ymap, const float* x,
                          int64_t ldx, float* y, int64_t ldy, int64_t n, int32_t ncols, const float* dot_with, float* dot_out,
                          void* dot_ws, void* stream);
int mgp_lap_spmm_pipe_f64(const int32_t* prowptr, const uint16_t* plcol, const double* ap, const double* diag,
                          const int32_t* halo_ptr, const int32_t* halo_col, int32_t tile_rows, int32_t lmax, int32_t pnzmax,
                          const double* shift, const double* post, const int32_t* xmap, const int32_t* ymap, const double* x,
                          int64_t ldx, double* y, int64_t ldy, int64_t n, int32_t ncols, const double* dot_with,
                          double* dot_out, void* dot_ws, void* stream);

/* Same product, v5 "warp-interleaved" kernel (lap_spmm_wi.cu) for 64-byte rows of X (ncols a multiple of 16 fp32 / 8 fp64
 * columns).  Tiles of 128 rows as above; the entry streams are laid out in consumption order: warp block b = 16 * tile + w
 * owns positions [wptr[b], wptr[b+1]) (whole 32-entry steps) and position wptr[b] + 32 t + lane holds nonzero
 * 4t + (lane & 3) of row 128 tile + 8 w + (lane >> 2); padding entries have value 0.  wcol / aw must be readable 64
 * entries past wptr[16 ntiles]; wnzmax = max entries of a tile.  Halo lists: hcol[hptr[t] .. hptr[t+1]) = the out-of-tile
 * rows of X tile t reads (tile-local column 128 + position), every list padded to a multiple of 4 ids; hmax = longest
 * padded list; lmax >= 128 + hmax.  The kernel bulk-copies its metadata in chunks of 32 tiles, so wptr must be readable
 * up to index 512 ceil(ntiles/32) + 4 and hptr up to 32 ceil(ntiles/32) + 4, and wptr / hptr / hcol / wcol / aw must be
 * 16-byte aligned.  mgp_lap_wi_values copies a CSR-ordered value array into the stream layout.
 * Multi-GPU (row-partitioned): peer_x = DEVICE array of npeers pointers to the ranks' X blocks (peer-mapped, same ldx);
 * halo ids are then (rank << 26) | row-in-that-rank's-X and the halo rows are fetched straight from the owners over
 * NVLink (no pack kernel, no all-to-all).  Producer and consumer launches must be separated by a cross-GPU sync: either
 * mgp_peer_barrier before the call, or the fused form -- sync_flags = DEVICE array of the ranks' flag arrays uint32[npeers]
 * (zeroed + barrier once per solve), sync_epoch -> a device scalar e: block 0 publishes e + 1 to every peer at kernel start
 * and the producer warps wait for all ranks' flags before fetching the first remote row (local copies start at once).
 * Replaces graph_laplacian_operator.py:117-119 / precision_matern_operator.py:28-32 like the kernels above.
 * Returns MGP_EUNSUPPORTED (nothing launched) when the call does not qualify. */
/* Single-column SpMV on the same warp-interleaved streams (lap_spmv_tile.cu): 6 bytes per nonzero streamed, the tile's slice
 * of x in shared memory.  Same operation and reference lines as above for one right-hand side (Lanczos, single-RHS CG).
 * x / y: column vectors with row strides ldx / ldy; optional fused dot_out[0] = dot_with^T y (dot_with strided like x; dot_ws
 * from mgp_lap_spmm_dot_ws_bytes).  MGP_EUNSUPPORTED when tile_rows != 128, the halo does not fit, or the dot workspace cannot
 * hold one partial per tile. */
int mgp_lap_spmv_tile_f32(const int32_t* wptr, const uint16_t* wcol, const float* aw, const float* diag, const int32_t* hptr,
                          const int32_t* hcol, int32_t tile_rows, int32_t hmax, const float* shift, const float* post,
                          const int32_t* xmap, const int32_t* ymap, const float* x, int64_t ldx, float* y, int64_t ldy, int64_t n,
                          const float* dot_with, float* dot_out, void* dot_ws, void* stream);
int mgp_lap_spmv_tile_f64(const int32_t* wptr, const uint16_t* wcol, const double* aw, const double* diag, const int32_t* hptr,
                          const int32_t* hcol, int32_t tile_rows, int32_t hmax, const double* shift, const double* post,
                          const int32_t* xmap, const int32_t* ymap, const double* x, int64_t ldx, double* y, int64_t ldy, int64_t n,
                          const double* dot_with, double* dot_out, void* dot_ws, void* stream);
int mgp_lap_wi_values_f32(const int32_t* rowptr, const int32_t* wptr, const float* a, int64_t n, float* aw, void* stream);
int mgp_lap_wi_values_f64(const int32_t* rowptr, const int32_t* wptr, const double* a, int64_t n, double* aw, void* stream);
int mgp_lap_spmm_wi_f32(const int32_t* wptr, const uint16_t* wcol, const float* aw, const float* diag, const int32_t* hptr,
                        const int32_t* hcol, int32_t tile_rows, int32_t lmax, int32_t wnzmax, int32_t hmax, const float* shift,
                        const float* post, const int32_t* xmap, const int32_t* ymap, const float* x, int64_t ldx, float* y,
                        int64_t ldy, int64_t n, int32_t ncols, const float* dot_with, float* dot_out, void* dot_ws,
                        const void* peer_x, int32_t npeers, int32_t rank, const void* sync_flags, const float* sync_epoch,
                        void* stream);
int mgp_lap_spmm_wi_f64(const int32_t* wptr, const uint16_t* wcol, const double* aw, const double* diag,
                        const int32_t* hptr, const int32_t* hcol, int32_t tile_rows, int32_t lmax, int32_t wnzmax, int32_t hmax,
                        const double* shift, const double* post, const int32_t* xmap, const int32_t* ymap, const double* x,
                        int64_t ldx, double* y, int64_t ldy, int64_t n, int32_t ncols, const double* dot_with, double* dot_out,
                        void* dot_ws, const void* peer_x, int32_t npeers, int32_t rank, const void* sync_flags,
                        const double* sync_epoch, void* stream);

/* Extended form of mgp_lap_spmm_wi for solver loops (CG past convergence, fused cross-GPU sync points).  All members optional
 * (NULL / 0 = off); the flag and reduction tables are DEVICE arrays of npeers pointers to the ranks' peer-mapped buffers:
 *   done_flag      device scalar (T): the launch is a no-op when it is non-zero (every CG kernel is, once converged);
 *   wait_flags     producers wait -- lazily, at the first halo row that lives on another rank -- until
 *                  wait_flags[rank][src] >= epoch for every src; nothing is published at kernel start;
 *   publish_at_start  with wait_flags: block 0 ALSO publishes wait_flags[dst][rank] = epoch at kernel start ("everything enqueued
 *                  before this launch is complete on my side", the semantics of mgp_lap_spmm_wi's sync_flags) -- the system-scope
 *                  fence of the publish then overlaps the launch's own work instead of ending the producing kernel (measured at
 *                  125k rows per rank: 4 us per launch);
 *   publish_flags  after ALL rows of Y are written (last block to finish, last column pass): publish_flags[dst][rank] = epoch
 *                  ("my Y is complete") -- the consumer of Y on another rank waits on it with wait_flags;
 *   ticket         device uint32 counter for that completion ticket (zeroed once; resets itself);
 *   red_ptrs/red_flags/ship_extra/ship_ncols  with dot_out: the block that completes the local dot products ships
 *                  dot_out[0..ship_ncols) (kind 0) and ship_extra[0..ship_ncols) (kind 1, may be NULL) to
 *                  red_ptrs[dst][kind][epoch & 1][rank][128] of every rank, then red_flags[dst][rank] = epoch
 *                  (first half of a fused all-reduce; mgp_cg_peer_cgstep is the second half).
 *   ep_coef / ep_add  epilogue algebra of the wrappers folded into the launch: Y <- coef * Y with coef = *ep_coef (device scalar;
 *                  the dot epilogue then sees the scaled Y), and with ep_add = 1: Y <- ADD + coef * Y where ADD is passed in the
 *                  dot_with argument (rows strided like X; no dot product on such a launch).  Replaces the elementwise passes of
 *                  scale_wrapper_operator.py:27-28 and noise_wrapper_operator.py:21-22 (x - s Q x).
 * epoch = (unsigned)*sync_epoch + 1 (the CG iteration counter in the solver state).  No reference counterpart. */
typedef struct mgp_wi_ext {
  const void* done_flag;
  const void* wait_flags;
  const void* publish_flags;
  void* ticket;
  const void* red_ptrs;
  const void* red_flags;
  const void* ship_extra;
  int32_t ship_ncols;
  int32_t ep_add;
  const void* ep_coef;
  int32_t publish_at_start;
  int32_t reserved;
  const uint8_t* pair_rows;   /* non-NULL (fp32 only): wptr / wcol / aw are the PAIRED-ROW streams of graph.pair_streams -- a slot of 8 lanes
                                 walks the union list of two spatially adjacent rows, aw holds two values per entry (mgp_lap_pair_values),
                                 pair_rows[128 tile + 2 pair + half] is the tile-local row that position outputs */
} mgp_wi_ext;
/* out[i] = src[i] >= 0 ? a[src[i]] : 0 -- the value stream of the paired layout (two sources per stream entry), once per bandwidth. */
int mgp_lap_pair_values_f32(const int32_t* src, const float* a, int64_t count, float* out, void* stream);
int mgp_lap_pair_values_f64(const int32_t* src, const double* a, int64_t count, double* out, void* stream);
int mgp_lap_spmm_wi_ex_f32(const int32_t* wptr, const uint16_t* wcol, const float* aw, const float* diag, const int32_t* hptr,
                           const int32_t* hcol, int32_t tile_rows, int32_t lmax, int32_t wnzmax, int32_t hmax, const float* shift,
                           const float* post, const int32_t* xmap, const int32_t* ymap, const float* x, int64_t ldx, float* y,
                           int64_t ldy, int64_t n, int32_t ncols, const float* dot_with, float* dot_out, void* dot_ws,
                           const void* peer_x, int32_t npeers, int32_t rank, const float* sync_epoch, const mgp_wi_ext* ext,
                           void* stream);
int mgp_lap_spmm_wi_ex_f64(const int32_t* wptr, const uint16_t* wcol, const double* aw, const double* diag, const int32_t* hptr,
                           const int32_t* hcol, int32_t tile_rows, int32_t lmax, int32_t wnzmax, int32_t hmax, const double* shift,
                           const double* post, const int32_t* xmap, const int32_t* ymap, const double* x, int64_t ldx, double* y,
                           int64_t ldy, int64_t n, int32_t ncols, const double* dot_with, double* dot_out, void* dot_ws,
                           const void* peer_x, int32_t npeers, int32_t rank, const double* sync_epoch, const mgp_wi_ext* ext,
                           void* stream);

/* ----------------------------------------------------------------------------------------------------------
 * Backward of the SpMM w.r.t. the matrix entries (what autograd through torch_sparse.spmm computes for `value`,
 * graph_laplacian_operator.py:117-119, reached via linear_operator's _bilinear_derivative):
 *      g_a[p]   = - sum_c L[i,c] * R[col_p,c]      (i = row of entry p)
 *      g_diag[i] =   sum_c L[i,c] * R[i,c]
 * with L = post .* grad_Y and R = pre .* X.   g_a [nnz], g_diag [n].
 * ---------------------------------------------------------------------------------------------------------- */
int mgp_lap_sddmm_f32(const int32_t* rowptr, const int32_t* col, const float* pre, const float* post,
                      const float* gy, int64_t ldgy, const float* x, int64_t ldx, int64_t n, int32_t ncols,
                      float* g_a, float* g_diag, void* stream);
int mgp_lap_sddmm_f64(const int32_t* rowptr, const int32_t* col, const double* pre, const double* post,
                      const double* gy, int64_t ldgy, const double* x, int64_t ldx, int64_t n, int32_t ncols,
                      double* g_a, double* g_diag, void* stream);

/* ----------------------------------------------------------------------------------------------------------
 * (a17) Batched CG vector kernels (mBCG of linear_operator.utils.linear_cg, third party; call sites
 *      utils/train_model.py:55,67-68, precision_matern_operator.py:53, schur_complement_operator.py:28).
 *      The host loop calls the operator's matvec (mgp_lap_spmm chain) and these fused updates; all scalars stay on
 *      the device in `state` (layout below, in elements of the value type, C = ncols):
 *        [0,C) rhs_norm  [C,2C) rz (=r^T r)  [2C,3C) pAp  [3C,4C) alpha  [4C,5C) beta  [5C,6C) resid_norm
 *        [6C,7C) flags (bit0 rhs_is_zero, bit1 has_converged)   [7C] mean residual norm  [7C+1] done flag
 *        [7C+2] iteration counter   [7C+3] tolerance  [7C+4] eps  [7C+5] stop_updating_after
 *        [7C+6] min_iter (= min(10, max_iter-1))  [7C+7] n_tridiag_min (iterations the tridiagonal still needs)
 *      hist [max_hist, 2, C]: alpha_k, beta_k per iteration (the Lanczos tridiagonals are assembled from these).
 *      Every kernel returns immediately when the done flag is set, so the host may enqueue iterations in chunks
 *      and poll the flag; reductions are deterministic (per-block partials reduced in fixed order by the last block).
 * ---------------------------------------------------------------------------------------------------------- */
size_t mgp_cg_state_elems(int32_t ncols);
size_t mgp_cg_ws_bytes(int64_t n, int32_t ncols);
/* r = b / |b|, p = r, x = 0, rz = |r|^2; fills state.  b [n,ncols] ld ldb; x,r,p [n,ncols] ld ld. */
int mgp_cg_init_f32(const float* b, int64_t ldb, float* x, float* r, float* p, int64_t ld, int64_t n, int32_t ncols,
                    float tolerance, float eps, float stop_updating_after, int32_t max_iter, int32_t n_tridiag_iter,
                    float* state, void* ws, void* stream);
int mgp_cg_init_f64(const double* b, int64_t ldb, double* x, double* r, double* p, int64_t ld, int64_t n, int32_t ncols,
                    double tolerance, double eps, double stop_updating_after, int32_t max_iter, int32_t n_tridiag_iter,
                    double* state, void* ws, void* stream);
/* pAp[c] = sum_i p*v (skipped when the SpMM epilogue already produced state.pAp: have_pap != 0), then
 * alpha = rz/pAp with the safe-division / convergence masks. */
int mgp_cg_alpha_f32(const float* p, const float* v, int64_t ld, int64_t n, int32_t ncols, int32_t have_pap,
                     float* state, void* ws, void* stream);
int mgp_cg_alpha_f64(const double* p, const double* v, int64_t ld, int64_t n, int32_t ncols, int32_t have_pap,
                     double* state, void* ws, void* stream);
/* x += alpha p; r -= alpha v; rz' = |r|^2; beta = rz'/rz; residual norms, convergence flags, done flag,
 * iteration counter, hist[k] = (alpha, beta). */
int mgp_cg_update_f32(float* x, float* r, const float* p, const float* v, int64_t ld, int64_t n, int32_t ncols,
                      float* state, float* hist, int32_t max_hist, void* ws, void* stream);
int mgp_cg_update_f64(double* x, double* r, const double* p, const double* v, int64_t ld, int64_t n, int32_t ncols,
                      double* state, double* hist, int32_t max_hist, void* ws, void* stream);

/* Split form of the update (8 instead of 9 vector passes per iteration; same arithmetic and scalars):
 *   mgp_cg_rupdate : r -= alpha v, rz' = |r|^2, then the scalar step (beta, norms, flags, history) -- or, with rbuf != NULL,
 *                    the column sums are exported to rbuf for an all-reduce and mgp_cg_dist_scalars(what=2) finishes;
 *   mgp_cg_pxupdate: x += alpha p, p = r + beta p (alpha, beta from state).  Runs once more after the iteration that set
 *                    the done flag (the x update of that iteration) and is a no-op afterwards.
 * Iteration: matvec (+ p^T A p) -> mgp_cg_alpha -> mgp_cg_rupdate -> mgp_cg_pxupdate. */
int mgp_cg_rupdate_f32(float* r, const float* v, int64_t ld, int64_t n, int32_t ncols, float* state, float* hist,
                       int32_t max_hist, float* rbuf, void* ws, void* stream);
int mgp_cg_rupdate_f64(double* r, const double* v, int64_t ld, int64_t n, int32_t ncols, double* state, double* hist,
                       int32_t max_hist, double* rbuf, void* ws, void* stream);
int mgp_cg_pxupdate_f32(float* x, float* p, const float* r, int64_t ld, int64_t n, int32_t ncols, const float* state, void* stream);
int mgp_cg_pxupdate_f64(double* x, double* p, const double* r, int64_t ld, int64_t n, int32_t ncols, const double* state, void* stream);
/* Multi-GPU split of the two reductions (rows partitioned across ranks): the kernels export their LOCAL column sums to
 * rbuf[ncols]; the host all-reduces rbuf (NCCL) and mgp_cg_dist_scalars finishes the scalar bookkeeping on every rank
 * (what = 0: right-hand-side norms, 1: initial residual / state, 2: one iteration's beta, norms, flags, done).
 * Per iteration: [halo exchange + SpMM chain, last launch with the p^T A p epilogue into state.pAp] -> all-reduce(pAp)
 * -> mgp_cg_dist_update -> all-reduce(rbuf) -> mgp_cg_dist_scalars(what=2) -> mgp_cg_pupdate. */
int mgp_cg_dist_norm2_f32(const float* b, int64_t ldb, int64_t n, int32_t ncols, float* state, float* rbuf, void* ws, void* stream);
int mgp_cg_dist_norm2_f64(const double* b, int64_t ldb, int64_t n, int32_t ncols, double* state, double* rbuf, void* ws, void* stream);
int mgp_cg_dist_init_f32(const float* b, int64_t ldb, float* x, float* r, float* p, int64_t ld, int64_t n, int32_t ncols,
                         float* state, float* rbuf, void* ws, void* stream);
int mgp_cg_dist_init_f64(const double* b, int64_t ldb, double* x, double* r, double* p, int64_t ld, int64_t n, int32_t ncols,
                         double* state, double* rbuf, void* ws, void* stream);
int mgp_cg_dist_update_f32(float* x, float* r, const float* p, const float* v, int64_t ld, int64_t n, int32_t ncols,
                           float* state, float* rbuf, void* ws, void* stream);
int mgp_cg_dist_update_f64(double* x, double* r, const double* p, const double* v, int64_t ld, int64_t n, int32_t ncols,
                           double* state, double* rbuf, void* ws, void* stream);
int mgp_cg_dist_scalars_f32(float* state, const float* rbuf, int32_t ncols, int32_t what, float tolerance, float eps,
                            float stop_updating_after, int32_t max_iter, int32_t n_tridiag_iter, float* hist, int32_t max_hist,
                            void* stream);
int mgp_cg_dist_scalars_f64(double* state, const double* rbuf, int32_t ncols, int32_t what, double tolerance, double eps,
                            double stop_updating_after, int32_t max_iter, int32_t n_tridiag_iter, double* hist, int32_t max_hist,
                            void* stream);
/* p = r + beta p */
/* Peer-memory (NVLink P2P) variants for the multi-GPU path.  red_ptrs / flag_ptrs: DEVICE arrays of `world` pointers to
 * every rank's reduction buffer T[2][world][128] and flag array uint32[world] (peer-mapped symmetric allocations, zeroed
 * once); epoch_ctr: this rank's private uint32 counter in device memory (zeroed once; advanced by the kernels, so a
 * captured CUDA graph can be replayed).  All ranks must enqueue the same sequence of these calls per (flag, counter) set.
 *   mgp_cg_peer_scalars: all-reduce of rbuf[ncols] over the ranks (stores into the peers' buffers, epoch flags, sum in rank
 *                        order => bit-identical totals everywhere) fused with the scalar step `what` (0/1/2 as
 *                        mgp_cg_dist_scalars, 3 = store the totals as p^T A p).  Replaces an NCCL all-reduce + a kernel.
 *   mgp_peer_barrier:    cross-GPU barrier on the stream (prior writes of every rank visible to every rank). */
int mgp_cg_peer_scalars_f32(float* state, const float* rbuf, int32_t ncols, int32_t what, float tolerance, float eps,
                            float stop_updating_after, int32_t max_iter, int32_t n_tridiag_iter, float* hist, int32_t max_hist,
                            void* red_ptrs, void* flag_ptrs, void* epoch_ctr, int32_t rank, int32_t world, void* stream);
int mgp_cg_peer_scalars_f64(double* state, const double* rbuf, int32_t ncols, int32_t what, double tolerance, double eps,
                            double stop_updating_after, int32_t max_iter, int32_t n_tridiag_iter, double* hist, int32_t max_hist,
                            void* red_ptrs, void* flag_ptrs, void* epoch_ctr, int32_t rank, int32_t world, void* stream);
/* Fully fused iteration tail of the peer-memory path (replaces mgp_cg_peer_scalars(3) + mgp_cg_rupdate + mgp_cg_peer_scalars(2)
 * + mgp_cg_pxupdate): sync points are keyed by the iteration number in `state` (epoch = iterations done + 1), so the flag
 * arrays flag2 / flag3 (uint32[world] per rank, peer-mapped) must be ZEROED on every rank, followed by a barrier, before each
 * solve.  red_ptrs[r] -> T[2][2][world][128].  pap_local: this rank's p^T A p partial sums (dot epilogue of the last SpMM).
 * ld must be a power of two <= 128 and the vectors 16-byte aligned, else MGP_EUNSUPPORTED (use the unfused calls). */
int mgp_cg_peer_rupdate_f32(float* r, const float* v, int64_t ld, int64_t n, int32_t ncols, float* state, const float* pap_local,
                            void* ws, void* red_ptrs, void* flag2_ptrs, void* flag3_ptrs, int32_t rank, int32_t world, void* stream);
int mgp_cg_peer_rupdate_f64(double* r, const double* v, int64_t ld, int64_t n, int32_t ncols, double* state, const double* pap_local,
                            void* ws, void* red_ptrs, void* flag2_ptrs, void* flag3_ptrs, int32_t rank, int32_t world, void* stream);
int mgp_cg_peer_pxupdate_f32(float* x, float* p, const float* r, int64_t ld, int64_t n, int32_t ncols, float* state, float* hist,
                             int32_t max_hist, void* ws, void* red_ptrs, void* flag3_ptrs, int32_t rank, int32_t world, void* stream);
int mgp_cg_peer_pxupdate_f64(double* x, double* p, const double* r, int64_t ld, int64_t n, int32_t ncols, double* state, double* hist,
                             int32_t max_hist, void* ws, void* red_ptrs, void* flag3_ptrs, int32_t rank, int32_t world, void* stream);
/* Single-reduction (Chronopoulos-Gear) iteration of the peer-memory path: an iteration is [mgp_lap_spmm_wi_ex x nu with r as the
 * source, the last launch shipping this rank's (r.w, |r|^2) partials] + this ONE vector kernel, which waits for all ranks'
 * partials (dflag), forms beta and alpha, applies p = r + beta p, s = w + beta s, x += alpha p, r -= alpha s, stores the local
 * |r_new|^2 in gamma_loc[ncols], advances the scalar state / history like mgp_cg_rupdate and publishes "r complete"
 * (iteration + 2) in rflag for the peers' next SpMM.  Same iterates, masks and stopping rules as linear_cg; convergence is noticed
 * one matvec later (the tested residual norm is the one entering the iteration) and no update is applied then.  s must be
 * zeroed, rflag published as 1 (mgp_peer_publish) and gamma_loc filled with the local |r_0|^2 before the first iteration.
 * delta_loc != NULL: block 0 of THIS kernel ships (delta_loc, gamma_loc) to the peers first (the SpMM launch then needs no
 * red_ptrs); rflag_ptrs == NULL: "r complete" is published by the next SpMM launch at its start (mgp_wi_ext.publish_at_start).
 * red_ptrs[r] -> T[2][2][world][128]; flags zeroed + barrier before each solve.  ld as mgp_cg_peer_rupdate. */
int mgp_cg_peer_cgstep_f32(float* x, float* r, float* p, float* s, const float* w, int64_t ld, int64_t n, int32_t ncols, float* state,
                           float* hist, int32_t max_hist, void* ws, float* gamma_loc, const float* delta_loc, void* red_ptrs, void* dflag_ptrs,
                           void* rflag_ptrs, int32_t rank, int32_t world, void* stream);
int mgp_cg_peer_cgstep_f64(double* x, double* r, double* p, double* s, const double* w, int64_t ld, int64_t n, int32_t ncols, double* state,
                           double* hist, int32_t max_hist, void* ws, double* gamma_loc, const double* delta_loc, void* red_ptrs, void* dflag_ptrs,
                           void* rflag_ptrs, int32_t rank, int32_t world, void* stream);
/* flags[dst][rank] = value on every rank, after everything enqueued before on this stream (no waiting). */
int mgp_peer_publish(void* flag_ptrs, uint32_t value, int32_t rank, int32_t world, void* stream);
int mgp_peer_barrier(void* flag_ptrs, void* epoch_ctr, int32_t rank, int32_t world, void* stream);
int mgp_cg_pupdate_f32(float* p, const float* r, int64_t ld, int64_t n, int32_t ncols, const float* state, void* stream);
int mgp_cg_pupdate_f64(double* p, const double* r, int64_t ld, int64_t n, int32_t ncols, const double* state, void* stream);
/* out[i,c] = x[i,c] * rhs_norm[c]  (un-normalise; out ld ldo) */
int mgp_cg_finalize_f32(const float* x, int64_t ld, float* out, int64_t ldo, int64_t n, int32_t ncols,
                        const float* state, void* stream);
int mgp_cg_finalize_f64(const double* x, int64_t ld, double* out, int64_t ldo, int64_t n, int32_t ncols,
                        const double* state, void* stream);

/* ----------------------------------------------------------------------------------------------------------
 * (a7,a17) Lanczos kernels (linear_operator.utils.lanczos.lanczos_tridiag, third party; call site
 *      graph_laplacian_operator.py:132-135).  Q is stored vector-major: Q[j] is a contiguous length-n vector (ldq >= n).
 *      mgp_lanczos_reorth:  c = Q[0:j]^T r ;  r -= Q[0:j] c   (one full re-orthogonalisation pass; c [j] is also
 *                           returned so the host can test max|c| against tol), then nrm2[0] = |r|^2.
 *      mgp_lanczos_dots:    c = Q[0:j]^T r only.
 *                           j = 0 is allowed (norm only).  Running one pass against ALL previous vectors also removes the
 *                           alpha_k q_k and beta_{k-1} q_{k-1} terms of the three-term recurrence: alpha_k = c[k].
 *      mgp_lanczos_normalize: q_out = r / sqrt(*nrm2); *beta_out = sqrt(*nrm2)   (beta_out may be NULL).
 * ---------------------------------------------------------------------------------------------------------- */
size_t mgp_lanczos_ws_bytes(int64_t n, int32_t j);
int mgp_lanczos_reorth_f32(const float* q, int64_t ldq, int32_t j, float* r, int64_t n, float* c, float* nrm2,
                           void* ws, void* stream);
int mgp_lanczos_reorth_f64(const double* q, int64_t ldq, int32_t j, double* r, int64_t n, double* c, double* nrm2,
                           void* ws, void* stream);
/* r -= Q[0:j] c (c given), nrm2[0] = |r|^2 -- the second half of mgp_lanczos_reorth. */
int mgp_lanczos_axpy_f32(const float* q, int64_t ldq, int32_t j, float* r, int64_t n, const float* c, float* nrm2,
                         void* ws, void* stream);
int mgp_lanczos_axpy_f64(const double* q, int64_t ldq, int32_t j, double* r, int64_t n, const double* c, double* nrm2,
                         void* ws, void* stream);
int mgp_lanczos_normalize_f32(const float* r, int64_t n, const float* nrm2, float* q_out, float* beta_out, void* stream);
int mgp_lanczos_normalize_f64(const double* r, int64_t n, const double* nrm2, double* q_out, double* beta_out, void* stream);
int mgp_lanczos_dots_f32(const float* q, int64_t ldq, int32_t j, const float* r, int64_t n, float* c, void* ws, void* stream);
int mgp_lanczos_dots_f64(const double* q, int64_t ldq, int32_t j, const double* r, int64_t n, double* c, void* ws, void* stream);

/* ----------------------------------------------------------------------------------------------------------
 * (a15) Out-of-sample (Nystrom) extension: an ELL SpMM with fixed row length k.
 *      Replaces GraphLaplacianOperator.out_of_sample -- graph_laplacian_operator.py:146-157:
 *        w = exp(-d2/(4 eps^2)); w /= Dt[idx] * rowsum(w); symmetric: w /= sqrt(D[idx]) * sqrt(rowsum(w));
 *        randomwalk: w /= rowsum(w);  out[q,:] = sum_k w[q,k] phi[idx[q,k],:]
 *      d2 [nq,k], idx [nq,k] int64, phi [n,m] ld ldphi, out [nq,m] ld ldo.  normalization: 0 symmetric, 1 randomwalk.
 * ---------------------------------------------------------------------------------------------------------- */
int mgp_out_of_sample_f32(const float* d2, const int64_t* idx, int64_t nq, int32_t k, const float* eps,
                          const float* deg_unnorm, const float* deg, int32_t normalization,
                          const float* phi, int64_t ldphi, int32_t m, float* out, int64_t ldo, void* stream);
int mgp_out_of_sample_f64(const double* d2, const int64_t* idx, int64_t nq, int32_t k, const double* eps,
                          const double* deg_unnorm, const double* deg, int32_t normalization,
                          const double* phi, int64_t ldphi, int32_t m, double* out, int64_t ldo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MGP_B200_H_ */
