/*
 * btcdet_b200.h — C ABI of the B200-native (sm_100a) hot path of BtcDet.
 *
 * This is the drop-in boundary.  The reference has no C ABI of
 * its own: its sparse path is reached through the `spconv` Python package
 * (spconv v1.2.1, un-vendored) and through torch index ops.  Every entry point
 * below names the reference call site / interface it replaces (paths relative
 * to the reference checkout).  The host side (the `spconv/` shim package and
 * `btcdet_b200/`) binds these with ctypes — see INTEGRATION.md.
 *
 * Conventions
 *   - extern "C", POD arguments only: device pointers, sizes, host geometry
 *     arrays, a `void* stream` (cudaStream_t).  No torch types.
 *   - The caller owns every buffer, including workspaces (`*_workspace_bytes`
 *     query functions).  The library never calls cudaMalloc / cudaFree /
 *     cudaDeviceSynchronize and is safe under CUDA-graph capture.
 *   - Counts produced on the device (number of voxels, number of output
 *     sites, pairs per offset) are written to device memory; the caller may
 *     copy them to the host when it needs exact tensor shapes.
 *   - Counts consumed by a kernel come as a capacity (`n_cap`, host int, sizes
 *     the grid) plus an optional device pointer `n_dev` holding the live
 *     count (NULL → the capacity is the count).
 *   - Coordinates are int32 rows (b, z, y, x) as in spconv.SparseConvTensor
 *     (.indices, btcdet/models/backbones_3d/spconv_backbone.py:155-160).
 *   - Kernel offsets are numbered row-major over (kz, ky, kx), kx fastest; a
 *     pair (in -> out) exists at offset k iff  out*stride - pad + k*dil == in
 *     per axis (transposed: out == in*stride - pad + k*dil).
 *   - "Neighbour tables" are int32 [N, K] arrays holding a row index or -1.
 *     `nbr_out[o][k]` = the input row feeding output row o through offset k
 *     (output-stationary, what the conv kernels consume); `nbr_in[i][k]` =
 *     the output row that input row i feeds through offset k.
 *   - Return value: 0 on success, negative BTC_E_* otherwise.  No exceptions,
 *     no exit().
 */
#ifndef BTCDET_B200_H
#define BTCDET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BTC_OK 0
#define BTC_E_BADARG (-1)
#define BTC_E_CAPACITY (-2)
#define BTC_E_CUDA (-3)
#define BTC_E_UNSUPPORTED (-4)

/* Library / build identification. Returns e.g. 100 for sm_100a builds. */
int btc_abi_version(void);
int btc_compiled_sm(void);
/* Last CUDA error string captured by a failing call (thread-local, static storage). */
const char* btc_last_error(void);

/* ------------------------------------------------------------------------- */
/* Point -> voxel grouping                                                    */
/* Replaces spconv.utils.VoxelGeneratorV2.generate (spconv 1.2.1             */
/* points_to_voxel_3d_np), call sites                                         */
/* btcdet/datasets/processor/data_processor.py:68-73,85 / :112-117,136 /      */
/* :165-170,177, plus the batch-index padding of                              */
/* btcdet/datasets/dataset.py:187-192 (collate_batch).                        */
/*                                                                            */
/* Semantics (bit-exact with the sequential reference): per scene, points in  */
/* input order; c = floor((p - range_min) / voxel_size) per axis in fp32;     */
/* dropped if outside [0, grid); voxel ids in first-come order; at most       */
/* `max_voxels` voxels per scene (later-first-seen voxels dropped); the first */
/* `max_points` points of a voxel kept in input order; padding slots zero.    */
/* ------------------------------------------------------------------------- */

/* Bytes of workspace btc_voxelize needs for up to n_points points. */
int64_t btc_voxelize_workspace_bytes(int64_t n_points, int n_scenes, int max_voxels, int max_points);

/*
 * points        [n_points, n_feat] f32 device; columns 0..2 are the coordinates
 *               that get quantised (x, y, z order of voxel_size / range).
 * scene_offsets [n_scenes + 1] i32 device: points of scene b are rows
 *               [scene_offsets[b], scene_offsets[b+1]).
 * voxel_size[3], range[6]  host, same order as the reference ctor kwargs.
 * grid[3]       host (x, y, z) = round((max - min) / voxel_size).
 * Outputs (capacity n_scenes * max_voxels rows):
 *   voxels     [cap, max_points, n_feat] f32   (zero padded)
 *   coords     [cap, 4] i32 (b, z, y, x)
 *   num_points [cap] i32
 *   voxel_mean [cap, n_feat] f32 or NULL: sum over valid slots / max(count,1)
 *              (MeanVFE, btcdet/models/backbones_3d/vfe/mean_vfe.py:27-44)
 *   n_voxels   [n_scenes + 1] i32 device: [0..n_scenes) per-scene counts,
 *              [n_scenes] the total (rows are scene-major, compact).
 */
int btc_voxelize(const float* points, int n_points, int n_feat,
                 const int* scene_offsets, int n_scenes,
                 const float* voxel_size, const float* range, const int* grid,
                 int max_points, int max_voxels,
                 float* voxels, int* coords, int* num_points, float* voxel_mean,
                 int* n_voxels,
                 void* workspace, int64_t workspace_bytes, void* stream);
/* The same call in two stream-ordered halves over the same arguments and workspace: _group = grouping (coords, n_voxels;
 * everything that only needs coordinates — the coordinate hash and the rulebooks — may start behind it), _fill = contents
 * (voxels, num_points, voxel_mean).  btc_voxelize == _group followed by _fill. */
int btc_voxelize_group(const float* points, int n_points, int n_feat,
                       const int* scene_offsets, int n_scenes,
                       const float* voxel_size, const float* range, const int* grid,
                       int max_points, int max_voxels,
                       float* voxels, int* coords, int* num_points, float* voxel_mean,
                       int* n_voxels,
                       void* workspace, int64_t workspace_bytes, void* stream);
/* After the grouping, the workspace holds a complete coordinate -> voxel-row hash (same hash function / probing as
 * btc_hash_build; keys over the voxeliser's own grid (grid_z, grid_y, grid_x), -1 = empty; vals = row or -1 for voxels
 * beyond max_voxels): byte offsets of the two arrays inside the workspace and the slot count, for
 * btc_rulebook_subm_hash(shape = (grid_z, grid_y, grid_x)). */
int btc_voxelize_hash_view(int64_t n_points, int n_scenes, int max_voxels, int max_points,
                           int64_t* keys_offset, int64_t* vals_offset, int64_t* n_slots);
int btc_voxelize_fill(const float* points, int n_points, int n_feat,
                      const int* scene_offsets, int n_scenes,
                      const float* voxel_size, const float* range, const int* grid,
                      int max_points, int max_voxels,
                      float* voxels, int* coords, int* num_points, float* voxel_mean,
                      int* n_voxels,
                      void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Dataset-side coordinate transform of the occupancy branch: absxyz_2_cylinxyz_np / absxyz_2_spherexyz_np
 * (btcdet/utils/coords_utils.py:268-292, called at btcdet/datasets/processor/data_processor.py:128-131) on device points,
 * numpy's float32 op order (norm = sqrt(x*x + y*y), angle = (atan2 * 180) / pi: two roundings).
 *   points [n_cap, n_feat] f32 -> out [n_cap, n_feat] f32: (rho, phi_deg, z, extra...) or, sphere = 1,
 *   (r, azimuth_deg, elevation_deg, extra...).  n_dev [1] i32 device live count or NULL.
 * The output feeds btc_voxelize with the cylindrical voxel size / range (the occ voxeliser).  (The augmentation-only
 * SAVE_PRE_ROT branch, data_processor.py:149-150, subtracts rot_z from the voxel CONTENTS after grouping and is not part
 * of this call.)
 */
int btc_points_to_cylinder(const float* points, int n_cap, const int* n_dev, int n_feat, int sphere, float* out,
                           void* stream);

/* ------------------------------------------------------------------------- */
/* Coordinate index ("rank bitmap")                                            */
/* Replaces the dense int32 `grid` / cuckoo hash of spconv 1.2.1              */
/* (SparseConvTensor.grid, ops.get_indice_pairs use_hash) as the coordinate → */
/* row lookup.  One 8-byte entry per 32 cells: low word = occupancy bits of   */
/* cells [32w, 32w+32) in flat (b,z,y,x) row-major order, high word = number  */
/* of occupied cells before the word.  rank(cell) is therefore the position   */
/* of the cell in ascending flat-key order — the order spconv's GPU path      */
/* gives to conv outputs (torch::_unique of flat indices).                    */
/* ------------------------------------------------------------------------- */

/* Number of 8-byte entries of an index over batch × shape[0..2] (z, y, x) cells. */
int64_t btc_index_entries(int batch, const int* shape);
/* Workspace bytes for btc_index_build / rulebook builds over that many entries. */
int64_t btc_index_workspace_bytes(int64_t n_entries);

/*
 * Build the index of `coords` (rows must be unique).  `index` must be zeroed
 * by the caller (cudaMemsetAsync) or cleared with btc_index_clear.
 *   perm  [n_cap] i32 or NULL: perm[rank(coords[i])] = i, needed when the rows
 *         are not already in ascending flat-key order.
 *   total [1] i32 device or NULL: number of occupied cells.
 */
int btc_index_build(const int* coords, int n_cap, const int* n_dev,
                    int batch, const int* shape,
                    uint64_t* index, int64_t n_entries, int* perm, int* total,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* Zero the index words touched by `coords` (cheaper than a full memset). */
int btc_index_clear(const int* coords, int n_cap, const int* n_dev,
                    int batch, const int* shape, uint64_t* index, int64_t n_entries,
                    void* stream);

/*
 * Coordinate hash: the lookup structure for UNSORTED site sets on huge grids (the voxeliser's
 * first-come rows on the 92 M-cell KITTI det grid), where a dense bitmap would cost more
 * traffic than the whole rulebook.  Open addressing over flat (b,z,y,x) keys; `keys` int64
 * [n_slots] (zeroed to -1 inside), `vals` int32 [n_slots]; n_slots = btc_hash_slots(n_cap).
 */
int64_t btc_hash_slots(int n_cap);
int btc_hash_build(const int* coords, int n_cap, const int* n_dev, int batch, const int* shape,
                   int64_t* keys, int* vals, int64_t n_slots, void* stream);

/* ------------------------------------------------------------------------- */
/* Rulebooks ("indice pairs")                                                  */
/* Replace spconv.ops.get_indice_pairs (spconv 1.2.1 src/spconv/indice.cu:    */
/* prepareSubMGridKernel/getSubMIndicePairsKernel, prepareIndicePairsKernel/  */
/* assignGridAndIndiceOutKernel/assignIndicePairsKernel) reached from         */
/* SparseConvolution.forward — reference call sites                           */
/* btcdet/models/backbones_3d/spconv_backbone.py:12-29,106-128,657-707.       */
/* ------------------------------------------------------------------------- */

/*
 * Submanifold rulebook: output sites == input sites, same row order.
 * `index`/`perm` must have been built over `coords` with btc_index_build.
 *   nbr_out [n_cap, K] i32: nbr_out[o][k] = input row at coords[o] + (k - K/2)*dil, or -1.
 * (nbr_in is the mirror image: nbr_in[i][k] == nbr_out[i][K-1-k].)
 */
int btc_rulebook_subm(const int* coords, int n_cap, const int* n_dev,
                      int batch, const int* shape, const int* ksize, const int* dilation,
                      const uint64_t* index, int64_t n_entries, const int* perm,
                      int* nbr_out, void* stream);

/* Same table, probing a coordinate hash built with btc_hash_build over `coords`. */
int btc_rulebook_subm_hash(const int* coords, int n_cap, const int* n_dev,
                           int batch, const int* shape, const int* ksize, const int* dilation,
                           const int64_t* keys, const int* vals, int64_t n_slots,
                           int* nbr_out, void* stream);

/*
 * Regular (strided) or transposed sparse convolution / pooling rulebook.
 *   out_index [out_entries] zeroed by the caller; on return it is the rank
 *             bitmap of the output sites (rows in ascending flat-key order,
 *             so no perm is needed for it).
 *   out_coords [out_cap, 4], n_out [1] device: the output sites, ascending.
 *   nbr_out [out_cap, K], nbr_in [n_in_cap, K] (either may be NULL).
 * Returns BTC_E_CAPACITY semantics on the device: if the number of output
 * sites exceeds out_cap, n_out still holds the true count and rows >= out_cap
 * are not written (the caller checks n_out when it reads it back).
 */
int btc_rulebook_conv(const int* coords_in, int n_in_cap, const int* n_in_dev,
                      int batch, const int* in_shape, const int* out_shape,
                      const int* ksize, const int* stride, const int* padding,
                      const int* dilation, int transposed,
                      uint64_t* out_index, int64_t out_entries,
                      int* out_coords, int out_cap, int* n_out,
                      int* nbr_out, int* nbr_in,
                      void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Sparse two-level build of a strided / transposed rulebook (same results as btc_rulebook_conv, bit for bit).
 * `summary` [btc_index_summary_words(out_entries)] u32 holds one bit per 32-cell word of `out_index`; BOTH bitmaps must
 * be all-zero on entry and are left populated (out_index is the rank bitmap of the output level, used by sub-manifold
 * rulebooks on that level).  After the last reader, btc_index_clear_sparse zeroes exactly the touched words from the
 * level's coordinate list, so no per-step memset of the grid is needed.  The build touches only occupied words:
 * mark -> one single-pass ranking scan (decoupled look-back over the summary) -> table fill -> tables + coordinates.
 */
int64_t btc_index_summary_words(int64_t n_entries);
int64_t btc_rulebook_conv_sparse_workspace_bytes(int64_t n_entries);
int btc_rulebook_conv_sparse(const int* coords_in, int n_in_cap, const int* n_in_dev, int batch, const int* in_shape,
                             const int* out_shape, const int* ksize, const int* stride, const int* padding,
                             const int* dilation, int transposed, uint64_t* out_index, int64_t out_entries,
                             uint32_t* summary, int* out_coords, int out_cap, int* n_out, int* nbr_out, int* nbr_in,
                             void* workspace, int64_t workspace_bytes, void* stream);
int btc_index_clear_sparse(const int* coords, int n_cap, const int* n_dev, int batch, const int* shape, uint64_t* index,
                           int64_t n_entries, uint32_t* summary, void* stream);

/* Workspace bytes for btc_rulebook_pairs. */
int64_t btc_rulebook_pairs_workspace_bytes(int n_in_cap, int K);

/*
 * spconv-1.2.1-format rulebook in canonical order:
 *   pairs    [2, K, n_in_cap] i32, -1 padded: pairs[0][k][s] = input row,
 *            pairs[1][k][s] = output row, sorted by input row within offset k.
 *   pair_num [K] i32.
 * Built from an input-major table nbr_in [n_in_cap, K]; `mirror` != 0 reads
 * nbr_in[i][k] as table[i][K-1-k] (submanifold tables, see btc_rulebook_subm).
 */
int btc_rulebook_pairs(const int* table, int n_in_cap, const int* n_in_dev, int K, int mirror,
                       int* pairs, int* pair_num,
                       void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- */
/* Sparse convolution arithmetic                                               */
/* Replaces spconv.ops.indice_conv / indice_conv_backward (spconv 1.2.1       */
/* src/spconv/spconv_ops.cc: per-offset gather -> torch::mm -> scatter-add,   */
/* reordering.cu gather/scatter kernels) behind SubMConv3d / SparseConv3d /   */
/* SparseConvTranspose3d / SparseInverseConv3d.forward.                       */
/* The replacement is one output-stationary gather-GEMM kernel: each output   */
/* row accumulates sum_k feat_in[nbr_out[o][k]] @ W[k] in ascending k, is     */
/* written once, no atomics (deterministic).                                  */
/* ------------------------------------------------------------------------- */

/*
 * feat_in  [n_in, c_in] f32, nbr_out [n_out_cap, K] i32,
 * weight   [K, c_in, c_out] f32 (spconv 1.2.1 layout [kD,kH,kW,Cin,Cout]),
 * bias     [c_out] or NULL,
 * scale/shift [c_out] or NULL: fused per-channel affine applied after bias
 *          (eval-mode BatchNorm1d folded: y = conv*scale + shift),
 * relu     != 0: fused ReLU after the affine,
 * feat_out [n_out_cap, c_out] f32.
 * algo: 0 / 1 = fp32 FFMA register tiles (exact fp32 products and sums).  The tcgen05
 *       tensor-core tile takes pre-packed weights: see btc_sparse_conv_fwd_tc below.
 */
int btc_sparse_conv_fwd(const float* feat_in, const int* nbr_out,
                        const float* weight, const float* bias,
                        const float* scale, const float* shift, int relu,
                        float* feat_out, int n_out_cap, const int* n_out_dev,
                        int K, int c_in, int c_out, int algo, void* stream);

/*
 * Tensor-core variant (tcgen05.mma kind::tf32 with the 3xTF32 split, accumulator in TMEM):
 * same contract as btc_sparse_conv_fwd, for c_in % 4 == 0, c_out % 4 == 0, c_out <= 128, K <= 33 (49 when
 * c_out <= 32: the two [128 x K] index tiles must fit in shared memory next to the operand rings;
 * btc_sparse_conv_tc_supported says),
 * K * c_in >= 32.  The reduction runs over the flattened (offset, input channel) axis in chunks
 * of 32, so thin layers (c_in = 4, 16) pack several offsets into one MMA stage.  The weights are
 * packed once per layer into the shared-memory image of the K-major, 128-byte-swizzled hi/lo
 * operand tiles (btc_sparse_conv_tc_pack).
 */
int btc_sparse_conv_tc_supported(int K, int c_in, int c_out);
/* Tile-variant knobs for A/B measurements and tests (process-wide; -1 keeps a setting): producer_warps 8 | 16
 * (16 needs c_out <= 64), dynamic_tiles 0 | 1 = persistent CTAs fetch 128-row tiles from a global counter instead of a
 * fixed round-robin (no result bit changes).  concat_b must be 0 or -1: the concatenated [B_hi|B_lo] variant of round 1
 * was measured on hardware (no gain, profiles/r2_battery.json) and removed; 1 returns BTC_E_UNSUPPORTED. */
int btc_sparse_conv_tc_config(int producer_warps, int concat_b, int dynamic_tiles);
/* Timing diagnostics for the profile write-ups (tools/step_breakdown.py --diag): a non-zero mask makes the tile skip a
 * part of its work — bit 0 the gather, bit 1 the smem read-back / hi-lo split / TMEM stores, bit 2 two of the three
 * MMAs per k-step, bit 3 the weight-tile copies — so the cost of each pipeline side can be read off a wall-clock difference.  RESULTS ARE WRONG
 * while the mask is non-zero; 0 (the default) restores the product path. */
int btc_sparse_conv_tc_diag(int mask);
/* Timeline diagnostics (tools/tc_timeline.py): with a non-null device buffer of 148 x 16 uint64 every CTA of the following
 * tcgen05 launches records clock64 / globaltimer stamps of its roles (entry, set-up done, first gather, first operand
 * stage, first MMA, last commit, last epilogue, exit) and the cycles its MMA issuer / producers / epilogue spent waiting.
 * Results are unchanged; null (the default) switches it off.  Process-wide, not thread-safe. */
int btc_sparse_conv_tc_trace(void* trace_u64);
/* Cap on the persistent grid of the tcgen05 tile (default 148 = one CTA per SM): a smaller grid leaves whole SMs to
 * kernels running concurrently on other streams (the rulebook chain of the engine).  Process-wide. */
int btc_sparse_conv_tc_grid(int max_ctas);
int64_t btc_sparse_conv_tc_packed_bytes(int K, int c_in, int c_out);
int btc_sparse_conv_tc_pack(const float* weight, int K, int c_in, int c_out, void* packed, void* stream);
int btc_sparse_conv_fwd_tc(const float* feat_in, const int* nbr_out, const void* packed_weight,
                           const float* bias, const float* scale, const float* shift, int relu,
                           float* feat_out, int n_out_cap, const int* n_out_dev,
                           int K, int c_in, int c_out, void* stream);
/*
 * Block skipping: per 128-row tile the kernel only gathers / multiplies the reduction chunks whose kernel offsets
 * have a valid neighbour in at least one row (the others contribute exact zeros).  btc_rulebook_sort_rows makes
 * that effective: it reorders the rows of an output-stationary table inside windows of 2048 rows by their
 * valid-offset bit mask (stable), so rows with the same neighbourhood pattern share tiles.
 *   nbr_out [n_cap, K] -> nbr_sorted [n_cap, K] (row i of the sorted table = row out_rows[i] of the input),
 *   out_rows [n_cap] i32.  K <= 64.  No workspace, one launch.
 * btc_sparse_conv_fwd_tc_rows consumes the sorted table and writes row i's result to feat_out[out_rows[i]]:
 * same results as btc_sparse_conv_fwd_tc on the unsorted table, bit for bit (per-row arithmetic is unchanged).
 */
int btc_rulebook_sort_rows(const int* nbr_out, int n_cap, const int* n_dev, int K, int* nbr_sorted, int* out_rows,
                           void* stream);
/*
 * Split feature format (round 2): the same 4*C bytes per row, every value stored as two bfloat16, x = hi + lo with
 * hi = bf16_rn(x), lo = bf16_rn(x - hi) (|x - hi - lo| <= 2^-17 |x|); per 32 channels the row holds [32 x hi | 32 x lo].
 * A layer whose INPUT is in split format (c_in % 32 == 0) gathers the packed operands as they are (no per-stage ALU work)
 * and runs three bf16 MMAs of K = 16 per k-step instead of three tf32 MMAs of K = 8; a layer whose OUTPUT is split
 * (c_out % 32 == 0) writes the format from its epilogue.  Weights for a split-input layer are packed by
 * btc_sparse_conv_tc_pack_split (bf16 hi / lo tiles per 32-element reduction chunk, SWIZZLE_64B rows).  btc_features_to_split / _from_split
 * convert whole feature matrices (tests, or a consumer that needs fp32).  Accuracy: ~2^-16 per product, inside the 1e-4
 * parity bar (tests/test_parity_gpu.py).  in_split / out_split = 0 reproduces btc_sparse_conv_fwd_tc_meta.
 */
int btc_sparse_conv_tc_split_supported(int K, int c_in, int c_out, int in_split, int out_split);
int64_t btc_sparse_conv_tc_split_packed_bytes(int K, int c_in, int c_out);
int btc_sparse_conv_tc_pack_split(const float* weight, int K, int c_in, int c_out, void* packed, void* stream);
int btc_features_to_split(const float* feat, int n_cap, const int* n_dev, int c, void* out, void* stream);
int btc_features_from_split(const void* feat_split, int n_cap, const int* n_dev, int c, float* out, void* stream);
int btc_sparse_conv_fwd_tc_split(const void* feat_in, const int* nbr_out, const void* packed_weight, const float* bias,
                                 const float* scale, const float* shift, int relu, void* feat_out, int n_out_cap,
                                 const int* n_out_dev, int K, int c_in, int c_out, int in_split, int out_split,
                                 const uint64_t* tile_mask, const int* tile_order, void* stream);
/*
 * Per-tile metadata of a neighbour table for the tcgen05 tile (tiles = 128 consecutive rows):
 *   tile_mask  [ceil(n_out_cap / 128)] u64: bit k set iff a live row of the tile has a neighbour through offset k
 *              (the block-skipping mask the tile otherwise derives from its staged index block, ~2.5 us per tile);
 *   tile_order [btc_rulebook_tile_order_ints(n_out_cap)] i32 or NULL: the live tiles bucketed by cost class (number of
 *              active offsets): 65 class counts followed by 65 buckets of tile ids.  The dynamic tile scheduler hands
 *              tiles out heaviest class first, so a launch ends on its cheapest tiles (no long tail).
 * Built once per rulebook (one launch, no CTA waits for another), shared by every layer that uses the table.  K <= 64.
 * btc_sparse_conv_fwd_tc_meta = btc_sparse_conv_fwd_tc with the two arrays (either may be NULL); results are identical.
 */
int64_t btc_rulebook_tile_order_ints(int n_out_cap);
int btc_rulebook_tile_meta(const int* nbr_out, int n_out_cap, const int* n_out_dev, int K, uint64_t* tile_mask,
                           int* tile_order, void* stream);
int btc_sparse_conv_fwd_tc_meta(const float* feat_in, const int* nbr_out, const void* packed_weight, const float* bias,
                                const float* scale, const float* shift, int relu, float* feat_out, int n_out_cap,
                                const int* n_out_dev, int K, int c_in, int c_out, const uint64_t* tile_mask,
                                const int* tile_order, void* stream);
int btc_sparse_conv_fwd_tc_rows(const float* feat_in, const int* nbr_sorted, const int* out_rows,
                                const void* packed_weight, const float* bias, const float* scale,
                                const float* shift, int relu, float* feat_out, int n_out_cap,
                                const int* n_out_dev, int K, int c_in, int c_out, void* stream);

/*
 * d feat_in [n_in_cap, c_in] = sum_k d_out[nbr_in[i][k]] @ W[k]^T.
 * `mirror` != 0 reads nbr_in[i][k] as table[i][K-1-k] (submanifold).
 * workspace: K * c_in * c_out floats (transposed weights).
 */
int64_t btc_sparse_conv_bwd_workspace_bytes(int K, int c_in, int c_out);
int btc_sparse_conv_bwd_data(const float* d_out, const int* table, int mirror,
                             const float* weight,
                             float* d_in, int n_in_cap, const int* n_in_dev,
                             int K, int c_in, int c_out,
                             void* workspace, int64_t workspace_bytes, void* stream);

/*
 * d weight [K, c_in, c_out] = sum_o feat_in[nbr_out[o][k]]^T (x) d_out[o];
 * d_bias [c_out] (or NULL) = sum_o d_out[o].   d_weight is overwritten.
 */
int btc_sparse_conv_bwd_weight(const float* feat_in, const float* d_out, const int* nbr_out,
                               float* d_weight, float* d_bias,
                               int n_out_cap, const int* n_out_dev,
                               int K, int c_in, int c_out, void* stream);

/* ------------------------------------------------------------------------- */
/* SparseMaxPool3d (spconv 1.2.1 src/spconv/maxpool.cu) —                     */
/* btcdet/models/backbones_3d/spconv_backbone.py:29,831-847.                  */
/* out is zero-initialised, out[o] = max(out[o], in[i]) over the pairs        */
/* (negative inputs clamp to 0).                             */
/* ------------------------------------------------------------------------- */
int btc_maxpool_fwd(const float* feat_in, const int* nbr_out, float* feat_out,
                    int n_out_cap, const int* n_out_dev, int K, int c, void* stream);
/* d_in[i] += d_out[o] where feat_in[i] == feat_out[o] over the pairs (d_in zeroed inside). */
int btc_maxpool_bwd(const float* feat_in, const float* feat_out, const float* d_out,
                    const int* nbr_out, float* d_in, int n_in,
                    int n_out_cap, const int* n_out_dev, int K, int c, void* stream);

/* ------------------------------------------------------------------------- */
/* SparseConvTensor.dense()  (spconv 1.2.1 SparseConvTensor.dense; call sites */
/* occ_head_3D.py:46,51, height_compression.py:21, spconv_backbone.py:890,930)*/
/* out [batch, c, D, H, W] f32 is zero-filled then rows are scattered.        */
/* ------------------------------------------------------------------------- */
int btc_to_dense(const float* feat, const int* coords, int n_cap, const int* n_dev,
                 int c, int batch, const int* shape, float* out, void* stream);
/* Compact copy of the live rows (count read on the device) of a capacity-sized row buffer;
 * row_bytes % 16 == 0.  Used to stage results for an asynchronous device->host copy. */
int btc_copy_rows(const void* src, void* dst, int n_cap, const int* n_dev, int row_bytes, void* stream);
/* Backward of dense(): d_feat[i][ch] = d_out[b, ch, z, y, x]. */
int btc_from_dense(const float* d_out, const int* coords, int n_cap, const int* n_dev,
                   int c, int batch, const int* shape, float* d_feat, void* stream);

/* ------------------------------------------------------------------------- */
/* Sorted re-voxelisation of labelled points                                   */
/* Replaces AddOccTemplate.combine_gt_occ_voxel_point + voxelize_pad          */
/* (btcdet/models/occ_pnt/add_occ_template.py:248-268): torch.unique(coords,  */
/* dim=0, sorted) + sort(inverse) + dense pad, without the host sync.         */
/* ------------------------------------------------------------------------- */
int64_t btc_revoxelize_workspace_bytes(int n_points, int64_t n_entries);
/*
 * pt_coords [n, 4] i32 (b,z,y,x), pt_feat [n, c] f32.
 * index [n_entries] zeroed by the caller (rank bitmap of the det grid).
 * Outputs: vox_coords [cap,4] ascending (b,z,y,x); vox_count [cap];
 *          slots [n] i32: position of point i inside its voxel (stable, by
 *          input order); pt_voxel [n] i32: voxel row of point i;
 *          n_voxels [1], max_count [1] device.
 * The caller pads with btc_revoxelize_fill once it knows / bounds max_count.
 */
int btc_revoxelize(const int* pt_coords, int n_cap, const int* n_dev,
                   int batch, const int* shape,
                   uint64_t* index, int64_t n_entries,
                   int* vox_coords, int vox_cap, int* vox_count,
                   int* slots, int* pt_voxel, int* n_voxels, int* max_count,
                   void* workspace, int64_t workspace_bytes, void* stream);
/* voxels [vox_rows, p_max, c] zero-filled, then voxels[pt_voxel[i]][slots[i]] = pt_feat[i] (slots >= p_max dropped). */
int btc_revoxelize_fill(const float* pt_feat, const int* pt_voxel, const int* slots,
                        int n_cap, const int* n_dev, int c, int p_max,
                        float* voxels, int vox_rows, void* stream);

/* ------------------------------------------------------------------------- */
/* Occupancy / occlusion masks (SURVEY §8 rows a5-a8 + the mask algebra of a12) */
/* Replaces, in one sync-free call, the torch index-op chain of               */
/* btcdet/models/occ_pnt/occ_training_targets/occ_targets_template.py:        */
/* get_valid/get_voxelwise_mask :194-202, create_predict_area3d :432-447,      */
/* occ_from_cylin_ocp :136-155 (+ :82-90, :110-134, :186-191), filter_occ      */
/* :249-255 and general_cls_loss_mask :333.  Bit-exact with that code run on   */
/* the same device (fp32 op order and CUDA libm calls reproduced, no FMA).     */
/*                                                                            */
/* voxels [m, P, C>=3] f32 cylindrical (rho, phi deg, z, ...); voxel_coords    */
/* [m, 4] i32 (b, z, y, x); num_points [m] i32; rot_z [batch] f32 or NULL.     */
/* geom_f[21]: voxel_size[3], range_min[3], range_max[3] of the occ grid       */
/*   (rho, phi, z order), sphere voxel_size[3], sphere_min[3], sphere_max[3]   */
/*   (r, az, el order), EMPT_SUR_THRESH, det z min, det z max.                 */
/* geom_i[11]: grid nx,ny,nz; sphere grid nx,ny,nz; DIST_KERN z,y,x;           */
/*   concede_x; use_empty_surround (EMPT_SUR_THRESH < 9).                      */
/* Outputs u8 [batch, nz, ny, nx]: voxelwise_mask, vcc_mask,                   */
/*   occ_voxelwise_mask (after filter_occ), general_cls_loss_mask (or NULL);   */
/*   sphere_map_out u8 [batch, snz, sny, snx] or NULL (before the empty-       */
/*   surround rewrite of range bin 0).                                         */
/* ------------------------------------------------------------------------- */
int64_t btc_occ_targets_workspace_bytes(int batch, const float* geom_f, const int* geom_i);
int btc_occ_targets(const float* voxels, int P, int C, const int* voxel_coords, const int* num_points,
                    int m_cap, const int* m_dev, int batch, const float* rot_z,
                    const float* geom_f, const int* geom_i,
                    uint8_t* voxelwise_mask, uint8_t* vcc_mask, uint8_t* occ_mask, uint8_t* general_mask,
                    uint8_t* sphere_map_out, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- */
/* Occupancy-point injection                         */
/* btc_occ_select replaces AddOccTemplate.filter_occ_points (without top-k),  */
/* occ_coords2absxyz, trans_voxel_grid and assemble_occ_points                */
/* (btcdet/models/occ_pnt/add_occ_template.py:78-165): cells with             */
/* prob > thresh, compacted in row-major (torch.nonzero) order per scene,     */
/* turned into pseudo points [x,y,z,inten,prob,1] and det-grid coordinates.   */
/* probs [batch,nz,ny,nx] f32, residuals [batch,3,nz,ny,nx] f32 or NULL,      */
/* grid (nx,ny,nz), geom_f[12] = occ voxel size[3], occ origin[3] (rho,phi,z), */
/* det voxel size[3], det range min[3] (x,y,z); det_grid (nx,ny,nz).          */
/* Outputs (capacity `cap` rows): occ_coords [cap,4] (b,z,y,x), occ_probs,    */
/* occ_xyz [cap,3], det_coords [cap,4], occ_points [cap,6];                   */
/* counts [batch+1] i32 device: per scene, then the total.                    */
/* btc_occ_vfe replaces OccVFE.forward (backbones_3d/vfe/occ_vfe.py:24-55).   */
/* ------------------------------------------------------------------------- */
int64_t btc_occ_select_workspace_bytes(int batch, const int* grid);
int btc_occ_select(const float* probs, const float* residuals, int batch, const int* grid, float thresh,
                   const float* rot_z, const float* geom_f, const int* det_grid, float inten, int cap,
                   int* occ_coords, float* occ_probs, float* occ_xyz, int* det_coords, float* occ_points,
                   int* counts, void* workspace, int64_t workspace_bytes, void* stream);
int btc_occ_vfe(const float* voxels, const int* num_points, int m_cap, const int* m_dev, int P, int C,
                int num_raw, float* feats, float* occ_feats, void* stream);
/* SURVEY §8(f) N4, occupancy-head half: the tail of OccHead3D.forward (occ_pnt/occ_dense_heads/occ_head_3D.py:46-49),
 * `softmax(conv_cls(x).dense(), dim=1)[:, -1] * general_cls_loss_mask`, straight from the head's sparse rows — the
 * [B, n_cls, nz, ny, nx] logits volume, the softmax pass over it and the mask product are never materialised.
 * logits [n_cap, n_cls], coords [n_cap, 4] (b,z,y,x), grid = (nx, ny, nz), mask u8 [B,nz,ny,nx] or NULL -> prob f32 [B,nz,ny,nx]
 * (cells without an active site: softmax of all-zero logits = 1/n_cls, times the mask, as dense() + softmax gives). */
int btc_occ_head_prob(const float* logits, const int* coords, int n_cap, const int* n_dev, int n_cls, int batch,
                      const int* grid, const unsigned char* mask, float* prob, void* stream);
/* MeanVFE of the occupancy branch (occ side): occ_targets_3d.py:45-47 (USE_ABSXYZ) rewrites every slot of
 * the cylindrical occ voxels to cylinder_uvd2absxyz(rho, phi, z) + extra columns, mean_vfe.py:27-44 then averages all
 * slots / clamp_min(count, 1).  voxels [m_cap, max_points, n_feat] (rho, phi_deg, z, ...) -> voxels_abs (same shape, may
 * be NULL) and voxel_mean [m_cap, n_feat] — the input features of VoxelBackBoneDeconv. */
int btc_occ_abs_mean_vfe(const float* voxels, int max_points, int n_feat, const int* num_points, int m_cap, const int* m_dev,
                         float* voxels_abs, float* voxel_mean, void* stream);

/* ------------------------------------------------------------------------- */
/* Box-driven occupancy targets                        */
/* btc_occ_box_targets replaces, per batch and without host synchronisation,   */
/*   get_fore_mirr_voxelwise_mask_res  (occ_targets_3d.py:146-171) over         */
/*     torch_points_and_sym_in_box_3d_batch (point_box_utils.py:70-97,252-306), */
/*   get_bm_voxelwise_mask_res (:95-119), get_mean_res (:122-130),              */
/*   get_voxel_center_xyz (:133-145) and the forebox-label loop (:70-86 over    */
/*   point_box_utils.py:198-250,332-365).                                       */
/* voxels/voxel_coords/num_points/rot_z/geom_*: as btc_occ_targets.             */
/* gt_boxes [batch, max_boxes, box_dim>=8] f32 (x,y,z,dx,dy,dz,heading,...,     */
/*   label last); gt_boxes_num [batch] i32; mirr_flag [batch, max_boxes] f32    */
/*   or NULL; bm_points [n_bm, 4] f32 (b,x,y,z) or NULL (16-byte aligned).      */
/* num_class == 1 binarises the label column (> 1e-2) as occ_targets_3d.py:52.  */
/* mirr_cap / bm_cap: upper bounds on the number of distinct cells hit by       */
/*   mirrored / template points (accumulator capacity); *status (device i32)    */
/*   is set to 1 if one overflowed.                                             */
/* Outputs: fore_mask, mirr_mask, bm_mask u8 [batch,nz,ny,nx] (raw, before the  */
/*   (1 - voxelwise) products); *_res f32 [batch,3,nz,ny,nx] (per-cell mean of  */
/*   the points minus the cell centre); forebox_label i8 [batch,nz,ny,nx] or    */
/*   NULL (BOX_WEIGHT == 1); point_label i8 [m_cap, P] or NULL.                 */
/* Box-frame coordinates use the analytic rigid inverse, not torch.inverse:     */
/*   see DESIGN.md for the face/bin-edge tolerance this implies.                */
/* btc_occ_loss_maps replaces prepare_cls_loss_map / prepare_reg_loss_map       */
/*   (occ_targets_template.py:330-401, dropout off) and the (1 - mask) products */
/*   of occ_targets_3d.py:58-65.  weights[8] = fore_cls, mirr_cls, bm_cls,      */
/*   neg_cls, fore_res, mirr_res, bm_res, (BOX_WEIGHT - neg_cls).               */
/*   grid = (nx, ny, nz).  bm_* and forebox_label, bm_voxelwise_mask and        */
/*   pos_all_num (device i32) may be NULL.                                      */
/* ------------------------------------------------------------------------- */
int64_t btc_occ_box_targets_workspace_bytes(int batch, int max_boxes, int mirr_cap, int bm_cap);
int btc_occ_box_targets(const float* voxels, int P, int C, const int* voxel_coords, const int* num_points,
                        int m_cap, const int* m_dev, int batch, const float* gt_boxes, int max_boxes, int box_dim,
                        const int* gt_boxes_num, const float* mirr_flag, const float* bm_points, int n_bm,
                        const float* rot_z, const float* geom_f, const int* geom_i, int num_class,
                        int mirr_cap, int bm_cap, uint8_t* fore_mask, float* fore_res, uint8_t* mirr_mask,
                        float* mirr_res, uint8_t* bm_mask, float* bm_res, int8_t* forebox_label,
                        int8_t* point_label, int* status, void* workspace, int64_t workspace_bytes, void* stream);
/* Same call with the reference's `voxel_centers["all_voxel_centers_2d"]` table (detector3d_template.py:52-63: the mean
 * over z of the voxel centres' xy, [ny * nx, 2] f32 device, computed by torch at model build) for the 2-D pre-filter of
 * the forebox label (occ_targets_3d.py:78-79).  With the table the forebox label is the reference's bit for bit; NULL uses
 * the un-averaged centre (differs from torch's 9-term mean by at most one ulp on a few columns). */
int btc_occ_box_targets_v2(const float* voxels, int P, int C, const int* voxel_coords, const int* num_points,
                           int m_cap, const int* m_dev, int batch, const float* gt_boxes, int max_boxes, int box_dim,
                           const int* gt_boxes_num, const float* mirr_flag, const float* bm_points, int n_bm,
                           const float* rot_z, const float* geom_f, const int* geom_i, int num_class,
                           int mirr_cap, int bm_cap, const float* centers2d, uint8_t* fore_mask, float* fore_res,
                           uint8_t* mirr_mask, float* mirr_res, uint8_t* bm_mask, float* bm_res, int8_t* forebox_label,
                           int8_t* point_label, int* status, void* workspace, int64_t workspace_bytes, void* stream);
int btc_occ_loss_maps(const uint8_t* voxelwise_mask, const uint8_t* general_mask, const uint8_t* fore_mask,
                      const uint8_t* mirr_mask, const uint8_t* bm_mask, const int8_t* forebox_label,
                      const float* fore_res, const float* mirr_res, const float* bm_res, const float* weights,
                      int batch, const int* grid, uint8_t* occ_fore_cls_mask, uint8_t* occ_mirr_cls_mask,
                      uint8_t* occ_bm_cls_mask, uint8_t* pos_mask, uint8_t* bm_voxelwise_mask,
                      float* cls_loss_mask_float, uint8_t* reg_loss_mask, float* reg_loss_mask_float,
                      float* res_mtrx, int* pos_all_num, void* stream);

/* ------------------------------------------------------------------------- */
/* Aliases under the names of SURVEY.md §8(b)'s minimum export set (thin forwards):                       */
/*   btc_voxelize_cuda          = btc_voxelize                                                            */
/*   btc_rulebook_pool          = btc_rulebook_conv with transposed = 0 (SparseMaxPool3d builds a regular   */
/*                                conv rulebook, spconv 1.2.1 pool.py / SURVEY App. A.6)                   */
/*   btc_occ_inject_revoxelize  = btc_revoxelize (combine_gt_occ_voxel_point + voxelize_pad,               */
/*                                btcdet/models/occ_pnt/occ_adder/add_occ_template.py:248-268)               */
/* ------------------------------------------------------------------------- */
int btc_voxelize_cuda(const float* points, int n_points, int n_feat,
                      const int* scene_offsets, int n_scenes,
                      const float* voxel_size, const float* range, const int* grid,
                      int max_points, int max_voxels,
                      float* voxels, int* coords, int* num_points, float* voxel_mean,
                      int* n_voxels,
                      void* workspace, int64_t workspace_bytes, void* stream);
int btc_rulebook_pool(const int* coords_in, int n_in_cap, const int* n_in_dev,
                      int batch, const int* in_shape, const int* out_shape,
                      const int* ksize, const int* stride, const int* padding,
                      const int* dilation,
                      uint64_t* out_index, int64_t out_entries,
                      int* out_coords, int out_cap, int* n_out,
                      int* nbr_out, int* nbr_in,
                      void* workspace, int64_t workspace_bytes, void* stream);
int btc_occ_inject_revoxelize(const int* pt_coords, int n_cap, const int* n_dev,
                              int batch, const int* shape,
                              uint64_t* index, int64_t n_entries,
                              int* vox_coords, int vox_cap, int* vox_count,
                              int* slots, int* pt_voxel, int* n_voxels, int* max_count,
                              void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- */
/* Rotated BEV overlap / IoU and rotated NMS.                */
/* Replace the reference's btcdet/ops/iou3d_nms extension:                      */
/*   boxes_overlap_bev_gpu / boxes_iou_bev_gpu  (iou3d_nms.cpp:58-100,           */
/*     iou3d_nms_kernel.cu:236-266)            -> btc_boxes_bev (mode 1 / 0)     */
/*   nms_gpu / nms_normal_gpu (iou3d_nms.cpp:103-188: N x N/64 mask on the       */
/*     device, copied to the host and scanned there, iou3d_nms_kernel.cu:269-413) */
/*                                              -> btc_nms (mask + greedy scan    */
/*     on the device, keep list and count stay on the device, no sync).           */
/* boxes: rows of 7 floats (x, y, z, dx, dy, dz, heading).  btc_nms expects the   */
/* boxes sorted by descending score (the reference's Python wrapper sorts) and    */
/* writes the kept row indices, ascending, to keep[0 .. *num_out).                */
/* ------------------------------------------------------------------------- */
int btc_boxes_bev(const float* boxes_a, int n, const float* boxes_b, int m, int mode /* 0 IoU, 1 overlap area */,
                  float* out /* [n, m] */, void* stream);
int64_t btc_nms_workspace_bytes(int n);
int btc_nms(const float* boxes, int n, float thresh, int normal /* 1: axis-aligned IoU (nms_normal_gpu) */,
            long long* keep /* device [n] */, int* num_out /* device */, void* workspace, int64_t workspace_bytes,
            void* stream);

/* ------------------------------------------------------------------------- */
/* RoI grid pooling of the second stage: the native ops      */
/* under ConvHead.roi_conv_pool (btcdet/models/roi_heads/conv_head.py:247-379). */
/*                                                                            */
/* btc_ball_query_stack replaces ball_query_wrapper_stack                     */
/*   (btcdet/ops/pointnet2/pointnet2_stack/src/ball_query.cpp:21-39, kernel    */
/*   ball_query_gpu.cu:16-60) for ALL radii of one StackSAModuleMSG            */
/*   (pointnet2_modules.py:10-108 calls it once per radius) in one pass:       */
/*   idx[r] is [M, nsamples[r]] i32: the first nsamples[r] points of the       */
/*   query's scene, in index order, with d2 < radii[r]^2 (d2 in the reference  */
/*   kernel's operation order); shorter rows repeat the first hit; an empty    */
/*   ball is (-1, 0, 0, ...).  `idx` is a HOST array of n_radii (1..4) device   */
/*   pointers.  xyz [sum N_b, 3], new_xyz [M, 3], *_batch_cnt [B] i32 device.   */
/* btc_group_points_stack(_grad) replace group_points_wrapper_stack /          */
/*   group_points_grad_wrapper_stack (group_points.cpp, kernels                */
/*   group_points_gpu.cu:16-95): out [M, C, nsample] = features[start_b +      */
/*   idx[m, s], c]; the gradient is scatter-added (fp32 atomics, like the      */
/*   reference) into grad_features [N, C], which the caller zeroes.            */
/* btc_trilinear_sparse_flag / _emit replace                                   */
/*   reverse_sparse_trilinear_interpolate_torch (btcdet/utils/common_utils.py: */
/*   247-311: feat.dense() + 8 corner gathers + 8 weights) and the non-zero    */
/*   row compaction of ConvHead.interpolate_from_3d_features                   */
/*   (conv_head.py:505-528) without the dense volume: `flag` builds a row-     */
/*   index volume of the sparse tensor (feats [n, C], coords [n, 4] b z y x,   */
/*   grid shape (Z, Y, X)), evaluates every target (zyx [T, 3] f32 = the       */
/*   reference's spatial_target_idxs; scene = b_target[t] or t / per_scene)    */
/*   and writes the number of rows with a non-zero channel to *count (device); */
/*   `emit` writes those rows, in target order, to out_feats [out_cap, C],     */
/*   their coordinates (t / P, unravel(t % P, local_shape)) to out_coords      */
/*   [out_cap, 4] i32 and (optional) the target index to out_target; rows      */
/*   beyond out_cap are dropped (the caller compares *count with out_cap).     */
/*   Products and sums follow the reference's expression order: bit-exact.     */
/*   Both calls take the SAME workspace (emit reads what flag left in it).     */
/* ------------------------------------------------------------------------- */
int btc_ball_query_stack(int B, int M, int n_radii, const float* radii /* host */, const int* nsamples /* host */,
                         const float* new_xyz, const int* new_xyz_batch_cnt, const float* xyz, const int* xyz_batch_cnt,
                         int* const* idx /* host array of device pointers */, void* stream);
int btc_group_points_stack(int B, int M, int C, int nsample, const float* features, const int* features_batch_cnt,
                           const int* idx, const int* idx_batch_cnt, float* out, void* stream);
int btc_group_points_stack_grad(int B, int M, int C, int N, int nsample, const float* grad_out, const int* idx,
                                const int* idx_batch_cnt, const int* features_batch_cnt, float* grad_features,
                                void* stream);
int64_t btc_trilinear_sparse_workspace_bytes(int64_t n_targets, int batch, const int* shape /* host (Z, Y, X) */);
int btc_trilinear_sparse_flag(const float* feats, const int* coords, int n_cap, const int* n_dev, int C, int batch,
                              const int* shape, const float* zyx, const long long* b_target /* or NULL */,
                              int64_t n_targets, int64_t per_scene, int normalize, int* count /* device */,
                              void* workspace, int64_t workspace_bytes, void* stream);
int btc_trilinear_sparse_emit(const float* feats, int C, int batch, const int* shape, const float* zyx,
                              const long long* b_target, int64_t n_targets, int64_t per_scene, int normalize, int P,
                              const int* local_shape /* host (lz, ly, lx), lz*ly*lx == P */, int out_cap,
                              float* out_feats, int* out_coords, long long* out_target /* or NULL */, void* workspace,
                              int64_t workspace_bytes, void* stream);
/* Adjoint of the emitted rows w.r.t. the sparse source features (the backward the reference gets from autograd through
 * im[b, :, z, y, x]): grad_feats [n, C] (zeroed by the caller) += w_j * grad_out[o] at the active corners of target
 * out_target[o]; same workspace as flag / emit (reads the row-index volume). */
int btc_trilinear_sparse_grad(const float* grad_out, const long long* out_target, int n_out, const int* n_out_dev, int C,
                              int batch, const int* shape, const float* zyx, const long long* b_target, int64_t n_targets,
                              int64_t per_scene, int normalize, float* grad_feats, void* workspace, int64_t workspace_bytes,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BTCDET_B200_H */
