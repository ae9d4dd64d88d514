/*
 * b200seg.h -- C ABI of the B200-native 3D instance-segmentation operator layer.
 *
 * One shared library (libb200seg.so, hand-written CUDA for sm_100a) exposes the hot path of
 * MeowMeowLady/InstanceSeg-Without-Voxelwise-Labeling behind plain-C entry points: raw pointers,
 * sizes and a CUDA stream, no torch types.  Each entry point names the reference interface it
 * replaces (path:line relative to the reference checkout).
 *
 * Conventions
 *   - Every function returns 0 on success, a positive cudaError_t value when the CUDA runtime
 *     reported an error, or a negative B200SEG_E* code for invalid arguments.  Nothing calls
 *     exit() (the reference launcher does: roi_align_kernel_3d.cu:165-169).
 *     b200seg_last_error() returns a thread-local human-readable message.
 *   - "_dev" entry points take DEVICE pointers, enqueue work on `stream` and never synchronise;
 *     data-dependent counts are written to device memory.
 *   - "_host" entry points take HOST pointers, perform the H2D/D2H copies themselves on an
 *     internal stream with grow-only device scratch, and return when the result is in host memory.
 *     They are what the reference's numpy call sites bind to.
 *   - Boxes are (x1,y1,z1,x2,y2,z2) with inclusive "+1" extents, as in the reference.
 *   - Tie rule for equal sort keys (scores / volumes) in NMS and in the visit order: key descending, then original
 *     index DESCENDING.  The reference takes `argsort()[::-1]` (cython_nms_3d.pyx:49, binarization_soma.py:60); numpy's
 *     default sort is stable below 17 elements (ties ascending, so the reversal puts the higher index first) and
 *     build dependent above -- the rule reproduces the reference wherever that sort is stable.
 */
#ifndef B200SEG_H_
#define B200SEG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SEG_EINVAL   (-1)   /* bad argument (null pointer, negative size, unsupported value) */
#define B200SEG_EWORKSPACE (-2) /* workspace too small */
#define B200SEG_EUNSUPPORTED (-3)

typedef void* b200seg_stream_t;   /* cudaStream_t */

/* dtype codes for feature / gradient tensors */
#define B200SEG_F32  0
#define B200SEG_BF16 1

const char* b200seg_last_error(void);
int b200seg_version(void);
/* Number of kernels launched by this library in the calling process so far (all entry points). */
long long b200seg_launch_count(void);
/* Profiling knobs (no effect on results unless stated).  "peaks_stop_after" = k: b200seg_peaks3d_dev returns after its
 * first k kernel launches (0 = memset only, 99 = the whole op, the default), so that a caller can time the op kernel by
 * kernel with CUDA events; outputs are incomplete while k < 99.  "host_batch_mode" (default 7) selects the transfer
 * scheme of b200seg_postproc_soma_host_batch, same results either way: bit 0 = label volumes come back compacted,
 * bit 1 = only the PRM crops of the NMS survivors are fetched (zero-copy gather), bit 2 = the raw volume is not
 * copied either: the NMS runs first, host threads pack the image crops of its survivors (the only voxels the chain reads)
 * into a pinned buffer and that buffer travels instead; bit 3 (with bit 2) = the PRM crops of those survivors are packed into
 * the same buffer by the host threads instead of being gathered over the link; bits 4 / 5 = the chain is launched for groups
 * of 2 / 4 volumes instead of one volume at a time (both measured slower on a 16-core host, see host_batch.cu; they pay
 * where host cores are plentiful); bit 6 / bit 8 = pinned label buffers always receive their lines through the staged
 * download and the host scatter / always in place from the GPU (default: in place when a rank has fewer than 8 host cores
 * to itself); bit 7 = the visit orders of the NMS share the bookkeeping download stream.  "host_batch_out" (default 0) tells that
 * entry point what the caller's label buffers hold ON ENTRY, again with identical results: 0 = anything (every buffer is
 * zero-filled, 2 bytes of host memory traffic per voxel -- the bound of the call), 1 = zeros (a fresh np.zeros / calloc
 * buffer, what tools/binarization_soma.py:57 allocates per volume), 2 = exactly what this entry point wrote into the same
 * buffer last time (only the 64-byte lines written then are cleared; a buffer the library has no record of is zero-filled).
 * Unknown names return B200SEG_EINVAL. */
int b200seg_set_option(const char* name, int value);

/* ----------------------------------------------------------------------------------------------
 * 3D NMS  -- replaces lib/utils/cython_nms_3d.pyx:39-96 (nms_3d) and :102-159 (nms_3d_volume),
 *            reached through lib/utils/boxes_3d.py:364-374.
 * Batched: `batch` independent detection sets stored back to back in dets[total,7]
 * (x1,y1,z1,x2,y2,z2,score fp32); set b owns rows [offsets[b], offsets[b+1]).
 *   keep        [total] int64: for set b, keep[offsets[b] + i], i < keep_count[b], are the kept
 *               ORIGINAL (set-local) indices in ascending order (== np.where(suppressed==0)[0]).
 *   keep_count  [batch] int32.
 *   rank_order  [total] int32 or NULL: the same kept indices in VISIT order (key descending),
 *               i.e. the order tools/binarization_soma.py:60-62 sorts them into.
 * n_max = upper bound of the per-set size (host-known, sizes the grid and the workspace).
 * ---------------------------------------------------------------------------------------------- */
size_t b200seg_nms3d_workspace_bytes(int batch, int n_max);
int b200seg_nms3d_dev(const float* dets, const int32_t* offsets, int batch, int n_max,
                      float thresh, int by_volume,
                      int64_t* keep, int32_t* keep_count, int32_t* rank_order,
                      void* workspace, size_t workspace_bytes, b200seg_stream_t stream);
/* numpy seam: dets on the host, keep[n] on the host, *n_keep returned. */
int b200seg_nms3d_host(const float* dets, int n, float thresh, int by_volume,
                       int64_t* keep, int* n_keep);

/* ----------------------------------------------------------------------------------------------
 * 3D box IoU matrix -- replaces lib/utils/cython_bbox_3d.pyx:32-80 (alias boxes_3d.py:55).
 * boxes [N,6], query [K,6] fp32 -> overlaps [N,K] fp32, with the reference's mixed fp32/fp64
 * arithmetic reproduced bit for bit.
 * ---------------------------------------------------------------------------------------------- */
int b200seg_iou3d_dev(const float* boxes, long long N, const float* query, long long K,
                      float* overlaps, b200seg_stream_t stream);
int b200seg_iou3d_host(const float* boxes, long long N, const float* query, long long K,
                       float* overlaps);

/* ----------------------------------------------------------------------------------------------
 * RoIAlign3D -- replaces roi_align_forward_cuda_3d / roi_align_backward_cuda_3d
 * (lib/modeling/roi_xfrom/roi_align_3d/src/roi_align_cuda_3d.h:1-5, kernels in
 * src/roi_align_kernel_3d.cu:81-151 and :238-338).
 * features [B,C,S,H,W] contiguous (fp32 or bf16), rois [R,7] fp32 (batch,x1,y1,z1,x2,y2,z2),
 * output [R,C,Ps,Ph,Pw] of the feature dtype.  Outputs need NOT be pre-zeroed.
 * layout: 0 = reference (forward emits bins in (H,W,S) order into the [.,.,Ps,Ph,Pw] tensor,
 *             backward reads grad_out in (S,H,W) order and uses the z guard -0.1; this is what the
 *             reference computes, quirks included),
 *         1 = "shw": forward emits (S,H,W) order and backward is its exact adjoint.
 * The backward is deterministic and atomics-free: every grad_in element is written exactly once.
 * ---------------------------------------------------------------------------------------------- */
/* workspace of the backward: per-RoI footprint boxes + per-axis adjoint tables (R x (S+H+W) x 8|16 floats) */
size_t b200seg_roialign3d_workspace_bytes(int R, int S, int H, int W, int P_max);
/* workspace that also enables the round-2 backward for pooled sizes <= 8 (per-RoI footprint gradients, one slot of
 * 8^3 voxels x 16 channels per (RoI, channel group): R x ceil(C/16) x 32 KB); b200seg_roialign3d_bwd_dev picks the path by
 * the size it is handed -- a workspace of the smaller size above keeps the round-1 kernel, same results to rounding */
size_t b200seg_roialign3d_bwd_workspace_bytes(int R, int C, int S, int H, int W, int P_max);
int b200seg_roialign3d_fwd_dev(const void* features, int dtype, const float* rois, void* output,
                               int B, int C, int S, int H, int W, int R,
                               int Ps, int Ph, int Pw, float spatial_scale, int sampling_ratio,
                               int layout, b200seg_stream_t stream);
int b200seg_roialign3d_bwd_dev(const void* grad_out, int dtype, const float* rois, void* grad_in,
                               int B, int C, int S, int H, int W, int R,
                               int Ps, int Ph, int Pw, float spatial_scale, int sampling_ratio,
                               int layout, void* workspace, size_t workspace_bytes,
                               b200seg_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * PRM peak stimulation -- replaces lib/prm/peak_stimulation_3d.py:9-41 (forward) with the
 * median threshold of lib/prm/peak_response_mapping_3d.py:45-49 fused in.
 * input [B,A,S,H,W] fp32.  A voxel is a peak when it is the arg-max of its own win^3 window under
 * ATen max_pool3d rules (first maximum in z,y,x scan order wins, -inf padding).
 * filter_mode: 0 none, 1 lower median of the (b,a) volume (torch.median), 2 thresholds given in
 *              thr_in[B*A] (the caller evaluated an arbitrary peak_filter callable).
 * peaks [cap,5] int64 rows (b,a,z,y,x) in lexicographic order (torch.nonzero order);
 * *n_peaks (device int32) = total number of peaks (may exceed cap; extra rows are dropped).
 * agg [B*A] fp32 or NULL: sum(input*peak)/sum(peak) (NaN when a map has no peak).
 * thr_out [B*A] fp32 or NULL: thresholds used.   win must be odd, 3 <= win <= 7.
 * ---------------------------------------------------------------------------------------------- */
size_t b200seg_peaks3d_workspace_bytes(int B, int A, int S, int H, int W);
int b200seg_peaks3d_dev(const float* input, int B, int A, int S, int H, int W, int win,
                        int filter_mode, const float* thr_in,
                        int64_t* peaks, int cap, int32_t* n_peaks, float* agg, float* thr_out,
                        void* workspace, size_t workspace_bytes, b200seg_stream_t stream);
/* backward of the aggregation (peak_stimulation_3d.py:43-48): grad_in = peak_map * grad_agg[b,a],
 * peaks as produced by the forward; grad_in [B,A,S,H,W] is fully written (zeros elsewhere). */
int b200seg_peaks3d_bwd_dev(const int64_t* peaks, const int32_t* n_peaks, int cap,
                            const float* grad_agg, float* grad_in,
                            int B, int A, int S, int H, int W, b200seg_stream_t stream);
/* Maps (since the library was loaded) whose sampled median interval missed and that took the slow exact selection inside
 * peaks3d (statistics; the result is exact either way).  "peaks_median_mode" option: 0 = automatic (sampled interval for
 * maps of 2^16..2^21 elements, level-1 histogram otherwise), 1 = histogram path only, 2 = force the miss. */
int b200seg_peaks3d_fallback_count(unsigned long long* count);

/* ----------------------------------------------------------------------------------------------
 * 2D-Otsu binarization -- replaces tools/otsu.py:199-284 (otsu_py_2d_fast, k = -1), batched over
 * instance crops.  Crop i owns samples [crop_off[i], crop_off[i+1]) of image/prm (uint16, the
 * dtype both binarization scripts pass).  Outputs per crop: mask (0/255 uint8, same packing),
 * b_max[i], g_info[i*4..] = {g_min, g_max, prm_min, prm_max}, status[i]: 0 ok, 1 = no b wins
 * (the reference raises on k_max, otsu.py:277; mask is then all 255), 2 = empty crop.
 * hist (optional, may be NULL): uint32 counts of hist.T ([prm_bin][img_bin], G*G per crop) at
 * hist_off[i] (int64 element offsets, caller sized from G = g_max-g_min+1).
 * ---------------------------------------------------------------------------------------------- */
int b200seg_otsu2d_dev(const uint16_t* image, const uint16_t* prm, const int64_t* crop_off,
                       int n_crops, uint8_t* mask, int32_t* b_max, int32_t* g_info, int32_t* status,
                       uint32_t* hist, const int64_t* hist_off, b200seg_stream_t stream);
/* numpy seam: one crop on the host (image, prm uint16[n]) -> mask uint8[n], *b_max; returns 0, or
 * 1 when no threshold exists (caller raises, like the reference). */
int b200seg_otsu2d_host(const uint16_t* image, const uint16_t* prm, long long n,
                        uint8_t* mask, int* b_max);

/* ----------------------------------------------------------------------------------------------
 * Per-instance binarization straight from the raw volume (tools/binarization_soma.py:78-94):
 * crop by the int()-truncated box, normalise image and PRM (:85-91), 2D-Otsu (:94); one launch for
 * every instance of a batch of equally shaped volumes.
 *   volumes [n_volumes,S,H,W] uint8; det_off [n_volumes+1] int32 (device): volume v owns instances
 *   [det_off[v], det_off[v+1]) of boxes / crop_off / order / b_max / status (may be NULL when
 *   n_volumes == 1: the single volume then owns instances [0, n_max)); n_max = upper bound of the
 *   per-volume instance count (host-known, sizes the grid);
 *   boxes [total,6] int32 inclusive voxel coords inside the volume;
 *   prm: raw uint8 PRM crops (box-shaped, [sz,sy,sx]), instance i at crop_off[i] (int64 [total+1]);
 *   order [total] int32 or NULL: volume v visits instance det_off[v] + order[det_off[v] + r] r-th
 *   (visit order from NMS); n_valid [n_volumes] int32 (device) or NULL: only the first n_valid[v]
 *   visits of volume v are processed.
 * Outputs as b200seg_otsu2d_dev, plus status 3 = PRM crop has no positive voxel (skipped,
 * binarization_soma.py:74-76, mask all 0).  Masks are written in the packing of `prm`.
 * ---------------------------------------------------------------------------------------------- */
int b200seg_soma_binarize_dev(const uint8_t* volumes, int n_volumes, int S, int H, int W,
                              const int32_t* det_off, int n_max,
                              const int32_t* boxes, const uint8_t* prm, const int64_t* crop_off,
                              const int32_t* order, const int32_t* n_valid,
                              uint8_t* mask, int32_t* b_max, int32_t* status,
                              b200seg_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * Largest connected component of every instance mask, in place -- replaces tools/binarization_soma.py:97-99
 * (skimage.measure.label with full = 26-connectivity, then the component with the largest voxel count; equal
 * counts: the component whose first voxel comes later in raster order, i.e. what a stable argsort()[-1] picks).
 * masks / crop_off / boxes / det_off / n_max / order / n_valid as in b200seg_soma_binarize_dev; mask bytes are
 * "set" when non-zero.  status (may be NULL): only instances with status 0 are processed; a mask without any
 * foreground (the reference raises IndexError there) gets status 5, a degenerate crop whose bookkeeping does not
 * fit the workspace gets status 6 and an empty mask.  workspace: b200seg_largest_cc_workspace_bytes(total mask bytes).
 * ---------------------------------------------------------------------------------------------- */
/* Diagnostics: how many instances took which path of b200seg_largest_cc_dev since the last reset (synchronises the
 * device): [0] filled by the warp-per-instance flood fill, [1] geometry outside its limits (rows > 64 voxels, > 32
 * planes), [2] no foreground, [3] empty centre row, [4] fill not converged, [5] seed component without the majority
 * ([1]..[5] are finished by the general union-find kernel), [6..7] reserved. */
int b200seg_largest_cc_path_counts(long long* counts, int reset);
size_t b200seg_largest_cc_workspace_bytes(long long total_mask_bytes);
int b200seg_largest_cc_dev(uint8_t* masks, const int64_t* crop_off, long long total_mask_bytes,
                           int n_volumes, const int32_t* det_off, int n_max, const int32_t* boxes,
                           const int32_t* order, const int32_t* n_valid, int32_t* status,
                           void* workspace, size_t workspace_bytes, b200seg_stream_t stream);
/* Same with the tie rule as a parameter: tie_first != 0 -> among equally large components the one whose first voxel comes
 * FIRST in raster order wins (np.argmax over ascending labels, tools/binarization_nuclei.py:126-130). */
int b200seg_largest_cc_ex_dev(uint8_t* masks, const int64_t* crop_off, long long total_mask_bytes,
                              int n_volumes, const int32_t* det_off, int n_max, const int32_t* boxes,
                              const int32_t* order, const int32_t* n_valid, int32_t* status, int tie_first,
                              void* workspace, size_t workspace_bytes, b200seg_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * Label paste-back -- replaces the inline numpy of tools/binarization_soma.py:100-104 (and
 * binarization_nuclei.py:141-149): instances are visited in order, instance i writes label
 * ids[i] where its mask is set and the label volume is still 0.  seg [n_volumes,S,H,W] uint16 is
 * written exactly once per voxel (zero-fill folded in; no pre-clear needed).  Batched over volumes
 * like b200seg_soma_binarize_dev (det_off / n_max / order / n_valid have the same meaning).
 *   ids [n_max] uint16 indexed by visit rank (shared by all volumes);
 *   survive (by visit rank, volume v at det_off[v]): 1 when the label is present in the volume
 *   (`mask_id in np.unique(seg)`, :103).  With det_off == NULL survive[n_max] is cleared by the
 *   call; with det_off given the caller clears survive[total] beforehand.
 * ---------------------------------------------------------------------------------------------- */
size_t b200seg_paste_labels_workspace_bytes(int n_volumes, int S, int H, int W, int n_max);
int b200seg_paste_labels_dev(uint16_t* seg, int n_volumes, int S, int H, int W,
                             const int32_t* det_off, int n_max,
                             const int32_t* boxes, const uint16_t* ids,
                             const uint8_t* masks, const int64_t* mask_off,
                             const int32_t* order, const int32_t* n_valid,
                             uint8_t* survive, void* workspace, size_t workspace_bytes,
                             b200seg_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * Post-processing chain of tools/binarization_soma.py:57-104 for a batch of volumes, device
 * resident: 3D NMS (:57) -> visit in descending score (:60-62) -> per-instance crop + normalise +
 * 2D-Otsu (:78-94) -> label paste-back + survivor test (:100-104).
 *   volumes [nv,S,H,W] uint8; detection set b = rows [det_off[b], det_off[b+1]) of dets[total,7];
 *   det_off is given twice: device copy (kernels) and host copy (launch geometry);
 *   boxes [total,6] int32 = int()-truncated detections clipped to the volume (:78);
 *   prm = raw uint8 PRM crops, crop i ([sz,sy,sx] of box i) at crop_off[i] (int64 [total+1], device).
 * Outputs: seg [nv,S,H,W] uint16 (label = visit rank + 1, :67), keep / keep_count / rank_order as
 * b200seg_nms3d_dev, masks (packing of prm), b_max[total], status[total] (-1 = suppressed by NMS,
 * else as b200seg_soma_binarize_dev), survive[total] (set b, visit rank r at det_off[b] + r).
 * keep_largest_cc != 0 inserts the largest-connected-component filter of :97-99 (b200seg_largest_cc_dev) between
 * the binarization and the paste (reference semantics); prm_bytes = crop_off[total] (host-known size of the
 * packed PRM / mask arrays); the workspace size takes cc_mask_bytes = prm_bytes when the filter is on, else 0.
 * status additionally reports 5 (mask without foreground) and 6 (degenerate crop) from that step.
 * ---------------------------------------------------------------------------------------------- */
size_t b200seg_postproc_soma_workspace_bytes(int n_volumes, int n_max, int S, int H, int W, long long cc_mask_bytes);
int b200seg_postproc_soma_dev(const uint8_t* volumes, int n_volumes, int S, int H, int W,
                              const float* dets, const int32_t* det_off_dev, const int32_t* det_off_host,
                              const int32_t* boxes, const uint8_t* prm, const int64_t* crop_off,
                              long long prm_bytes, float nms_thresh, int keep_largest_cc,
                              uint16_t* seg, int64_t* keep, int32_t* keep_count, int32_t* rank_order,
                              uint8_t* masks, int32_t* b_max, int32_t* status, uint8_t* survive,
                              void* workspace, size_t workspace_bytes, b200seg_stream_t stream);
/* One volume, HOST buffers in and out (H2D / D2H inside; returns when seg is in host memory). */
int b200seg_postproc_soma_host(const uint8_t* volume, int S, int H, int W,
                               const float* dets, int n, const int32_t* boxes,
                               const uint8_t* prm, const int64_t* crop_off, float nms_thresh,
                               int keep_largest_cc,
                               uint16_t* seg, int* n_keep, int32_t* rank_order,
                               int32_t* b_max, int32_t* status, uint8_t* survive);

/* ----------------------------------------------------------------------------------------------
 * Run-length codec of binary 3D masks -- replaces lib/utils/mask_3d.py:15-71 (binary_mask_to_rle,
 * rle_to_binary_mask) and lib/utils/cython_mask_3d.pyx:19-77.  mask [S,H,W] uint8 C-contiguous (set = non-zero);
 * counts = lengths of the alternating runs of the FORTRAN-order ravel (first axis fastest), starting with a run
 * of zeros (length 0 when the first element is set); an all-zero mask is the single count S*H*W.
 * encode: at most `cap` counts are written, *n_counts (device int64) always holds the true number.
 * decode: writes the whole mask; *sum_out (device int64, may be NULL) = sum of the counts, which the caller
 * compares with S*H*W like mask_3d.py:53.  workspace: b200seg_rle3d_workspace_bytes(S, H, W, cap or n_counts).
 * ---------------------------------------------------------------------------------------------- */
size_t b200seg_rle3d_workspace_bytes(int S, int H, int W, long long cap);
int b200seg_rle3d_encode_dev(const uint8_t* mask, int S, int H, int W, int64_t* counts, long long cap,
                             int64_t* n_counts, void* workspace, size_t workspace_bytes, b200seg_stream_t stream);
int b200seg_rle3d_decode_dev(const int64_t* counts, long long n_counts, uint8_t* mask, int S, int H, int W,
                             int64_t* sum_out, void* workspace, size_t workspace_bytes, b200seg_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * Pairwise instance-mask overlaps between two label volumes -- replaces tools/evaluation/mask_iou.py:49-109
 * (mask_iou_fast, mask_ios_fast, mask_iog_fast) as called by eval_instance_segmentation_soma.py:186-197, where the
 * boolean stacks are cut out of label volumes: pred_masks[i] = (pred == id_i), gt_masks[k] = (gt == id_k).
 *   pred, gt [n_voxels] uint16 label volumes; lut_pred / lut_gt [65536] int32: label id -> row (0-based) or -1;
 *   iou / ios / iog [n_pred, n_gt] fp32 (any may be NULL): I/(A+B-I), I/A, I/B with A = pred area, B = gt area, each
 *   computed as float32(double / double); 0/0 gives NaN (the numba reference raises ZeroDivisionError there).
 *   inter [(n_pred+1),(n_gt+1)] int64 (row / column 0 = voxels whose id is unlisted or background; cell (0,0)
 *   excludes the voxels that are background in both volumes), area_pred [n_pred+1], area_gt [n_gt+1]: optional.
 * ---------------------------------------------------------------------------------------------- */
size_t b200seg_mask_overlaps_workspace_bytes(int n_pred, int n_gt);
int b200seg_mask_overlaps_dev(const uint16_t* pred, const uint16_t* gt, long long n_voxels,
                              const int32_t* lut_pred, const int32_t* lut_gt, int n_pred, int n_gt,
                              float* iou, float* ios, float* iog, int64_t* inter, int64_t* area_pred, int64_t* area_gt,
                              void* workspace, size_t workspace_bytes, b200seg_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * Whole-volume prefilters and normalisations.
 *
 * b200seg_gaussian3d_dev replaces  scipy.ndimage.gaussian_filter(img, sigma)  as called on the raw integer volume by
 *   tools/binarization_nuclei.py:43 : three separable passes (axis 0, 1, 2), 'reflect' borders, fp64 accumulation in
 *   scipy's order, every pass truncated to the volume's dtype -- bit exact.  `weights` (HOST pointer) holds the first
 *   radius+1 taps of the symmetric kernel, weights[radius] = centre; the host side computes them exactly as scipy
 *   does (radius = int(truncate*sigma + 0.5) <= 8).  elem_bytes: 1 = uint8, 2 = uint16.  in != out.
 * b200seg_median3d_dev replaces  scipy.ndimage.median_filter(img, size=3)  (tools/binarization_nuclei.py:44).
 * b200seg_zscore_norm_dev replaces  (im - mean(im[im>0])) / std(im[im>0])  (tools/infer_simple.py:180-183,
 *   lib/utils/blob.py:180-184).  elem_bytes 1 / 2 = uint8 / uint16 (exact integer moments), 4 = float32 (two-pass
 *   fp64 moments, deterministic).  out [n] float32 (may be NULL: statistics only); stats [3] double on the device =
 *   {mean, std, count of non-zero voxels}.
 * b200seg_prm_to_u8_dev replaces the per-channel  fm -= min; fm /= max; fm *= 255; astype(uint8)  of
 *   tools/infer_simple.py:233-238 (fp32, same operation order; a constant map gives 0).  in [n_maps, per_map].
 * ---------------------------------------------------------------------------------------------- */
int b200seg_gaussian3d_dev(const void* in, void* out, int elem_bytes, int S, int H, int W, const double* weights, int radius,
                           b200seg_stream_t stream);
int b200seg_median3d_dev(const void* in, void* out, int elem_bytes, int S, int H, int W, b200seg_stream_t stream);
size_t b200seg_zscore_workspace_bytes(void);
int b200seg_zscore_norm_dev(const void* in, int elem_bytes, float* out, long long n, double* stats,
                            void* workspace, size_t workspace_bytes, b200seg_stream_t stream);
size_t b200seg_prm_to_u8_workspace_bytes(int n_maps);
int b200seg_prm_to_u8_dev(const float* in, uint8_t* out, int n_maps, long long per_map,
                          void* workspace, size_t workspace_bytes, b200seg_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * RPN proposal generation -- replaces GenerateProposalsOp_3d.forward /
 * proposals_for_one_image / _filter_boxes_3d (lib/modeling/generate_proposals_3d.py:20-192) with bbox_transform_3d
 * and clip_tiled_boxes_3d (lib/utils/boxes_3d.py:144-225): top pre_nms_topN scores -> decode -> clip -> filter ->
 * NMS -> first post_nms_topN, all on the device.
 *   scores [n_images, A, S, H, W], deltas [n_images, 6A, S, H, W] fp32 on the device;
 *   im_info [n_images, 4] (slices, height, width, scale) and anchors [A, 6] (cell anchors) on the HOST;
 *   pre_nms_topN <= 0 or >= A*S*H*W: every anchor is a candidate (:125-126); nms_thresh <= 0: no NMS and no
 *   post_nms_topN cut (:163); post_nms_topN <= 0: no cut (:166).
 * Outputs, `cap` = b200seg_generate_proposals_capacity(...) rows per image: rois [n_images*cap, 7]
 *   (image, x1, y1, z1, x2, y2, z2), probs [n_images*cap], keep_idx [n_images*cap] int64 = index into the image's score
 *   map flattened in (S,H,W,A) order; image i fills rows [i*cap, i*cap + counts[i]).
 * Equal scores: the smaller (S,H,W,A) index ranks first (the reference's argpartition/argsort are unstable there).
 * ---------------------------------------------------------------------------------------------- */
size_t b200seg_generate_proposals_workspace_bytes(int A, int S, int H, int W, int pre_nms_topN, int post_nms_topN, float nms_thresh);
int b200seg_generate_proposals_capacity(int A, int S, int H, int W, int pre_nms_topN, int post_nms_topN, float nms_thresh);
int b200seg_generate_proposals_dev(const float* scores, const float* deltas, const float* im_info, int n_images,
                                   int A, int S, int H, int W, const float* anchors, float feat_stride,
                                   int pre_nms_topN, int post_nms_topN, float nms_thresh, float min_size,
                                   float* rois, float* probs, int64_t* keep_idx, int32_t* counts,
                                   void* workspace, size_t workspace_bytes, b200seg_stream_t stream);

/* A batch of equally shaped volumes, HOST buffers in and out (the per-volume loop of tools/my_subprocess.py:56 over
 * tools/binarization_soma.py:57-104).  Pipelined over twelve device slots: the NMS runs ahead and host threads pack the
 * image crops of its survivors into a pinned buffer that travels by DMA (the raw volume does not), the PRM crops of
 * the same survivors are pulled by a gather kernel straight from the caller's buffer when it is pinned and 16-byte
 * aligned (else the whole packed array is copied), and the label volume comes back as its non-zero 64-byte lines,
 * which a pool of host threads (B200SEG_HOST_THREADS, default 3/4 of cores / LOCAL_WORLD_SIZE, at most 32) writes into
 * the caller's volume after zero-filling it -- or which the GPU writes in place when seg[v] is pinned and the host is
 * shared by several ranks (a volume with more than 1/4 of its lines labelled is copied densely).
 * seg[v] may be pageable; pass pinned volumes / prm / seg for full overlap.  "host_batch_mode" (b200seg_set_option)
 * selects the transfer scheme, "host_batch_out" tells what seg[v] holds on entry; results are identical either way.
 * Every array argument has n_volumes entries; per-volume meanings as in b200seg_postproc_soma_host.
 * b200seg_postproc_soma_host_batch_traffic reports the bytes the last successful call moved over the link. */
void b200seg_postproc_soma_host_batch_traffic(unsigned long long* h2d_bytes, unsigned long long* d2h_bytes);
int b200seg_postproc_soma_host_batch(int n_volumes, int S, int H, int W,
                                     const uint8_t* const* volumes, const float* const* dets, const int32_t* n_dets,
                                     const int32_t* const* boxes, const uint8_t* const* prm,
                                     const int64_t* const* crop_off, float nms_thresh, int keep_largest_cc,
                                     uint16_t* const* seg, int32_t* n_keep, int32_t* const* rank_order,
                                     int32_t* const* b_max, int32_t* const* status, uint8_t* const* survive);

/* ----------------------------------------------------------------------------------------------
 * Mask R-CNN mask paste-back (second row) -- replaces the per-detection body of segm_results
 * (lib/core/test.py:902-938): zero-pad the M^3 mask block to (M+2)^3, skimage.transform.resize(padded, (s,h,w),
 * mode='reflect', anti_aliasing=True) (:919), `> THRESH_BINARIZE` (:920), paste the part inside the volume (:921-931).
 * The resize arithmetic is scipy.ndimage's (gaussian_filter 'mirror' with sigma = max(0, (in/out-1)/2), then
 * zoom(order=1, 'mirror', grid_mode=True)), reproduced in fp64 in scipy's operation order; see segm_paste.cu.
 *   masks       fp32 blocks of M*M*M values; detection d uses block mask_index[d]
 *               (= d*C + class for MRCNN.CLS_SPECIFIC_MASK, d*C otherwise, :903-906)
 *   ref_boxes   [n,6] int32: boxes after expand_boxes(ref_boxes, (M+2)/M) and .astype(int32) (:895-898) -- host logic
 *   gauss_w     [(M+2)][2*(M+2)] fp64 (b200seg_segm_gauss_table_size(M) values): row o = the first R+1 taps of the
 *               normalised Gaussian scipy builds for an axis resized from M+2 to o < M+2 samples
 *               (sigma = ((M+2)/o - 1)/2, R = int(4*sigma + 0.5), phi = exp(-0.5/sigma^2 * x^2), x = -R..0, / sum(phi));
 *               rows with R = 0 are unused.  Built by the caller so that exp() is the caller's (numpy's in the Python host).
 *   crops       packed uint8 {0,1}: detection d owns crop_off[d] .. crop_off[d+1], shape (z1-z0, y1-y0, x1-x0) in C order
 *               with x0 = max(bx1,0), x1 = min(bx2+1, im_w) ... (:923-928); boxes that miss the volume own 0 bytes
 *               (the reference's negative slice bounds are undefined there).
 * b200seg_segm_expand_dev writes the reference's output format: volumes [n, im_s, im_h, im_w] uint8, zero outside the boxes.
 * ---------------------------------------------------------------------------------------------- */
int b200seg_segm_gauss_table_size(int M);
int b200seg_segm_paste_dev(const float* masks, const int32_t* mask_index, const int32_t* ref_boxes, int n, int M,
                           const double* gauss_w, float thresh, int im_s, int im_h, int im_w,
                           uint8_t* crops, const int64_t* crop_off, b200seg_stream_t stream);
int b200seg_segm_expand_dev(const uint8_t* crops, const int64_t* crop_off, const int32_t* ref_boxes, int n,
                            int im_s, int im_h, int im_w, uint8_t* volumes, b200seg_stream_t stream);
/* numpy seam: every pointer on the host (masks holds n_mask_blocks blocks); crops come back packed. */
int b200seg_segm_paste_host(const float* masks, long long n_mask_blocks, const int32_t* mask_index, const int32_t* ref_boxes,
                            int n, int M, const double* gauss_w, float thresh, int im_s, int im_h, int im_w,
                            uint8_t* crops, const int64_t* crop_off);

/* ----------------------------------------------------------------------------------------------
 * Per-instance chain of tools/binarization_nuclei.py:92-149 for ONE volume, for the n instances that survived the
 * detection filters (:72-87: edge filter, nms_3d_volume -> b200seg_nms3d_dev(by_volume = 1), score > 0.4), in visit order:
 * crop + normalise (:98-121) -> 2D-Otsu (:124) -> largest 26-connected component (:126-130) -> hole filling = all but the
 * largest 26-connected component of the complement (:132-137) -> binary closing, 6-neighbour cross (:139) -> first-come
 * label paste with label = visit rank + 1 (:141-147) -> survivor test (:148-149).
 *   volume [S,H,W] uint8 (elem_bytes 1) or uint16 (2): the prefiltered image (:43-44, b200seg_gaussian3d_dev / median3d_dev);
 *   boxes [n,6] int32: the clamped boxes (:101-106) in VOLUME coordinates (tile offset added), inclusive;
 *   prm: uint8 PRM crops (box-shaped [sz,sy,sx]), instance i at crop_off[i] (int64 [n+1], device); total_voxels = crop_off[n].
 * Outputs: seg [S,H,W] uint16 (fully written), masks (final per-instance masks 0/255, packing of prm), b_max[n], status[n]:
 *   0 ok; 1 Otsu found no threshold; 2 box outside the volume or crop size mismatch; 4 more than 2048 gray levels;
 *   5 no foreground / no background left (the reference raises); 6 degenerate crop (see b200seg_largest_cc_dev);
 *   (where the script's normalisation divides 0 by 0 -- gray_max == 0 under the stretch, constant PRM crop -- numpy's NaN
 *   becomes 0 in the uint16 cast and the script carries on; the kernels do the same, status 7 of round 1 is gone);
 *   survive[n]: label present in seg.  Instances with status != 0 paste nothing.
 * Connected components and closing are cc3d / skimage in the reference (unversioned, not vendored): parity is pinned
 * against scipy.ndimage (label with the 3x3x3 structure, binary_dilation / binary_erosion(border_value=1)).
 * ---------------------------------------------------------------------------------------------- */
size_t b200seg_binarize_nuclei_workspace_bytes(int n, long long total_voxels, int S, int H, int W);
int b200seg_binarize_nuclei_dev(const void* volume, int elem_bytes, int S, int H, int W,
                                const int32_t* boxes, const uint8_t* prm, const int64_t* crop_off, int n, long long total_voxels,
                                uint16_t* seg, uint8_t* masks, int32_t* b_max, int32_t* status, uint8_t* survive,
                                void* workspace, size_t workspace_bytes, b200seg_stream_t stream);
/* numpy seam: every pointer on the host; masks may be NULL. */
int b200seg_binarize_nuclei_host(const void* volume, int elem_bytes, int S, int H, int W,
                                 const int32_t* boxes, const uint8_t* prm, const int64_t* crop_off, int n,
                                 uint16_t* seg, uint8_t* masks, int32_t* b_max, int32_t* status, uint8_t* survive);

/* ----------------------------------------------------------------------------------------------
 * Evaluation helpers: the volume-sized parts of tools/evaluation/eval_instance_segmentation_soma.py
 * and tools/evaluation/evaluation_nuclei_f1score_seg.py; the greedy matching on the (few hundred) rows stays on the host.
 * b200seg_label_presence_dev: present[65536] uint8 (device), present[v] = 1 <=> label v occurs in labels[n] -- replaces
 *   np.unique(label volume) (eval_instance_segmentation_soma.py:177-181).
 * b200seg_eval_voxel_counts_dev: counts[3] uint64 (device) = { tp_pixel, gt_pixel, pre_pixel } of
 *   evaluation_nuclei_f1score_seg.py:86-89 / :124-130: gt_pixel = #(gt > 0), pre_pixel = #(pred > 0),
 *   tp_pixel = #(pred > 0 and gt > 0 and inside the box of a matched detection); boxes [n_boxes,6] int32 = the matched
 *   detections' bbox.astype(int), inclusive, clipped to the volume here (device).
 * ---------------------------------------------------------------------------------------------- */
int b200seg_label_presence_dev(const uint16_t* labels, long long n, uint8_t* present, b200seg_stream_t stream);
size_t b200seg_eval_voxel_counts_workspace_bytes(long long n_voxels);
int b200seg_eval_voxel_counts_dev(const uint16_t* pred, const uint16_t* gt, int S, int H, int W,
                                  const int32_t* boxes, int n_boxes, unsigned long long* counts,
                                  void* workspace, size_t workspace_bytes, b200seg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SEG_H_ */
