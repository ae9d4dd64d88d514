// pipeline.cu -- the per-volume post-processing chain of tools/binarization_soma.py:57-104 as one
// device-resident sequence: cross-tile 3D NMS (:57) -> visit order by descending score (:60-62) ->
// per-instance crop + normalisation + 2D-Otsu (:78-94) -> label paste-back with survivor test
// (:100-104).  No host round trip between the steps: the NMS survivor count stays on the device
// and sizes the later launches through `n_valid` pointers.
//
// The reference keeps only the largest connected component of each Otsu mask (skimage.measure.label, :97-99)
// before pasting: largest_cc.cu, switched by `keep_largest_cc` (1 = reference semantics).
#include "common.cuh"

#include <mutex>
#include <stdlib.h>

namespace b200seg {

__global__ void iota_u16_kernel(uint16_t* p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint16_t)(i + 1);
}

}  // namespace b200seg

namespace b200seg {

SomaChainWs soma_chain_ws(void* workspace, size_t workspace_bytes, int n_volumes, int n_max, int S, int H, int W, long long cc_bytes) {
    SomaChainWs L;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    L.ids = (uint16_t*)ws;
    L.paste_ws = ws + align_up((size_t)(n_max > 0 ? n_max : 1) * 2, 256);
    L.paste_ws_bytes = align_up(b200seg_paste_labels_workspace_bytes(n_volumes, S, H, W, n_max), 256);
    L.cc_ws = L.paste_ws + L.paste_ws_bytes;
    L.cc_ws_bytes = cc_bytes > 0 ? align_up(b200seg_largest_cc_workspace_bytes(cc_bytes), 256) : 0;
    L.nms_ws = L.cc_ws + L.cc_ws_bytes;
    L.nms_ws_bytes = workspace_bytes - (size_t)(L.nms_ws - (char*)workspace);
    return L;
}

// Everything of the chain that follows the NMS (shared with the pipelined host entry point of host_batch.cu, which runs
// the NMS first and then fetches only the PRM crops of the survivors).
int postproc_soma_after_nms(const uint8_t* volumes, int n_volumes, int S, int H, int W, const int32_t* det_off_dev, int n_max,
                            int total, const int32_t* boxes, const uint8_t* prm, const int64_t* crop_off, long long prm_bytes,
                            int keep_largest_cc, uint16_t* seg, const int32_t* keep_count, const int32_t* rank_order,
                            uint8_t* masks, int32_t* b_max, int32_t* status, uint8_t* survive, uint16_t* ids,
                            void* paste_ws, size_t paste_ws_bytes, void* cc_ws, size_t cc_ws_bytes, cudaStream_t stream,
                            const PasteLines* lines) {
    int e = 0;
    if (n_max > 0) {
        iota_u16_kernel<<<(n_max + 255) / 256, 256, 0, stream>>>(ids, n_max);
        B200_LAUNCH_CHECK("iota_u16_kernel");
        B200_CUDA(cudaMemsetAsync(status, 0xFF, sizeof(int32_t) * (size_t)total, stream));   // -1 = not visited (suppressed)
        B200_CUDA(cudaMemsetAsync(survive, 0, (size_t)total, stream));
        B200_CUDA(cudaMemsetAsync(b_max, 0, sizeof(int32_t) * (size_t)total, stream));            // suppressed detections report 0
        // one launch for the instances of every volume: grid (n_max, n_volumes)
        e = b200seg_soma_binarize_dev(volumes, n_volumes, S, H, W, det_off_dev, n_max, boxes, prm, crop_off,
                                      rank_order, keep_count, masks, b_max, status, stream);
        if (e) return e;
        if (keep_largest_cc) {                               // binarization_soma.py:97-99
            e = b200seg_largest_cc_dev(masks, crop_off, prm_bytes, n_volumes, det_off_dev, n_max, boxes, rank_order, keep_count,
                                       status, cc_ws, cc_ws_bytes, stream);
            if (e) return e;
        }
    }
    return paste_labels_launch(seg, n_volumes, S, H, W, det_off_dev, n_max, boxes, ids, masks, crop_off,
                               rank_order, keep_count, survive, paste_ws, paste_ws_bytes, stream, lines);
}

}  // namespace b200seg

using namespace b200seg;

extern "C" size_t b200seg_postproc_soma_workspace_bytes(int n_volumes, int n_max, int S, int H, int W, long long cc_mask_bytes) {
    return align_up(b200seg_nms3d_workspace_bytes(n_volumes, n_max), 256) + align_up((size_t)(n_max > 0 ? n_max : 1) * 2, 256) +
           align_up(b200seg_paste_labels_workspace_bytes(n_volumes, S, H, W, n_max), 256) +
           (cc_mask_bytes > 0 ? align_up(b200seg_largest_cc_workspace_bytes(cc_mask_bytes), 256) : 0) + 512;
}

extern "C" int b200seg_postproc_soma_dev(const uint8_t* volumes, int n_volumes, int S, int H, int W,
                                         const float* dets, const int32_t* det_off_dev, const int32_t* det_off_host,
                                         const int32_t* boxes, const uint8_t* prm, const int64_t* crop_off,
                                         long long prm_bytes, float nms_thresh, int keep_largest_cc,
                                         uint16_t* seg, int64_t* keep, int32_t* keep_count, int32_t* rank_order,
                                         uint8_t* masks, int32_t* b_max, int32_t* status, uint8_t* survive,
                                         void* workspace, size_t workspace_bytes, b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(n_volumes >= 0 && S > 0 && H > 0 && W > 0, "postproc_soma: bad sizes");
    if (n_volumes == 0) return 0;
    B200_CHECK_ARG(volumes && det_off_dev && det_off_host && seg && keep_count && workspace, "postproc_soma: null pointer");
    int n_max = 0;
    for (int b = 0; b < n_volumes; ++b) {
        const int n = det_off_host[b + 1] - det_off_host[b];
        B200_CHECK_ARG(n >= 0, "postproc_soma: det offsets must be non-decreasing");
        if (n > n_max) n_max = n;
    }
    B200_CHECK_ARG(n_max < 65535, "postproc_soma: more than 65534 instances per volume do not fit uint16 labels");
    const int total = det_off_host[n_volumes];
    B200_CHECK_ARG(total == 0 || (dets && boxes && prm && crop_off && keep && rank_order && masks && b_max && status && survive),
                   "postproc_soma: null pointer");
    B200_CHECK_ARG(prm_bytes >= 0, "postproc_soma: negative prm_bytes");
    const long long cc_bytes = keep_largest_cc ? prm_bytes : 0;
    if (workspace_bytes < b200seg_postproc_soma_workspace_bytes(n_volumes, n_max, S, H, W, cc_bytes)) {
        set_error("postproc_soma: workspace too small");
        return B200SEG_EWORKSPACE;
    }
    const SomaChainWs L = soma_chain_ws(workspace, workspace_bytes, n_volumes, n_max, S, H, W, cc_bytes);
    uint16_t* ids = L.ids;
    char* paste_ws = L.paste_ws; const size_t paste_ws_bytes = L.paste_ws_bytes;
    char* cc_ws = L.cc_ws; const size_t cc_ws_bytes = L.cc_ws_bytes;
    char* nms_ws = L.nms_ws; const size_t nms_ws_bytes = L.nms_ws_bytes;

    int e = b200seg_nms3d_dev(dets, det_off_dev, n_volumes, n_max, nms_thresh, 0, keep, keep_count, rank_order,
                              nms_ws, nms_ws_bytes, stream);
    if (e) return e;
    return postproc_soma_after_nms(volumes, n_volumes, S, H, W, det_off_dev, n_max, total, boxes, prm, crop_off, prm_bytes,
                                   keep_largest_cc, seg, keep_count, rank_order, masks, b_max, status, survive, ids,
                                   paste_ws, paste_ws_bytes, cc_ws, cc_ws_bytes, stream);
}

// HOST-buffer entry point for ONE volume: what a drop-in binarization call site binds to.
// Copies the volume, detections, boxes and PRM crops to the device, runs the chain, and returns the
// label volume and the per-detection bookkeeping in host memory.
extern "C" int b200seg_postproc_soma_host(const uint8_t* volume, int S, int H, int W,
                                          const float* dets, int n, const int32_t* boxes,
                                          const uint8_t* prm, const int64_t* crop_off, float nms_thresh,
                                          int keep_largest_cc,
                                          uint16_t* seg, int* n_keep, int32_t* rank_order,
                                          int32_t* b_max, int32_t* status, uint8_t* survive) {
    B200_CHECK_ARG(S > 0 && H > 0 && W > 0 && n >= 0 && volume && seg && n_keep, "postproc_soma_host: bad arguments");
    B200_CHECK_ARG(n == 0 || (dets && boxes && prm && crop_off && rank_order && b_max && status && survive),
                   "postproc_soma_host: null pointer");
    HostCtx& hc = host_ctx();
    std::lock_guard<std::mutex> lock(hc.mu);
    const size_t V = (size_t)S * H * W;
    const size_t nn = n > 0 ? n : 1;
    const size_t prm_bytes = n > 0 ? (size_t)crop_off[n] : 0;
    const size_t ws_bytes = b200seg_postproc_soma_workspace_bytes(1, n, S, H, W, keep_largest_cc ? (long long)prm_bytes : 0);
    size_t total = Carver::need(V) + Carver::need(V * 2) + Carver::need(nn * 28) + Carver::need(8) + Carver::need(nn * 24) +
                   2 * Carver::need(prm_bytes + 16) + Carver::need((nn + 1) * 8) + Carver::need(nn * 8) + Carver::need(4) +
                   3 * Carver::need(nn * 4) + Carver::need(nn) + ws_bytes;
    int e = hc.ensure(total);
    if (e) return e;
    Carver cv(hc.buf);
    uint8_t* d_vol = cv.take<uint8_t>(V);
    uint16_t* d_seg = cv.take<uint16_t>(V);
    float* d_dets = cv.take<float>(nn * 7);
    int32_t* d_off = cv.take<int32_t>(2);
    int32_t* d_boxes = cv.take<int32_t>(nn * 6);
    uint8_t* d_prm = cv.take<uint8_t>(prm_bytes + 16);
    uint8_t* d_mask = cv.take<uint8_t>(prm_bytes + 16);
    int64_t* d_coff = cv.take<int64_t>(nn + 1);
    int64_t* d_keep = cv.take<int64_t>(nn);
    int32_t* d_cnt = cv.take<int32_t>(1);
    int32_t* d_rank = cv.take<int32_t>(nn);
    int32_t* d_bmax = cv.take<int32_t>(nn);
    int32_t* d_stat = cv.take<int32_t>(nn);
    uint8_t* d_surv = cv.take<uint8_t>(nn);
    void* d_ws = cv.p;
    cudaStream_t st = hc.stream;
    const int32_t off[2] = {0, n};
    B200_CUDA(cudaMemcpyAsync(d_vol, volume, V, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(d_off, off, 8, cudaMemcpyHostToDevice, st));
    if (n > 0) {
        B200_CUDA(cudaMemcpyAsync(d_dets, dets, (size_t)n * 28, cudaMemcpyHostToDevice, st));
        B200_CUDA(cudaMemcpyAsync(d_boxes, boxes, (size_t)n * 24, cudaMemcpyHostToDevice, st));
        B200_CUDA(cudaMemcpyAsync(d_prm, prm, prm_bytes, cudaMemcpyHostToDevice, st));
        B200_CUDA(cudaMemcpyAsync(d_coff, crop_off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    }
    e = b200seg_postproc_soma_dev(d_vol, 1, S, H, W, d_dets, d_off, off, d_boxes, d_prm, d_coff, (long long)prm_bytes, nms_thresh,
                                  keep_largest_cc, d_seg, d_keep,
                                  d_cnt, d_rank, d_mask, d_bmax, d_stat, d_surv, d_ws, ws_bytes, st);
    if (e) return e;
    int32_t cnt = 0;
    B200_CUDA(cudaMemcpyAsync(seg, d_seg, V * 2, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost, st));
    if (n > 0) {
        B200_CUDA(cudaMemcpyAsync(rank_order, d_rank, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaMemcpyAsync(b_max, d_bmax, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaMemcpyAsync(status, d_stat, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaMemcpyAsync(survive, d_surv, (size_t)n, cudaMemcpyDeviceToHost, st));
    }
    B200_CUDA(cudaStreamSynchronize(st));
    *n_keep = cnt;
    return 0;
}

