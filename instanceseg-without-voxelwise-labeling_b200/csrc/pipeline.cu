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
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    uint16_t* ids = (uint16_t*)ws;
    char* paste_ws = ws + align_up((size_t)(n_max > 0 ? n_max : 1) * 2, 256);
    const size_t paste_ws_bytes = align_up(b200seg_paste_labels_workspace_bytes(n_volumes, S, H, W, n_max), 256);
    char* cc_ws = paste_ws + paste_ws_bytes;
    const size_t cc_ws_bytes = cc_bytes > 0 ? align_up(b200seg_largest_cc_workspace_bytes(cc_bytes), 256) : 0;
    char* nms_ws = cc_ws + cc_ws_bytes;
    const size_t nms_ws_bytes = workspace_bytes - (size_t)(nms_ws - (char*)workspace);

    int e = b200seg_nms3d_dev(dets, det_off_dev, n_volumes, n_max, nms_thresh, 0, keep, keep_count, rank_order,
                              nms_ws, nms_ws_bytes, stream);
    if (e) return e;
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
    return b200seg_paste_labels_dev(seg, n_volumes, S, H, W, det_off_dev, n_max, boxes, ids, masks, crop_off,
                                    rank_order, keep_count, survive, paste_ws, paste_ws_bytes, stream);
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

// HOST-buffer entry point for a BATCH of equally shaped volumes: the same chain as above, software-pipelined
// over three streams (H2D | kernels | D2H) and three device slots, so that the upload of volume v+1, the
// kernels of volume v and the download of volume v-1 overlap (PCIe is full duplex; the chain itself takes a
// few percent of a transfer).  Pass pinned host buffers for real overlap; pageable memory still works
// (the copies then serialise inside the driver).  Returns when every output is in host memory.
namespace b200seg {
struct BatchStreams {
    cudaStream_t in = nullptr, out = nullptr;
    cudaEvent_t in_done[3] = {nullptr, nullptr, nullptr}, comp_done[3] = {nullptr, nullptr, nullptr},
                out_done[3] = {nullptr, nullptr, nullptr};
    int device = -1;
    char* pinned = nullptr;           // grow-only pinned staging for the small per-volume outputs
    size_t pinned_cap = 0;
    int ensure_pinned(size_t bytes) {
        if (bytes <= pinned_cap) return 0;
        if (pinned) { cudaFreeHost(pinned); pinned = nullptr; pinned_cap = 0; }
        const size_t want = align_up(bytes + (bytes >> 1), 1 << 16);
        B200_CUDA(cudaHostAlloc((void**)&pinned, want, cudaHostAllocDefault));
        pinned_cap = want;
        return 0;
    }
    int ensure(int dev) {
        if (device == dev && in) return 0;
        if (in) {
            cudaStreamDestroy(in); cudaStreamDestroy(out);
            for (int k = 0; k < 3; ++k) { cudaEventDestroy(in_done[k]); cudaEventDestroy(comp_done[k]); cudaEventDestroy(out_done[k]); }
        }
        B200_CUDA(cudaStreamCreateWithFlags(&in, cudaStreamNonBlocking));
        B200_CUDA(cudaStreamCreateWithFlags(&out, cudaStreamNonBlocking));
        for (int k = 0; k < 3; ++k) {
            B200_CUDA(cudaEventCreateWithFlags(&in_done[k], cudaEventDisableTiming));
            B200_CUDA(cudaEventCreateWithFlags(&comp_done[k], cudaEventDisableTiming));
            B200_CUDA(cudaEventCreateWithFlags(&out_done[k], cudaEventDisableTiming));
        }
        device = dev;
        return 0;
    }
};
static BatchStreams g_batch_dev[64];      // one set of streams / events / pinned staging per device ordinal (guarded by the host context's mutex)
}  // namespace b200seg

extern "C" int b200seg_postproc_soma_host_batch(int n_volumes, int S, int H, int W,
                                                const uint8_t* const* volumes, const float* const* dets, const int32_t* n_dets,
                                                const int32_t* const* boxes, const uint8_t* const* prm,
                                                const int64_t* const* crop_off, float nms_thresh, int keep_largest_cc,
                                                uint16_t* const* seg, int32_t* n_keep, int32_t* const* rank_order,
                                                int32_t* const* b_max, int32_t* const* status, uint8_t* const* survive) {
    B200_CHECK_ARG(n_volumes >= 0 && S > 0 && H > 0 && W > 0, "postproc_soma_host_batch: bad sizes");
    if (n_volumes == 0) return 0;
    B200_CHECK_ARG(volumes && n_dets && seg && n_keep, "postproc_soma_host_batch: null pointer");
    int n_max = 0;
    size_t prm_max = 0;
    for (int v = 0; v < n_volumes; ++v) {
        const int n = n_dets[v];
        B200_CHECK_ARG(n >= 0 && volumes[v] && seg[v], "postproc_soma_host_batch: bad volume %d", v);
        B200_CHECK_ARG(n == 0 || (dets && boxes && prm && crop_off && rank_order && b_max && status && survive &&
                                  dets[v] && boxes[v] && prm[v] && crop_off[v] && rank_order[v] && b_max[v] && status[v] && survive[v]),
                       "postproc_soma_host_batch: null pointer for volume %d", v);
        if (n > n_max) n_max = n;
        if (n > 0 && (size_t)crop_off[v][n] > prm_max) prm_max = (size_t)crop_off[v][n];
    }
    B200_CHECK_ARG(n_max < 65535, "postproc_soma_host_batch: more than 65534 instances per volume do not fit uint16 labels");
    HostCtx& hc = host_ctx();
    std::lock_guard<std::mutex> lock(hc.mu);
    const size_t V = (size_t)S * H * W;
    const size_t nn = n_max > 0 ? n_max : 1;
    const size_t ws_bytes = b200seg_postproc_soma_workspace_bytes(1, n_max, S, H, W, keep_largest_cc ? (long long)prm_max : 0);
    const int NB = n_volumes < 3 ? n_volumes : 3;
    const size_t slot_bytes = Carver::need(V) + Carver::need(V * 2) + Carver::need(nn * 28) + Carver::need(8) + Carver::need(nn * 24) +
                              2 * Carver::need(prm_max + 16) + Carver::need((nn + 1) * 8) + Carver::need(nn * 8) + Carver::need(4) +
                              3 * Carver::need(nn * 4) + Carver::need(nn) + Carver::need(ws_bytes);
    int e = hc.ensure(slot_bytes * NB);
    if (e) return e;
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_CHECK_ARG(dev >= 0 && dev < 64, "postproc_soma_host_batch: device ordinal out of range");
    BatchStreams& g_batch = g_batch_dev[dev];
    e = g_batch.ensure(dev);
    if (e) return e;
    struct Slot {
        uint8_t* vol; uint16_t* seg; float* dets; int32_t* off; int32_t* boxes; uint8_t* prm; uint8_t* mask; int64_t* coff;
        int64_t* keep; int32_t* cnt; int32_t* rank; int32_t* bmax; int32_t* stat; uint8_t* surv; void* ws;
    } slot[3];
    for (int k = 0; k < NB; ++k) {
        Carver cv(hc.buf + slot_bytes * k);
        Slot& s = slot[k];
        s.vol = cv.take<uint8_t>(V); s.seg = cv.take<uint16_t>(V); s.dets = cv.take<float>(nn * 7); s.off = cv.take<int32_t>(2);
        s.boxes = cv.take<int32_t>(nn * 6); s.prm = cv.take<uint8_t>(prm_max + 16); s.mask = cv.take<uint8_t>(prm_max + 16);
        s.coff = cv.take<int64_t>(nn + 1); s.keep = cv.take<int64_t>(nn); s.cnt = cv.take<int32_t>(1); s.rank = cv.take<int32_t>(nn);
        s.bmax = cv.take<int32_t>(nn); s.stat = cv.take<int32_t>(nn); s.surv = cv.take<uint8_t>(nn); s.ws = cv.p;
    }
    // small outputs go through pinned staging: a device-to-host copy into pageable memory would block the host
    // thread until the volume is finished and serialise the whole pipeline
    size_t small_bytes = 0;
    for (int v = 0; v < n_volumes; ++v) small_bytes += align_up(16 + 13 * (size_t)n_dets[v], 16);
    e = g_batch.ensure_pinned(small_bytes);
    if (e) return e;
    cudaStream_t s_in = g_batch.in, s_comp = hc.stream, s_out = g_batch.out;
    size_t small_off = 0;
    // per-volume {0, n} offset pairs must stay alive until their asynchronous copies have run
    int32_t* offs = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)n_volumes);
    if (!offs) { set_error("postproc_soma_host_batch: out of host memory"); return B200SEG_EINVAL; }
    for (int v = 0; v < n_volumes; ++v) { offs[2 * v] = 0; offs[2 * v + 1] = n_dets[v]; }
    int rc = 0;
#define B200_BATCH(call) do { int _e = ::b200seg::check_cuda((call), #call); if (_e) { rc = _e; goto done; } } while (0)
    for (int v = 0; v < n_volumes; ++v) {
        const int k = v % NB;
        Slot& s = slot[k];
        const int n = n_dets[v];
        if (v >= NB) B200_BATCH(cudaStreamWaitEvent(s_in, g_batch.out_done[k], 0));     // slot free again
        B200_BATCH(cudaMemcpyAsync(s.vol, volumes[v], V, cudaMemcpyHostToDevice, s_in));
        B200_BATCH(cudaMemcpyAsync(s.off, offs + 2 * v, 8, cudaMemcpyHostToDevice, s_in));
        if (n > 0) {
            B200_BATCH(cudaMemcpyAsync(s.dets, dets[v], (size_t)n * 28, cudaMemcpyHostToDevice, s_in));
            B200_BATCH(cudaMemcpyAsync(s.boxes, boxes[v], (size_t)n * 24, cudaMemcpyHostToDevice, s_in));
            B200_BATCH(cudaMemcpyAsync(s.prm, prm[v], (size_t)crop_off[v][n], cudaMemcpyHostToDevice, s_in));
            B200_BATCH(cudaMemcpyAsync(s.coff, crop_off[v], (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, s_in));
        }
        B200_BATCH(cudaEventRecord(g_batch.in_done[k], s_in));
        B200_BATCH(cudaStreamWaitEvent(s_comp, g_batch.in_done[k], 0));
        {
            const int32_t off[2] = {0, n};
            const int ce = b200seg_postproc_soma_dev(s.vol, 1, S, H, W, s.dets, s.off, off, s.boxes, s.prm, s.coff,
                                                     n > 0 ? (long long)crop_off[v][n] : 0, nms_thresh, keep_largest_cc, s.seg,
                                                     s.keep, s.cnt, s.rank, s.mask, s.bmax, s.stat, s.surv, s.ws, ws_bytes, s_comp);
            if (ce) { rc = ce; goto done; }
        }
        B200_BATCH(cudaEventRecord(g_batch.comp_done[k], s_comp));
        B200_BATCH(cudaStreamWaitEvent(s_out, g_batch.comp_done[k], 0));
        B200_BATCH(cudaMemcpyAsync(seg[v], s.seg, V * 2, cudaMemcpyDeviceToHost, s_out));
        {
            char* st = g_batch.pinned + small_off;          // [cnt | pad to 16 | rank n*4 | b_max n*4 | status n*4 | survive n]
            B200_BATCH(cudaMemcpyAsync(st, s.cnt, 4, cudaMemcpyDeviceToHost, s_out));
            if (n > 0) {
                B200_BATCH(cudaMemcpyAsync(st + 16, s.rank, (size_t)n * 4, cudaMemcpyDeviceToHost, s_out));
                B200_BATCH(cudaMemcpyAsync(st + 16 + (size_t)n * 4, s.bmax, (size_t)n * 4, cudaMemcpyDeviceToHost, s_out));
                B200_BATCH(cudaMemcpyAsync(st + 16 + (size_t)n * 8, s.stat, (size_t)n * 4, cudaMemcpyDeviceToHost, s_out));
                B200_BATCH(cudaMemcpyAsync(st + 16 + (size_t)n * 12, s.surv, (size_t)n, cudaMemcpyDeviceToHost, s_out));
            }
            small_off += align_up(16 + 13 * (size_t)n, 16);
        }
        B200_BATCH(cudaEventRecord(g_batch.out_done[k], s_out));
    }
done:
#undef B200_BATCH
    {
        // drain all three streams even on error: the slots and `offs` must not be reused while copies are in flight
        const cudaError_t e1 = cudaStreamSynchronize(s_in), e2 = cudaStreamSynchronize(s_comp), e3 = cudaStreamSynchronize(s_out);
        free(offs);
        if (rc == 0) {
            if (e1 != cudaSuccess) rc = check_cuda(e1, "cudaStreamSynchronize(in)");
            else if (e2 != cudaSuccess) rc = check_cuda(e2, "cudaStreamSynchronize(compute)");
            else if (e3 != cudaSuccess) rc = check_cuda(e3, "cudaStreamSynchronize(out)");
        }
    }
    if (rc == 0) {                                          // hand the staged small outputs to the caller
        size_t o = 0;
        for (int v = 0; v < n_volumes; ++v) {
            const size_t n = (size_t)n_dets[v];
            const char* st = g_batch.pinned + o;
            memcpy(&n_keep[v], st, 4);
            if (n > 0) {
                memcpy(rank_order[v], st + 16, n * 4);
                memcpy(b_max[v], st + 16 + n * 4, n * 4);
                memcpy(status[v], st + 16 + n * 8, n * 4);
                memcpy(survive[v], st + 16 + n * 12, n);
            }
            o += align_up(16 + 13 * n, 16);
        }
    }
    return rc;
}
