// api.cu -- error reporting, launch accounting and the HOST-buffer entry points of the C ABI.
//
// The "_host" functions are what the reference's numpy call sites bind to (boxes_3d.nms_3d,
// bbox_overlaps_3d, otsu_py_2d_fast): they copy host -> device, run the same kernels as the
// "_dev" entry points on an internal stream, copy the result back and synchronise.  Device
// scratch is a grow-only arena (one per process, mutex protected).  There is no CPU fallback:
// without a usable GPU every entry point fails with the CUDA error.
#include "common.cuh"

#include <atomic>
#include <mutex>
#include <stdarg.h>

namespace b200seg {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int HostCtx::ensure(size_t bytes) {
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    if (dev != device) {           // the arena belongs to one device; rebuild on switch
        if (buf) cudaFree(buf);
        if (stream) cudaStreamDestroy(stream);
        buf = nullptr; cap = 0; stream = nullptr; device = dev;
    }
    if (!stream) B200_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (bytes > cap) {
        if (buf) { B200_CUDA(cudaStreamSynchronize(stream)); B200_CUDA(cudaFree(buf)); buf = nullptr; cap = 0; }
        size_t want = align_up(bytes + (bytes >> 2), 1 << 20);
        B200_CUDA(cudaMalloc((void**)&buf, want));
        cap = want;
    }
    return 0;
}
static HostCtx g_host;
HostCtx& host_ctx() { return g_host; }

// profiling knobs (b200seg_set_option)
static std::atomic<int> g_peaks_stop_after{99};
int opt_peaks_stop_after() { return g_peaks_stop_after.load(std::memory_order_relaxed); }
void opt_set_peaks_stop_after(int v) { g_peaks_stop_after.store(v, std::memory_order_relaxed); }

static std::atomic<int> g_peaks_median_mode{0};
int opt_peaks_median_mode() { return g_peaks_median_mode.load(std::memory_order_relaxed); }
static std::atomic<int> g_host_batch_out{0};
int opt_host_batch_out() { return g_host_batch_out.load(std::memory_order_relaxed); }
static std::atomic<int> g_host_batch_mode{7};
int opt_host_batch_mode() { return g_host_batch_mode.load(std::memory_order_relaxed); }

}  // namespace b200seg

using namespace b200seg;

extern "C" const char* b200seg_last_error(void) { return g_err; }
extern "C" int b200seg_version(void) { return B200SEG_VERSION; }
extern "C" long long b200seg_launch_count(void) { return g_launches.load(); }

extern "C" int b200seg_set_option(const char* name, int value) {
    B200_CHECK_ARG(name, "set_option: null name");
    if (!strcmp(name, "peaks_stop_after")) { opt_set_peaks_stop_after(value); return 0; }
    // bit 0: compacted label download (else dense copies), bit 1: zero-copy gather of the surviving PRM crops (else whole array)
    if (!strcmp(name, "peaks_median_mode")) { g_peaks_median_mode.store(value, std::memory_order_relaxed); return 0; }
    if (!strcmp(name, "host_batch_out")) {
        if (value < 0 || value > 2) { set_error("set_option: host_batch_out must be 0, 1 or 2"); return B200SEG_EINVAL; }
        g_host_batch_out.store(value, std::memory_order_relaxed);
        return 0;
    }
    if (!strcmp(name, "host_batch_mode")) { g_host_batch_mode.store(value, std::memory_order_relaxed); return 0; }
    set_error("set_option: unknown option '%s'", name);
    return B200SEG_EINVAL;
}

extern "C" int b200seg_nms3d_host(const float* dets, int n, float thresh, int by_volume, int64_t* keep, int* n_keep) {
    B200_CHECK_ARG(n >= 0 && n_keep, "nms3d_host: bad arguments");
    *n_keep = 0;
    if (n == 0) return 0;
    B200_CHECK_ARG(dets && keep, "nms3d_host: null pointer");
    HostCtx& hc = host_ctx(); std::lock_guard<std::mutex> lock(hc.mu);
    const size_t ws_bytes = b200seg_nms3d_workspace_bytes(1, n);
    const size_t total = Carver::need((size_t)n * 7 * 4) + Carver::need(8) + Carver::need((size_t)n * 8) + Carver::need(4) + ws_bytes;
    int e = hc.ensure(total);
    if (e) return e;
    Carver cv(hc.buf);
    float* d_dets = cv.take<float>((size_t)n * 7);
    int32_t* d_off = cv.take<int32_t>(2);
    int64_t* d_keep = cv.take<int64_t>(n);
    int32_t* d_cnt = cv.take<int32_t>(1);
    void* d_ws = cv.p;
    cudaStream_t st = hc.stream;
    const int32_t off[2] = {0, n};
    B200_CUDA(cudaMemcpyAsync(d_dets, dets, (size_t)n * 7 * 4, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(d_off, off, 8, cudaMemcpyHostToDevice, st));
    e = b200seg_nms3d_dev(d_dets, d_off, 1, n, thresh, by_volume, d_keep, d_cnt, nullptr, d_ws, ws_bytes, st);
    if (e) return e;
    int32_t cnt = 0;
    B200_CUDA(cudaMemcpyAsync(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(keep, d_keep, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    *n_keep = cnt;
    return 0;
}

extern "C" int b200seg_iou3d_host(const float* boxes, long long N, const float* query, long long K, float* overlaps) {
    B200_CHECK_ARG(N >= 0 && K >= 0, "iou3d_host: negative size");
    if (N == 0 || K == 0) return 0;
    B200_CHECK_ARG(boxes && query && overlaps, "iou3d_host: null pointer");
    HostCtx& hc = host_ctx(); std::lock_guard<std::mutex> lock(hc.mu);
    const size_t total = Carver::need((size_t)N * 24) + Carver::need((size_t)K * 24) + Carver::need((size_t)N * K * 4);
    int e = hc.ensure(total);
    if (e) return e;
    Carver cv(hc.buf);
    float* d_b = cv.take<float>((size_t)N * 6);
    float* d_q = cv.take<float>((size_t)K * 6);
    float* d_o = cv.take<float>((size_t)N * K);
    cudaStream_t st = hc.stream;
    B200_CUDA(cudaMemcpyAsync(d_b, boxes, (size_t)N * 24, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(d_q, query, (size_t)K * 24, cudaMemcpyHostToDevice, st));
    e = b200seg_iou3d_dev(d_b, N, d_q, K, d_o, st);
    if (e) return e;
    B200_CUDA(cudaMemcpyAsync(overlaps, d_o, (size_t)N * K * 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int b200seg_otsu2d_host(const uint16_t* image, const uint16_t* prm, long long n, uint8_t* mask, int* b_max) {
    B200_CHECK_ARG(n > 0 && image && prm && mask && b_max, "otsu2d_host: bad arguments");
    HostCtx& hc = host_ctx(); std::lock_guard<std::mutex> lock(hc.mu);
    const size_t total = 2 * Carver::need((size_t)n * 2) + Carver::need(16) + Carver::need((size_t)n) + 3 * Carver::need(16);
    int e = hc.ensure(total);
    if (e) return e;
    Carver cv(hc.buf);
    uint16_t* d_i = cv.take<uint16_t>(n);
    uint16_t* d_p = cv.take<uint16_t>(n);
    int64_t* d_off = cv.take<int64_t>(2);
    uint8_t* d_m = cv.take<uint8_t>(n);
    int32_t* d_b = cv.take<int32_t>(1);
    int32_t* d_g = cv.take<int32_t>(4);
    int32_t* d_s = cv.take<int32_t>(1);
    cudaStream_t st = hc.stream;
    const int64_t off[2] = {0, (int64_t)n};
    B200_CUDA(cudaMemcpyAsync(d_i, image, (size_t)n * 2, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(d_p, prm, (size_t)n * 2, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(d_off, off, 16, cudaMemcpyHostToDevice, st));
    e = b200seg_otsu2d_dev(d_i, d_p, d_off, 1, d_m, d_b, d_g, d_s, nullptr, nullptr, st);
    if (e) return e;
    int32_t hb = 0, hs = 0;
    B200_CUDA(cudaMemcpyAsync(&hb, d_b, 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(&hs, d_s, 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(mask, d_m, (size_t)n, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    *b_max = hb;
    if (hs == 4) { set_error("otsu2d_host: gray range exceeds %d levels", 2048); return B200SEG_EUNSUPPORTED; }
    return hs == 1 ? 1 : 0;
}
