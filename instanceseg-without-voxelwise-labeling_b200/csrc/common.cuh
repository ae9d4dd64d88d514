// common.cuh -- shared helpers for the b200seg CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <atomic>

#include "../../include/b200seg.h"

#define B200SEG_VERSION 100

namespace b200seg {

// thread-local error text + global launch counter (api.cu)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_cuda(cudaError_t e, const char* what);
int num_sms();
int opt_host_batch_mode();            // b200seg_set_option("host_batch_mode"): bit 0 compacted label download, bit 1 PRM gather, bit 2 packed image crops, bit 3 packed PRM crops, bits 4 / 5 groups of 2 / 4 volumes per chain launch
int opt_host_batch_out();             // what the caller's label buffers hold on entry: 0 unknown, 1 zeros, 2 this entry point's previous result
int opt_peaks_median_mode();          // 0 auto, 1 histogram path only, 2 sampled interval forced to miss (tests the fallback)
int opt_peaks_stop_after();          // profiling knob: peaks3d stops after its first k kernels (99 = run all)

#define B200_CHECK_ARG(cond, ...)                         \
    do {                                                  \
        if (!(cond)) {                                    \
            ::b200seg::set_error(__VA_ARGS__);            \
            return B200SEG_EINVAL;                        \
        }                                                 \
    } while (0)

#define B200_CUDA(call)                                              \
    do {                                                             \
        int _e = ::b200seg::check_cuda((call), #call);               \
        if (_e) return _e;                                           \
    } while (0)

#define B200_LAUNCH_CHECK(name)                                      \
    do {                                                             \
        ::b200seg::count_launch();                                   \
        int _e = ::b200seg::check_cuda(cudaGetLastError(), name);    \
        if (_e) return _e;                                           \
    } while (0)

// "do this once per device" flag for per-function attributes (cudaFuncSetAttribute is per device): one bit per device ordinal,
// lock-free; a lost race only repeats an idempotent call.
struct OncePerDevice {
    std::atomic<unsigned long long> done{0ull};
    bool needed(int* dev_out) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { *dev_out = -1; return true; }
        *dev_out = dev;
        return !(done.load(std::memory_order_acquire) & (1ull << dev));
    }
    void mark(int dev) { if (dev >= 0) done.fetch_or(1ull << dev, std::memory_order_release); }
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// grow-only device arena + private stream used by the HOST-buffer entry points (api.cu owns it)
struct HostCtx {
    std::mutex mu;
    cudaStream_t stream = nullptr;
    char* buf = nullptr;
    size_t cap = 0;
    int device = -1;
    int ensure(size_t bytes);
};
HostCtx& host_ctx();

struct Carver {
    char* p;
    explicit Carver(char* base) : p(base) {}
    template <typename T> T* take(size_t n) { T* r = (T*)p; p += align_up(n * sizeof(T), 256); return r; }
    static size_t need(size_t bytes) { return align_up(bytes, 256); }
};

// carving of the workspace of b200seg_postproc_soma_dev (pipeline.cu) and the part of the chain that follows the NMS
struct SomaChainWs {
    uint16_t* ids;
    char* paste_ws; size_t paste_ws_bytes;
    char* cc_ws; size_t cc_ws_bytes;
    char* nms_ws; size_t nms_ws_bytes;
};
// optional second output of the paste kernel: the non-zero 64-byte lines of the (single) label volume (paste.cu)
struct PasteLines { uint32_t* idx; uint4* val; uint32_t* count; uint32_t cap; };
bool paste_lines_supported(const uint16_t* seg, int n_volumes, int S, int H, int W);
int paste_labels_launch(uint16_t* seg, int n_volumes, int S, int H, int W, const int32_t* det_off, int n_max, const int32_t* boxes,
                        const uint16_t* ids, const uint8_t* masks, const int64_t* mask_off, const int32_t* order,
                        const int32_t* n_valid, uint8_t* survive, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                        const PasteLines* lines);
SomaChainWs soma_chain_ws(void* workspace, size_t workspace_bytes, int n_volumes, int n_max, int S, int H, int W, long long cc_bytes);
int postproc_soma_after_nms(const uint8_t* volumes, int n_volumes, int S, int H, int W, const int32_t* det_off_dev, int n_max,
                            int total, const int32_t* boxes, const uint8_t* prm, const int64_t* crop_off, long long prm_bytes,
                            int keep_largest_cc, uint16_t* seg, const int32_t* keep_count, const int32_t* rank_order,
                            uint8_t* masks, int32_t* b_max, int32_t* status, uint8_t* survive, uint16_t* ids,
                            void* paste_ws, size_t paste_ws_bytes, void* cc_ws, size_t cc_ws_bytes, cudaStream_t stream,
                            const PasteLines* lines = nullptr);

// ---- device helpers -------------------------------------------------------------------------

// Monotone map float -> uint32 such that a < b  <=>  key(a) < key(b); -0.0 and +0.0 map to the
// same key; every NaN maps to the largest key.
__device__ __forceinline__ uint32_t ordered_key(float f) {
    if (f != f) return 0xFFFFFFFFu;
    if (f == 0.0f) f = 0.0f;                       // canonicalise -0
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(u);
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// streaming (evict-first) 128-bit global accesses for data touched exactly once
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 8 bytes starting at p (any alignment) as a little-endian 64-bit value, fetched with one or two ALIGNED
// 64-bit loads.  Only the first `nbytes` (1..8) bytes are meaningful: the second word is read only when
// one of them lives there.  `safe_end` = end of the underlying buffer rounded DOWN to 8 bytes; windows
// that would touch an aligned word crossing it fall back to byte loads (never reads out of bounds).
__device__ __forceinline__ unsigned long long load8_unaligned(const uint8_t* p, int nbytes, const uint8_t* safe_end) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const unsigned mis = (unsigned)(a & 7);
    const unsigned long long* q = reinterpret_cast<const unsigned long long*>(a - mis);
    const bool need_hi = mis + (unsigned)nbytes > 8u;
    if (reinterpret_cast<const uint8_t*>(q) + (need_hi ? 16 : 8) <= safe_end) {
        const unsigned long long lo = __ldg(q);
        if (!need_hi) return lo >> (mis * 8);
        const unsigned long long hi = __ldg(q + 1);
        return (lo >> (mis * 8)) | (hi << (64 - mis * 8));       // mis >= 1 here
    }
    unsigned long long v = 0ull;
    for (int k = 0; k < nbytes; ++k) v |= (unsigned long long)__ldg(p + k) << (8 * k);
    return v;
}

}  // namespace b200seg
