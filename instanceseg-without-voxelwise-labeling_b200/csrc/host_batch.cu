// host_batch.cu -- HOST-buffer entry point of the binarization chain for a BATCH of equally shaped volumes
// (tools/binarization_soma.py:57-104 per volume; the per-volume loop of tools/my_subprocess.py:56).
//
// The reference hands numpy arrays in and gets the uint16 label volume back, so the PCIe link, the host memory and the
// latency of a dozen small launches per volume bound this entry point, not the kernels.  Layout, per volume:
//   stage A  (five volumes ahead) the small per-detection arrays travel by DMA, the NMS runs on its own high-priority stream
//            and its visit order comes straight back (a few hundred bytes);
//   stage P  (three steps before the chain) a pool of host threads packs the image crops of the NMS survivors that have a
//            positive PRM voxel -- the only voxels of the raw volume the chain ever reads -- into a pinned buffer
//            ("host_batch_mode" bit 2; without it the whole 33.5 MB volume travels by DMA);
//   stage B  the packed crops travel by DMA and an unpack kernel restores their rows in the device volume; the PRM crops of
//            the same survivors are pulled straight out of the caller's pinned buffer by a gather kernel (zero copy; all-zero
//            crops are zero-filled on the device instead; pageable or unaligned buffers fall back to a plain copy); then
//            binarize -> largest component -> paste, and the label volume is compacted into its non-zero 64-byte lines
//            (32 voxels): line index + payload, a few percent of the volume.  Volumes rotate over three compute streams;
//   down     only the compacted lines and the per-detection bookkeeping travel; the pool of host threads prepares the
//            caller's label volumes (zero fill, or nothing / an undo list: "host_batch_out") and writes the lines with
//            full-line non-temporal stores -- or, for pinned label volumes of ranks that share a host, a small kernel on its
//            own stream writes the lines in place over the link and only their indices are downloaded.  A volume whose labels
//            cover more than a quarter of the lines is copied densely.  "host_batch_mode" bits 3-8 select measured
//            alternatives (PRM crops packed by the host, chain launches for groups of 2 / 4 volumes, ...): include/b200seg.h.
// Twelve device slots; the host thread sizes the download of a volume (it needs the line count) two volumes after it has
// enqueued its chain and hands finished downloads to the pool one volume later.
#include "common.cuh"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <sched.h>
#include <stdlib.h>
#include <thread>
#include <unistd.h>
#include <unordered_map>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

namespace b200seg {

// ---- device side -------------------------------------------------------------------------------------------------
// non-zero 64-byte lines of the label volume (4 consecutive 16-byte groups = 32 labels), unordered: idx_out[k] = line index,
// val_out[4k .. 4k+3] = its groups (groups past the end of the volume read as zero).  A line is the unit the host writes with
// full-line non-temporal stores: no read-for-ownership of the destination, which is what bounded the 16-byte scatter of
// round 2.  Loads stay coalesced (lane = group); the four lanes of a quad own one line (base and the 256-group stride are
// multiples of 4).  `count` keeps counting past `cap` (the host then takes the dense path).
__global__ void __launch_bounds__(256)
seg_compact_kernel(const uint4* __restrict__ seg, unsigned int ngroups, unsigned int cap,
                   uint32_t* __restrict__ idx_out, uint4* __restrict__ val_out, uint32_t* __restrict__ count) {
    __shared__ unsigned int s_tot[8], s_base;
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int per_cta = 256u * 4u;
    for (unsigned int base = blockIdx.x * per_cta; base < ngroups; base += gridDim.x * per_cta) {      // CTA-uniform trip count
        uint4 v[4];
        unsigned int gi[4], lm[4], tot = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            gi[j] = base + j * 256u + threadIdx.x;
            v[j] = gi[j] < ngroups ? ld_stream_u4(seg + gi[j]) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool nz = (v[j].x | v[j].y | v[j].z | v[j].w) != 0u;
            const unsigned int m = __ballot_sync(0xffffffffu, nz);
            lm[j] = (m | (m >> 1) | (m >> 2) | (m >> 3)) & 0x11111111u;     // bit 4q: quad q holds a non-zero line
            tot += __popc(lm[j]);
        }
        // one global atomic per CTA and trip (a volume holds ~10^5 lines: one atomic per warp and row of groups made the
        // single counter the bound of the kernel); most CTAs see only zeros and skip it
        if (lane == 0) s_tot[warp] = tot;
        __syncthreads();
        unsigned int before = 0u, all = 0u;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const unsigned int t = s_tot[w]; all += t; before += w < (int)warp ? t : 0u; }
        if (all != 0u) {                                            // CTA-uniform
            if (threadIdx.x == 0) s_base = atomicAdd(count, all);
            __syncthreads();
            unsigned int b = s_base + before;
            const unsigned int q0 = lane & ~3u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (lm[j] == 0u) continue;                          // warp-uniform
                const unsigned int pos = b + __popc(lm[j] & ((1u << q0) - 1u));
                b += __popc(lm[j]);
                if (((lm[j] >> q0) & 1u) && pos < cap) {
                    if ((lane & 3u) == 0u) idx_out[pos] = gi[j] >> 2;
                    val_out[(size_t)pos * 4 + (lane & 3u)] = v[j];
                }
            }
        }
        __syncthreads();                                            // s_tot / s_base are reused by the next trip
    }
}

// Compacted lines -> their place in the caller's label volume (host-mapped pinned memory, 16-byte aligned): posted 64-byte
// writes over the link instead of a staged download that host threads scatter.  The host side of the staged scheme costs
// three passes over the line data in host DRAM (DMA write, read, non-temporal write) and a third of the pool's time; here it
// is one write.  The kernel is bound by the link, not by the SMs, so it is launched with a handful of CTAs on a stream of
// its own: it runs beside the chains of the following volumes instead of filling the machine with stalled CTAs (a
// full-grid version that wrote the lines straight from the volume walk took 131 us per volume on the compute stream).
// The line count stays on the device.  The caller's volume must already be zero (cleared / zero-filled by the pool).
constexpr int HB_SCATTER_CTAS = 32;
__global__ void __launch_bounds__(256)
lines_to_host_kernel(const uint32_t* __restrict__ idx, const uint4* __restrict__ val, const uint32_t* __restrict__ count,
                     unsigned int cap, unsigned int ngroups, uint4* __restrict__ host_dst) {
    const unsigned int n = min(*count, cap);
    const unsigned int q = threadIdx.x & 3u;
    for (unsigned int i = blockIdx.x * 64u + (threadIdx.x >> 2); i < n; i += gridDim.x * 64u) {
        const unsigned int g = idx[i] * 4u + q;
        if (g < ngroups) host_dst[g] = val[(size_t)i * 4 + q];      // the four stores of a quad are one 64-byte line
    }
}

// PRM crops of the NMS survivors: host-mapped (pinned) source -> device copy with the same packing.
// grid (n_max): CTA r moves the crop of visit rank r.  src and dst share the offset, both bases are 16-byte aligned.
__global__ void __launch_bounds__(256)
prm_gather_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int64_t* __restrict__ crop_off, long long shift,
                  const int32_t* __restrict__ rank_order, const int32_t* __restrict__ keep_count,
                  const uint8_t* __restrict__ all_zero, const int64_t* __restrict__ src_off) {
    const int r = blockIdx.x;
    if (r >= keep_count[0]) return;
    const int inst = rank_order[r];
    const long long a = crop_off[inst] - shift, b = crop_off[inst + 1] - shift;   // the device copy of the offsets counts from the group's base
    // src_off: the crops sit back to back in visit order (packed by the host, already on the device), crop r at src_off[r],
    // which is congruent to `a` modulo 16; else src mirrors dst
    if (src_off && !(all_zero && all_zero[r])) src += src_off[r] - a;
    long long a16 = (a + 15) & ~15ll, b16 = b & ~15ll;
    if (a16 > b16) { a16 = b; b16 = b; }                            // crop shorter than one aligned group
    if (all_zero && all_zero[r]) {                                  // the host saw no positive voxel in this crop: nothing to fetch
        for (long long i = a + threadIdx.x; i < a16; i += 256) dst[i] = 0;
        uint4* z4 = reinterpret_cast<uint4*>(dst + a16);
        for (long long i = threadIdx.x; i < ((b16 - a16) >> 4); i += 256) z4[i] = make_uint4(0u, 0u, 0u, 0u);
        for (long long k = b16 + threadIdx.x; k < b; k += 256) dst[k] = 0;
        return;
    }
    for (long long i = a + threadIdx.x; i < a16; i += 256) dst[i] = src[i];
    const uint4* s4 = reinterpret_cast<const uint4*>(src + a16);
    uint4* d4 = reinterpret_cast<uint4*>(dst + a16);
    const long long n4 = (b16 - a16) >> 4;
    long long i = threadIdx.x;
    for (; i + 3 * 256 < n4; i += 4 * 256) {                        // four 16-byte reads over the link in flight per thread
        const uint4 v0 = ld_stream_u4(s4 + i), v1 = ld_stream_u4(s4 + i + 256), v2 = ld_stream_u4(s4 + i + 512), v3 = ld_stream_u4(s4 + i + 768);
        d4[i] = v0; d4[i + 256] = v1; d4[i + 512] = v2; d4[i + 768] = v3;
    }
    for (; i < n4; i += 256) d4[i] = ld_stream_u4(s4 + i);
    for (long long k = b16 + threadIdx.x; k < b; k += 256) dst[k] = src[k];
}

// Image crops of the NMS survivors, packed by host threads (box-shaped, visit order, back to back) -> the rows of the device
// copy of the volume they came from.  The chain reads the image only inside the boxes of the instances it binarizes
// (binarization_soma.py:78-94), so with "host_batch_mode" bit 2 the 33.5 MB volume never crosses the link: the NMS runs
// first, its visit order comes back (a few hundred bytes), host threads copy the survivors' box rows into one pinned
// buffer (about 2.5 MB per volume) and that buffer travels by DMA.  Fetching the rows in place over the link instead (a
// gather kernel reading the pinned volume) was measured: 29-byte rows become 32-byte read requests and the link delivers
// 9 GB/s of them, less than the DMA of the whole volume.  The rest of the device volume keeps stale bytes nobody reads.
// grid (n_max, HB_UNPACK_PARTS): CTAs (r, 0..parts-1) move the box of visit rank r, a quarter of its rows each (the kernel lasts as
// long as its largest box: 41 us per volume with one CTA per box); one warp per row.
constexpr int HB_UNPACK_PARTS = 4;
__global__ void __launch_bounds__(256)
img_unpack_kernel(const uint8_t* __restrict__ pack, const int64_t* __restrict__ pk_off, const uint8_t* __restrict__ all_zero,
                  uint8_t* __restrict__ dst, int H, int W, const int32_t* __restrict__ boxes,
                  const int32_t* __restrict__ rank_order, const int32_t* __restrict__ keep_count) {
    const int r = blockIdx.x;
    if (r >= keep_count[0]) return;
    // an instance without a positive PRM voxel is skipped by the script before it touches the image (:74-76): the host did
    // not pack its crop (the bytes of its slot are stale) and it must not overwrite rows it shares with other boxes
    if (all_zero[r]) return;
    const int inst = rank_order[r];
    const int32_t* bx = boxes + (size_t)inst * 6;
    const int x1 = bx[0], y1 = bx[1], z1 = bx[2];
    const int sx = bx[3] - x1 + 1, sy = bx[4] - y1 + 1, sz = bx[5] - z1 + 1;
    if (sx <= 0 || sy <= 0 || sz <= 0) return;
    const uint8_t* src = pack + pk_off[r];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t HW = (size_t)H * W;
    const int all_rows = sz * sy;
    const int r_lo = (int)((long long)all_rows * blockIdx.y / gridDim.y), rows = (int)((long long)all_rows * (blockIdx.y + 1) / gridDim.y);
    for (int row0 = r_lo + warp; row0 < rows; row0 += 8 * 4) { // four rows in flight per warp
        uint8_t v[4];
        size_t off[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int row = min(row0 + 8 * u, rows - 1);
            const int z = row / sy, y = row - z * sy;
            off[u] = (size_t)(z1 + z) * HW + (size_t)(y1 + y) * W + x1;
            v[u] = lane < sx ? src[(size_t)row * sx + lane] : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) if (lane < sx && row0 + 8 * u < rows) dst[off[u] + lane] = v[u];
        if (sx > 32) {                                          // wide boxes: the rest of the rows
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int row = row0 + 8 * u;
                if (row < rows) for (int x = 32 + lane; x < sx; x += 32) dst[off[u] + x] = src[(size_t)row * sx + x];
            }
        }
    }
}

// Zero fill with non-temporal stores: the label volumes are far larger than the caches and are not read again by this
// thread, so the lines need not be fetched before they are overwritten (half the memory traffic of a plain memset
// whenever the C library's own streaming threshold is not reached).
static void zero_stream(char* dst, size_t bytes) {
#if defined(__x86_64__)
    static const bool plain = getenv("B200SEG_HB_PLAIN_MEMSET") != nullptr;
    if (!plain && bytes >= 4096) {
        const size_t head = (size_t)((64 - ((uintptr_t)dst & 63)) & 63);
        if (head) { memset(dst, 0, head); dst += head; bytes -= head; }
        const __m128i z = _mm_setzero_si128();
        size_t i = 0;
        for (; i + 64 <= bytes; i += 64) {
            _mm_stream_si128((__m128i*)(dst + i), z); _mm_stream_si128((__m128i*)(dst + i + 16), z);
            _mm_stream_si128((__m128i*)(dst + i + 32), z); _mm_stream_si128((__m128i*)(dst + i + 48), z);
        }
        _mm_sfence();
        if (i < bytes) memset(dst + i, 0, bytes - i);
        return;
    }
#endif
    memset(dst, 0, bytes);
}

// ---- host side: worker pool -------------------------------------------------------------------------------------
// One pool, two queues.  The crop packing jobs (short, on the critical path of the uploads) go to the urgent queue and
// are always taken first; zero fills, clears and scatters of the label volumes go to the normal queue.  Every thread
// serves both, so no core idles while the other kind of work is queued (round 2 had two pools of half the cores each).
struct HostPool {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<void()>> q, q_urgent;
    int n_threads = 0;
    int cap = 32;
    pid_t pid = 0;
    void ensure() {
        std::lock_guard<std::mutex> lk(mu);
        if (n_threads > 0 && pid == getpid()) return;              // a forked child starts its own threads
        pid = getpid();
        q.clear();
        q_urgent.clear();
        int want = 0;
        if (const char* e = getenv("B200SEG_HOST_THREADS")) want = atoi(e);
        if (want <= 0) {
            int hw = (int)std::thread::hardware_concurrency();
            if (hw <= 0) hw = 4;
            int share = 1;
            if (const char* e = getenv("LOCAL_WORLD_SIZE")) share = atoi(e) > 0 ? atoi(e) : 1;   // ranks of one box share its cores
            // One core stays with the thread that drives the GPU.  Beyond a handful of cores a quarter is left idle: with every
            // core streaming label lines the zero-copy PRM reads of the GPU slow down (15 threads of a 16-core box: 66 Gvox/s,
            // 8 threads: 73 Gvox/s through this entry point).
            const int per = hw / share;
            want = per <= 4 ? per - 1 : per * 3 / 4;
            if (want > cap) want = cap;
            if (want < 2) want = 2;
        }
        n_threads = want;
        for (int i = 0; i < want; ++i) std::thread([this] { run(); }).detach();
    }
    void run() {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [this] { return !q.empty() || !q_urgent.empty(); });
                std::deque<std::function<void()>>& src = q_urgent.empty() ? q : q_urgent;
                job = std::move(src.front());
                src.pop_front();
            }
            job();
        }
    }
    void push(std::function<void()> f, bool urgent = false) {
        { std::lock_guard<std::mutex> lk(mu); (urgent ? q_urgent : q).push_back(std::move(f)); }
        cv.notify_one();
    }
};
// never destroyed: its threads are detached and wait on the condition variable until the process ends (destroying a
// condition variable with waiters blocks in glibc)
static HostPool& g_pool = *new HostPool;

// ---- host side: the three loops the pool threads spend their time in ---------------------------------------------------
// The packing touches one or two cache lines at a time at addresses the hardware prefetchers cannot guess (box rows of a
// 33 MB volume), so it requests its lines a fixed distance ahead; the label loops write whole lines past the caches.
constexpr int HB_PF_ROWS = 12;         // image rows requested ahead of the copy

static inline void copy_row(uint8_t* d, const uint8_t* s, int n) {
#if defined(__x86_64__)
    if (n >= 16 && n <= 32) {                                       // the common box widths: two overlapping 16-byte moves
        const __m128i a = _mm_loadu_si128((const __m128i*)s), b = _mm_loadu_si128((const __m128i*)(s + n - 16));
        _mm_storeu_si128((__m128i*)d, a); _mm_storeu_si128((__m128i*)(d + n - 16), b);
        return;
    }
    if (n > 32 && n <= 64) {
        const __m128i a = _mm_loadu_si128((const __m128i*)s), b = _mm_loadu_si128((const __m128i*)(s + 16));
        const __m128i c = _mm_loadu_si128((const __m128i*)(s + n - 32)), e = _mm_loadu_si128((const __m128i*)(s + n - 16));
        _mm_storeu_si128((__m128i*)d, a); _mm_storeu_si128((__m128i*)(d + 16), b);
        _mm_storeu_si128((__m128i*)(d + n - 32), c); _mm_storeu_si128((__m128i*)(d + n - 16), e);
        return;
    }
    if (n >= 8 && n < 16) {
        uint64_t a, b;
        memcpy(&a, s, 8); memcpy(&b, s + n - 8, 8);
        memcpy(d, &a, 8); memcpy(d + n - 8, &b, 8);
        return;
    }
#endif
    memcpy(d, s, (size_t)n);
}

// box rows (sz x sy rows of sx bytes, origin `base`, row pitch W, plane pitch HW) -> d, back to back
static void pack_box_rows(uint8_t* d, const uint8_t* base, int sx, int sy, int sz, size_t W, size_t HW) {
    int pz = 0, py = 0;                                             // cursor of the prefetch, HB_PF_ROWS rows ahead of the copy
    auto request = [&] {
        if (pz >= sz) return;
        const uint8_t* r = base + (size_t)pz * HW + (size_t)py * W;
        __builtin_prefetch(r, 0, 3);
        __builtin_prefetch(r + sx - 1, 0, 3);
        if (++py == sy) { py = 0; ++pz; }
    };
    for (int i = 0; i < HB_PF_ROWS; ++i) request();
    for (int z = 0; z < sz; ++z) {
        const uint8_t* row = base + (size_t)z * HW;
        for (int y = 0; y < sy; ++y, row += W, d += sx) { request(); copy_row(d, row, sx); }
    }
}

// line `li` of a label volume of `vbytes` bytes: 64 bytes, except the last one of a volume that is not a multiple of 64 bytes
static inline void put_line(char* dst, size_t vbytes, uint32_t li, const char* src, bool aligned) {
    const size_t off = (size_t)li * 64;
    if (off >= vbytes) return;
    const size_t n = vbytes - off < 64 ? vbytes - off : 64;
#if defined(__x86_64__)
    if (aligned && n == 64) {
        const __m128i a = _mm_loadu_si128((const __m128i*)src), b = _mm_loadu_si128((const __m128i*)(src + 16));
        const __m128i c = _mm_loadu_si128((const __m128i*)(src + 32)), d = _mm_loadu_si128((const __m128i*)(src + 48));
        __m128i* o = (__m128i*)(dst + off);
        _mm_stream_si128(o, a); _mm_stream_si128(o + 1, b); _mm_stream_si128(o + 2, c); _mm_stream_si128(o + 3, d);
        return;
    }
#endif
    memcpy(dst + off, src, n);
}
static inline void fence_lines() {
#if defined(__x86_64__)
    _mm_sfence();
#endif
}
// The label volume is "zeros + the listed lines", so a listed line is cleared / written as a whole.
static void clear_lines(char* dst, size_t vbytes, const uint32_t* li, uint32_t nl) {
    alignas(16) static const char zeros[64] = {0};
    const bool aligned = (((uintptr_t)dst) & 63) == 0;
    for (uint32_t i = 0; i < nl; ++i) put_line(dst, vbytes, li[i], zeros, aligned);
    fence_lines();
}
static void scatter_lines(char* dst, size_t vbytes, const uint32_t* li, const char* lv, uint32_t nl) {
    const bool aligned = (((uintptr_t)dst) & 63) == 0;
    for (uint32_t i = 0; i < nl; ++i) put_line(dst, vbytes, li[i], lv + (size_t)i * 64, aligned);
    fence_lines();
}

constexpr int HB_SLOTS = 12;
constexpr int HB_LAG_A = 2;            // the download of volume v is sized and enqueued while volume v + HB_LAG_A is being enqueued
constexpr int HB_LAG_B = 3;            // ... and handed to the pool one step later
constexpr int HB_ZERO_PARTS = 4;
constexpr int HB_GROUP = 4;            // largest number of volumes per chain launch ("host_batch_mode" bits 4 / 5); HB_SLOTS is a multiple of it
constexpr int HB_LAG_P = 2;            // packed-image mode: the crops of volume v are handed to the pool while the NMS of volume v + HB_LAG_P is enqueued
constexpr int HB_LAG_N = 5;            // packed-image mode: the chain of volume v is enqueued while the NMS of volume v + HB_LAG_N is

struct BatchStreams {
    cudaStream_t in = nullptr, out = nullptr, out2 = nullptr;   // uploads | bookkeeping downloads | label downloads
    cudaStream_t out4 = nullptr;                                // line scatter kernels (link bound, a few CTAs)
    cudaEvent_t scat_done[HB_SLOTS] = {};
    cudaStream_t out3 = nullptr;                                // visit orders of the NMS (must not queue behind the bookkeeping of older volumes)
    cudaStream_t comp2 = nullptr, comp3 = nullptr;              // more compute streams: consecutive volumes overlap on the GPU
    cudaStream_t nms = nullptr;                                 // high priority: the NMS of a volume must not queue behind the chains of others
    cudaEvent_t in_done[HB_SLOTS] = {}, comp_done[HB_SLOTS] = {}, cnt_done[HB_SLOTS] = {}, out_done[HB_SLOTS] = {};
    cudaEvent_t nmsc_done[HB_SLOTS] = {}, nms_done[HB_SLOTS] = {}, pack_done[HB_SLOTS] = {};
    int device = -1;
    char* pinned = nullptr;           // grow-only pinned staging: per-volume bookkeeping + per-slot compacted groups
    size_t pinned_cap = 0;
    int ensure_pinned(size_t bytes) {
        if (bytes <= pinned_cap) return 0;
        if (pinned) { cudaFreeHost(pinned); pinned = nullptr; pinned_cap = 0; }
        const size_t want = align_up(bytes + (bytes >> 2), 1 << 16);
        B200_CUDA(cudaHostAlloc((void**)&pinned, want, cudaHostAllocDefault));
        pinned_cap = want;
        return 0;
    }
    int ensure(int dev) {
        if (device == dev && in) return 0;
        if (in) {
            cudaStreamDestroy(in); cudaStreamDestroy(out); cudaStreamDestroy(out2); cudaStreamDestroy(out3); cudaStreamDestroy(out4);
            for (int k = 0; k < HB_SLOTS; ++k) cudaEventDestroy(scat_done[k]); cudaStreamDestroy(comp2); cudaStreamDestroy(comp3); cudaStreamDestroy(nms);
            for (int k = 0; k < HB_SLOTS; ++k) {
                cudaEventDestroy(in_done[k]); cudaEventDestroy(comp_done[k]); cudaEventDestroy(cnt_done[k]); cudaEventDestroy(out_done[k]);
                cudaEventDestroy(nmsc_done[k]); cudaEventDestroy(nms_done[k]); cudaEventDestroy(pack_done[k]);
            }
        }
        B200_CUDA(cudaStreamCreateWithFlags(&in, cudaStreamNonBlocking));
        B200_CUDA(cudaStreamCreateWithFlags(&out, cudaStreamNonBlocking));
        B200_CUDA(cudaStreamCreateWithFlags(&out2, cudaStreamNonBlocking));
        B200_CUDA(cudaStreamCreateWithFlags(&out3, cudaStreamNonBlocking));
        B200_CUDA(cudaStreamCreateWithFlags(&out4, cudaStreamNonBlocking));
        for (int k = 0; k < HB_SLOTS; ++k) B200_CUDA(cudaEventCreateWithFlags(&scat_done[k], cudaEventDisableTiming));
        B200_CUDA(cudaStreamCreateWithFlags(&comp2, cudaStreamNonBlocking));
        B200_CUDA(cudaStreamCreateWithFlags(&comp3, cudaStreamNonBlocking));
        {
            int lo_p = 0, hi_p = 0;
            B200_CUDA(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
            B200_CUDA(cudaStreamCreateWithPriority(&nms, cudaStreamNonBlocking, hi_p));
        }
        for (int k = 0; k < HB_SLOTS; ++k) {
            B200_CUDA(cudaEventCreateWithFlags(&in_done[k], cudaEventDisableTiming));
            B200_CUDA(cudaEventCreateWithFlags(&nmsc_done[k], cudaEventDisableTiming));
            B200_CUDA(cudaEventCreateWithFlags(&nms_done[k], cudaEventDisableTiming));
            B200_CUDA(cudaEventCreateWithFlags(&pack_done[k], cudaEventDisableTiming));
            B200_CUDA(cudaEventCreateWithFlags(&comp_done[k], cudaEventDisableTiming));
            B200_CUDA(cudaEventCreateWithFlags(&cnt_done[k], cudaEventDisableTiming));
            B200_CUDA(cudaEventCreateWithFlags(&out_done[k], cudaEventDisableTiming));
        }
        device = dev;
        return 0;
    }
};
static BatchStreams g_batch_dev[64];      // one set per device ordinal (guarded by the host context's mutex)
static std::atomic<unsigned long long> g_last_h2d{0}, g_last_d2h{0};

// "host_batch_out" = 2: the caller's label buffer still holds the result this entry point wrote into it last time, so
// instead of zero-filling all of it (2 bytes per voxel of host memory traffic, the bound of the whole call) only the
// non-zero 64-byte lines written last time are cleared.  Keyed by the buffer address; an entry is dropped whenever the
// content of the buffer is not exactly "zeros + the listed lines" (dense download, error, different size).
struct PrevLabels { size_t bytes = 0; std::vector<uint32_t> groups; };
static std::mutex g_prev_mu;
static std::unordered_map<const void*, PrevLabels>& g_prev = *new std::unordered_map<const void*, PrevLabels>;

// device pointer of a host buffer the GPU can read in place (pinned / registered and 16-byte aligned), else null
static const uint8_t* mapped_device_pointer(const void* host) {
    if (!host || (((uintptr_t)host) & 15)) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return (const uint8_t*)at.devicePointer;
}

}  // namespace b200seg

using namespace b200seg;

extern "C" void b200seg_postproc_soma_host_batch_traffic(unsigned long long* h2d_bytes, unsigned long long* d2h_bytes) {
    if (h2d_bytes) *h2d_bytes = g_last_h2d.load();
    if (d2h_bytes) *d2h_bytes = g_last_d2h.load();
}

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
    const size_t V = (size_t)S * H * W;
    B200_CHECK_ARG(V < (1ull << 34), "postproc_soma_host_batch: volume too large");
    HostCtx& hc = host_ctx();
    std::lock_guard<std::mutex> lock(hc.mu);
    const size_t nn = n_max > 0 ? n_max : 1;
    const size_t ws_bytes = b200seg_postproc_soma_workspace_bytes(1, n_max, S, H, W, keep_largest_cc ? (long long)prm_max : 0);
    const int NB = n_volumes < HB_SLOTS ? n_volumes : HB_SLOTS;
    // compacted form: only when the volume splits into whole 16-byte groups (the kernel loads them as uint4)
    const int mode = opt_host_batch_mode();
    const bool sparse = (V % 8) == 0 && (mode & 1);
    const size_t ngroups = V / 8;
    const size_t nlines = (ngroups + 3) / 4;                  // 64-byte lines (the last one may be partial)
    size_t cap = nlines / 4;                                  // lines the compacted form may hold
    if (cap < 16384) cap = 16384;
    if (cap > nlines) cap = nlines;
    if (!sparse) cap = 0;
    const bool pack_mode = (mode & 4) != 0;                   // image crops of the NMS survivors packed by host threads, no DMA of the volume
    // volumes per chain launch; consecutive slots of a group are adjacent in every array the chain kernels index by volume
    // ("host_batch_mode" bits 4 / 5: groups of 2 / 4 volumes.  Measured on 64 volumes, same box, alternating runs: 80.6 Gvox/s
    // with single-volume launches, 78.3 in pairs, 72.6 in fours -- the call is bound by the host's memory traffic and by the
    // upload direction of the link, not by the kernels, and a group delays the downloads of its first volumes; default 1.)
    const int G = (mode & 32) ? HB_GROUP : (mode & 16) ? 2 : 1;
    const int n_groups_dev = (NB + G - 1) / G;
    const size_t nd = nn + 1;                                 // per-volume stride of the per-detection arrays (one spare entry: crop_off[n])
    const size_t P = align_up(prm_max + 16, 256);            // per-volume stride of the PRM crops and of the masks
    const size_t ws_bytes_g = b200seg_postproc_soma_workspace_bytes(G, (int)nd, S, H, W, keep_largest_cc ? (long long)(P * G) : 0);
    // "host_batch_mode" bit 3 (with bit 2): the PRM crops of the survivors with a positive voxel travel in the same pinned buffer,
    // behind the image crops, instead of being pulled over the link by the zero-copy gather kernel (16-byte reads reach about a
    // quarter of the DMA rate).  It takes the wait for the GPU out of the call (12 -> 1-3 ms of 25) and puts the same time into
    // the host pool, which copies 2.5 MB more per volume: 69.6 against 78.6 Gvox/s on the 16-core box, so it is off by default;
    // it is the better choice where host cores are plentiful and for pageable PRM buffers (no whole-array copy).
    const bool prm_packed = pack_mode && (mode & 8);
    const size_t packbuf_bytes = align_up(prm_max + 16, 256) + (prm_packed ? align_up(prm_max + 32 * nn + 64, 256) : 0);
    const size_t slot_bytes = (pack_mode ? Carver::need(packbuf_bytes) + 2 * Carver::need(nn * 8) + Carver::need(nn) : 0) + Carver::need(nn * 28) + Carver::need(8) +
                              Carver::need(nn * 8) + Carver::need(cap * 4 + 16) + Carver::need(cap * 64 + 64) + Carver::need(ws_bytes);
    const size_t shared_bytes = Carver::need(V * NB) + Carver::need(V * 2 * NB) + 2 * Carver::need(P * NB) + Carver::need(nd * 24 * NB) + Carver::need(nd * 8 * NB) +
                                3 * Carver::need(nd * 4 * NB) + Carver::need(nd * NB) + 2 * Carver::need(4 * (size_t)HB_SLOTS) + Carver::need(4 * (size_t)(HB_SLOTS + 1)) +
                                (size_t)n_groups_dev * Carver::need(ws_bytes_g);
    int e = hc.ensure(slot_bytes * NB + shared_bytes);
    if (e) return e;
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_CHECK_ARG(dev >= 0 && dev < 64, "postproc_soma_host_batch: device ordinal out of range");
    BatchStreams& g_batch = g_batch_dev[dev];
    e = g_batch.ensure(dev);
    if (e) return e;
    struct Slot {
        uint8_t* pack; int64_t* pk_off; int64_t* pk_poff; uint8_t* pk_zero; uint8_t* vol; uint16_t* seg; float* dets; int32_t* off; int32_t* boxes; uint8_t* prm; uint8_t* mask; int64_t* coff;
        int64_t* keep; int32_t* cnt; uint32_t* lines; int32_t* rank; int32_t* bmax; int32_t* stat; uint8_t* surv; uint32_t* gidx; uint4* gval; void* ws;
    } slot[HB_SLOTS];
    Carver cs(hc.buf + slot_bytes * NB);
    uint8_t* const vol_all = cs.take<uint8_t>(V * NB);
    uint16_t* const seg_all = cs.take<uint16_t>(V * NB);
    uint8_t* const prm_all = cs.take<uint8_t>(P * NB);
    uint8_t* const mask_all = cs.take<uint8_t>(P * NB);
    int32_t* const boxes_all = cs.take<int32_t>(nd * 6 * NB);
    int64_t* const coff_all = cs.take<int64_t>(nd * NB);
    int32_t* const rank_all = cs.take<int32_t>(nd * NB);
    int32_t* const bmax_all = cs.take<int32_t>(nd * NB);
    int32_t* const stat_all = cs.take<int32_t>(nd * NB);
    uint8_t* const surv_all = cs.take<uint8_t>(nd * NB);
    int32_t* const keepcnt_all = cs.take<int32_t>(HB_SLOTS);        // NMS survivors per slot: the chain's n_valid[volume]
    uint32_t* const linecnt_all = cs.take<uint32_t>(HB_SLOTS);      // compacted lines per slot
    int32_t* const det_off_tab = cs.take<int32_t>(HB_SLOTS + 1);    // {0, nd, 2 nd, ...}: detection offsets inside a group
    void* group_ws[HB_SLOTS];
    for (int q = 0; q < n_groups_dev; ++q) group_ws[q] = cs.take<char>(ws_bytes_g);
    for (int k = 0; k < NB; ++k) {
        Carver cv(hc.buf + slot_bytes * k);
        Slot& s = slot[k];
        s.pack = nullptr; s.pk_off = nullptr; s.pk_poff = nullptr; s.pk_zero = nullptr;
        if (pack_mode) { s.pack = cv.take<uint8_t>(packbuf_bytes); s.pk_off = cv.take<int64_t>(nn); s.pk_poff = cv.take<int64_t>(nn); s.pk_zero = cv.take<uint8_t>(nn); }
        s.dets = cv.take<float>(nn * 7); s.off = cv.take<int32_t>(2); s.keep = cv.take<int64_t>(nn);
        s.gidx = cv.take<uint32_t>(cap + 4); s.gval = cv.take<uint4>(cap * 4 + 4); s.ws = cv.p;          // (s.ws: workspace of this volume's NMS)
        s.vol = vol_all + V * k; s.seg = seg_all + V * k; s.prm = prm_all + P * k; s.mask = mask_all + P * k;
        s.boxes = boxes_all + nd * 6 * k; s.coff = coff_all + nd * k; s.rank = rank_all + nd * k; s.bmax = bmax_all + nd * k;
        s.stat = stat_all + nd * k; s.surv = surv_all + nd * k; s.cnt = keepcnt_all + k; s.lines = linecnt_all + k;
    }
    // pinned staging: [per-volume bookkeeping] [per-slot line indices | line payloads]
    size_t small_bytes = 0;
    for (int v = 0; v < n_volumes; ++v) small_bytes += align_up(16 + 13 * (size_t)n_dets[v], 16);
    small_bytes = align_up(small_bytes, 256);
    const size_t stage_bytes = align_up(cap * 4, 256) + align_up(cap * 64, 256);
    // packed-image mode, per slot: [keep count | pad to 16 | visit order n*4] [offsets n*8 | all-zero-PRM flags n] [packed crops]
    const size_t nmsst_bytes = align_up(16 + nn * 4, 256), pkoff_bytes = 2 * align_up(nn * 8, 256) + align_up(nn, 256);
    const size_t pack_bytes = pack_mode ? nmsst_bytes + pkoff_bytes + packbuf_bytes : 0;
    const size_t coffst_bytes = align_up(nd * 8, 256);       // per slot: the volume's crop offsets counted from its group's PRM base
    e = g_batch.ensure_pinned(small_bytes + (stage_bytes + pack_bytes + coffst_bytes) * NB + 256);
    if (e) return e;
    g_pool.ensure();
    char* const pack_base = g_batch.pinned + small_bytes + stage_bytes * NB;
    char* const coffst_base = pack_base + pack_bytes * NB;
    int32_t* const det_tab_host = (int32_t*)(coffst_base + coffst_bytes * NB);

    cudaStream_t s_in = g_batch.in, s_out = g_batch.out, s_out2 = g_batch.out2;
    // the kernels of one volume are small (a dozen launches of a few hundred CTAs): volumes rotate over three compute streams
    // so that the chains of consecutive volumes overlap on the GPU
    const cudaStream_t comp_streams[3] = {hc.stream, g_batch.comp2, g_batch.comp3};
    // host-side state shared with the pool (kept alive until every job has run)
    struct Shared {
        std::vector<std::atomic<int>> zero_left;                  // per volume: zero-fill parts still running
        std::atomic<int> slot_busy[HB_SLOTS];                     // staging slot still being scattered
        std::atomic<int> pending{0};                              // jobs pushed and not finished
        explicit Shared(int nv) : zero_left(nv) {}
    };
    Shared* sh = new Shared(n_volumes);
    for (int k = 0; k < HB_SLOTS; ++k) sh->slot_busy[k].store(0);
    // zero-fill of the label volumes starts now (it does not depend on the GPU)
    const int out_state = sparse ? opt_host_batch_out() : 0;
    for (int v = 0; v < n_volumes; ++v) {
        if (out_state == 1) { sh->zero_left[v].store(0); continue; }            // the caller vouches for zeros
        if (out_state == 2) {
            std::vector<uint32_t>* prev = nullptr;
            {
                std::lock_guard<std::mutex> lk(g_prev_mu);
                auto it = g_prev.find(seg[v]);
                if (it != g_prev.end() && it->second.bytes == V * 2) prev = new std::vector<uint32_t>(std::move(it->second.groups));
                if (it != g_prev.end()) g_prev.erase(it);                       // re-registered by the scatter job of this call
            }
            if (prev) {
                sh->zero_left[v].store(1);
                char* dst = (char*)seg[v];
                sh->pending.fetch_add(1);
                const size_t vbytes0 = V * 2;
                g_pool.push([sh, v, dst, prev, vbytes0] {
                    clear_lines(dst, vbytes0, prev->data(), (uint32_t)prev->size());
                    delete prev;
                    sh->zero_left[v].fetch_sub(1, std::memory_order_release);
                    sh->pending.fetch_sub(1, std::memory_order_release);
                });
                continue;
            }
        }
        sh->zero_left[v].store(HB_ZERO_PARTS);
        const size_t bytes = V * 2, part = align_up((bytes + HB_ZERO_PARTS - 1) / HB_ZERO_PARTS, 4096);
        for (int p = 0; p < HB_ZERO_PARTS; ++p) {
            const size_t lo = (size_t)p * part < bytes ? (size_t)p * part : bytes;
            const size_t hi = lo + part < bytes ? lo + part : bytes;
            char* dst = (char*)seg[v] + lo;
            sh->pending.fetch_add(1);
            g_pool.push([sh, v, dst, lo, hi] {
                if (hi > lo) zero_stream(dst, hi - lo);
                sh->zero_left[v].fetch_sub(1, std::memory_order_release);
                sh->pending.fetch_sub(1, std::memory_order_release);
            });
        }
    }
    static const bool trace = getenv("B200SEG_HB_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
    const auto t_start = now();
    double w_cnt = 0, w_slot = 0, w_out = 0, w_zero = 0, w_tail = 0, w_nms = 0, w_pack = 0;
    std::vector<int32_t> offs(2 * (size_t)n_volumes);            // per-volume {0, n} pairs, alive until the copies have run
    for (int v = 0; v < n_volumes; ++v) { offs[2 * v] = 0; offs[2 * v + 1] = n_dets[v]; }
    std::vector<size_t> small_off(n_volumes);
    { size_t o = 0; for (int v = 0; v < n_volumes; ++v) { small_off[v] = o; o += align_up(16 + 13 * (size_t)n_dets[v], 16); } }
    std::vector<uint32_t> n_groups(n_volumes, 0u);               // compacted groups per volume (0xFFFFFFFF = dense copy)
    // Lines written in place by the GPU (pinned label buffers only) or staged download + host scatter?  The staged form is a
    // little faster while the host has cores and memory bandwidth to spare (one rank on a 16-core box: 74.6 against 70.8 Gvox/s);
    // when the ranks of a box share the host it is the host's DRAM traffic that bounds the call and the in-place form wins
    // (8 ranks on 32 cores, 8 volumes each: 16.3 against 20.4 ms per call, 132 against 105 Gvox/s in aggregate).  Default: in
    // place when a rank has fewer than 8 cores to itself; "host_batch_mode" bit 6 forces the staged form, bit 8 the in-place one.
    bool want_direct = (mode & 256) != 0;
    if (!(mode & (64 | 256))) {
        int hw = (int)std::thread::hardware_concurrency(), share = 1;
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) share = atoi(e) > 0 ? atoi(e) : 1;
        want_direct = hw > 0 && hw / share < 8;
    }
    std::vector<uint4*> direct(n_volumes, nullptr);             // device view of the caller's label volume where the GPU can write it in place
    for (int v = 0; v < n_volumes && sparse && want_direct; ++v) direct[v] = (uint4*)mapped_device_pointer(seg[v]);
    unsigned long long h2d = 0, d2h = 0;
    int rc = 0;
#define B200_BATCH(call) do { int _e = ::b200seg::check_cuda((call), #call); if (_e) { rc = _e; goto done; } } while (0)
    const int lag_n = pack_mode ? HB_LAG_N : 0;               // stage A (NMS) runs lag_n volumes ahead of stage B, stage P lag_n - HB_LAG_P
    std::atomic<int> pack_left[HB_SLOTS];                       // packing jobs of the slot's volume still running
    std::atomic<int> scan_left[HB_SLOTS];                       // ... and the PRM scans that precede them
    int pack_kc[HB_SLOTS] = {};
    size_t pack_total[HB_SLOTS] = {};
    for (int k = 0; k < HB_SLOTS; ++k) { pack_left[k].store(0); scan_left[k].store(0); }
    // chains in flight on the GPU: the host sizes the download of volume v (it needs its line count) while it enqueues volume
    // v + lag_a, so lag_a + 1 chains can be queued before the host has to wait for one of them
    static const int lag_env = getenv("B200SEG_HB_LAG") ? atoi(getenv("B200SEG_HB_LAG")) : 0;
    // (a chain is launched when the last volume of its group arrives: G - 1 more steps for the first one)
    const int lag_a = (lag_env >= 1 && lag_env <= 2 ? lag_env : HB_LAG_A) + G - 1, lag_b = lag_a + (HB_LAG_B - HB_LAG_A);
    static_assert(HB_LAG_N + 2 + HB_GROUP - 1 + (HB_LAG_B - HB_LAG_A) + 1 <= HB_SLOTS, "a slot must be free again before its next volume arrives");
    static_assert(HB_SLOTS % HB_GROUP == 0, "groups must not wrap around the slot ring");
    for (int j = 0; j <= HB_SLOTS; ++j) det_tab_host[j] = (int32_t)(j * nd);
    B200_BATCH(cudaMemcpyAsync(det_off_tab, det_tab_host, 4 * (size_t)(HB_SLOTS + 1), cudaMemcpyHostToDevice, s_in));
    for (int step = 0; step < n_volumes + lag_n + lag_b; ++step) {
        // ---- stage A, volume `step`: uploads of the small arrays (and of the volume unless it travels packed), NMS -------
        if (step < n_volumes) {
            const int v = step, k = v % NB;
            const cudaStream_t s_comp = pack_mode ? g_batch.nms : comp_streams[(v / G) % 3];   // (stage A only enqueues the NMS; else the group's stream)
            Slot& s = slot[k];
            const int n = n_dets[v];
            const size_t pbytes = n > 0 ? (size_t)crop_off[v][n] : 0;
            const uint8_t* prm_mapped = (n > 0 && (mode & 2)) ? mapped_device_pointer(prm[v]) : nullptr;
            if (v >= NB) B200_BATCH(cudaStreamWaitEvent(s_in, g_batch.out_done[k], 0));     // slot free again
            if (!pack_mode && n > 0) { B200_BATCH(cudaMemcpyAsync(s.vol, volumes[v], V, cudaMemcpyHostToDevice, s_in)); h2d += V; }
            B200_BATCH(cudaMemcpyAsync(s.off, offs.data() + 2 * v, 8, cudaMemcpyHostToDevice, s_in));
            h2d += 8;
            if (n > 0) {
                B200_BATCH(cudaMemcpyAsync(s.dets, dets[v], (size_t)n * 28, cudaMemcpyHostToDevice, s_in));
                B200_BATCH(cudaMemcpyAsync(s.boxes, boxes[v], (size_t)n * 24, cudaMemcpyHostToDevice, s_in));
                {   // crop offsets counted from the PRM base of the volume's group
                    int64_t* cst = (int64_t*)(coffst_base + coffst_bytes * k);
                    if (v >= NB) B200_BATCH(cudaEventSynchronize(g_batch.in_done[k]));     // (the copy out of this staging for volume v - NB ran long ago)
                    const int64_t shift = (int64_t)(P * (size_t)(k % G));
                    for (int i = 0; i <= n; ++i) cst[i] = crop_off[v][i] + shift;
                    B200_BATCH(cudaMemcpyAsync(s.coff, cst, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, s_in));
                }
                h2d += (size_t)n * 52 + (size_t)(n + 1) * 8;
                if (!prm_mapped && !prm_packed) { B200_BATCH(cudaMemcpyAsync(s.prm, prm[v], pbytes, cudaMemcpyHostToDevice, s_in)); h2d += pbytes; }
            }
            B200_BATCH(cudaEventRecord(g_batch.in_done[k], s_in));
            B200_BATCH(cudaStreamWaitEvent(s_comp, g_batch.in_done[k], 0));
            {
                const long long cc_bytes = keep_largest_cc ? (long long)pbytes : 0;
                const SomaChainWs L = soma_chain_ws(s.ws, ws_bytes, 1, n, S, H, W, cc_bytes);
                int ce = b200seg_nms3d_dev(s.dets, s.off, 1, n, nms_thresh, 0, s.keep, s.cnt, s.rank, L.nms_ws, L.nms_ws_bytes, s_comp);
                if (ce) { rc = ce; goto done; }
            }
            if (pack_mode) B200_BATCH(cudaEventRecord(g_batch.nmsc_done[k], s_comp));
            if (pack_mode && n > 0) {                         // the visit order comes back right away: the host packs by it
                char* nst = pack_base + pack_bytes * k;
                // on its own stream: on the bookkeeping stream it would queue behind the downloads of older volumes, which wait for
                // their chains -- the NMS would then run only as far ahead as the chains let it ("host_batch_mode" bit 7: as before)
                const cudaStream_t s_nmsout = (mode & 128) ? s_out : g_batch.out3;
                B200_BATCH(cudaStreamWaitEvent(s_nmsout, g_batch.nmsc_done[k], 0));
                B200_BATCH(cudaMemcpyAsync(nst, s.cnt, 4, cudaMemcpyDeviceToHost, s_nmsout));
                B200_BATCH(cudaMemcpyAsync(nst + 16, s.rank, (size_t)n * 4, cudaMemcpyDeviceToHost, s_nmsout));
                B200_BATCH(cudaEventRecord(g_batch.nms_done[k], s_nmsout));
                d2h += 4 + (size_t)n * 4;
            }
        }
        // ---- stage P, volume `step - HB_LAG_P`: its visit order is back -> hand the packing of its image crops to the pool ----
        const int sp = step - HB_LAG_P;                          // (the copies get lag_n - HB_LAG_P steps before stage B needs them)
        if (pack_mode && sp >= 0 && sp < n_volumes && n_dets[sp] > 0) {
            const int v = sp, k = v % NB;
            const int n = n_dets[v];
            char* nst = pack_base + pack_bytes * k;
            int64_t* pko = (int64_t*)(nst + nmsst_bytes);
            int64_t* ppo = (int64_t*)(nst + nmsst_bytes + align_up(nn * 8, 256));
            uint8_t* pkz = (uint8_t*)(nst + nmsst_bytes + 2 * align_up(nn * 8, 256));
            uint8_t* pk = (uint8_t*)(nst + nmsst_bytes + pkoff_bytes);
            // the previous user of this slot's pinned buffers was volume v - NB: its copies were issued long ago; make sure they are done
            if (v >= NB) B200_BATCH(cudaEventSynchronize(g_batch.pack_done[k]));
            { const auto t = now(); B200_BATCH(cudaEventSynchronize(g_batch.nms_done[k])); w_nms += ms_since(t); }
            int32_t kc = 0;
            memcpy(&kc, nst, 4);
            if (kc > n) kc = n;
            if (kc < 0) kc = 0;
            const int32_t* ro = (const int32_t*)(nst + 16);
            pack_kc[k] = kc; pack_total[k] = 0;
            // Two rounds of jobs.  Round 1 looks for a positive voxel in the PRM crop of every survivor (an instance without one
            // is skipped by the script before it touches the image, binarization_soma.py:74-76: nothing of it travels).  The
            // job that finishes last lays out the buffer -- image crops back to back, then the PRM crops, each starting at its
            // own offset modulo 16 so that the device copy moves aligned 16-byte words -- and hands out round 2, the copies.
            const int njobs = kc < 1 ? 0 : (kc < 4 * g_pool.n_threads ? (kc + 3) / 4 : g_pool.n_threads * 2);
            pack_left[k].store(njobs > 0 ? 1 : 0);                 // round 2 not handed out yet
            scan_left[k].store(njobs);
            const uint8_t* vol = volumes[v];
            const uint8_t* prm_h = prm[v];
            const int64_t* co = crop_off[v];
            const int32_t* bxs = boxes[v];
            std::atomic<int>* left = &pack_left[k];
            std::atomic<int>* sleft = &scan_left[k];
            size_t* total_out = &pack_total[k];
            HostPool* pool = &g_pool;
            for (int j = 0; j < njobs; ++j) {
                const int r0 = (int)((long long)kc * j / njobs), r1 = (int)((long long)kc * (j + 1) / njobs);
                g_pool.push([=] {
                    for (int r = r0; r < r1; ++r) {
                        const int i = ro[r];
                        const int32_t* bx = bxs + (size_t)i * 6;
                        const int sx = bx[3] - bx[0] + 1, sy = bx[4] - bx[1] + 1, sz = bx[5] - bx[2] + 1;
                        pkz[r] = 1;
                        if (sx <= 0 || sy <= 0 || sz <= 0) continue;
                        const uint8_t* pc = prm_h + co[i];
                        const size_t pn = (size_t)(co[i + 1] - co[i]);
                        size_t q = 0;
                        unsigned long long acc = 0;
                        for (; q + 8 <= pn && !acc; q += 8) { unsigned long long w; memcpy(&w, pc + q, 8); acc |= w; }
                        for (; q < pn && !acc; ++q) acc |= pc[q];
                        pkz[r] = acc ? 0 : 1;
                    }
                    if (sleft->fetch_sub(1, std::memory_order_acq_rel) != 1) return;
                    // last scan job: layout + round 2
                    size_t total = 0;
                    for (int r = 0; r < kc; ++r) {
                        pko[r] = (int64_t)total;
                        if (!pkz[r]) total += (size_t)(co[ro[r] + 1] - co[ro[r]]);
                    }
                    if (prm_packed)
                        for (int r = 0; r < kc; ++r) {
                            ppo[r] = 0;
                            if (pkz[r]) continue;
                            total = ((total + 15) & ~(size_t)15) + ((size_t)co[ro[r]] & 15);
                            ppo[r] = (int64_t)total;
                            total += (size_t)(co[ro[r] + 1] - co[ro[r]]);
                        }
                    *total_out = total;
                    left->fetch_add(njobs, std::memory_order_relaxed);
                    for (int j2 = 0; j2 < njobs; ++j2) {
                        const int q0 = (int)((long long)kc * j2 / njobs), q1 = (int)((long long)kc * (j2 + 1) / njobs);
                        pool->push([=] {
                            for (int r = q0; r < q1; ++r) {
                                if (pkz[r]) continue;
                                const int i = ro[r];
                                const int32_t* bx = bxs + (size_t)i * 6;
                                const int sx = bx[3] - bx[0] + 1, sy = bx[4] - bx[1] + 1, sz = bx[5] - bx[2] + 1;
                                pack_box_rows(pk + pko[r], vol + ((size_t)bx[2] * H + bx[1]) * W + bx[0], sx, sy, sz, (size_t)W, (size_t)H * W);
                                if (prm_packed) memcpy(pk + ppo[r], prm_h + co[i], (size_t)(co[i + 1] - co[i]));
                            }
                            left->fetch_sub(1, std::memory_order_release);
                        }, true);
                    }
                    left->fetch_sub(1, std::memory_order_release);             // the round-2 marker
                }, true);
            }
        }
        const int sb = step - lag_n;
        // ---- stage B, volume `sb`: (packed image crops,) PRM gather; with the last volume of a group: chain of the group,
        //      compaction and download of the bookkeeping of each of its volumes ---------------------------------------------
        if (sb >= 0 && sb < n_volumes) {
            const int v = sb, k = v % NB;
            const int g0 = v - v % G, g1 = (g0 + G < n_volumes ? g0 + G : n_volumes) - 1;      // first / last volume of the group
            const int k0 = g0 % NB, ng_vol = g1 - g0 + 1;
            const cudaStream_t s_comp = comp_streams[(v / G) % 3];
            Slot& s = slot[k];
            const int n = n_dets[v];
            const size_t pbytes = n > 0 ? (size_t)crop_off[v][n] : 0;
            const uint8_t* prm_mapped = (n > 0 && (mode & 2)) ? mapped_device_pointer(prm[v]) : nullptr;
            B200_BATCH(cudaStreamWaitEvent(s_comp, pack_mode ? g_batch.nmsc_done[k] : g_batch.in_done[k], 0));   // the NMS ran on its own stream
            const uint8_t* zero_flags = nullptr;
            if (pack_mode && n > 0) {
                char* nst = pack_base + pack_bytes * k;
                int64_t* pko = (int64_t*)(nst + nmsst_bytes);
                int64_t* ppo = (int64_t*)(nst + nmsst_bytes + align_up(nn * 8, 256));
            uint8_t* pkz = (uint8_t*)(nst + nmsst_bytes + 2 * align_up(nn * 8, 256));
                uint8_t* pk = (uint8_t*)(nst + nmsst_bytes + pkoff_bytes);
                { const auto tp = now(); while (pack_left[k].load(std::memory_order_acquire) > 0) sched_yield(); w_pack += ms_since(tp); }
                const int kc = pack_kc[k];
                if (kc > 0) {
                    B200_BATCH(cudaMemcpyAsync(s.pk_off, pko, (size_t)kc * 8, cudaMemcpyHostToDevice, s_in));
                    B200_BATCH(cudaMemcpyAsync(s.pk_zero, pkz, (size_t)kc, cudaMemcpyHostToDevice, s_in));
                    if (prm_packed) { B200_BATCH(cudaMemcpyAsync(s.pk_poff, ppo, (size_t)kc * 8, cudaMemcpyHostToDevice, s_in)); h2d += (size_t)kc * 8; }
                    if (pack_total[k] > 0) B200_BATCH(cudaMemcpyAsync(s.pack, pk, pack_total[k], cudaMemcpyHostToDevice, s_in));
                    h2d += (size_t)kc * 9 + pack_total[k];
                    zero_flags = s.pk_zero;
                }
                B200_BATCH(cudaEventRecord(g_batch.pack_done[k], s_in));
                B200_BATCH(cudaStreamWaitEvent(s_comp, g_batch.pack_done[k], 0));
                if (kc > 0) {
                    img_unpack_kernel<<<dim3(n, HB_UNPACK_PARTS), 256, 0, s_comp>>>(s.pack, s.pk_off, s.pk_zero, s.vol, H, W, s.boxes, s.rank, s.cnt);
                    count_launch();
                    B200_BATCH(cudaGetLastError());
                }
            }
            if (prm_packed) {                                 // PRM crops of the survivors: packed buffer (on the device by now) -> their own offsets
                if (n > 0 && pack_kc[k] > 0) {
                    prm_gather_kernel<<<n, 256, 0, s_comp>>>(s.pack, s.prm, s.coff, (long long)(P * (size_t)(k % G)), s.rank, s.cnt, s.pk_zero, s.pk_poff);
                    count_launch();
                    B200_BATCH(cudaGetLastError());
                }
            } else if (prm_mapped) {                          // PRM crops of the survivors, fetched in place (zero fill where the host saw only zeros)
                prm_gather_kernel<<<n, 256, 0, s_comp>>>(prm_mapped, s.prm, s.coff, (long long)(P * (size_t)(k % G)), s.rank, s.cnt, zero_flags, nullptr);
                count_launch();
                B200_BATCH(cudaGetLastError());
            }
            if (v == g1) {
                // every array the chain indexes by volume is strided per slot, so the group is one batched launch: volume j of the
                // group = slot k0 + j, its detections start at j * nd, its PRM crops / masks at j * P
                int n_any = 0;
                for (int u = g0; u <= g1; ++u) n_any = n_dets[u] > n_any ? n_dets[u] : n_any;
                const long long mask_bytes = (long long)(P * (size_t)ng_vol);
                const SomaChainWs L = soma_chain_ws(group_ws[k0 / G], ws_bytes_g, ng_vol, (int)nd, S, H, W, keep_largest_cc ? mask_bytes : 0);
                Slot& s0 = slot[k0];
                if (sparse) B200_BATCH(cudaMemsetAsync(s0.lines, 0, 4 * (size_t)ng_vol, s_comp));
                // The paste kernel can emit the compacted lines itself (B200SEG_HB_FUSED_LINES=1, single-volume groups only) instead of
                // a second pass over the volume.  Measured through this entry point (64 volumes, same box, alternating runs): 73.9
                // Gvox/s fused against 76.4 with the separate pass -- the extra ballots / atomics lengthen the paste kernel by more
                // than the 50 us streaming kernel they replace, so the separate pass stays the default.
                static const bool fused_on = getenv("B200SEG_HB_FUSED_LINES") != nullptr;
                const bool fused_lines = sparse && fused_on && ng_vol == 1 && paste_lines_supported(s0.seg, 1, S, H, W);
                const PasteLines pl{s0.gidx, s0.gval, s0.lines, (uint32_t)cap};
                int ce = postproc_soma_after_nms(s0.vol, ng_vol, S, H, W, det_off_tab, n_any > 0 ? (int)nd : 0, (int)(nd * ng_vol), s0.boxes, s0.prm, s0.coff,
                                                 mask_bytes, keep_largest_cc, s0.seg, s0.cnt, s0.rank, s0.mask, s0.bmax, s0.stat, s0.surv, L.ids,
                                                 L.paste_ws, L.paste_ws_bytes, L.cc_ws, L.cc_ws_bytes, s_comp, fused_lines ? &pl : nullptr);
                if (ce) { rc = ce; goto done; }
                for (int u = g0; u <= g1 && sparse && !fused_lines; ++u) {
                    Slot& su = slot[u % NB];
                    const unsigned int grid = (unsigned int)((ngroups + 1023) / 1024);          // one trip per CTA (the loop covers grids beyond 2^31 groups)
                    seg_compact_kernel<<<grid, 256, 0, s_comp>>>((const uint4*)su.seg, (unsigned int)ngroups, (unsigned int)cap,
                                                                 su.gidx, su.gval, su.lines);
                    count_launch();
                    B200_BATCH(cudaGetLastError());
                }
                B200_BATCH(cudaEventRecord(g_batch.comp_done[k], s_comp));
                B200_BATCH(cudaStreamWaitEvent(s_out, g_batch.comp_done[k], 0));
                for (int u = g0; u <= g1 && sparse; ++u) {
                    if (!direct[u]) continue;                      // the lines go straight into the caller's (pinned) volume: it must be zero by now
                    Slot& su = slot[u % NB];
                    { const auto t = now(); while (sh->zero_left[u].load(std::memory_order_acquire) > 0) sched_yield(); w_zero += ms_since(t); }
                    B200_BATCH(cudaStreamWaitEvent(g_batch.out4, g_batch.comp_done[k], 0));
                    lines_to_host_kernel<<<HB_SCATTER_CTAS, 256, 0, g_batch.out4>>>(su.gidx, su.gval, su.lines, (unsigned int)cap, (unsigned int)ngroups, direct[u]);
                    count_launch();
                    B200_BATCH(cudaGetLastError());
                    B200_BATCH(cudaEventRecord(g_batch.scat_done[u % NB], g_batch.out4));
                }
                for (int u = g0; u <= g1; ++u) {
                    Slot& su = slot[u % NB];
                    const int nu = n_dets[u];
                    char* st = g_batch.pinned + small_off[u];      // [keep count | line count | pad to 16 | rank n*4 | b_max n*4 | status n*4 | survive n]
                    B200_BATCH(cudaMemcpyAsync(st, su.cnt, 4, cudaMemcpyDeviceToHost, s_out));
                    B200_BATCH(cudaMemcpyAsync(st + 4, su.lines, 4, cudaMemcpyDeviceToHost, s_out));
                    d2h += 8;
                    if (nu > 0) {
                        B200_BATCH(cudaMemcpyAsync(st + 16, su.rank, (size_t)nu * 4, cudaMemcpyDeviceToHost, s_out));
                        B200_BATCH(cudaMemcpyAsync(st + 16 + (size_t)nu * 4, su.bmax, (size_t)nu * 4, cudaMemcpyDeviceToHost, s_out));
                        B200_BATCH(cudaMemcpyAsync(st + 16 + (size_t)nu * 8, su.stat, (size_t)nu * 4, cudaMemcpyDeviceToHost, s_out));
                        B200_BATCH(cudaMemcpyAsync(st + 16 + (size_t)nu * 12, su.surv, (size_t)nu, cudaMemcpyDeviceToHost, s_out));
                        d2h += (size_t)nu * 13;
                    }
                    B200_BATCH(cudaEventRecord(g_batch.cnt_done[u % NB], s_out));
                }
            }
        }
        // ---- volume `sb - HB_LAG_A`: its group count is known -> size and enqueue the download of the label data --------
        if (sb >= lag_a && sb - lag_a < n_volumes) {
            const int v = sb - lag_a, k = v % NB;
            Slot& s = slot[k];
            { const auto t = now(); B200_BATCH(cudaEventSynchronize(g_batch.cnt_done[k])); w_cnt += ms_since(t); }
            uint32_t ng = 0xFFFFFFFFu;
            if (sparse) { memcpy(&ng, g_batch.pinned + small_off[v] + 4, 4); if (ng > cap) ng = 0xFFFFFFFFu; }
            n_groups[v] = ng;
            if (ng == 0xFFFFFFFFu) {                               // dense: the DMA must not race the zero fill
                { const auto t = now(); while (sh->zero_left[v].load(std::memory_order_acquire) > 0) sched_yield(); w_zero += ms_since(t); }
                B200_BATCH(cudaMemcpyAsync(seg[v], s.seg, V * 2, cudaMemcpyDeviceToHost, s_out2));
                d2h += V * 2;
            } else if (ng > 0) {
                { const auto t = now(); while (sh->slot_busy[k].load(std::memory_order_acquire)) sched_yield(); w_slot += ms_since(t); }   // staging slot scattered by now
                char* stage = g_batch.pinned + small_bytes + stage_bytes * k;
                B200_BATCH(cudaMemcpyAsync(stage, s.gidx, (size_t)ng * 4, cudaMemcpyDeviceToHost, s_out2));
                if (!direct[v]) B200_BATCH(cudaMemcpyAsync(stage + align_up(cap * 4, 256), s.gval, (size_t)ng * 64, cudaMemcpyDeviceToHost, s_out2));
                d2h += (size_t)ng * 68;                            // (direct: the 64-byte payloads crossed the link as the kernel's own writes)
            }
            if (direct[v]) B200_BATCH(cudaStreamWaitEvent(s_out2, g_batch.scat_done[k], 0));   // the slot's line buffers are free once the scatter kernel is done
            B200_BATCH(cudaEventRecord(g_batch.out_done[k], s_out2));   // (the kernels of this volume finished before cnt_done)
            // traffic of the gathered PRM crops: the crops of the survivors (known from the visit order now on the host)
            const int n = n_dets[v];
            if (n > 0 && !prm_packed && (mode & 2) && mapped_device_pointer(prm[v])) {
                int32_t kc = 0;
                memcpy(&kc, g_batch.pinned + small_off[v], 4);
                const int32_t* ro = (const int32_t*)(g_batch.pinned + small_off[v] + 16);
                const uint8_t* pkz = pack_mode ? (const uint8_t*)(pack_base + pack_bytes * k + nmsst_bytes + 2 * align_up(nn * 8, 256)) : nullptr;
                for (int r = 0; r < kc && r < n; ++r)
                    if (!pkz || !pkz[r]) h2d += (unsigned long long)(crop_off[v][ro[r] + 1] - crop_off[v][ro[r]]);
            }
        }
        // ---- volume `sb - HB_LAG_B`: its download has been enqueued one step ago -> wait for it, hand it to the pool ----
        if (sb >= lag_b && sb - lag_b < n_volumes) {
            const int v = sb - lag_b, k = v % NB;
            const uint32_t ng = n_groups[v];
            if (ng == 0 && out_state == 2) {                                    // all-zero result: nothing to clear next time
                std::lock_guard<std::mutex> lk(g_prev_mu);
                PrevLabels pl; pl.bytes = V * 2;
                g_prev[seg[v]] = std::move(pl);
            }
            if (ng != 0xFFFFFFFFu && ng > 0) {
                { const auto t = now(); B200_BATCH(cudaEventSynchronize(g_batch.out_done[k])); w_out += ms_since(t); }
                const char* stage = g_batch.pinned + small_bytes + stage_bytes * k;
                const uint32_t* gi = (const uint32_t*)stage;
                const char* gv = stage + align_up(cap * 4, 256);
                char* dst = (char*)seg[v];
                sh->slot_busy[k].store(1, std::memory_order_release);
                sh->pending.fetch_add(1);
                const size_t vbytes = V * 2;
                const bool in_place = direct[v] != nullptr;
                g_pool.push([sh, v, k, gi, gv, dst, ng, out_state, vbytes, in_place] {
                    while (sh->zero_left[v].load(std::memory_order_acquire) > 0) sched_yield();
                    if (!in_place) scatter_lines(dst, vbytes, gi, gv, ng);
                    if (out_state == 2) {                                       // remember what has to be cleared next time
                        PrevLabels pl;
                        pl.bytes = vbytes;
                        pl.groups.assign(gi, gi + ng);
                        std::lock_guard<std::mutex> lk(g_prev_mu);
                        g_prev[dst] = std::move(pl);
                    }
                    sh->slot_busy[k].store(0, std::memory_order_release);
                    sh->pending.fetch_sub(1, std::memory_order_release);
                });
            }
        }
    }
done:
#undef B200_BATCH
    {
        // drain all three streams even on error: the slots, `offs` and the staging must not be reused while copies are in flight
        cudaError_t e2 = cudaStreamSynchronize(comp_streams[0]);
        for (int q = 1; q < 3; ++q) { const cudaError_t eq = cudaStreamSynchronize(comp_streams[q]); if (e2 == cudaSuccess) e2 = eq; }
        { const cudaError_t eq = cudaStreamSynchronize(g_batch.nms); if (e2 == cudaSuccess) e2 = eq; }
        { const cudaError_t eq = cudaStreamSynchronize(g_batch.out3); if (e2 == cudaSuccess) e2 = eq; }
        { const cudaError_t eq = cudaStreamSynchronize(g_batch.out4); if (e2 == cudaSuccess) e2 = eq; }
        const cudaError_t e1 = cudaStreamSynchronize(s_in), e3 = cudaStreamSynchronize(s_out),
                          e4 = cudaStreamSynchronize(s_out2);
        const auto t_tail = now();
        for (int k = 0; k < HB_SLOTS; ++k) while (pack_left[k].load(std::memory_order_acquire) > 0) sched_yield();   // (only after an error)
        while (sh->pending.load(std::memory_order_acquire) > 0) sched_yield();
        w_tail = ms_since(t_tail);
        delete sh;
        if (rc != 0 || e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess) {
            std::lock_guard<std::mutex> lk(g_prev_mu);                         // buffer contents are undefined after an error
            for (int v = 0; v < n_volumes; ++v) g_prev.erase(seg[v]);
        }
        if (rc == 0) {
            if (e1 != cudaSuccess) rc = check_cuda(e1, "cudaStreamSynchronize(in)");
            else if (e2 != cudaSuccess) rc = check_cuda(e2, "cudaStreamSynchronize(compute)");
            else if (e3 != cudaSuccess) rc = check_cuda(e3, "cudaStreamSynchronize(out)");
            else if (e4 != cudaSuccess) rc = check_cuda(e4, "cudaStreamSynchronize(out2)");
        }
    }
    if (rc == 0) {                                          // hand the staged small outputs to the caller
        for (int v = 0; v < n_volumes; ++v) {
            const size_t n = (size_t)n_dets[v];
            const char* st = g_batch.pinned + small_off[v];
            memcpy(&n_keep[v], st, 4);
            if (n > 0) {
                memcpy(rank_order[v], st + 16, n * 4);
                memcpy(b_max[v], st + 16 + n * 4, n * 4);
                memcpy(status[v], st + 16 + n * 8, n * 4);
                memcpy(survive[v], st + 16 + n * 12, n);
            }
        }
        g_last_h2d.store(h2d);
        g_last_d2h.store(d2h);
        if (trace)
            fprintf(stderr, "[b200seg host_batch] %d volumes, %d pool threads: %.2f ms; host waits: group count %.2f, staging slot %.2f, "
                            "download %.2f, zero fill (dense) %.2f, pool tail %.2f, visit order %.2f, packing %.2f ms; up %.1f MB, down %.1f MB\n",
                    n_volumes, g_pool.n_threads, ms_since(t_start), w_cnt, w_slot, w_out, w_zero, w_tail, w_nms, w_pack, h2d / 1e6, d2h / 1e6);
    }
    return rc;
}
