// rle3d.cu -- run-length codec of binary 3D masks (replaces lib/utils/mask_3d.py:15-71 and
// lib/utils/cython_mask_3d.pyx:19-77; first piece of the "eval kernels + RLE codec" next row).
//
// Reference format: ravel the mask in FORTRAN order (first axis fastest), `counts` = lengths of the alternating
// runs starting with a run of zeros (of length 0 when the first element is set); an all-zero mask is the single
// count [size].  For a C-contiguous [S,H,W] mask the Fortran index is f = s + S*(h + H*w): the S elements of a
// column (h,w) are consecutive in f, columns follow each other with h fastest, then w.
//
// encode:  rle_count_kernel    one thread per column walks its S elements (adjacent threads = adjacent w, so every
//                              step is a coalesced row read) and counts the run starts inside the column, including
//                              the one at its first element (compared with the last element of the previous column);
//          rle_scan_kernels    exclusive prefix sum of the per-column counts in column order (two levels);
//          rle_emit_kernel     same walk, writes the Fortran index of every run start in order;
//          rle_counts_kernel   counts[j] = start[j+1] - start[j] (+ the leading 0 when the mask starts with a 1).
// decode:  rle_cum_kernels     inclusive prefix sum of the counts; rle_fill_kernel: one thread per column finds the run
//                              holding its first element by binary search and walks from there; writes are coalesced
//                              across the warp like the reads of the encoder.
#include "common.cuh"

namespace b200seg {

constexpr int RLE_THREADS = 256;
constexpr int RLE_SCAN = 1024;            // elements per block in the first scan level

// column c (Fortran order: c = w * H + h) -> its (h, w)
__device__ __forceinline__ void rle_col(int c, int H, int& h, int& w) { w = c / H; h = c - w * H; }

// MODE 0: count run starts per column; MODE 1: emit their Fortran indices at offs[c]
template <int MODE>
__global__ void __launch_bounds__(RLE_THREADS)
rle_walk_kernel(const uint8_t* __restrict__ mask, int S, int H, int W, int32_t* __restrict__ cnt,
                const int64_t* __restrict__ offs, int64_t* __restrict__ starts, long long cap) {
    // thread -> (h, w) with w fastest across the warp (coalesced reads); the column index is w * H + h
    const long long t = (long long)blockIdx.x * RLE_THREADS + threadIdx.x;
    if (t >= (long long)H * W) return;
    const int h = (int)(t / W), w = (int)(t - (long long)h * W);
    const int c = w * H + h;
    const size_t HW = (size_t)H * W;
    const uint8_t* p = mask + (size_t)h * W + w;
    // value just before this column in Fortran order: last element of column (h-1, w), or of (H-1, w-1); none for c == 0
    int prev = -1;
    if (c > 0) {
        const int ph = h > 0 ? h - 1 : H - 1, pw = h > 0 ? w : w - 1;
        prev = mask[(size_t)(S - 1) * HW + (size_t)ph * W + pw] != 0;
    }
    int n = 0;
    long long pos = MODE == 1 ? (long long)offs[c] : 0;
    const long long f0 = (long long)c * S;
    // eight independent loads in flight per thread (the walk is latency bound: one byte per step and column)
    for (int s0 = 0; s0 < S; s0 += 8) {
        uint8_t buf[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) buf[k] = s0 + k < S ? p[(size_t)(s0 + k) * HW] : (uint8_t)0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (s0 + k >= S) break;
            const int v = buf[k] != 0;
            if (prev >= 0 && v != prev) {
                if (MODE == 1) { if (pos < cap) starts[pos] = f0 + s0 + k; ++pos; }
                ++n;
            }
            prev = v;
        }
    }
    if (MODE == 0) cnt[c] = n;
}

// level 1: exclusive scan inside blocks of RLE_SCAN elements (int32 in, int64 out) + block totals
template <typename TIN>
__global__ void __launch_bounds__(RLE_THREADS)
rle_scan1_kernel(const TIN* __restrict__ in, long long n, int64_t* __restrict__ out, int64_t* __restrict__ block_sum, int inclusive) {
    __shared__ long long s_w[RLE_THREADS / 32];
    constexpr int PER = RLE_SCAN / RLE_THREADS;
    const long long base = (long long)blockIdx.x * RLE_SCAN + (long long)threadIdx.x * PER;
    long long v[PER], sum = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) { v[i] = base + i < n ? (long long)in[base + i] : 0; sum += v[i]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const long long a = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += a; }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    long long before = 0, total = 0;
#pragma unroll
    for (int k = 0; k < RLE_THREADS / 32; ++k) { if (k < warp) before += s_w[k]; total += s_w[k]; }
    long long run = before + incl - sum;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        if (base + i < n) out[base + i] = inclusive ? run + v[i] : run;
        run += v[i];
    }
    if (threadIdx.x == 0) block_sum[blockIdx.x] = total;
}
// level 2: one block scans the block totals in place (exclusive) and writes the grand total
__global__ void __launch_bounds__(RLE_THREADS)
rle_scan2_kernel(int64_t* __restrict__ block_sum, int nb, int64_t* __restrict__ total) {
    __shared__ long long s_carry;
    __shared__ long long s_w[RLE_THREADS / 32];
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b0 = 0; b0 < nb; b0 += RLE_THREADS) {
        const int b = b0 + threadIdx.x;
        const long long v = b < nb ? block_sum[b] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long a = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += a; }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        long long before = 0, tot = 0;
#pragma unroll
        for (int k = 0; k < RLE_THREADS / 32; ++k) { if (k < warp) before += s_w[k]; tot += s_w[k]; }
        if (b < nb) block_sum[b] = s_carry + before + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}
// level 3: add the block offsets
__global__ void __launch_bounds__(RLE_THREADS)
rle_scan3_kernel(int64_t* __restrict__ out, long long n, const int64_t* __restrict__ block_sum) {
    const long long i = (long long)blockIdx.x * RLE_THREADS + threadIdx.x;
    if (i < n) out[i] += block_sum[i / RLE_SCAN];
}

// counts from the run starts: runs = changes + 1 (the run that begins at f = 0).
__global__ void __launch_bounds__(RLE_THREADS)
rle_counts_kernel(const uint8_t* __restrict__ mask, const int64_t* __restrict__ starts, const int64_t* __restrict__ n_changes,
                  long long N, int64_t* __restrict__ counts, long long cap, int64_t* __restrict__ n_counts) {
    const long long R = *n_changes + 1;                       // number of runs
    const int lead = mask[0] != 0 ? 1 : 0;                    // first element set: a zero-length run of zeros comes first
    const long long j = (long long)blockIdx.x * RLE_THREADS + threadIdx.x;
    if (j == 0) { *n_counts = R + lead; if (lead && cap > 0) counts[0] = 0; }
    if (j < R) {
        const long long a = j == 0 ? 0 : starts[j - 1];
        const long long b = j + 1 < R ? starts[j] : N;        // starts[] holds the R-1 change positions
        if (j + lead < cap) counts[j + lead] = b - a;
    }
}

// decode: cum[j] = counts[0] + ... + counts[j]; column thread finds its first run by binary search
__global__ void __launch_bounds__(RLE_THREADS)
rle_fill_kernel(const int64_t* __restrict__ cum, long long n_counts, uint8_t* __restrict__ mask, int S, int H, int W) {
    const long long t = (long long)blockIdx.x * RLE_THREADS + threadIdx.x;
    if (t >= (long long)H * W) return;
    const int h = (int)(t / W), w = (int)(t - (long long)h * W);
    const size_t HW = (size_t)H * W;
    uint8_t* p = mask + (size_t)h * W + w;
    const long long f0 = ((long long)w * H + h) * S;
    // first run j with cum[j] > f0
    long long lo = 0, hi = n_counts;
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (cum[mid] > f0) hi = mid; else lo = mid + 1; }
    long long j = lo;
    long long end = j < n_counts ? cum[j] : (long long)0x7fffffffffffffffll;
    for (int s = 0; s < S; ++s) {
        const long long f = f0 + s;
        while (f >= end) { ++j; end = j < n_counts ? cum[j] : (long long)0x7fffffffffffffffll; }
        p[(size_t)s * HW] = (uint8_t)(n_counts == 1 ? 0 : (j & 1));      // runs alternate 0,1,0,...; a single count = blank mask
    }
}

static int rle_scan(const void* in, bool in_is_i32, long long n, int64_t* out, int64_t* block_sum, int64_t* total,
                    int inclusive, cudaStream_t stream) {
    const int nb = (int)((n + RLE_SCAN - 1) / RLE_SCAN);
    if (in_is_i32) rle_scan1_kernel<int32_t><<<nb, RLE_THREADS, 0, stream>>>((const int32_t*)in, n, out, block_sum, inclusive);
    else rle_scan1_kernel<int64_t><<<nb, RLE_THREADS, 0, stream>>>((const int64_t*)in, n, out, block_sum, inclusive);
    B200_LAUNCH_CHECK("rle_scan1_kernel");
    rle_scan2_kernel<<<1, RLE_THREADS, 0, stream>>>(block_sum, nb, total);
    B200_LAUNCH_CHECK("rle_scan2_kernel");
    rle_scan3_kernel<<<(unsigned)((n + RLE_THREADS - 1) / RLE_THREADS), RLE_THREADS, 0, stream>>>(out, n, block_sum);
    B200_LAUNCH_CHECK("rle_scan3_kernel");
    return 0;
}

}  // namespace b200seg

using namespace b200seg;

extern "C" size_t b200seg_rle3d_workspace_bytes(int S, int H, int W, long long cap) {
    if (S <= 0 || H <= 0 || W <= 0 || cap < 0) return 256;
    const size_t cols = (size_t)H * W;
    const size_t m = cols > (size_t)cap ? cols : (size_t)cap;
    return align_up(cols * 4, 256) + align_up(cols * 8, 256) + align_up((m / RLE_SCAN + 2) * 8, 256) +
           align_up(((size_t)cap + 1) * 8, 256) + 512;
}

extern "C" int b200seg_rle3d_encode_dev(const uint8_t* mask, int S, int H, int W, int64_t* counts, long long cap,
                                        int64_t* n_counts, void* workspace, size_t workspace_bytes,
                                        b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(S > 0 && H > 0 && W > 0 && cap >= 1, "rle3d_encode: bad sizes");
    B200_CHECK_ARG(mask && counts && n_counts && workspace, "rle3d_encode: null pointer");
    B200_CHECK_ARG((long long)H * W < (1ll << 31), "rle3d_encode: too many columns");
    if (workspace_bytes < b200seg_rle3d_workspace_bytes(S, H, W, cap)) { set_error("rle3d_encode: workspace too small"); return B200SEG_EWORKSPACE; }
    const size_t cols = (size_t)H * W;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    int32_t* cnt = (int32_t*)ws; ws += align_up(cols * 4, 256);
    int64_t* offs = (int64_t*)ws; ws += align_up(cols * 8, 256);
    const size_t m = cols > (size_t)cap ? cols : (size_t)cap;
    int64_t* block_sum = (int64_t*)ws; ws += align_up((m / RLE_SCAN + 2) * 8, 256);
    int64_t* starts = (int64_t*)ws; ws += align_up(((size_t)cap + 1) * 8, 256);
    int64_t* total = (int64_t*)ws;
    const unsigned gb = (unsigned)((cols + RLE_THREADS - 1) / RLE_THREADS);
    rle_walk_kernel<0><<<gb, RLE_THREADS, 0, stream>>>(mask, S, H, W, cnt, nullptr, nullptr, 0);
    B200_LAUNCH_CHECK("rle_walk_kernel<0>");
    int e = rle_scan(cnt, true, (long long)cols, offs, block_sum, total, 0, stream);
    if (e) return e;
    rle_walk_kernel<1><<<gb, RLE_THREADS, 0, stream>>>(mask, S, H, W, nullptr, offs, starts, cap);
    B200_LAUNCH_CHECK("rle_walk_kernel<1>");
    // at most cap counts are written; *n_counts always reports the true number
    const long long max_runs = cap + 1;
    rle_counts_kernel<<<(unsigned)((max_runs + RLE_THREADS - 1) / RLE_THREADS), RLE_THREADS, 0, stream>>>(
        mask, starts, total, (long long)S * H * W, counts, cap, n_counts);
    B200_LAUNCH_CHECK("rle_counts_kernel");
    return 0;
}

extern "C" int b200seg_rle3d_decode_dev(const int64_t* counts, long long n_counts, uint8_t* mask, int S, int H, int W,
                                        int64_t* sum_out, void* workspace, size_t workspace_bytes, b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(S > 0 && H > 0 && W > 0 && n_counts >= 1, "rle3d_decode: bad sizes");
    B200_CHECK_ARG(counts && mask && workspace, "rle3d_decode: null pointer");
    if (workspace_bytes < b200seg_rle3d_workspace_bytes(S, H, W, n_counts)) { set_error("rle3d_decode: workspace too small"); return B200SEG_EWORKSPACE; }
    const size_t cols = (size_t)H * W;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    ws += align_up(cols * 4, 256); ws += align_up(cols * 8, 256);
    const size_t m = cols > (size_t)n_counts ? cols : (size_t)n_counts;
    int64_t* block_sum = (int64_t*)ws; ws += align_up((m / RLE_SCAN + 2) * 8, 256);
    int64_t* cum = (int64_t*)ws; ws += align_up(((size_t)n_counts + 1) * 8, 256);
    int64_t* total = (int64_t*)ws;
    int e = rle_scan(counts, false, n_counts, cum, block_sum, total, 1, stream);
    if (e) return e;
    if (sum_out) B200_CUDA(cudaMemcpyAsync(sum_out, total, 8, cudaMemcpyDeviceToDevice, stream));   // caller checks == S*H*W (mask_3d.py:53)
    rle_fill_kernel<<<(unsigned)((cols + RLE_THREADS - 1) / RLE_THREADS), RLE_THREADS, 0, stream>>>(cum, n_counts, mask, S, H, W);
    B200_LAUNCH_CHECK("rle_fill_kernel");
    return 0;
}
