// nms3d.cu -- batched 3D greedy NMS on the GPU (replaces lib/utils/cython_nms_3d.pyx:39-159).
//
// Three launches per batch of detection sets (one for sets of at most 64 boxes, nms_small_kernel), no host round trip:
//   1. nms_rank_kernel   rank sort: rank(i) = #{j : key_j > key_i or (key_j == key_i and j > i)}  (equal keys: higher index first =
//                        what the reference's `argsort()[::-1]` yields wherever numpy's sort is stable)
//                        (O(n^2) compares, trivially parallel, deterministic, stable tie rule) and
//                        scatter of {box, volume, original index} into visit order.
//   2. nms_mask_kernel   upper-triangular 64x64 tiles of the suppression relation; one warp owns a
//                        32-row strip, lane = column box, one __ballot_sync per row builds the 32-bit
//                        word (warp-ballot bitmask).  IoU uses the exact fp32 operation order of the
//                        reference (cython_nms_3d.pyx:82-93) through __f*_rn intrinsics (never
//                        contracted into FMA).
//   3. nms_reduce_kernel one CTA per set walks the 64-row chunks: warp 0 resolves the chunk's
//                        diagonal tile in registers (ffs over the alive word), then all threads OR the
//                        kept rows' mask words into the removed bitset held in shared memory.  Emits the
//                        kept ORIGINAL indices ascending (np.where(suppressed==0)) and, optionally, in
//                        visit order.
#include "common.cuh"

namespace b200seg {

struct __align__(16) SortedBox {
    float x1, y1, z1, x2;
    float y2, z2, vol;
    int orig;
};

__device__ __forceinline__ float box_volume_f32(const float* d) {
    // numpy fp32: (x2 - x1 + 1) * (y2 - y1 + 1) * (z2 - z1 + 1)      cython_nms_3d.pyx:48
    float a = __fadd_rn(__fsub_rn(d[3], d[0]), 1.0f);
    float b = __fadd_rn(__fsub_rn(d[4], d[1]), 1.0f);
    float c = __fadd_rn(__fsub_rn(d[5], d[2]), 1.0f);
    return __fmul_rn(__fmul_rn(a, b), c);
}

__device__ __forceinline__ float ref_max(float a, float b) { return a >= b ? a : b; }   // pyx:30-31
__device__ __forceinline__ float ref_min(float a, float b) { return a <= b ? a : b; }   // pyx:33-34

// true when box j must be suppressed by box i (cython_nms_3d.pyx:82-95)
__device__ __forceinline__ bool suppresses(const SortedBox& bi, const SortedBox& bj, float thresh) {
    float xx1 = ref_max(bi.x1, bj.x1);
    float yy1 = ref_max(bi.y1, bj.y1);
    float zz1 = ref_max(bi.z1, bj.z1);
    float xx2 = ref_min(bi.x2, bj.x2);
    float yy2 = ref_min(bi.y2, bj.y2);
    float zz2 = ref_min(bi.z2, bj.z2);
    float w = ref_max(0.0f, __fadd_rn(__fsub_rn(xx2, xx1), 1.0f));
    float h = ref_max(0.0f, __fadd_rn(__fsub_rn(yy2, yy1), 1.0f));
    float s = ref_max(0.0f, __fadd_rn(__fsub_rn(zz2, zz1), 1.0f));
    float inter = __fmul_rn(__fmul_rn(w, h), s);
    float uni = __fsub_rn(__fadd_rn(bi.vol, bj.vol), inter);
    float ovr = __fdiv_rn(inter, uni);
    return ovr >= thresh;
}

constexpr int RANK_THREADS = 128;

__global__ void __launch_bounds__(RANK_THREADS)
nms_rank_kernel(const float* __restrict__ dets, const int32_t* __restrict__ offsets,
                int by_volume, int n_max, SortedBox* __restrict__ sorted_all) {
    const int b = blockIdx.y;
    const int off = offsets[b];
    const int n = offsets[b + 1] - off;
    const int i0 = blockIdx.x * RANK_THREADS;
    if (i0 >= n) return;
    const float* d = dets + (size_t)off * 7;
    SortedBox* sorted = sorted_all + (size_t)b * n_max;

    __shared__ uint32_t s_key[RANK_THREADS];
    const int i = i0 + threadIdx.x;
    float my[7];
    uint32_t ki = 0;
    float vi = 0.f;
    if (i < n) {
#pragma unroll
        for (int k = 0; k < 7; ++k) my[k] = d[(size_t)i * 7 + k];
        vi = box_volume_f32(my);
        ki = ordered_key(by_volume ? vi : my[6]);
    }
    int rank = 0;
    for (int j0 = 0; j0 < n; j0 += RANK_THREADS) {
        const int j = j0 + threadIdx.x;
        uint32_t kj = 0;
        if (j < n) {
            if (by_volume) {
                float t[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) t[k] = d[(size_t)j * 7 + k];
                kj = ordered_key(box_volume_f32(t));
            } else {
                kj = ordered_key(d[(size_t)j * 7 + 6]);
            }
        }
        __syncthreads();
        s_key[threadIdx.x] = kj;
        __syncthreads();
        const int lim = min(RANK_THREADS, n - j0);
        if (i < n) {
            for (int t = 0; t < lim; ++t) {
                const uint32_t k = s_key[t];
                rank += (k > ki) || (k == ki && (j0 + t) > i);
            }
        }
    }
    if (i < n) {
        SortedBox sb;
        sb.x1 = my[0]; sb.y1 = my[1]; sb.z1 = my[2]; sb.x2 = my[3]; sb.y2 = my[4]; sb.z2 = my[5];
        sb.vol = vi; sb.orig = i;
        sorted[rank] = sb;
    }
}

// grid (W64, W64, batch), 64 threads: warp w owns rows rb*64 + 32*w .. +31, both warps share the
// 64 column boxes staged in shared memory.  mask[row][cb] (uint64) bit c set <=> row suppresses
// column cb*64+c (only columns later in visit order).
__global__ void __launch_bounds__(64)
nms_mask_kernel(const SortedBox* __restrict__ sorted_all, const int32_t* __restrict__ offsets,
                int n_max, int w64, float thresh, unsigned long long* __restrict__ mask_all) {
    const int b = blockIdx.z;
    const int n = offsets[b + 1] - offsets[b];
    const int rb = blockIdx.y, cb = blockIdx.x;
    if (cb < rb || rb * 64 >= n || cb * 64 >= n) return;
    const SortedBox* sorted = sorted_all + (size_t)b * n_max;
    unsigned long long* mask = mask_all + (size_t)b * n_max * w64;

    __shared__ SortedBox s_row[64];
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const int col = cb * 64 + t;
    const int row = rb * 64 + t;
    SortedBox cbox0, cbox1;       // lane's two column boxes: columns cb*64+lane and cb*64+32+lane
    {
        SortedBox z; z.x1 = z.y1 = z.z1 = 0.f; z.x2 = z.y2 = z.z2 = -1.f; z.vol = 0.f; z.orig = -1;
        s_row[t] = row < n ? sorted[row] : z;
        const int c0 = cb * 64 + lane, c1 = c0 + 32;
        cbox0 = c0 < n ? sorted[c0] : z;
        cbox1 = c1 < n ? sorted[c1] : z;
        (void)col;
    }
    __syncthreads();
    const int c0 = cb * 64 + lane, c1 = c0 + 32;
    unsigned long long my_word = 0ull;
    // each warp walks its 32 rows; the row box is a shared-memory broadcast, the predicate is
    // evaluated by 32 lanes at once and gathered with one ballot per 32 columns
    for (int r = 0; r < 32; ++r) {
        const int rl = warp * 32 + r;
        const int rg = rb * 64 + rl;
        if (rg >= n) break;                                   // warp-uniform
        const SortedBox rbox = s_row[rl];
        const bool p0 = (c0 < n) && (c0 > rg) && suppresses(rbox, cbox0, thresh);
        const bool p1 = (c1 < n) && (c1 > rg) && suppresses(rbox, cbox1, thresh);
        const unsigned lo = __ballot_sync(0xffffffffu, p0);
        const unsigned hi = __ballot_sync(0xffffffffu, p1);
        if (lane == r) my_word = ((unsigned long long)hi << 32) | lo;
    }
    const int my_row = rb * 64 + warp * 32 + lane;
    if (my_row < n) mask[(size_t)my_row * w64 + cb] = my_word;
}

constexpr int REDUCE_THREADS = 256;
constexpr size_t NMS_PRELOAD_BYTES = 160 * 1024;     // whole suppression mask staged in shared memory when it fits

// One CTA per detection set.  dynamic smem: removed[w64] u64 | flags[w64] u64 | mask copy (PRELOAD) or diag words.
// PRELOAD: the set's whole bitmask (n x w64 words) is copied to shared memory first with coalesced loads, so the
// 64-row chunk walk never waits on L2.  Otherwise only the diagonal words are staged and the kept rows' words are
// read from global memory chunk by chunk.
// Per chunk, warp 0 resolves the diagonal tile: only ALIVE rows with a NON-ZERO diagonal word form the serial
// chain (a row with an empty diagonal word suppresses nobody inside the chunk), everything else is bit arithmetic.
template <bool PRELOAD>
__global__ void __launch_bounds__(REDUCE_THREADS)
nms_reduce_kernel(const SortedBox* __restrict__ sorted_all, const unsigned long long* __restrict__ mask_all,
                  const int32_t* __restrict__ offsets, int n_max, int w64,
                  int64_t* __restrict__ keep, int32_t* __restrict__ keep_count,
                  int32_t* __restrict__ rank_order) {
    extern __shared__ unsigned long long s_dyn[];
    unsigned long long* removed = s_dyn;
    unsigned long long* flags = s_dyn + w64;
    unsigned long long* s_mask = s_dyn + 2 * w64;         // PRELOAD: [n][w64]; else diag[n]
    __shared__ int s_rows[64];
    __shared__ int s_nk;
    __shared__ int s_scan[REDUCE_THREADS];

    const int b = blockIdx.x;
    const int off = offsets[b];
    const int n = offsets[b + 1] - off;
    const SortedBox* sorted = sorted_all + (size_t)b * n_max;
    const unsigned long long* mask = mask_all + (size_t)b * n_max * w64;
    const int nw = (n + 63) >> 6;
    const int tid = threadIdx.x;

    for (int w = tid; w < nw; w += REDUCE_THREADS) { removed[w] = 0ull; flags[w] = 0ull; }
    int* s_orig = reinterpret_cast<int*>(s_mask + (PRELOAD ? (size_t)n_max * w64 : (size_t)n_max));   // original index per row
    for (int r = tid; r < n; r += REDUCE_THREADS) s_orig[r] = sorted[r].orig;
    if (PRELOAD) {
        // stage the set's whole bitmask with flat coalesced loads, 8 independent words in flight per thread.
        // Words left of the diagonal were never written by nms_mask_kernel and are never read here either.
        const int total = n * w64;
        int i = tid;
        for (; i + 7 * REDUCE_THREADS < total; i += 8 * REDUCE_THREADS) {
            unsigned long long v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = mask[i + j * REDUCE_THREADS];
#pragma unroll
            for (int j = 0; j < 8; ++j) s_mask[i + j * REDUCE_THREADS] = v[j];
        }
        for (; i < total; i += REDUCE_THREADS) s_mask[i] = mask[i];
    } else {
        for (int r = tid; r < n; r += REDUCE_THREADS) s_mask[r] = mask[(size_t)r * w64 + (r >> 6)];
    }
    __syncthreads();

    int n_kept = 0;    // uniform across the CTA
    for (int c = 0; c < nw; ++c) {
        if (tid < 32) {
            const int r0 = c * 64 + tid, r1 = r0 + 32;
            const unsigned long long d0 = r0 < n ? (PRELOAD ? s_mask[(size_t)r0 * w64 + c] : s_mask[r0]) : 0ull;
            const unsigned long long d1 = r1 < n ? (PRELOAD ? s_mask[(size_t)r1 * w64 + c] : s_mask[r1]) : 0ull;
            const unsigned long long nz = (unsigned long long)__ballot_sync(0xffffffffu, d0 != 0ull) |
                                          ((unsigned long long)__ballot_sync(0xffffffffu, d1 != 0ull) << 32);
            const int rows = min(64, n - c * 64);
            unsigned long long alive = ~removed[c];
            if (rows < 64) alive &= ((1ull << rows) - 1ull);
            unsigned long long cand = alive & nz;
            while (cand) {                                     // warp-uniform serial chain: suppressing rows only
                const int r = __ffsll((long long)cand) - 1;
                const unsigned long long lo = __shfl_sync(0xffffffffu, d0, r & 31);
                const unsigned long long hi = __shfl_sync(0xffffffffu, d1, r & 31);
                alive &= ~((r < 32) ? lo : hi);
                cand = alive & nz & ~((2ull << r) - 1ull);     // candidates strictly after r (2<<63 wraps to 0: all cleared)
                if (r == 63) cand = 0ull;
            }
            // kept rows of this chunk, in order: lane l owns rows l and l+32, position = popcount of lower kept bits
            if (tid == 0) s_nk = __popcll(alive);
            if ((alive >> tid) & 1ull) s_rows[__popcll(alive & ((1ull << tid) - 1ull))] = c * 64 + tid;
            if ((alive >> (tid + 32)) & 1ull) s_rows[__popcll(alive & ((1ull << (tid + 32)) - 1ull))] = c * 64 + tid + 32;
        }
        __syncthreads();
        const int nk = s_nk;
        const int rest = nw - c - 1;
        for (int item = tid; item < nk * rest; item += REDUCE_THREADS) {
            const int k = item / rest, w = c + 1 + item % rest;
            const size_t idx = (size_t)s_rows[k] * w64 + w;
            const unsigned long long m = PRELOAD ? s_mask[idx] : mask[idx];
            if (m) atomicOr(&removed[w], m);
        }
        for (int k = tid; k < nk; k += REDUCE_THREADS) {
            const int orig = s_orig[s_rows[k]];
            atomicOr(&flags[orig >> 6], 1ull << (orig & 63));
            if (rank_order) rank_order[off + n_kept + k] = orig;
        }
        n_kept += nk;
        __syncthreads();
    }
    if (tid == 0) keep_count[b] = n_kept;

    // ascending original indices: exclusive scan of popcounts over the flag words
    int base = 0;
    for (int w0 = 0; w0 < nw; w0 += REDUCE_THREADS) {
        const int w = w0 + tid;
        const unsigned long long f = w < nw ? flags[w] : 0ull;
        const int cnt = __popcll(f);
        s_scan[tid] = cnt;
        __syncthreads();
        for (int d = 1; d < REDUCE_THREADS; d <<= 1) {         // Hillis-Steele inclusive scan
            const int v = tid >= d ? s_scan[tid - d] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        int pos = base + s_scan[tid] - cnt;
        unsigned long long g = f;
        while (g) {
            const int bit = __ffsll((long long)g) - 1;
            g &= g - 1;
            keep[off + pos++] = (int64_t)(w * 64 + bit);
        }
        base += s_scan[REDUCE_THREADS - 1];
        __syncthreads();
    }
}

// Sets of at most 64 boxes (the per-class NMS of box_results, config 1's 50 boxes): the three steps in ONE launch, one
// 64-thread CTA per set -- rank by counting, scatter into visit order in shared memory, one 64-bit suppression word per row,
// and a 64-step serial resolve by thread 0.  Same arithmetic, tie rule and outputs as the three-kernel path.
__global__ void __launch_bounds__(64)
nms_small_kernel(const float* __restrict__ dets, const int32_t* __restrict__ offsets, int by_volume, float thresh,
                 int64_t* __restrict__ keep, int32_t* __restrict__ keep_count, int32_t* __restrict__ rank_order) {
    __shared__ uint32_t s_key[64];
    __shared__ SortedBox s_box[64];
    __shared__ unsigned long long s_mask[64];
    const int b = blockIdx.x, t = threadIdx.x;
    const int off = offsets[b];
    const int n = min(offsets[b + 1] - off, 64);            // n_max <= 64 is the caller's promise
    float my[7];
    uint32_t ki = 0;
    float vi = 0.f;
    if (t < n) {
#pragma unroll
        for (int k = 0; k < 7; ++k) my[k] = dets[((size_t)off + t) * 7 + k];
        vi = box_volume_f32(my);
        ki = ordered_key(by_volume ? vi : my[6]);
    }
    s_key[t] = ki;
    __syncthreads();
    if (t < n) {
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const uint32_t k = s_key[j];
            rank += (k > ki) || (k == ki && j > t);
        }
        SortedBox sb;
        sb.x1 = my[0]; sb.y1 = my[1]; sb.z1 = my[2]; sb.x2 = my[3]; sb.y2 = my[4]; sb.z2 = my[5];
        sb.vol = vi; sb.orig = t;
        s_box[rank] = sb;
    }
    __syncthreads();
    if (t < n) {
        const SortedBox rbox = s_box[t];
        unsigned long long word = 0ull;
        for (int c = t + 1; c < n; ++c)
            if (suppresses(rbox, s_box[c], thresh)) word |= 1ull << c;
        s_mask[t] = word;
    }
    __syncthreads();
    if (t == 0) {
        unsigned long long removed = 0ull, flags = 0ull;
        int nk = 0;
        for (int r = 0; r < n; ++r) {
            if ((removed >> r) & 1ull) continue;
            const int orig = s_box[r].orig;
            flags |= 1ull << orig;
            if (rank_order) rank_order[off + nk] = orig;
            ++nk;
            removed |= s_mask[r];
        }
        keep_count[b] = nk;
        int pos = 0;
        while (flags) {                                       // kept ORIGINAL indices, ascending (np.where(suppressed == 0))
            const int bit = __ffsll((long long)flags) - 1;
            flags &= flags - 1ull;
            keep[off + pos++] = (int64_t)bit;
        }
    }
}

}  // namespace b200seg

using namespace b200seg;

extern "C" size_t b200seg_nms3d_workspace_bytes(int batch, int n_max) {
    if (batch <= 0 || n_max <= 0) return 256;
    const size_t w64 = (size_t)(n_max + 63) / 64;
    size_t per = align_up((size_t)n_max * sizeof(SortedBox), 256) + align_up((size_t)n_max * w64 * 8, 256);
    return per * batch + 256;
}

extern "C" int b200seg_nms3d_dev(const float* dets, const int32_t* offsets, int batch, int n_max,
                                 float thresh, int by_volume, int64_t* keep, int32_t* keep_count,
                                 int32_t* rank_order, void* workspace, size_t workspace_bytes,
                                 b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(batch >= 0 && n_max >= 0, "nms3d: negative batch/n_max");
    if (batch == 0) return 0;
    B200_CHECK_ARG(offsets && keep_count, "nms3d: null offsets/keep_count");
    if (n_max == 0) {
        B200_CUDA(cudaMemsetAsync(keep_count, 0, sizeof(int32_t) * batch, stream));
        return 0;
    }
    B200_CHECK_ARG(dets && keep && workspace, "nms3d: null pointer");
    if (workspace_bytes < b200seg_nms3d_workspace_bytes(batch, n_max)) {
        set_error("nms3d: workspace too small (%zu < %zu)", workspace_bytes,
                  b200seg_nms3d_workspace_bytes(batch, n_max));
        return B200SEG_EWORKSPACE;
    }
    if (n_max <= 64) {                                        // small sets: everything in one launch
        nms_small_kernel<<<batch, 64, 0, stream>>>(dets, offsets, by_volume, thresh, keep, keep_count, rank_order);
        B200_LAUNCH_CHECK("nms_small_kernel");
        return 0;
    }
    const int w64 = (n_max + 63) / 64;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    SortedBox* sorted = (SortedBox*)ws;
    unsigned long long* mask = (unsigned long long*)(ws + align_up((size_t)n_max * sizeof(SortedBox), 256) * batch);

    dim3 g1((n_max + RANK_THREADS - 1) / RANK_THREADS, batch);
    nms_rank_kernel<<<g1, RANK_THREADS, 0, stream>>>(dets, offsets, by_volume, n_max, sorted);
    B200_LAUNCH_CHECK("nms_rank_kernel");
    dim3 g2(w64, w64, batch);
    nms_mask_kernel<<<g2, 64, 0, stream>>>(sorted, offsets, n_max, w64, thresh, mask);
    B200_LAUNCH_CHECK("nms_mask_kernel");
    const size_t pre_bytes = (size_t)n_max * w64 * 8;
    const bool preload = pre_bytes <= NMS_PRELOAD_BYTES;
    const size_t smem = 2 * (size_t)w64 * 8 + (preload ? pre_bytes : (size_t)n_max * 8) + (size_t)n_max * 4;
    B200_CHECK_ARG(smem <= 220 * 1024, "nms3d: n_max=%d too large for the reduce kernel", n_max);
    if (preload) {
        if (smem > 48 * 1024)
            B200_CUDA(cudaFuncSetAttribute(nms_reduce_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        nms_reduce_kernel<true><<<batch, REDUCE_THREADS, smem, stream>>>(sorted, mask, offsets, n_max, w64, keep, keep_count, rank_order);
    } else {
        if (smem > 48 * 1024)
            B200_CUDA(cudaFuncSetAttribute(nms_reduce_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        nms_reduce_kernel<false><<<batch, REDUCE_THREADS, smem, stream>>>(sorted, mask, offsets, n_max, w64, keep, keep_count, rank_order);
    }
    B200_LAUNCH_CHECK("nms_reduce_kernel");
    return 0;
}
