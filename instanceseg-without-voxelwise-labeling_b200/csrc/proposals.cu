// proposals.cu -- RPN proposal generation on the device: replaces
// GenerateProposalsOp_3d.forward / proposals_for_one_image / _filter_boxes_3d
// (lib/modeling/generate_proposals_3d.py:20-192) together with bbox_transform_3d and clip_tiled_boxes_3d
// (lib/utils/boxes_3d.py:144-225), which the reference runs in numpy after copying the score and delta maps to the host.
//
// Per image:  top pre_nms_topN scores of the [A,S,H,W] map (exact radix select on the composite key
// (score, smaller (S,H,W,A) index first), 11 bits per pass, histograms in shared memory)  ->  rank the selected
// candidates (score descending)  ->  anchors + deltas -> boxes in the reference's float32 operation order, clip to the
// image, the reference's min-size / centre filter (x extent only, :186-192)  ->  order-preserving compaction  ->
// bitmask NMS (nms3d.cu)  ->  first post_nms_topN survivors, as rois [batch, x1, y1, z1, x2, y2, z2], their scores and
// their indices into the flattened (S,H,W,A) score map.  Nothing but the per-image counts ever has to leave the GPU.
//
// exp(): the reference calls numpy's float32 exp (<= 1 ulp, not always correctly rounded); the kernel rounds the fp64
// exp to float32.  The two agree except for a vanishing fraction of arguments, where the box edge differs by one ulp
// (parity tolerance 1e-6 relative on coordinates, exact on the selected indices).
#include "common.cuh"

namespace b200seg {

constexpr int GP_BINS = 2048, GP_DIGIT = 11, GP_MAXPASS = 6, GP_THREADS = 256, GP_AMAX = 128;

struct GpState { unsigned long long prefix; unsigned rank; unsigned pad; };
struct GpAnchors { float a[GP_AMAX * 6]; };

struct GpGeom {
    int A, S, H, W, SHW;
    int idx_bits, TB, npass;          // composite key = ordered_key(score) << idx_bits | (idx_mask - idx); TB = npass * 11 >= 32 + idx_bits
    unsigned idx_mask;
};

__device__ __forceinline__ unsigned long long gp_composite(float score, unsigned idx, const GpGeom& g) {
    return ((unsigned long long)ordered_key(score) << g.idx_bits) | (unsigned long long)(g.idx_mask - idx);
}

// Which digit holds the rank-th largest element, given the histogram of that digit over the elements that match the
// prefix found so far?  Whole CTA (GP_THREADS threads, 8 bins each, top bins first); result broadcast through `out`.
__device__ void gp_resolve(const unsigned* __restrict__ hist, GpState prev, GpState* out, unsigned* scratch) {
    const int t = threadIdx.x;
    unsigned h[8], local = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { h[k] = hist[GP_BINS - 1 - (t * 8 + k)]; local += h[k]; }
    scratch[t] = local;
    __syncthreads();
    for (int o = 1; o < GP_THREADS; o <<= 1) {            // inclusive scan over threads (descending bins)
        const unsigned v = t >= o ? scratch[t - o] : 0u;
        __syncthreads();
        scratch[t] += v;
        __syncthreads();
    }
    unsigned before = scratch[t] - local;
    if (before < prev.rank && prev.rank <= before + local) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (before < prev.rank && prev.rank <= before + h[k]) {
                out->prefix = (prev.prefix << GP_DIGIT) | (unsigned)(GP_BINS - 1 - (t * 8 + k));
                out->rank = prev.rank - before;
            }
            before += h[k];
        }
    }
    __syncthreads();
}

// pass p: histogram of digit p over the elements whose first p digits equal the prefix resolved so far
__global__ void __launch_bounds__(GP_THREADS) gp_hist_kernel(const float* __restrict__ scores, GpGeom g, int p, unsigned k_target,
                                                             unsigned* __restrict__ hist_all, GpState* __restrict__ states) {
    __shared__ unsigned sh[GP_BINS];
    __shared__ unsigned scratch[GP_THREADS];
    __shared__ GpState st;
    if (p == 0) {
        if (threadIdx.x == 0) { st.prefix = 0ull; st.rank = k_target; }
    } else {
        GpState prev;
        if (p == 1) { prev.prefix = 0ull; prev.rank = k_target; } else prev = states[p - 2];
        gp_resolve(hist_all + (size_t)(p - 1) * GP_BINS, prev, &st, scratch);
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) states[p - 1] = st;
    }
    for (int b = threadIdx.x; b < GP_BINS; b += GP_THREADS) sh[b] = 0u;
    __syncthreads();
    const unsigned long long prefix = st.prefix;
    const int sh_digit = g.TB - GP_DIGIT * (p + 1), sh_prefix = g.TB - GP_DIGIT * p;
    const int a = blockIdx.y;
    const float* sc = scores + (size_t)a * g.SHW;
    for (int cell = blockIdx.x * GP_THREADS + threadIdx.x; cell < g.SHW; cell += gridDim.x * GP_THREADS) {
        const unsigned long long c = gp_composite(sc[cell], (unsigned)cell * g.A + a, g);
        if (p == 0 || (c >> sh_prefix) == prefix) atomicAdd(&sh[(unsigned)(c >> sh_digit) & (GP_BINS - 1)], 1u);
    }
    __syncthreads();
    unsigned* hist = hist_all + (size_t)p * GP_BINS;
    for (int b = threadIdx.x; b < GP_BINS; b += GP_THREADS)
        if (sh[b]) atomicAdd(hist + b, sh[b]);
}

// candidates = every element whose composite key is >= the k-th largest (exactly k of them; all when take_all)
__global__ void __launch_bounds__(GP_THREADS) gp_collect_kernel(const float* __restrict__ scores, GpGeom g, int take_all, unsigned k_target,
                                                                const unsigned* __restrict__ hist_all, const GpState* __restrict__ states,
                                                                unsigned long long* __restrict__ cand, unsigned* __restrict__ counter) {
    __shared__ unsigned scratch[GP_THREADS];
    __shared__ GpState st;
    if (take_all) {
        if (threadIdx.x == 0) st.prefix = 0ull;
    } else {
        GpState prev;
        if (g.npass == 1) { prev.prefix = 0ull; prev.rank = k_target; } else prev = states[g.npass - 2];
        gp_resolve(hist_all + (size_t)(g.npass - 1) * GP_BINS, prev, &st, scratch);
    }
    __syncthreads();
    const unsigned long long thr = st.prefix;
    const int a = blockIdx.y;
    const float* sc = scores + (size_t)a * g.SHW;
    for (int cell = blockIdx.x * GP_THREADS + threadIdx.x; cell < g.SHW; cell += gridDim.x * GP_THREADS) {
        const unsigned long long c = gp_composite(sc[cell], (unsigned)cell * g.A + a, g);
        if (c >= thr) {
            const unsigned pos = atomicAdd(counter, 1u);
            if (pos < k_target) cand[pos] = c;
        }
    }
}

// rank by counting (composite keys are distinct), then decode the candidate at its rank
__global__ void __launch_bounds__(GP_THREADS) gp_rank_decode_kernel(const unsigned long long* __restrict__ cand, int K, GpGeom g,
                                                                    const float* __restrict__ scores, const float* __restrict__ deltas,
                                                                    GpAnchors anchors, double feat_stride, float xform_clip,
                                                                    float im_s, float im_h, float im_w, float min_size_scaled,
                                                                    float* __restrict__ dets /*[K,7]*/, int* __restrict__ src_idx, uint8_t* __restrict__ flag) {
    __shared__ unsigned long long tile[GP_THREADS];
    const int i = blockIdx.x * GP_THREADS + threadIdx.x;
    const unsigned long long mine = i < K ? cand[i] : 0ull;
    int rank = 0;
    for (int base = 0; base < K; base += GP_THREADS) {
        const int j = base + threadIdx.x;
        tile[threadIdx.x] = j < K ? cand[j] : 0ull;
        __syncthreads();
        const int lim = min(GP_THREADS, K - base);
#pragma unroll 8
        for (int k = 0; k < lim; ++k) rank += tile[k] > mine;
        __syncthreads();
    }
    if (i >= K) return;
    const unsigned idx = g.idx_mask - (unsigned)(mine & g.idx_mask);
    const int a = idx % g.A, cell = idx / g.A;
    const int w = cell % g.W, h = (cell / g.W) % g.H, s = cell / (g.W * g.H);
    // all_anchors = anchors + shifts in float64, cast to float32 by bbox_transform_3d (boxes_3d.py:176)
    const float x1 = (float)((double)anchors.a[a * 6 + 0] + w * feat_stride), y1 = (float)((double)anchors.a[a * 6 + 1] + h * feat_stride);
    const float z1 = (float)((double)anchors.a[a * 6 + 2] + s * feat_stride), x2 = (float)((double)anchors.a[a * 6 + 3] + w * feat_stride);
    const float y2 = (float)((double)anchors.a[a * 6 + 4] + h * feat_stride), z2 = (float)((double)anchors.a[a * 6 + 5] + s * feat_stride);
    float d[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) d[c] = deltas[(size_t)(a * 6 + c) * g.SHW + cell];
    const float lo[3] = {x1, y1, z1}, hi[3] = {x2, y2, z2}, lim[3] = {__fsub_rn(im_w, 1.0f), __fsub_rn(im_h, 1.0f), __fsub_rn(im_s, 1.0f)};
    float b1[3], b2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {                                          // boxes_3d.py:178-222, float32, no contraction
        const float ext = __fadd_rn(__fsub_rn(hi[c], lo[c]), 1.0f);
        const float ctr = __fadd_rn(lo[c], __fmul_rn(0.5f, ext));
        const float dl = fminf(d[3 + c], xform_clip);
        const float pc = __fadd_rn(__fmul_rn(d[c], ext), ctr);
        const float pe = __fmul_rn((float)exp((double)dl), ext);
        const float half = __fmul_rn(0.5f, pe);
        b1[c] = fmaxf(fminf(__fsub_rn(pc, half), lim[c]), 0.0f);           // clip_tiled_boxes_3d, boxes_3d.py:144-164
        b2[c] = fmaxf(fminf(__fsub_rn(__fadd_rn(pc, half), 1.0f), lim[c]), 0.0f);
    }
    const float ss = __fadd_rn(__fsub_rn(b2[0], b1[0]), 1.0f), hs = __fmul_rn(ss, 0.5f);     // generate_proposals_3d.py:180-192
    const bool keep = ss >= min_size_scaled && __fadd_rn(b1[0], hs) < im_w && __fadd_rn(b1[1], hs) < im_h && __fadd_rn(b1[2], hs) < im_s;
    float* o = dets + (size_t)rank * 7;
    o[0] = b1[0]; o[1] = b1[1]; o[2] = b1[2]; o[3] = b2[0]; o[4] = b2[1]; o[5] = b2[2];
    o[6] = scores[(size_t)a * g.SHW + cell];
    src_idx[rank] = (int)idx;
    flag[rank] = keep ? 1 : 0;
}

// order-preserving compaction of the flagged candidates (one CTA; K <= a few thousand)
__global__ void __launch_bounds__(1024) gp_compact_kernel(const float* __restrict__ dets, const int* __restrict__ src_idx, const uint8_t* __restrict__ flag, int K,
                                                          float* __restrict__ dets_kept, int* __restrict__ src_kept, int32_t* __restrict__ offsets) {
    __shared__ int warp_cnt[32];
    __shared__ int base_s;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int start = 0; start < K; start += 1024) {
        const int i = start + threadIdx.x;
        const bool f = i < K && flag[i];
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, f);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int before = base_s;
        for (int k = 0; k < warp; ++k) before += warp_cnt[k];
        if (f) {
            const int pos = before + __popc(bal & ((1u << lane) - 1u));
#pragma unroll
            for (int c = 0; c < 7; ++c) dets_kept[(size_t)pos * 7 + c] = dets[(size_t)i * 7 + c];
            src_kept[pos] = src_idx[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int k = 0; k < 32; ++k) tot += warp_cnt[k];
            base_s += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { offsets[0] = 0; offsets[1] = base_s; }
}

__global__ void gp_gather_kernel(const float* __restrict__ dets_kept, const int* __restrict__ src_kept, const int32_t* __restrict__ offsets,
                                 const int64_t* __restrict__ keep, const int32_t* __restrict__ keep_count, int use_nms, int post_top_n, int cap,
                                 float image, float* __restrict__ rois, float* __restrict__ probs, int64_t* __restrict__ keep_idx, int32_t* __restrict__ count_out) {
    int n = use_nms ? keep_count[0] : offsets[1];
    if (use_nms && post_top_n > 0 && n > post_top_n) n = post_top_n;
    if (n > cap) n = cap;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *count_out = n;
    if (i >= n) return;
    const int src = use_nms ? (int)keep[i] : i;
    rois[(size_t)i * 7] = image;
#pragma unroll
    for (int c = 0; c < 6; ++c) rois[(size_t)i * 7 + 1 + c] = dets_kept[(size_t)src * 7 + c];
    probs[i] = dets_kept[(size_t)src * 7 + 6];
    keep_idx[i] = src_kept[src];
}

struct GpPlan {
    GpGeom g;
    int K, cap, take_all, use_nms;
    size_t off_hist, off_states, off_counter, off_cand, off_dets, off_src, off_flag, off_dets_kept, off_src_kept, off_offsets, off_keep, off_keep_count,
        off_nms, nms_bytes, total;
};

static int gp_plan(int A, int S, int H, int W, int pre_nms_topN, int post_nms_topN, float nms_thresh, GpPlan* p) {
    const long long n = (long long)A * S * H * W;
    if (A < 1 || A > GP_AMAX || S < 1 || H < 1 || W < 1 || n >= (1ll << 31)) return B200SEG_EINVAL;
    GpGeom& g = p->g;
    g.A = A; g.S = S; g.H = H; g.W = W; g.SHW = S * H * W;
    g.idx_bits = 1;
    while ((1ll << g.idx_bits) < n) ++g.idx_bits;
    g.idx_mask = (unsigned)((1ull << g.idx_bits) - 1ull);
    g.npass = (32 + g.idx_bits + GP_DIGIT - 1) / GP_DIGIT;
    g.TB = g.npass * GP_DIGIT;
    p->take_all = pre_nms_topN <= 0 || pre_nms_topN >= n;
    p->K = p->take_all ? (int)n : pre_nms_topN;
    p->use_nms = nms_thresh > 0.0f;
    p->cap = (p->use_nms && post_nms_topN > 0 && post_nms_topN < p->K) ? post_nms_topN : p->K;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t r = o; o += align_up(bytes, 256); return r; };
    p->off_hist = take((size_t)GP_MAXPASS * GP_BINS * 4);
    p->off_states = take(GP_MAXPASS * sizeof(GpState));
    p->off_counter = take(4);
    p->off_cand = take((size_t)p->K * 8);
    p->off_dets = take((size_t)p->K * 7 * 4);
    p->off_src = take((size_t)p->K * 4);
    p->off_flag = take((size_t)p->K);
    p->off_dets_kept = take((size_t)p->K * 7 * 4);
    p->off_src_kept = take((size_t)p->K * 4);
    p->off_offsets = take(8);
    p->off_keep = take((size_t)p->K * 8);
    p->off_keep_count = take(4);
    p->nms_bytes = p->use_nms ? b200seg_nms3d_workspace_bytes(1, p->K) : 0;
    p->off_nms = take(p->nms_bytes);
    p->total = o + 256;
    return 0;
}

}  // namespace b200seg

using namespace b200seg;

extern "C" size_t b200seg_generate_proposals_workspace_bytes(int A, int S, int H, int W, int pre_nms_topN, int post_nms_topN, float nms_thresh) {
    GpPlan p;
    if (gp_plan(A, S, H, W, pre_nms_topN, post_nms_topN, nms_thresh, &p)) return 0;
    return p.total;
}

extern "C" int b200seg_generate_proposals_capacity(int A, int S, int H, int W, int pre_nms_topN, int post_nms_topN, float nms_thresh) {
    GpPlan p;
    if (gp_plan(A, S, H, W, pre_nms_topN, post_nms_topN, nms_thresh, &p)) return -1;
    return p.cap;
}

extern "C" int b200seg_generate_proposals_dev(const float* scores, const float* deltas, const float* im_info, int n_images,
                                              int A, int S, int H, int W, const float* anchors, float feat_stride,
                                              int pre_nms_topN, int post_nms_topN, float nms_thresh, float min_size,
                                              float* rois, float* probs, int64_t* keep_idx, int32_t* counts,
                                              void* workspace, size_t workspace_bytes, b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(n_images >= 0, "generate_proposals: negative image count");
    if (n_images == 0) return 0;
    GpPlan p;
    B200_CHECK_ARG(gp_plan(A, S, H, W, pre_nms_topN, post_nms_topN, nms_thresh, &p) == 0,
                   "generate_proposals: bad geometry A=%d (1..%d) S=%d H=%d W=%d", A, GP_AMAX, S, H, W);
    B200_CHECK_ARG(scores && deltas && im_info && anchors && rois && probs && keep_idx && counts && workspace, "generate_proposals: null pointer");
    if (workspace_bytes < p.total) { set_error("generate_proposals: workspace too small (%zu < %zu)", workspace_bytes, p.total); return B200SEG_EWORKSPACE; }
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    unsigned* hist = (unsigned*)(ws + p.off_hist);
    GpState* states = (GpState*)(ws + p.off_states);
    unsigned* counter = (unsigned*)(ws + p.off_counter);
    unsigned long long* cand = (unsigned long long*)(ws + p.off_cand);
    float* dets = (float*)(ws + p.off_dets);
    int* src = (int*)(ws + p.off_src);
    uint8_t* flag = (uint8_t*)(ws + p.off_flag);
    float* dets_kept = (float*)(ws + p.off_dets_kept);
    int* src_kept = (int*)(ws + p.off_src_kept);
    int32_t* offsets = (int32_t*)(ws + p.off_offsets);
    int64_t* keep = (int64_t*)(ws + p.off_keep);
    int32_t* keep_count = (int32_t*)(ws + p.off_keep_count);
    GpAnchors an;
    for (int i = 0; i < A * 6; ++i) an.a[i] = anchors[i];
    const float xform_clip = (float)log(1000.0 / 16.0);                   // cfg.BBOX_XFORM_CLIP (lib/core/config.py:947) at float32
    const GpGeom& g = p.g;
    int bx = (g.SHW + GP_THREADS * 4 - 1) / (GP_THREADS * 4);
    const int capx = (num_sms() * 4 + A - 1) / A;
    if (bx > capx) bx = capx;
    if (bx < 1) bx = 1;
    const dim3 grid(bx, A);
    const size_t n_per_image = (size_t)A * g.SHW;
    for (int im = 0; im < n_images; ++im) {
        const float* sc = scores + (size_t)im * n_per_image;
        const float* dl = deltas + (size_t)im * n_per_image * 6;
        const float* info = im_info + (size_t)im * 4;
        B200_CUDA(cudaMemsetAsync(ws + p.off_hist, 0, p.off_cand - p.off_hist, stream));       // histograms, states, counter
        if (!p.take_all) {
            for (int pass = 0; pass < g.npass; ++pass) {
                gp_hist_kernel<<<grid, GP_THREADS, 0, stream>>>(sc, g, pass, (unsigned)p.K, hist, states);
                B200_LAUNCH_CHECK("gp_hist_kernel");
            }
        }
        gp_collect_kernel<<<grid, GP_THREADS, 0, stream>>>(sc, g, p.take_all, (unsigned)p.K, hist, states, cand, counter);
        B200_LAUNCH_CHECK("gp_collect_kernel");
        const float min_scaled = min_size * info[3];
        gp_rank_decode_kernel<<<(p.K + GP_THREADS - 1) / GP_THREADS, GP_THREADS, 0, stream>>>(cand, p.K, g, sc, dl, an, (double)feat_stride, xform_clip,
                                                                                         info[0], info[1], info[2], min_scaled, dets, src, flag);
        B200_LAUNCH_CHECK("gp_rank_decode_kernel");
        gp_compact_kernel<<<1, 1024, 0, stream>>>(dets, src, flag, p.K, dets_kept, src_kept, offsets);
        B200_LAUNCH_CHECK("gp_compact_kernel");
        if (p.use_nms) {
            const int e = b200seg_nms3d_dev(dets_kept, offsets, 1, p.K, nms_thresh, 0, keep, keep_count, nullptr, ws + p.off_nms, p.nms_bytes, stream_);
            if (e) return e;
        }
        gp_gather_kernel<<<(p.cap + 255) / 256, 256, 0, stream>>>(dets_kept, src_kept, offsets, keep, keep_count, p.use_nms, post_nms_topN, p.cap, (float)im,
                                                                  rois + (size_t)im * p.cap * 7, probs + (size_t)im * p.cap, keep_idx + (size_t)im * p.cap,
                                                                  counts + im);
        B200_LAUNCH_CHECK("gp_gather_kernel");
    }
    return 0;
}
