// eval.cu -- the volume-sized parts of the two evaluation scripts:
//   label presence   np.unique(label_volume) minus background (tools/evaluation/eval_instance_segmentation_soma.py:177-181)
//                    as a 65536-entry presence table filled in one streaming pass;
//   voxel counts     tp_pixel / gt_pixel / pre_pixel of tools/evaluation/evaluation_nuclei_f1score_seg.py:86-89, :124-130:
//                    keep_pred_mask = (pred > 0) inside the boxes of the matched detections; tp = keep & (gt > 0).
//                    The matched boxes are rasterised into a 1 bit/voxel volume (one CTA per box, 32-bit ORs per row
//                    segment), then one pass over both label volumes (128-bit loads, 8 voxels per thread) counts the three
//                    sums with warp reductions and one 64-bit atomic per warp.
// The greedy matching itself (a few hundred rows, sequential by construction) stays on the host (evaluation.py).
#include "common.cuh"

namespace b200seg {

// Each CTA collects the labels it sees in a 65536-bit shared-memory bitmap (test before set: the same few labels are seen
// thousands of times) and publishes the set bits once at the end, so global memory sees at most one byte store per
// (CTA, label) instead of one per sighting.
__global__ void __launch_bounds__(256) label_presence_kernel(const uint16_t* __restrict__ lab, long long n, uint8_t* present) {
    __shared__ uint32_t s_bits[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) s_bits[i] = 0u;
    __syncthreads();
    auto mark = [&](uint32_t a) {
        const uint32_t bit = 1u << (a & 31);
        if (!(s_bits[a >> 5] & bit)) atomicOr(&s_bits[a >> 5], bit);
    };
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nvec = n >> 3;
    const bool aligned = (reinterpret_cast<uintptr_t>(lab) & 15) == 0;
    if (aligned) {
        for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < nvec; g += stride) {
            const uint4 v = ld_stream_u4(lab + g * 8);
            if ((v.x | v.y | v.z | v.w) == 0u) continue;                  // background dominates
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            uint32_t last = 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t a = w[k] & 0xFFFFu, b = w[k] >> 16;
                if (a != last) { mark(a); last = a; }
                if (b != last) { mark(b); last = b; }
            }
        }
    }
    const long long tail0 = aligned ? nvec * 8 : 0;
    for (long long i = tail0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) mark(lab[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += 256) {
        uint32_t m = s_bits[i];
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            present[i * 32 + bit] = 1;                                    // benign race: every writer stores 1
        }
    }
}

// one CTA per matched box: set the bits of its voxels (volume-clipped) in `bits` (1 bit per voxel, flat index)
__global__ void __launch_bounds__(128) eval_mark_boxes_kernel(const int32_t* __restrict__ boxes, int S, int H, int W, uint32_t* __restrict__ bits) {
    const int32_t* b = boxes + 6 * (size_t)blockIdx.x;
    const int x1 = max(b[0], 0), y1 = max(b[1], 0), z1 = max(b[2], 0);
    const int x2 = min(b[3], W - 1), y2 = min(b[4], H - 1), z2 = min(b[5], S - 1);
    if (x2 < x1 || y2 < y1 || z2 < z1) return;
    const int sy = y2 - y1 + 1, rows = sy * (z2 - z1 + 1);
    for (int r = threadIdx.x; r < rows; r += blockDim.x) {
        const int z = z1 + r / sy, y = y1 + r % sy;
        const long long f0 = ((long long)z * H + y) * W + x1, f1 = f0 + (x2 - x1);        // inclusive flat range of the row
        for (long long w = f0 >> 5; w <= (f1 >> 5); ++w) {
            uint32_t m = 0xFFFFFFFFu;
            if (w == (f0 >> 5)) m &= 0xFFFFFFFFu << (f0 & 31);
            if (w == (f1 >> 5)) m &= 0xFFFFFFFFu >> (31 - (f1 & 31));
            atomicOr(&bits[w], m);
        }
    }
}

// counts[0] += #(pred>0 & gt>0 & bit), counts[1] += #(gt>0), counts[2] += #(pred>0)
__global__ void __launch_bounds__(256) eval_voxel_counts_kernel(const uint16_t* __restrict__ pred, const uint16_t* __restrict__ gt, long long n,
                                                                const uint8_t* __restrict__ bits, unsigned long long* __restrict__ counts) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(gt)) & 15) == 0;
    const long long nvec = aligned ? n >> 3 : 0;
    unsigned int tp = 0, ng = 0, np_ = 0;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < nvec; g += stride) {
        const uint4 p = ld_stream_u4(pred + g * 8), q = ld_stream_u4(gt + g * 8);
        const uint32_t pw[4] = {p.x, p.y, p.z, p.w}, qw[4] = {q.x, q.y, q.z, q.w};
        unsigned pm = 0u, gm = 0u;                                          // bit k = voxel k of the group is set
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            pm |= ((pw[k] & 0xFFFFu) ? 1u : 0u) << (2 * k) | ((pw[k] >> 16) ? 1u : 0u) << (2 * k + 1);
            gm |= ((qw[k] & 0xFFFFu) ? 1u : 0u) << (2 * k) | ((qw[k] >> 16) ? 1u : 0u) << (2 * k + 1);
        }
        np_ += __popc(pm); ng += __popc(gm);
        if (pm & gm) tp += __popc(pm & gm & (unsigned)bits[g]);            // 8 voxels = one byte of the bit volume
    }
    for (long long i = nvec * 8 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const bool p = pred[i] != 0, q = gt[i] != 0;
        np_ += p; ng += q;
        tp += p && q && ((bits[i >> 3] >> (i & 7)) & 1);
    }
    tp = __reduce_add_sync(0xFFFFFFFFu, tp); ng = __reduce_add_sync(0xFFFFFFFFu, ng); np_ = __reduce_add_sync(0xFFFFFFFFu, np_);
    if ((threadIdx.x & 31) == 0) {
        if (tp) atomicAdd(&counts[0], (unsigned long long)tp);
        if (ng) atomicAdd(&counts[1], (unsigned long long)ng);
        if (np_) atomicAdd(&counts[2], (unsigned long long)np_);
    }
}

}  // namespace b200seg

using namespace b200seg;

extern "C" int b200seg_label_presence_dev(const uint16_t* labels, long long n, uint8_t* present, b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(n >= 0 && present && (n == 0 || labels), "label_presence: bad arguments");
    B200_CUDA(cudaMemsetAsync(present, 0, 65536, stream));
    if (n == 0) return 0;
    long long grid = (n / 8 + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    grid = grid < 1 ? 1 : (grid > cap ? cap : grid);
    label_presence_kernel<<<(unsigned)grid, 256, 0, stream>>>(labels, n, present);
    B200_LAUNCH_CHECK("label_presence_kernel");
    return 0;
}

extern "C" size_t b200seg_eval_voxel_counts_workspace_bytes(long long n_voxels) {
    return align_up((size_t)(n_voxels > 0 ? n_voxels : 0) / 8 + 8, 256) + 256;
}

extern "C" int b200seg_eval_voxel_counts_dev(const uint16_t* pred, const uint16_t* gt, int S, int H, int W,
                                             const int32_t* boxes, int n_boxes, unsigned long long* counts,
                                             void* workspace, size_t workspace_bytes, b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(S > 0 && H > 0 && W > 0 && n_boxes >= 0 && pred && gt && counts && workspace, "eval_voxel_counts: bad arguments");
    B200_CHECK_ARG(n_boxes == 0 || boxes, "eval_voxel_counts: null boxes");
    const long long n = (long long)S * H * W;
    if (workspace_bytes < b200seg_eval_voxel_counts_workspace_bytes(n)) { set_error("eval_voxel_counts: workspace too small"); return B200SEG_EWORKSPACE; }
    uint32_t* bits = (uint32_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    B200_CUDA(cudaMemsetAsync(bits, 0, align_up((size_t)n / 8 + 8, 256), stream));
    B200_CUDA(cudaMemsetAsync(counts, 0, 3 * sizeof(unsigned long long), stream));
    if (n_boxes > 0) {
        eval_mark_boxes_kernel<<<n_boxes, 128, 0, stream>>>(boxes, S, H, W, bits);
        B200_LAUNCH_CHECK("eval_mark_boxes_kernel");
    }
    long long grid = (n / 8 + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    grid = grid < 1 ? 1 : (grid > cap ? cap : grid);
    eval_voxel_counts_kernel<<<(unsigned)grid, 256, 0, stream>>>(pred, gt, n, (const uint8_t*)bits, counts);
    B200_LAUNCH_CHECK("eval_voxel_counts_kernel");
    return 0;
}
