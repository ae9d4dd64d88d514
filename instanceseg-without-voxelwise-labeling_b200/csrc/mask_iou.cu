// mask_iou.cu -- pairwise instance-mask overlaps between two label volumes
// (replaces tools/evaluation/mask_iou.py:49-109: mask_iou_fast / mask_ios_fast / mask_iog_fast, numba loops over
//  N x K x V voxels; second piece of the "eval kernels + RLE codec" next row).
//
// The reference is handed boolean stacks pred_masks[i] = (pred == id_i), gt_masks[k] = (gt == id_k) that its callers
// cut out of two label volumes (tools/evaluation/eval_instance_segmentation_soma.py:186-197).  Instances of a label
// volume are disjoint, so ONE pass over the two volumes with a joint histogram over (pred row, gt row) gives every
// intersection, and the row / column sums give the areas:
//   mask_joint_hist_kernel   8 voxels per thread (128-bit loads of both uint16 volumes), ids -> rows through two
//                            64 K-entry tables, equal consecutive keys merged in registers, one L2 reduction per run;
//                            voxels that are background in both volumes are skipped.
//   mask_sums_kernel         row sums (area of every pred instance) and column sums (area of every gt instance).
//   mask_ratios_kernel       iou = I / (A + B - I), ios = I / A, iog = I / B, each as float32(double / double) like the
//                            reference's float accumulators stored into a float32 array.
#include "common.cuh"

namespace b200seg {

constexpr int MI_THREADS = 256;

// table[(rp + 1) * (Ng + 1) + (rg + 1)], rp / rg = row of the id or -1 (index 0) when the id is not listed
__global__ void __launch_bounds__(MI_THREADS)
mask_joint_hist_kernel(const uint16_t* __restrict__ pred, const uint16_t* __restrict__ gt, long long V,
                       const int32_t* __restrict__ lut_p, const int32_t* __restrict__ lut_g, int Ng1,
                       unsigned long long* __restrict__ table, int vec_ok) {
    const long long ngroups = (V + 7) >> 3;
    for (long long gidx = (long long)blockIdx.x * MI_THREADS + threadIdx.x; gidx < ngroups;
         gidx += (long long)gridDim.x * MI_THREADS) {
        const long long j0 = gidx << 3;
        unsigned short p[8], g[8];
        if (vec_ok && j0 + 7 < V) {
            const uint4 a = ld_stream_u4(pred + j0), b = ld_stream_u4(gt + j0);
            p[0] = a.x & 0xFFFF; p[1] = a.x >> 16; p[2] = a.y & 0xFFFF; p[3] = a.y >> 16;
            p[4] = a.z & 0xFFFF; p[5] = a.z >> 16; p[6] = a.w & 0xFFFF; p[7] = a.w >> 16;
            g[0] = b.x & 0xFFFF; g[1] = b.x >> 16; g[2] = b.y & 0xFFFF; g[3] = b.y >> 16;
            g[4] = b.z & 0xFFFF; g[5] = b.z >> 16; g[6] = b.w & 0xFFFF; g[7] = b.w >> 16;
            if ((a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) == 0u) continue;      // all background: nothing to count
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) { p[k] = j0 + k < V ? pred[j0 + k] : 0; g[k] = j0 + k < V ? gt[j0 + k] : 0; }
        }
        int run_key = -1;
        unsigned run_len = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int key = -1;
            if (p[k] | g[k]) {
                const int rp = p[k] ? lut_p[p[k]] : -1, rg = g[k] ? lut_g[g[k]] : -1;
                if (rp >= 0 || rg >= 0) key = (rp + 1) * Ng1 + (rg + 1);
            }
            if (key != run_key) {
                if (run_key >= 0) atomicAdd(&table[run_key], (unsigned long long)run_len);
                run_key = key; run_len = 0;
            }
            ++run_len;
        }
        if (run_key >= 0) atomicAdd(&table[run_key], (unsigned long long)run_len);
    }
}

// grid = (Np1 + Ng1 + 255) / 256: thread t < Np1 sums row t, thread Np1 + k sums column k
__global__ void mask_sums_kernel(const unsigned long long* __restrict__ table, int Np1, int Ng1,
                                 unsigned long long* __restrict__ area_p, unsigned long long* __restrict__ area_g) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < Np1) {
        unsigned long long s = 0;
        for (int k = 0; k < Ng1; ++k) s += table[(size_t)t * Ng1 + k];
        area_p[t] = s;
    } else if (t < Np1 + Ng1) {
        const int k = t - Np1;
        unsigned long long s = 0;
        for (int n = 0; n < Np1; ++n) s += table[(size_t)n * Ng1 + k];
        area_g[k] = s;
    }
}

__global__ void mask_ratios_kernel(const unsigned long long* __restrict__ table, const unsigned long long* __restrict__ area_p,
                                   const unsigned long long* __restrict__ area_g, int Np, int Ng,
                                   float* __restrict__ iou, float* __restrict__ ios, float* __restrict__ iog) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)Np * Ng) return;
    const int n = (int)(i / Ng), k = (int)(i - (long long)n * Ng);
    const double I = (double)table[(size_t)(n + 1) * (Ng + 1) + (k + 1)];
    const double A = (double)area_p[n + 1], B = (double)area_g[k + 1];
    if (iou) iou[i] = (float)(I / (A + B - I));               // mask_iou.py:66-67 (0 / 0 -> NaN; numba would raise)
    if (ios) ios[i] = (float)(I / A);                          // :87-88
    if (iog) iog[i] = (float)(I / B);                          // :108-109
}

}  // namespace b200seg

using namespace b200seg;

extern "C" size_t b200seg_mask_overlaps_workspace_bytes(int n_pred, int n_gt) {
    if (n_pred < 0 || n_gt < 0) return 256;
    const size_t cells = (size_t)(n_pred + 1) * (n_gt + 1);
    return align_up(cells * 8, 256) + align_up((size_t)(n_pred + 1) * 8, 256) + align_up((size_t)(n_gt + 1) * 8, 256) + 256;
}

extern "C" int b200seg_mask_overlaps_dev(const uint16_t* pred, const uint16_t* gt, long long n_voxels,
                                         const int32_t* lut_pred, const int32_t* lut_gt, int n_pred, int n_gt,
                                         float* iou, float* ios, float* iog, int64_t* inter, int64_t* area_pred, int64_t* area_gt,
                                         void* workspace, size_t workspace_bytes, b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(n_voxels >= 0 && n_pred >= 0 && n_gt >= 0, "mask_overlaps: bad sizes");
    B200_CHECK_ARG(pred && gt && lut_pred && lut_gt && workspace, "mask_overlaps: null pointer");
    if (workspace_bytes < b200seg_mask_overlaps_workspace_bytes(n_pred, n_gt)) { set_error("mask_overlaps: workspace too small"); return B200SEG_EWORKSPACE; }
    const int Np1 = n_pred + 1, Ng1 = n_gt + 1;
    const size_t cells = (size_t)Np1 * Ng1;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    unsigned long long* table = (unsigned long long*)ws; ws += align_up(cells * 8, 256);
    unsigned long long* ap = (unsigned long long*)ws; ws += align_up((size_t)Np1 * 8, 256);
    unsigned long long* ag = (unsigned long long*)ws;
    B200_CUDA(cudaMemsetAsync(table, 0, cells * 8, stream));
    if (n_voxels > 0) {
        const long long groups = (n_voxels + 7) / 8;
        long long blocks = (groups + MI_THREADS - 1) / MI_THREADS;
        const long long capb = (long long)num_sms() * 16;
        if (blocks > capb) blocks = capb;
        const int vec_ok = ((((uintptr_t)pred) | ((uintptr_t)gt)) & 15) == 0;
        mask_joint_hist_kernel<<<(unsigned)blocks, MI_THREADS, 0, stream>>>(pred, gt, n_voxels, lut_pred, lut_gt, Ng1, table, vec_ok);
        B200_LAUNCH_CHECK("mask_joint_hist_kernel");
    }
    mask_sums_kernel<<<(Np1 + Ng1 + 255) / 256, 256, 0, stream>>>(table, Np1, Ng1, ap, ag);
    B200_LAUNCH_CHECK("mask_sums_kernel");
    if ((long long)n_pred * n_gt > 0 && (iou || ios || iog)) {
        const long long n = (long long)n_pred * n_gt;
        mask_ratios_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(table, ap, ag, n_pred, n_gt, iou, ios, iog);
        B200_LAUNCH_CHECK("mask_ratios_kernel");
    }
    // optional raw counts (int64, same bits as the unsigned accumulators)
    if (inter) B200_CUDA(cudaMemcpyAsync(inter, table, cells * 8, cudaMemcpyDeviceToDevice, stream));     // [(n_pred+1), (n_gt+1)], row/col 0 = unlisted / background
    if (area_pred) B200_CUDA(cudaMemcpyAsync(area_pred, ap, (size_t)Np1 * 8, cudaMemcpyDeviceToDevice, stream));
    if (area_gt) B200_CUDA(cudaMemcpyAsync(area_gt, ag, (size_t)Ng1 * 8, cudaMemcpyDeviceToDevice, stream));
    return 0;
}
