// segm_paste.cu -- Mask R-CNN mask paste-back: the per-detection body of segm_results (lib/core/test.py:902-938).
//
// For every detection the reference zero-pads the M^3 mask logit block to (M+2)^3 (:906-908), resizes it to the
// (expanded, int32-truncated) box with skimage.transform.resize(padded, (s,h,w), mode='reflect', anti_aliasing=True)
// (:919), thresholds (:920) and pastes the part inside the volume into a fresh uint8 volume (:921-931).
// skimage's resize delegates its arithmetic to scipy.ndimage, and that is what is reproduced here op by op:
//   1. ndimage.gaussian_filter(padded, sigma=max(0, (in/out - 1)/2) per axis, mode='mirror'): axis 0, 1, 2 in turn,
//      only where out < in; symmetric correlate1d in fp64 ( acc = w[R]*x0; acc += (x[-j] + x[+j]) * w[R-j], j = R..1,
//      separate multiply and add ), float32 stored after every pass;
//   2. ndimage.zoom(order=1, mode='mirror', grid_mode=True): source coordinate (o + 0.5) * in/out - 0.5, mirrored at 0,
//      taps floor(c) and floor(c)+1 (mirrored at the far edge), weights w0 = 1 - frac, w1 = 1 - w0, and the eight-tap
//      sum  t += ((v * wz) * wy) * wx  in z-major tap order, fp64, rounded to float32 once;
//   3. clip to [min, max] of the padded block (skimage's _clip_warp_output), compare  > thresh  in float32.
// Only the voxels inside the volume are computed; they are written as packed uint8 crops (one per detection, C order),
// from which segm_expand_kernel builds the reference's one-volume-per-detection output where that is wanted.
//
// One CTA per (detection, slice of the crop's (y, x) columns): the padded block lives in shared memory (two float buffers for the Gaussian
// ping-pong), each thread owns (y, x) columns of the crop -- consecutive threads write consecutive bytes -- and walks the
// column in z with the z taps staged in shared memory.  All fp64 arithmetic uses the _rn intrinsics so that nothing is
// contracted into an FMA (scipy's C code is not).
#include "common.cuh"

namespace b200seg {

constexpr int SG_THREADS = 256;
constexpr int SG_ZT = 128;                                  // z taps staged per round

struct AxisTap { int i0, i1; double w0, w1; };

// scipy NI_ZoomShift (grid_mode) + map_coordinate('mirror') + order-1 spline weights for output index o.
// -0.5 < cc < n_in - 0.5 always, so the mirror map reduces to |cc| and to folding tap n_in back to n_in - 2.
__device__ __forceinline__ AxisTap axis_tap(int o, double zoom, int n_in) {
    double cc = __dadd_rn((double)o, 0.5);
    cc = __dmul_rn(cc, zoom);
    cc = __dadd_rn(cc, -0.5);
    if (cc < 0.0) cc = -cc;
    const double f = floor(cc);
    const double x = __dsub_rn(cc, f);
    AxisTap t;
    t.w0 = __dsub_rn(1.0, x);
    t.w1 = __dsub_rn(1.0, t.w0);
    t.i0 = (int)f;
    t.i1 = t.i0 + 1;
    if (t.i1 >= n_in) t.i1 = 2 * n_in - 2 - t.i1;
    return t;
}

__device__ __forceinline__ int mirror_index(int i, int n) {   // scipy 'mirror' (d c b | a b c d | c b a), any offset
    const int s2 = 2 * n - 2;
    int m = i % s2;
    if (m < 0) m += s2;
    return m < n ? m : s2 - m;
}

// sigma and radius of the anti-aliasing Gaussian for one axis, as skimage / scipy compute them in fp64
__device__ __forceinline__ int gauss_radius(int n_in, int n_out) {
    if (n_out >= n_in) return 0;
    const double factor = __ddiv_rn((double)n_in, (double)n_out);
    const double sigma = __ddiv_rn(__dsub_rn(factor, 1.0), 2.0);
    return (int)__dadd_rn(__dmul_rn(4.0, sigma), 0.5);
}

__global__ void __launch_bounds__(SG_THREADS) segm_resize_paste_kernel(const float* __restrict__ masks, const int32_t* __restrict__ mask_index,
                                                                       const int32_t* __restrict__ boxes, int M, const double* __restrict__ gauss_w,
                                                                       int gauss_stride, float thresh, int im_s, int im_h, int im_w,
                                                                       uint8_t* __restrict__ out, const int64_t* __restrict__ out_off, int zsplit) {
    extern __shared__ __align__(16) unsigned char sg_smem[];
    const int M2 = M + 2, M2sq = M2 * M2, M2c = M2sq * M2;
    float* buf_a = reinterpret_cast<float*>(sg_smem);
    float* buf_b = buf_a + M2c;
    double* s_w = reinterpret_cast<double*>(buf_b + M2c);                      // 2*M2 Gaussian taps (2*M2c floats: 8-byte aligned)
    double* s_zw = s_w + 2 * M2;                                                 // SG_ZT x {w0, w1}
    int* s_zi = reinterpret_cast<int*>(s_zw + 2 * SG_ZT);                        // SG_ZT x {i0, i1}
    __shared__ float s_red[2][SG_THREADS / 32];

    const int d = blockIdx.x, tid = threadIdx.x;
    const int32_t* b = boxes + (size_t)d * 6;
    const int bx0 = b[0], by0 = b[1], bz0 = b[2], bx1 = b[3], by1 = b[4], bz1 = b[5];
    const int ow = max(bx1 - bx0 + 1, 1), oh = max(by1 - by0 + 1, 1), os = max(bz1 - bz0 + 1, 1);     // core/test.py:911-916
    const int x0 = max(bx0, 0), x1 = min(bx1 + 1, im_w), y0 = max(by0, 0), y1 = min(by1 + 1, im_h);   // :923-928
    const int z0 = max(bz0, 0), z1 = min(bz1 + 1, im_s);
    const int cw = x1 - x0, ch = y1 - y0, cs = z1 - z0;
    if (cw <= 0 || ch <= 0 || cs <= 0) return;
    // the CTAs of one detection split the (y, x) columns of the crop, not its z extent: a thread's per-column setup (two axis
    // taps, one division) is then amortised over the whole column
    const int plane = ch * cw;
    const int col_lo = (int)((long long)plane * blockIdx.y / zsplit), col_hi = (int)((long long)plane * (blockIdx.y + 1) / zsplit);
    if (col_hi <= col_lo) return;
    const int zlo = z0, zhi = z1;

    // padded block (core/test.py:899, :906-908) + its min / max for the final clip
    const float* src = masks + (size_t)mask_index[d] * (size_t)(M * M * M);
    float mn = 0.0f, mx = 0.0f;                                                 // the zero border is part of the block
    for (int e = tid; e < M2c; e += SG_THREADS) {
        const int z = e / M2sq, y = (e / M2) % M2, x = e % M2;
        float v = 0.0f;
        if (z >= 1 && z <= M && y >= 1 && y <= M && x >= 1 && x <= M) v = __ldg(src + ((size_t)(z - 1) * M + (y - 1)) * M + (x - 1));
        buf_a[e] = v;
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = mn; s_red[1][tid >> 5] = mx; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_THREADS / 32; ++k) { mn = fminf(mn, s_red[0][k]); mx = fmaxf(mx, s_red[1][k]); }

    // anti-aliasing Gaussian, axis 0 (z), 1 (y), 2 (x); float32 between the passes like scipy's output array
    float* cur = buf_a;
    float* nxt = buf_b;
    const int n_out[3] = {os, oh, ow};
    for (int axis = 0; axis < 3; ++axis) {
        const int R = gauss_radius(M2, n_out[axis]);
        if (R < 1) continue;                                                    // uniform over the CTA
        const double* wrow = gauss_w + (size_t)n_out[axis] * gauss_stride;
        __syncthreads();
        for (int k = tid; k <= R; k += SG_THREADS) s_w[k] = wrow[k];
        __syncthreads();
        const int stride = axis == 0 ? M2sq : (axis == 1 ? M2 : 1);
        for (int e = tid; e < M2c; e += SG_THREADS) {
            const int l = (e / stride) % M2, base = e - l * stride;
            double acc = __dmul_rn(s_w[R], (double)cur[e]);
            for (int jj = -R; jj < 0; ++jj) {
                const double lo = (double)cur[base + mirror_index(l + jj, M2) * stride];
                const double hi = (double)cur[base + mirror_index(l - jj, M2) * stride];
                acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(lo, hi), s_w[jj + R]));
            }
            nxt[e] = (float)acc;
        }
        float* t = cur; cur = nxt; nxt = t;
    }
    __syncthreads();

    const double zoom_z = __ddiv_rn((double)M2, (double)os), zoom_y = __ddiv_rn((double)M2, (double)oh), zoom_x = __ddiv_rn((double)M2, (double)ow);
    uint8_t* dst = out + out_off[d];
    for (int zc = zlo; zc < zhi; zc += SG_ZT) {
        const int nz = min(SG_ZT, zhi - zc);
        __syncthreads();
        for (int k = tid; k < nz; k += SG_THREADS) {
            const AxisTap t = axis_tap(zc + k - bz0, zoom_z, M2);
            s_zw[2 * k] = t.w0; s_zw[2 * k + 1] = t.w1;
            s_zi[2 * k] = t.i0 * M2sq; s_zi[2 * k + 1] = t.i1 * M2sq;
        }
        __syncthreads();
        for (int col = col_lo + tid; col < col_hi; col += SG_THREADS) {
            const int yy = col / cw, xx = col - yy * cw;
            const AxisTap ty = axis_tap(y0 + yy - by0, zoom_y, M2), tx = axis_tap(x0 + xx - bx0, zoom_x, M2);
            const int o00 = ty.i0 * M2 + tx.i0, o01 = ty.i0 * M2 + tx.i1, o10 = ty.i1 * M2 + tx.i0, o11 = ty.i1 * M2 + tx.i1;
            uint8_t* q = dst + (size_t)(zc - z0) * plane + col;
            for (int k = 0; k < nz; ++k) {
                double t = 0.0;
#pragma unroll
                for (int tz = 0; tz < 2; ++tz) {
                    const float* p = cur + s_zi[2 * k + tz];
                    const double wz = s_zw[2 * k + tz];
                    t = __dadd_rn(t, __dmul_rn(__dmul_rn(__dmul_rn((double)p[o00], wz), ty.w0), tx.w0));
                    t = __dadd_rn(t, __dmul_rn(__dmul_rn(__dmul_rn((double)p[o01], wz), ty.w0), tx.w1));
                    t = __dadd_rn(t, __dmul_rn(__dmul_rn(__dmul_rn((double)p[o10], wz), ty.w1), tx.w0));
                    t = __dadd_rn(t, __dmul_rn(__dmul_rn(__dmul_rn((double)p[o11], wz), ty.w1), tx.w1));
                }
                float v = (float)t;
                v = fminf(fmaxf(v, mn), mx);                                     // np.clip(out, min, max)
                q[(size_t)k * plane] = v > thresh ? (uint8_t)1 : (uint8_t)0;
            }
        }
    }
}

// crops -> the reference's im_mask volumes (core/test.py:921-931): vols[d] = zeros(S,H,W); vols[d][box ∩ volume] = crop d.
// One thread per 16 output bytes; rows outside the box are pure zero stores.
__global__ void __launch_bounds__(256) segm_expand_kernel(const uint8_t* __restrict__ crops, const int64_t* __restrict__ out_off,
                                                          const int32_t* __restrict__ boxes, int im_s, int im_h, int im_w,
                                                          uint8_t* __restrict__ vols, long long groups_per_vol) {
    const int d = blockIdx.y;
    const int32_t* b = boxes + (size_t)d * 6;
    const int x0 = max(b[0], 0), x1 = min(b[3] + 1, im_w), y0 = max(b[1], 0), y1 = min(b[4] + 1, im_h), z0 = max(b[2], 0), z1 = min(b[5] + 1, im_s);
    const bool empty = x1 <= x0 || y1 <= y0 || z1 <= z0;
    const int cw = x1 - x0, ch = y1 - y0;
    const uint8_t* crop = crops + out_off[d];
    const size_t V = (size_t)im_s * im_h * im_w;
    uint8_t* vol = vols + (size_t)d * V;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups_per_vol; g += (long long)gridDim.x * blockDim.x) {
        const size_t e0 = (size_t)g * 16;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (!empty) {
            size_t row = e0 / im_w;
            int x = (int)(e0 - row * im_w);
            int y = (int)(row % im_h), z = (int)(row / im_h);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (e0 + k < V && z >= z0 && z < z1 && y >= y0 && y < y1 && x >= x0 && x < x1) {
                    const uint32_t v = crop[((size_t)(z - z0) * ch + (y - y0)) * cw + (x - x0)];
                    w[k >> 2] |= v << (8 * (k & 3));
                }
                if (++x == im_w) { x = 0; if (++y == im_h) { y = 0; ++z; } }
            }
        }
        if (e0 + 16 <= V) st_stream_u4(vol + e0, make_uint4(w[0], w[1], w[2], w[3]));
        else for (size_t k = 0; e0 + k < V; ++k) vol[e0 + k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
    }
}

static size_t segm_smem_bytes(int M) {
    const size_t M2 = (size_t)M + 2, M2c = M2 * M2 * M2;
    return 2 * M2c * 4 + (2 * M2 + 2 * SG_ZT) * 8 + 2 * SG_ZT * 4;
}

}  // namespace b200seg

using namespace b200seg;

extern "C" int b200seg_segm_gauss_table_size(int M) { return M < 1 ? 0 : (M + 2) * 2 * (M + 2); }

extern "C" int b200seg_segm_paste_dev(const float* masks, const int32_t* mask_index, const int32_t* ref_boxes, int n, int M,
                                      const double* gauss_w, float thresh, int im_s, int im_h, int im_w,
                                      uint8_t* crops, const int64_t* crop_off, b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(n >= 0 && M >= 1 && M <= 26 && im_s > 0 && im_h > 0 && im_w > 0, "segm_paste: bad sizes (1 <= M <= 26)");
    if (n == 0) return 0;
    B200_CHECK_ARG(masks && mask_index && ref_boxes && gauss_w && crops && crop_off, "segm_paste: null pointer");
    const size_t smem = segm_smem_bytes(M);
    static std::mutex mu;
    static size_t smem_set[64] = {0};
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 0 && dev < 64 && smem_set[dev] < smem) {
            B200_CUDA(cudaFuncSetAttribute(segm_resize_paste_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            smem_set[dev] = smem;
        }
    }
    int zsplit = (4 * num_sms() + n - 1) / n;
    zsplit = zsplit < 1 ? 1 : (zsplit > 32 ? 32 : zsplit);
    segm_resize_paste_kernel<<<dim3((unsigned)n, (unsigned)zsplit), SG_THREADS, smem, stream>>>(masks, mask_index, ref_boxes, M, gauss_w, 2 * (M + 2), thresh,
                                                                                               im_s, im_h, im_w, crops, crop_off, zsplit);
    B200_LAUNCH_CHECK("segm_resize_paste_kernel");
    return 0;
}

extern "C" int b200seg_segm_expand_dev(const uint8_t* crops, const int64_t* crop_off, const int32_t* ref_boxes, int n,
                                       int im_s, int im_h, int im_w, uint8_t* volumes, b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(n >= 0 && im_s > 0 && im_h > 0 && im_w > 0, "segm_expand: bad sizes");
    if (n == 0) return 0;
    B200_CHECK_ARG(crops && crop_off && ref_boxes && volumes, "segm_expand: null pointer");
    B200_CHECK_ARG(n <= 65535, "segm_expand: more than 65535 detections");
    const size_t V = (size_t)im_s * im_h * im_w;
    B200_CHECK_ARG((V % 16 == 0 || n == 1) && (reinterpret_cast<uintptr_t>(volumes) & 15) == 0,
                   "segm_expand: volumes must be 16-byte aligned and S*H*W a multiple of 16 when n > 1");
    const long long groups = (long long)((V + 15) / 16);
    long long gx = (groups + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (gx > cap) gx = cap;
    segm_expand_kernel<<<dim3((unsigned)gx, (unsigned)n), 256, 0, stream>>>(crops, crop_off, ref_boxes, im_s, im_h, im_w, volumes, groups);
    B200_LAUNCH_CHECK("segm_expand_kernel");
    return 0;
}

// numpy seam: everything on the host; crops [crop_off[n]] bytes come back packed
extern "C" int b200seg_segm_paste_host(const float* masks, long long n_mask_blocks, const int32_t* mask_index, const int32_t* ref_boxes, int n, int M,
                                       const double* gauss_w, float thresh, int im_s, int im_h, int im_w,
                                       uint8_t* crops, const int64_t* crop_off) {
    B200_CHECK_ARG(n >= 0 && M >= 1 && M <= 26 && n_mask_blocks >= 0, "segm_paste_host: bad sizes");
    if (n == 0) return 0;
    B200_CHECK_ARG(masks && mask_index && ref_boxes && gauss_w && crop_off, "segm_paste_host: null pointer");
    for (int d = 0; d < n; ++d) B200_CHECK_ARG(mask_index[d] >= 0 && mask_index[d] < n_mask_blocks, "segm_paste_host: mask_index[%d] out of range", d);
    const size_t total_out = (size_t)crop_off[n];
    B200_CHECK_ARG(total_out == 0 || crops, "segm_paste_host: null output");
    HostCtx& hc = host_ctx(); std::lock_guard<std::mutex> lock(hc.mu);
    const size_t mb = (size_t)n_mask_blocks * M * M * M * 4, tb = (size_t)b200seg_segm_gauss_table_size(M) * 8;
    int e = hc.ensure(Carver::need(mb) + Carver::need((size_t)n * 4) + Carver::need((size_t)n * 24) + Carver::need(tb) +
                      Carver::need(((size_t)n + 1) * 8) + Carver::need(total_out + 16));
    if (e) return e;
    Carver cv(hc.buf);
    float* d_m = cv.take<float>(mb / 4);
    int32_t* d_i = cv.take<int32_t>(n);
    int32_t* d_b = cv.take<int32_t>((size_t)n * 6);
    double* d_t = cv.take<double>(tb / 8);
    int64_t* d_o = cv.take<int64_t>((size_t)n + 1);
    uint8_t* d_c = cv.take<uint8_t>(total_out + 16);
    cudaStream_t st = hc.stream;
    B200_CUDA(cudaMemcpyAsync(d_m, masks, mb, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(d_i, mask_index, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(d_b, ref_boxes, (size_t)n * 24, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(d_t, gauss_w, tb, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(d_o, crop_off, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
    e = b200seg_segm_paste_dev(d_m, d_i, d_b, n, M, d_t, thresh, im_s, im_h, im_w, d_c, d_o, st);
    if (e) return e;
    if (total_out) B200_CUDA(cudaMemcpyAsync(crops, d_c, total_out, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return 0;
}
