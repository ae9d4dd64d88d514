// nuclei.cu -- per-instance chain of tools/binarization_nuclei.py:92-149 for the instances that survived the
// detection filters (:72-87, host logic + 3D NMS by volume), visited in the order given:
//   crop image and PRM by the clamped box (:98-109), normalise both into one intensity range (:111-121),
//   2D-Otsu (:124, otsu2d.cu), keep the largest 26-connected component (:126-130, cc3d), fill holes = everything but the
//   largest 26-connected component of the complement (:132-137), binary closing with the 6-neighbour cross (:139,
//   skimage.morphology.binary_closing = dilation with an unset border, then erosion with a set border), first-come
//   label paste (:141-147) and the survivor test (:148-149).
// New kernels here: the normalisation (min/max pass, then lookup tables built in fp64 in numpy's operation order), the mask
// complement and the two halves of the closing, each on a grid of (instance, row chunk) CTAs; Otsu, connected components and paste are the kernels of the soma chain.
//
// Normalisation, per crop (numpy semantics of :111-121, image dtype uint8 or uint16, PRM uint8):
//   if (gray_max - gray_min + 1 < 400):  img' = uint16(trunc(img / gray_max * 400)) + gray_min
//   prm' = uint16(rint((prm - prm_min) / (prm_max - prm_min) * (gray_max' - gray_min') + gray_min'))
// Both are monotone maps of the raw value, so gray_min' / gray_max' are the images of the raw extremes.
// status 7: the reference divides 0 by 0 here (gray_max == 0 with a stretch, or a constant PRM crop) and continues with
// NaN-derived garbage; the instance is skipped instead.  status 2: box outside the volume / inconsistent crop size.
#include "otsu_common.cuh"

namespace b200seg {

constexpr int NU_THREADS = 256;
constexpr int NU_NW = NU_THREADS / 32;

// crop geometry shared by the kernels below; false -> status 2
struct NuCrop { int bx, by, bz, sx, sy, sz; long long off; };
__device__ __forceinline__ bool nu_crop(const int32_t* __restrict__ boxes, const int64_t* __restrict__ crop_off, int inst, int S, int H, int W, NuCrop& c) {
    const int32_t* bb = boxes + 6 * (size_t)inst;
    c.bx = bb[0]; c.by = bb[1]; c.bz = bb[2];
    c.sx = bb[3] - c.bx + 1; c.sy = bb[4] - c.by + 1; c.sz = bb[5] - c.bz + 1;
    c.off = crop_off[inst];
    const long long n = crop_off[inst + 1] - c.off;
    return !(c.sx <= 0 || c.sy <= 0 || c.sz <= 0 || c.bx < 0 || c.by < 0 || c.bz < 0 || bb[3] >= W || bb[4] >= H || bb[5] >= S ||
             n != (long long)c.sx * c.sy * c.sz || n >= (1ll << 31));
}

// grid (instances, row chunks): min / max of the image crop and of the PRM crop -> mm[inst] = {gmin, pmin, gmax, pmax}
// (mins pre-set to 0x7F7F7F7F, maxes to 0 by the launcher)
template <typename T>
__global__ void __launch_bounds__(NU_THREADS) nuclei_minmax_kernel(const T* __restrict__ vol, int S, int H, int W,
                                                                    const int32_t* __restrict__ boxes, const uint8_t* __restrict__ prm,
                                                                    const int64_t* __restrict__ crop_off, int* __restrict__ mins,
                                                                    int* __restrict__ maxs, int32_t* __restrict__ nstat) {
    __shared__ int s_red[4][NU_NW];
    const int inst = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    NuCrop c;
    if (!nu_crop(boxes, crop_off, inst, S, H, W, c)) {
        if (tid == 0 && blockIdx.y == 0) nstat[inst] = 2;
        return;
    }
    const int rows = c.sy * c.sz;
    const int r0 = (int)((long long)rows * blockIdx.y / gridDim.y), r1 = (int)((long long)rows * (blockIdx.y + 1) / gridDim.y);
    const size_t HW = (size_t)H * W;
    int mn_i = 0x7FFFFFFF, mx_i = 0, mn_p = 255, mx_p = 0;
    for (int r = r0 + warp; r < r1; r += NU_NW) {
        const int z = r / c.sy, y = r - z * c.sy;
        const T* src = vol + (size_t)(c.bz + z) * HW + (size_t)(c.by + y) * W + c.bx;
        const uint8_t* ps = prm + c.off + (size_t)r * c.sx;
        for (int x = lane; x < c.sx; x += 32) {
            const int v = (int)src[x], p = (int)ps[x];
            mn_i = min(mn_i, v); mx_i = max(mx_i, v); mn_p = min(mn_p, p); mx_p = max(mx_p, p);
        }
    }
    mn_i = warp_min(mn_i); mx_i = warp_max(mx_i); mn_p = warp_min(mn_p); mx_p = warp_max(mx_p);
    if (lane == 0) { s_red[0][warp] = mn_i; s_red[1][warp] = mx_i; s_red[2][warp] = mn_p; s_red[3][warp] = mx_p; }
    __syncthreads();
    if (tid == 0) {
        int a = s_red[0][0], b = s_red[1][0], cc = s_red[2][0], d = s_red[3][0];
        for (int w = 1; w < NU_NW; ++w) { a = min(a, s_red[0][w]); b = max(b, s_red[1][w]); cc = min(cc, s_red[2][w]); d = max(d, s_red[3][w]); }
        if (r1 > r0) {
            atomicMin(&mins[2 * inst], a); atomicMin(&mins[2 * inst + 1], cc);
            atomicMax(&maxs[2 * inst], b); atomicMax(&maxs[2 * inst + 1], d);
        }
    }
}

// grid (instances, row chunks): the two normalisation maps as shared-memory tables, then two lookups per voxel
template <typename T>
__global__ void __launch_bounds__(NU_THREADS) nuclei_normalise_kernel(const T* __restrict__ vol, int S, int H, int W,
                                                                       const int32_t* __restrict__ boxes, const uint8_t* __restrict__ prm,
                                                                       const int64_t* __restrict__ crop_off, const int* __restrict__ mins,
                                                                       const int* __restrict__ maxs,
                                                                       uint16_t* __restrict__ img16, uint16_t* __restrict__ prm16,
                                                                       int32_t* __restrict__ nstat) {
    __shared__ unsigned short s_lut_i[400], s_lut_p[256];
    const int inst = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    NuCrop c;
    if (!nu_crop(boxes, crop_off, inst, S, H, W, c)) return;                 // status 2 already reported
    const int gmin = mins[2 * inst], pmin = mins[2 * inst + 1], gmax = maxs[2 * inst], pmax = maxs[2 * inst + 1];
    // gray_range = gray_max - gray_min + 1 with numpy 1.x scalar semantics (the sum is promoted, no wrap-around)
    const bool stretch = gmax - gmin + 1 < 400;
    // Where the script divides 0 by 0 -- gray_max == 0 under the stretch (an all-zero crop), or a constant PRM crop -- numpy
    // yields NaN and the following .astype(np.uint16) turns it into 0 (x86-64 / aarch64); the script carries on with the
    // zeros and so do these tables (round 1 reported status 7 here).
    const bool img_nan = stretch && gmax == 0, prm_nan = pmax == pmin;
    auto f_img = [&](int v) -> int {                           // (box_img / gray_max * 400).astype(uint16) + gray_min
        if (!stretch) return v;
        if (img_nan) return gmin & 0xFFFF;                     // uint16(NaN) + gray_min
        const double q = __dmul_rn(__ddiv_rn((double)v, (double)gmax), 400.0);
        return ((int)q + gmin) & 0xFFFF;
    };
    const int gmin2 = f_img(gmin), gmax2 = f_img(gmax);
    if (stretch) for (int k = tid; k < 400 && gmin + k <= gmax; k += NU_THREADS) s_lut_i[k] = (unsigned short)f_img(gmin + k);
    {
        const double num = (double)(tid - pmin), den = (double)(pmax - pmin), span = (double)((gmax2 - gmin2) & 0xFFFF);
        const double t = __dadd_rn(__dmul_rn(__ddiv_rn(num, den), span), (double)gmin2);
        s_lut_p[tid] = (!prm_nan && tid >= pmin && tid <= pmax) ? (unsigned short)(int)rint(t) : (unsigned short)0;
    }
    __syncthreads();
    const int rows = c.sy * c.sz;
    const int r0 = (int)((long long)rows * blockIdx.y / gridDim.y), r1 = (int)((long long)rows * (blockIdx.y + 1) / gridDim.y);
    const size_t HW = (size_t)H * W;
    for (int r = r0 + warp; r < r1; r += NU_NW) {
        const int z = r / c.sy, y = r - z * c.sy;
        const T* src = vol + (size_t)(c.bz + z) * HW + (size_t)(c.by + y) * W + c.bx;
        const size_t o = (size_t)c.off + (size_t)r * c.sx;
        for (int x = lane; x < c.sx; x += 32) {
            const int v = (int)src[x];
            img16[o + x] = stretch ? s_lut_i[v - gmin] : (unsigned short)v;
            prm16[o + x] = s_lut_p[prm[o + x]];
        }
    }
}

// mask byte -> its complement (0 -> 255, set -> 0)
__global__ void __launch_bounds__(256) mask_complement_kernel(uint8_t* __restrict__ m, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nvec = (reinterpret_cast<uintptr_t>(m) & 15) == 0 ? n >> 4 : 0;                 // 16 mask bytes per thread
    auto inv = [](unsigned x) -> unsigned {                  // per byte: 0 -> 0xFF, non-zero -> 0
        const unsigned t = (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;                    // bit 7 = byte != 0
        return ~((t >> 7) * 0xFFu);
    };
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < nvec; g += stride) {
        uint4 v = *reinterpret_cast<const uint4*>(m + g * 16);
        v.x = inv(v.x); v.y = inv(v.y); v.z = inv(v.z); v.w = inv(v.w);
        *reinterpret_cast<uint4*>(m + g * 16) = v;
    }
    for (long long i = nvec * 16 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) m[i] = m[i] ? (uint8_t)0 : (uint8_t)255;
}

// one half of the binary closing with the 6-neighbour cross, grid (instances, row chunks):
// ERODE = false: dst = dilate(src) (outside the crop = unset); ERODE = true: dst = erode(src) (outside = set).
// The erosion pass also empties the masks of instances whose status is not 0 (they paste nothing).
template <bool ERODE>
__global__ void __launch_bounds__(NU_THREADS) nuclei_morph_kernel(const uint8_t* __restrict__ src_all, uint8_t* __restrict__ dst_all,
                                                                   const int32_t* __restrict__ boxes, const int64_t* __restrict__ crop_off,
                                                                   const int32_t* __restrict__ status) {
    const int inst = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t off = crop_off[inst];
    const long long n = crop_off[inst + 1] - off;
    const uint8_t* s = src_all + off;
    uint8_t* d = dst_all + off;
    if (status[inst] != 0) {
        if (ERODE) {
            const long long j0 = n * blockIdx.y / gridDim.y, j1 = n * (blockIdx.y + 1) / gridDim.y;
            for (long long j = j0 + tid; j < j1; j += NU_THREADS) d[j] = 0;
        }
        return;
    }
    const int32_t* bb = boxes + 6 * (size_t)inst;
    const int sx = bb[3] - bb[0] + 1, sy = bb[4] - bb[1] + 1, sz = bb[5] - bb[2] + 1;
    const int rows = sy * sz, plane = sy * sx;
    const int r0 = (int)((long long)rows * blockIdx.y / gridDim.y), r1 = (int)((long long)rows * (blockIdx.y + 1) / gridDim.y);
    for (int r = r0 + warp; r < r1; r += NU_NW) {
        const int z = r / sy, y = r - z * sy;
        const uint8_t* c = s + (size_t)r * sx;
        for (int x = lane; x < sx; x += 32) {
            unsigned v = c[x];
            if (ERODE) {
                if (x > 0) v &= c[x - 1];
                if (x + 1 < sx) v &= c[x + 1];
                if (y > 0) v &= c[x - sx];
                if (y + 1 < sy) v &= c[x + sx];
                if (z > 0) v &= c[x - plane];
                if (z + 1 < sz) v &= c[x + plane];
            } else {
                if (x > 0) v |= c[x - 1];
                if (x + 1 < sx) v |= c[x + 1];
                if (y > 0) v |= c[x - sx];
                if (y + 1 < sy) v |= c[x + sx];
                if (z > 0) v |= c[x - plane];
                if (z + 1 < sz) v |= c[x + plane];
            }
            d[(size_t)r * sx + x] = v ? (uint8_t)255 : (uint8_t)0;
        }
    }
}

__global__ void nuclei_iota_u16_kernel(uint16_t* p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint16_t)(i + 1);
}

// rejected by the normalisation (status 2 / 7): that verdict overrides whatever Otsu said about the zero-filled crop
__global__ void nuclei_merge_status_kernel(const int32_t* __restrict__ nstat, int32_t* __restrict__ status, int32_t* __restrict__ b_max, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && nstat[i] != 0) { status[i] = nstat[i]; b_max[i] = 0; }
}

struct NucleiWs { size_t img16, prm16, tmp, ginfo, nstat, mins, maxs, ids, cc, cc_bytes, paste, paste_bytes, total; };
static NucleiWs nuclei_ws(int n, long long total_vox, int S, int H, int W) {
    NucleiWs w;
    const size_t tv = (size_t)(total_vox > 0 ? total_vox : 0), nn = (size_t)(n > 0 ? n : 1);
    size_t o = 0;
    w.img16 = o; o += align_up(tv * 2 + 16, 256);
    w.prm16 = o; o += align_up(tv * 2 + 16, 256);
    w.tmp = o; o += align_up(tv + 16, 256);
    w.ginfo = o; o += align_up(nn * 16, 256);
    w.nstat = o; o += align_up(nn * 4, 256);
    w.mins = o; o += align_up(nn * 8, 256);
    w.maxs = o; o += align_up(nn * 8, 256);
    w.ids = o; o += align_up(nn * 2, 256);
    w.cc_bytes = align_up(b200seg_largest_cc_workspace_bytes((long long)tv), 256);
    w.cc = o; o += w.cc_bytes;
    w.paste_bytes = align_up(b200seg_paste_labels_workspace_bytes(1, S, H, W, n), 256);
    w.paste = o; o += w.paste_bytes;
    w.total = o + 256;
    return w;
}

}  // namespace b200seg

using namespace b200seg;

extern "C" size_t b200seg_binarize_nuclei_workspace_bytes(int n, long long total_voxels, int S, int H, int W) {
    return nuclei_ws(n, total_voxels, S, H, W).total;
}

extern "C" int b200seg_binarize_nuclei_dev(const void* volume, int elem_bytes, int S, int H, int W,
                                           const int32_t* boxes, const uint8_t* prm, const int64_t* crop_off, int n, long long total_voxels,
                                           uint16_t* seg, uint8_t* masks, int32_t* b_max, int32_t* status, uint8_t* survive,
                                           void* workspace, size_t workspace_bytes, b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(S > 0 && H > 0 && W > 0 && n >= 0 && total_voxels >= 0 && (elem_bytes == 1 || elem_bytes == 2), "binarize_nuclei: bad sizes");
    B200_CHECK_ARG(n < 65535, "binarize_nuclei: more than 65534 instances do not fit uint16 labels");
    B200_CHECK_ARG(volume && seg && workspace, "binarize_nuclei: null pointer");
    B200_CHECK_ARG(n == 0 || (boxes && prm && crop_off && masks && b_max && status && survive), "binarize_nuclei: null pointer");
    const NucleiWs w = nuclei_ws(n, total_voxels, S, H, W);
    if (workspace_bytes < w.total) { set_error("binarize_nuclei: workspace too small"); return B200SEG_EWORKSPACE; }
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    uint16_t* img16 = (uint16_t*)(ws + w.img16);
    uint16_t* prm16 = (uint16_t*)(ws + w.prm16);
    uint8_t* tmp = (uint8_t*)(ws + w.tmp);
    int32_t* ginfo = (int32_t*)(ws + w.ginfo);
    int32_t* nstat = (int32_t*)(ws + w.nstat);
    uint16_t* ids = (uint16_t*)(ws + w.ids);
    if (n > 0) {
        B200_CUDA(cudaMemsetAsync(nstat, 0, sizeof(int32_t) * (size_t)n, stream));
        B200_CUDA(cudaMemsetAsync(img16, 0, (size_t)total_voxels * 2, stream));     // rejected crops stay constant
        B200_CUDA(cudaMemsetAsync(prm16, 0, (size_t)total_voxels * 2, stream));
        int* mins = (int*)(ws + w.mins);
        int* maxs = (int*)(ws + w.maxs);
        B200_CUDA(cudaMemsetAsync(mins, 0x7F, sizeof(int) * 2 * (size_t)n, stream));
        B200_CUDA(cudaMemsetAsync(maxs, 0, sizeof(int) * 2 * (size_t)n, stream));
        int nch = (4 * num_sms() + n - 1) / n;                // row chunks per instance: ~4 CTAs per SM in total
        nch = nch < 1 ? 1 : (nch > 64 ? 64 : nch);
        const dim3 cgrid2((unsigned)n, (unsigned)nch);
        if (elem_bytes == 1) {
            nuclei_minmax_kernel<uint8_t><<<cgrid2, NU_THREADS, 0, stream>>>((const uint8_t*)volume, S, H, W, boxes, prm, crop_off, mins, maxs, nstat);
            B200_LAUNCH_CHECK("nuclei_minmax_kernel");
            nuclei_normalise_kernel<uint8_t><<<cgrid2, NU_THREADS, 0, stream>>>((const uint8_t*)volume, S, H, W, boxes, prm, crop_off, mins, maxs, img16, prm16, nstat);
        } else {
            nuclei_minmax_kernel<uint16_t><<<cgrid2, NU_THREADS, 0, stream>>>((const uint16_t*)volume, S, H, W, boxes, prm, crop_off, mins, maxs, nstat);
            B200_LAUNCH_CHECK("nuclei_minmax_kernel");
            nuclei_normalise_kernel<uint16_t><<<cgrid2, NU_THREADS, 0, stream>>>((const uint16_t*)volume, S, H, W, boxes, prm, crop_off, mins, maxs, img16, prm16, nstat);
        }
        B200_LAUNCH_CHECK("nuclei_normalise_kernel");
        int e = b200seg_otsu2d_dev(img16, prm16, crop_off, n, masks, b_max, ginfo, status, nullptr, nullptr, stream);      // :124
        if (e) return e;
        nuclei_merge_status_kernel<<<(n + 255) / 256, 256, 0, stream>>>(nstat, status, b_max, n);
        B200_LAUNCH_CHECK("nuclei_merge_status_kernel");
        const unsigned cgrid = (unsigned)((total_voxels + 256 * 16 - 1) / (256 * 16) > 0 ? (total_voxels + 256 * 16 - 1) / (256 * 16) : 1);
        // :126-130 largest component; :132-137 hole filling = the same on the complement; ties -> first label (np.argmax)
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) {
                mask_complement_kernel<<<cgrid, 256, 0, stream>>>(masks, total_voxels);
                B200_LAUNCH_CHECK("mask_complement_kernel");
            }
            e = b200seg_largest_cc_ex_dev(masks, crop_off, total_voxels, 1, nullptr, n, boxes, nullptr, nullptr, status, 1, ws + w.cc, w.cc_bytes, stream);
            if (e) return e;
        }
        mask_complement_kernel<<<cgrid, 256, 0, stream>>>(masks, total_voxels);
        B200_LAUNCH_CHECK("mask_complement_kernel");
        nuclei_morph_kernel<false><<<cgrid2, NU_THREADS, 0, stream>>>(masks, tmp, boxes, crop_off, status);               // :139 dilation
        B200_LAUNCH_CHECK("nuclei_morph_kernel<dilate>");
        nuclei_morph_kernel<true><<<cgrid2, NU_THREADS, 0, stream>>>(tmp, masks, boxes, crop_off, status);                // :139 erosion
        B200_LAUNCH_CHECK("nuclei_morph_kernel<erode>");
        nuclei_iota_u16_kernel<<<(n + 255) / 256, 256, 0, stream>>>(ids, n);                                                 // mask_id, :93-94
        B200_LAUNCH_CHECK("nuclei_iota_u16_kernel");
    }
    return b200seg_paste_labels_dev(seg, 1, S, H, W, nullptr, n, boxes, ids, masks, crop_off, nullptr, nullptr, survive,  // :141-149
                                    ws + w.paste, w.paste_bytes, stream);
}

// numpy seam: one volume, every pointer on the host
extern "C" int b200seg_binarize_nuclei_host(const void* volume, int elem_bytes, int S, int H, int W,
                                            const int32_t* boxes, const uint8_t* prm, const int64_t* crop_off, int n,
                                            uint16_t* seg, uint8_t* masks, int32_t* b_max, int32_t* status, uint8_t* survive) {
    B200_CHECK_ARG(S > 0 && H > 0 && W > 0 && n >= 0 && (elem_bytes == 1 || elem_bytes == 2) && volume && seg, "binarize_nuclei_host: bad arguments");
    B200_CHECK_ARG(n == 0 || (boxes && prm && crop_off && b_max && status && survive), "binarize_nuclei_host: null pointer");
    HostCtx& hc = host_ctx(); std::lock_guard<std::mutex> lock(hc.mu);
    const size_t V = (size_t)S * H * W, nn = n > 0 ? n : 1;
    const long long tv = n > 0 ? (long long)crop_off[n] : 0;
    B200_CHECK_ARG(tv >= 0, "binarize_nuclei_host: bad crop offsets");
    const size_t ws_bytes = b200seg_binarize_nuclei_workspace_bytes(n, tv, S, H, W);
    int e = hc.ensure(Carver::need(V * elem_bytes) + Carver::need(V * 2) + Carver::need(nn * 24) + 2 * Carver::need((size_t)tv + 16) +
                      Carver::need((nn + 1) * 8) + 2 * Carver::need(nn * 4) + Carver::need(nn) + ws_bytes);
    if (e) return e;
    Carver cv(hc.buf);
    uint8_t* d_vol = cv.take<uint8_t>(V * elem_bytes);
    uint16_t* d_seg = cv.take<uint16_t>(V);
    int32_t* d_boxes = cv.take<int32_t>(nn * 6);
    uint8_t* d_prm = cv.take<uint8_t>((size_t)tv + 16);
    uint8_t* d_mask = cv.take<uint8_t>((size_t)tv + 16);
    int64_t* d_off = cv.take<int64_t>(nn + 1);
    int32_t* d_bmax = cv.take<int32_t>(nn);
    int32_t* d_stat = cv.take<int32_t>(nn);
    uint8_t* d_surv = cv.take<uint8_t>(nn);
    void* d_ws = cv.p;
    cudaStream_t st = hc.stream;
    B200_CUDA(cudaMemcpyAsync(d_vol, volume, V * elem_bytes, cudaMemcpyHostToDevice, st));
    if (n > 0) {
        B200_CUDA(cudaMemcpyAsync(d_boxes, boxes, (size_t)n * 24, cudaMemcpyHostToDevice, st));
        B200_CUDA(cudaMemcpyAsync(d_prm, prm, (size_t)tv, cudaMemcpyHostToDevice, st));
        B200_CUDA(cudaMemcpyAsync(d_off, crop_off, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
    }
    e = b200seg_binarize_nuclei_dev(d_vol, elem_bytes, S, H, W, d_boxes, d_prm, d_off, n, tv, d_seg, d_mask, d_bmax, d_stat, d_surv, d_ws, ws_bytes, st);
    if (e) return e;
    B200_CUDA(cudaMemcpyAsync(seg, d_seg, V * 2, cudaMemcpyDeviceToHost, st));
    if (n > 0) {
        if (masks) B200_CUDA(cudaMemcpyAsync(masks, d_mask, (size_t)tv, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaMemcpyAsync(b_max, d_bmax, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaMemcpyAsync(status, d_stat, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaMemcpyAsync(survive, d_surv, (size_t)n, cudaMemcpyDeviceToHost, st));
    }
    B200_CUDA(cudaStreamSynchronize(st));
    return 0;
}
