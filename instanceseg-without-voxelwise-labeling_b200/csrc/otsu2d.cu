// otsu2d.cu -- batched per-instance 2D-Otsu binarization on uint16 crops
// (replaces tools/otsu.py:199-284 `otsu_py_2d_fast`, k = -1; the fused soma variant that starts from the
//  raw uint8 volume lives in soma_binarize.cu).
//
// One CTA per instance crop; all crops go in one launch.
// The reference's O(G) Python scan over b, each step summing histogram cells, collapses to a
// closed form: the background region of line y = -x + b is { (r,c) : r + c < b - 2*g_min,
// r <= G-2, c <= G-2 } (r = PRM bin, c = image bin; derived from otsu.py:232-235,251-253), so the
// criterion only needs ANTI-DIAGONAL sums of the joint histogram.  Per crop we therefore build,
// in shared memory, two integer histograms over s = r + c: the count and the sum of c.  Exact
// integer prefix sums over s then give p0 and both first moments for every b; the between-class
// criterion is evaluated in fp64 for all b in parallel and reduced with the reference's
// first-strictly-greater rule (otsu.py:247-250,271-274) -- otsu_common.cuh:otsu_scan_b.
// The G x G joint histogram itself is never materialised (optional debug output only).
//
// Data movement: pass A streams the crop ONCE from global memory, records min/max and parks the
// samples in a shared-memory crop cache; the histogram and mask passes then run out of shared
// memory (samples beyond the cache capacity fall back to L2).  HBM sees the crop once and the
// mask once.
//
// numpy.histogram2d binning is reproduced exactly (fp64 linspace edges i*step+start with the last
// edge pinned, searchsorted-right, right edge inclusive; each axis over its own [min,max] split in
// G bins) through per-gray-level lookup tables.
#include "otsu_common.cuh"

namespace b200seg {

template <int GMAX>
struct OtsuShared {
    unsigned int cnt[2 * GMAX - 1];   // count per anti-diagonal s = r + c (r,c <= G-2)
    unsigned int sumc[2 * GMAX - 1];  // sum of image bin c per anti-diagonal
    unsigned short bin_i[GMAX];       // gray level (v - g_min) -> bin
    unsigned short bin_p[GMAX];
    int red[4][OTSU_NW];
    int bcast[4];
    OtsuScanShared<OTSU_THREADS> scan;
};

// image/prm are uint16 sample arrays (crop i at crop_off[i]).
template <int GMAX>
__global__ void __launch_bounds__(OTSU_THREADS, 2)
otsu2d_kernel(const uint16_t* __restrict__ image, const uint16_t* __restrict__ prm,
              const int64_t* __restrict__ crop_off, int n_crops,
              uint8_t* __restrict__ mask, int32_t* __restrict__ b_max_out, int32_t* __restrict__ g_info,
              int32_t* __restrict__ status_out, uint32_t* __restrict__ hist, const int64_t* __restrict__ hist_off,
              int cache_vox) {
    __shared__ OtsuShared<GMAX> sh;
    extern __shared__ __align__(16) unsigned char s_cache[];
    uint16_t* c_img = reinterpret_cast<uint16_t*>(s_cache);
    uint16_t* c_prm = c_img + cache_vox;

    const int inst = blockIdx.x;
    if (inst >= n_crops) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int64_t off = crop_off[inst];
    const int n = (int)(crop_off[inst + 1] - off);               // samples of this crop (< 2^31)
    const uint16_t* g_img = image + off;
    const uint16_t* g_prm = prm + off;
    uint8_t* mout = mask + off;

    if (n <= 0) {
        if (tid == 0) { status_out[inst] = 2; b_max_out[inst] = 0; }
        return;
    }
    const int ncache = min(n, cache_vox);

    // ---- pass A: stream the crop once: min/max + fill the shared-memory cache ----------------------
    int g_min = 0x7fffffff, g_max = -1, p_min = 0x7fffffff, p_max = -1;
    {
        constexpr int U = 4;
        for (int j0 = tid; j0 < n; j0 += U * OTSU_THREADS) {
            int vi[U], vp[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * OTSU_THREADS;
                vi[u] = -1; vp[u] = -1;
                if (j < n) { vi[u] = (int)g_img[j]; vp[u] = (int)g_prm[j]; }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * OTSU_THREADS;
                if (j < n) {
                    g_min = min(g_min, vi[u]); g_max = max(g_max, vi[u]);
                    p_min = min(p_min, vp[u]); p_max = max(p_max, vp[u]);
                    if (j < ncache) { c_img[j] = (uint16_t)vi[u]; c_prm[j] = (uint16_t)vp[u]; }
                }
            }
        }
        g_min = warp_min(g_min); g_max = warp_max(g_max); p_min = warp_min(p_min); p_max = warp_max(p_max);
        if (lane == 0) { sh.red[0][warp] = g_min; sh.red[1][warp] = g_max; sh.red[2][warp] = p_min; sh.red[3][warp] = p_max; }
        __syncthreads();
        if (warp == 0) {
            int a = lane < OTSU_NW ? sh.red[0][lane] : 0x7fffffff, b = lane < OTSU_NW ? sh.red[1][lane] : -1;
            int c = lane < OTSU_NW ? sh.red[2][lane] : 0x7fffffff, d = lane < OTSU_NW ? sh.red[3][lane] : -1;
            a = warp_min(a); b = warp_max(b); c = warp_min(c); d = warp_max(d);
            if (lane == 0) { sh.bcast[0] = a; sh.bcast[1] = b; sh.bcast[2] = c; sh.bcast[3] = d; }
        }
        __syncthreads();
        g_min = sh.bcast[0]; g_max = sh.bcast[1]; p_min = sh.bcast[2]; p_max = sh.bcast[3];
    }
    const int G = g_max - g_min + 1;                    // otsu.py:201-202
    const int PR = p_max - p_min + 1;
    if (tid == 0 && g_info) {
        g_info[inst * 4 + 0] = g_min; g_info[inst * 4 + 1] = g_max;
        g_info[inst * 4 + 2] = p_min; g_info[inst * 4 + 3] = p_max;
    }
    if (G > GMAX || PR > GMAX || (unsigned long long)n * (unsigned long long)(G > 1 ? G - 1 : 1) >= 0xFFFFFFFFull) {
        for (int j = tid; j < n; j += OTSU_THREADS) mout[j] = 255;
        if (tid == 0) { status_out[inst] = 4; b_max_out[inst] = 0; }
        return;
    }

    // ---- binning LUTs + clear diagonal histograms ---------------------------------------------------
    for (int v = tid; v < G; v += OTSU_THREADS) sh.bin_i[v] = (unsigned short)np_axis_bin(g_min + v, g_min, g_max, G);
    for (int v = tid; v < PR; v += OTSU_THREADS) sh.bin_p[v] = (unsigned short)np_axis_bin(p_min + v, p_min, p_max, G);
    const int ndiag = 2 * G - 1;
    for (int s = tid; s < ndiag; s += OTSU_THREADS) { sh.cnt[s] = 0u; sh.sumc[s] = 0u; }
    __syncthreads();

    // ---- pass B: anti-diagonal histograms in shared memory ------------------------------------------
    // Warp aggregation: when all 32 lanes hit the same diagonal (background runs -- the contended case)
    // one lane adds the warp's count and redux-summed column index; otherwise lanes issue their own
    // shared-memory atomics (spread addresses).
    unsigned int tot_c = 0u, tot_r = 0u;      // fit 32 bits: n * (G-1) < 2^32 is checked above
    uint32_t* hist_i = hist ? hist + hist_off[inst] : nullptr;
    auto add_sample = [&](bool in, int vi, int vp) {
        int c = 0, r = 0;
        if (in) {
            c = (int)sh.bin_i[vi - g_min];
            r = (int)sh.bin_p[vp - p_min];
            tot_c += (unsigned)c; tot_r += (unsigned)r;
            if (hist_i) atomicAdd(&hist_i[(size_t)r * G + c], 1u);
        }
        const bool ok = in && c <= G - 2 && r <= G - 2;
        const unsigned key = ok ? (unsigned)(r + c) : 0xFFFFFFFFu;
        const unsigned k0 = __shfl_sync(0xffffffffu, key, 0);
        if (__all_sync(0xffffffffu, key == k0)) {                    // warp-uniform branch
            if (k0 != 0xFFFFFFFFu) {
                const unsigned csum = __reduce_add_sync(0xffffffffu, (unsigned)c);
                if (lane == 0) { atomicAdd(&sh.cnt[k0], 32u); atomicAdd(&sh.sumc[k0], csum); }
            }
        } else if (ok) {
            atomicAdd(&sh.cnt[key], 1u);
            atomicAdd(&sh.sumc[key], (unsigned)c);
        }
    };
    {
        const int n_round = (n + OTSU_THREADS - 1) / OTSU_THREADS * OTSU_THREADS;
        for (int j = tid; j < n_round; j += OTSU_THREADS) {
            const bool in = j < n;
            int vi = 0, vp = 0;
            if (in) {
                if (j < ncache) { vi = c_img[j]; vp = c_prm[j]; }
                else { vi = g_img[j]; vp = g_prm[j]; }
            }
            add_sample(in, vi, vp);
        }
    }
    // ---- scan over b (otsu.py:226-274, closed form); the per-thread totals are reduced inside --------
    int b_max, found;
    otsu_scan_b<OTSU_THREADS>(sh.cnt, sh.sumc, ndiag, tot_c, tot_r, true, n, g_min, g_max, p_min, p_max, sh.scan, b_max, found);
    if (tid == 0) { b_max_out[inst] = b_max; status_out[inst] = found ? 0 : 1; }

    // ---- pass C: mask (otsu.py:276-282, closed form) -----------------------------------------------
    // background <=> I < min(b_max - g_min, g_max)  and  P < min(b_max - I, g_max + 1);
    // when no b wins the reference raises (otsu.py:277): the mask is then all 255.
    const int x_hi = min(b_max - g_min, g_max);
    for (int j = tid; j < n; j += OTSU_THREADS) {
        int I, Pv;
        if (j < ncache) { I = c_img[j]; Pv = c_prm[j]; } else { I = g_img[j]; Pv = g_prm[j]; }
        mout[j] = (found && I < x_hi && Pv < min(b_max - I, g_max + 1)) ? (uint8_t)0 : (uint8_t)255;
    }
}

constexpr int OTSU_GMAX_GENERIC = 2048;     // gray range supported by the generic entry (reference: G*G fp64 histogram)
constexpr int OTSU_CACHE_BYTES_GENERIC = 64 * 1024;

}  // namespace b200seg

using namespace b200seg;

extern "C" int b200seg_otsu2d_dev(const uint16_t* image, const uint16_t* prm, const int64_t* crop_off,
                                  int n_crops, uint8_t* mask, int32_t* b_max, int32_t* g_info, int32_t* status,
                                  uint32_t* hist, const int64_t* hist_off, b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(n_crops >= 0, "otsu2d: negative n_crops");
    if (n_crops == 0) return 0;
    B200_CHECK_ARG(image && prm && crop_off && mask && b_max && status, "otsu2d: null pointer");
    B200_CHECK_ARG(!hist || hist_off, "otsu2d: hist given without hist_off");
    auto kern = otsu2d_kernel<OTSU_GMAX_GENERIC>;
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, OTSU_CACHE_BYTES_GENERIC));
    kern<<<n_crops, OTSU_THREADS, OTSU_CACHE_BYTES_GENERIC, stream>>>(image, prm, crop_off, n_crops, mask, b_max, g_info, status,
                                                                      hist, hist_off, OTSU_CACHE_BYTES_GENERIC / 4);
    B200_LAUNCH_CHECK("otsu2d_kernel");
    return 0;
}
