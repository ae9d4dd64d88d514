// soma_binarize.cu -- per-instance binarization straight from the raw uint8 volume
// (tools/binarization_soma.py:78-94: crop by the int()-truncated box, normalise image and PRM (:85-91),
//  2D-Otsu tools/otsu.py:199-284 with k = -1), all instances of all volumes of a batch in one launch.
//
// One CTA per instance.  Work items are 8-voxel groups of x-rows of the box (item = (row, group)), so that
// every global access is one or two aligned 64-bit loads per 8 voxels and the only integer divisions happen
// once per pass.  Three passes over the items:
//   A  stream the crop once from global memory (image rows of the volume + the packed PRM crop), SIMD
//      byte min/max of both attributes, park the raw bytes in a shared-memory cache (items beyond the cache
//      are re-read from L2 by the later passes);
//   B  anti-diagonal histograms.  The criterion of otsu.py:226-274 only needs, per anti-diagonal s = r + c of
//      the joint histogram (r = PRM bin, c = image bin, both <= G-2), the count and the sum of c (see
//      otsu2d.cu).  The caller's normalisation and numpy's histogram2d binning are monotone maps of the raw
//      byte, folded into 256-entry lookup tables; samples in the last bin of an axis are diverted to three
//      side regions of the same shared-memory arrays through an offset baked into the table, so the inner
//      loop is two table reads, one add and two shared-memory atomics per voxel, with no predicate and no
//      running totals (the global first moments follow from the four regions afterwards);
//   C  mask from the closed form of otsu.py:276-282: background <=> raw_prm < thr[raw_image], a 256-entry
//      table built once b_max is known; evaluated two voxels per SIMD compare.
// HBM sees the crop once and the mask once; everything else is shared memory.
#include "otsu_common.cuh"

namespace b200seg {

constexpr int SB_THREADS = 256;
constexpr int SB_NW = SB_THREADS / 32;
constexpr int SB_MIN_CTAS = 4;               // 64 registers per thread; 4 x (16 KB tables + 40 KB cache) of shared memory
constexpr int SB_GMAX = 304;                 // soma levels live in [30,330] -> G <= 301
// shared-memory histogram layout (indices into cnt[] / sumc[]), r = PRM bin, c = image bin:
//   [0, 2G-3)            r,c <= G-2 : anti-diagonal s = r + c              (<= 599 entries)
//   [SB_OFF1 + r]        c == G-1   : lut_i holds SB_OFF1 instead of c     (r <= G-2)
//   [SB_OFF2 + c]        r == G-1   : lut_p holds SB_OFF2 instead of r     (c <= G-2)
//   [SB_OFF1 + SB_OFF2]  both
//   [SB_TRASH]           swallows the skipped bytes of a shifted / partial group
constexpr int SB_OFF1 = 608, SB_OFF2 = 912;
constexpr int SB_TRASH = SB_OFF1 + SB_OFF2 + 1;
constexpr int SB_HIST = SB_TRASH + 7;        // 1528 entries
constexpr int SB_CACHE_ITEMS = 2560;         // 8-voxel items per attribute kept in shared memory (20 KB each)

struct SomaShared {
    unsigned int cnt[SB_HIST];
    unsigned int sumc[SB_HIST];
    unsigned short lut_i[256];               // raw image byte -> image bin (SB_OFF1 when it is the last bin)
    unsigned short lut_p[256];               // raw PRM byte   -> PRM bin   (SB_OFF2 when it is the last bin)
    unsigned short norm_i[256];              // soma normalisation (raw uint8 -> level in [30,330])
    unsigned short norm_p[256];
    unsigned short thr[256];                 // raw image byte -> 0x8000 | number of raw PRM bytes that are background
    int red[4][SB_NW];
    int bcast[4];
    OtsuScanShared<SB_THREADS> scan;
};

// exact n / d for n * d < 2^32 with a precomputed m = 2^32 / d + 1 (d >= 2); d == 1 handled by the caller's select
__device__ __forceinline__ unsigned int fast_div(unsigned int n, unsigned int m, unsigned int d) {
    return d == 1u ? n : __umulhi(n, m);
}

// 8 bytes starting at base + off (any alignment): three aligned 32-bit loads + two funnel shifts.
// May touch up to 3 bytes past the 8 requested ones (the caller guarantees they are inside the buffer).
__device__ __forceinline__ uint2 load8_fs(const uint8_t* __restrict__ base, unsigned int off) {
    const uint8_t* a = base + off;
    const unsigned int mis = (unsigned int)(reinterpret_cast<uintptr_t>(a) & 3);
    const unsigned int* q = reinterpret_cast<const unsigned int*>(a - mis);
    const unsigned int w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
    return make_uint2(__funnelshift_r(w0, w1, mis * 8), __funnelshift_r(w1, w2, mis * 8));
}
// same result with byte loads of exactly the bytes [u_lo, u_hi) (buffer ends, rows shorter than 8 voxels)
__device__ __forceinline__ uint2 load8_bytes(const uint8_t* __restrict__ base, unsigned int off, int u_lo, int u_hi) {
    unsigned int w[2] = {0u, 0u};
#pragma unroll
    for (int u = 0; u < 8; ++u)
        if (u >= u_lo && u < u_hi) w[u >> 2] |= (unsigned int)__ldg(base + off + u) << (8 * (u & 3));
    return make_uint2(w[0], w[1]);
}

// per-16-bit-lane sign replication of bytes 1,3 of a and 1,3 of b (prmt with the replicate-sign selector bit)
__device__ __forceinline__ unsigned int prmt_signs(unsigned int a, unsigned int b) {
    unsigned int d;
    asm("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// SMALL = rows shorter than 8 voxels (one partial group per row); otherwise every group holds 8 valid voxels
// and the last group of a row is shifted left to end at the row end (it overlaps its neighbour by u_lo bytes,
// which only the histogram pass has to skip).  FAST = the 3 spare bytes of load8_fs stay inside both buffers.
template <bool SMALL, bool FAST>
__device__ __forceinline__ void soma_binarize_cta(SomaShared& sh, uint2* c_img, uint2* c_prm,
                                                   const uint8_t* __restrict__ img_vol, const uint8_t* __restrict__ prm_crop,
                                                   uint8_t* __restrict__ mout, int n, int sx, int sy, int H, int W,
                                                   unsigned int img_base, int inst,
                                                   int32_t* __restrict__ b_max_out, int32_t* __restrict__ status_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ng = SMALL ? 1 : (sx + 7) >> 3;                     // 8-voxel groups per row
    const int n_rows = n / sx;                                    // sz * sy
    const int n_items = n_rows * ng;
    const int ncache = min(n_items, SB_CACHE_ITEMS);
    const unsigned int m_ng = 0xFFFFFFFFu / (unsigned)ng + 1u, m_sy = 0xFFFFFFFFu / (unsigned)sy + 1u;
    const unsigned int HW = (unsigned)H * (unsigned)W;

    // item idx -> first voxel x0 of the group, byte offsets into the volume / the crop, first valid byte
    auto decode = [&](int idx, unsigned int& ioff, unsigned int& poff, int& u_lo) {
        const unsigned int rho = fast_div((unsigned)idx, m_ng, (unsigned)ng);
        const int g = idx - (int)rho * ng;
        const unsigned int z = fast_div(rho, m_sy, (unsigned)sy);
        const unsigned int y = rho - z * (unsigned)sy;
        const int x0 = SMALL ? 0 : min(g << 3, sx - 8);
        u_lo = SMALL ? 0 : (g << 3) - x0;
        ioff = img_base + z * HW + y * (unsigned)W + (unsigned)x0;
        poff = rho * (unsigned)sx + (unsigned)x0;
    };
    auto fetch = [&](unsigned int ioff, unsigned int poff, uint2& wi, uint2& wp) {
        if (FAST && !SMALL) { wi = load8_fs(img_vol, ioff); wp = load8_fs(prm_crop, poff); }
        else { const int hi = SMALL ? sx : 8; wi = load8_bytes(img_vol, ioff, 0, hi); wp = load8_bytes(prm_crop, poff, 0, hi); }
    };

    // ---- pass A: stream the crop once: raw min/max + fill the shared-memory cache ------------------
    // The PRM crop comes first: an instance without a positive PRM voxel is skipped by the script before it touches the
    // image (binarization_soma.py:74-76) -- about half of the NMS survivors of a volume (duplicates / false boxes) -- so the
    // image rows of such a crop are never read from HBM.
    int rmin_i, rmax_i, rmin_p, rmax_p;
    auto minmax8 = [&](uint2 w, unsigned int& mn, unsigned int& mx) {
        if (SMALL) {                                               // replicate byte 0 into the invalid bytes
            const unsigned int f0 = (w.x & 0xFFu) * 0x01010101u;
            const unsigned int v0 = sx >= 4 ? 0xFFFFFFFFu : ((1u << (8 * sx)) - 1u);
            const unsigned int v1 = sx <= 4 ? 0u : ((1u << (8 * (sx - 4))) - 1u);
            w.x = (w.x & v0) | (f0 & ~v0); w.y = (w.y & v1) | (f0 & ~v1);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const unsigned int a = h ? w.y : w.x;
            const unsigned int ae = __byte_perm(a, 0u, 0x4240), ao = __byte_perm(a, 0u, 0x4341);   // even / odd bytes
            mn = __vimin3_u16x2(mn, ae, ao); mx = __vimax3_u16x2(mx, ae, ao);
        }
    };
    auto block_minmax = [&](unsigned int mn2, unsigned int mx2, int slot, int& mn_out, int& mx_out) {
        const int mn = warp_min((int)min(mn2 & 0xFFFFu, mn2 >> 16)), mx = warp_max((int)max(mx2 & 0xFFFFu, mx2 >> 16));
        if (lane == 0) { sh.red[slot][warp] = mn; sh.red[slot + 1][warp] = mx; }
        __syncthreads();
        if (warp == 0) {
            int a = lane < SB_NW ? sh.red[slot][lane] : 255, b = lane < SB_NW ? sh.red[slot + 1][lane] : 0;
            a = warp_min(a); b = warp_max(b);
            if (lane == 0) { sh.bcast[slot] = a; sh.bcast[slot + 1] = b; }
        }
        __syncthreads();
        mn_out = sh.bcast[slot]; mx_out = sh.bcast[slot + 1];
    };
    {
        unsigned int mn_p = 0x00FF00FFu, mx_p = 0u;                // two 16-bit lanes each
#pragma unroll 2
        for (int idx = tid; idx < n_items; idx += SB_THREADS) {
            unsigned int ioff, poff; int u_lo;
            decode(idx, ioff, poff, u_lo);
            uint2 wp;
            if (FAST && !SMALL) wp = load8_fs(prm_crop, poff);
            else wp = load8_bytes(prm_crop, poff, 0, SMALL ? sx : 8);
            if (idx < ncache) c_prm[idx] = wp;
            minmax8(wp, mn_p, mx_p);
        }
        block_minmax(mn_p, mx_p, 2, rmin_p, rmax_p);
    }
    if (rmax_p == 0) {                                    // no positive PRM voxel: instance skipped (:74-76)
        for (int j = tid; j < n; j += SB_THREADS) mout[j] = 0;
        if (tid == 0) { status_out[inst] = 3; b_max_out[inst] = 0; }
        return;
    }
    {
        unsigned int mn_i = 0x00FF00FFu, mx_i = 0u;
#pragma unroll 2
        for (int idx = tid; idx < n_items; idx += SB_THREADS) {
            unsigned int ioff, poff; int u_lo;
            decode(idx, ioff, poff, u_lo);
            uint2 wi;
            if (FAST && !SMALL) wi = load8_fs(img_vol, ioff);
            else wi = load8_bytes(img_vol, ioff, 0, SMALL ? sx : 8);
            if (idx < ncache) c_img[idx] = wi;
            minmax8(wi, mn_i, mx_i);
        }
        block_minmax(mn_i, mx_i, 0, rmin_i, rmax_i);
    }

    // ---- normalisation tables (binarization_soma.py:85-91) and value range (otsu.py:201) -----------
    const int gray_max = rmax_i, prm_max = rmax_p;
    {
        // np.clip(v / gray_max * 300, 0, 300) + 30 -> astype(uint16) (truncation)
        double f = gray_max > 0 ? __dmul_rn(__ddiv_rn((double)tid, (double)gray_max), 300.0) : 0.0;
        f = fmin(fmax(f, 0.0), 300.0);
        sh.norm_i[tid] = gray_max > 0 ? (unsigned short)(int)__dadd_rn(f, 30.0) : (unsigned short)0;
        // np.round(p / max * 300 + 30) -> uint16   (round half to even)
        const double p = __dadd_rn(__dmul_rn(__ddiv_rn((double)tid, (double)prm_max), 300.0), 30.0);
        sh.norm_p[tid] = (unsigned short)(int)rint(p);
    }
    for (int s = tid; s < SB_HIST; s += SB_THREADS) { sh.cnt[s] = 0u; sh.sumc[s] = 0u; }
    __syncthreads();
    // both maps are monotone non-decreasing, so the normalised range is the image of the raw range
    const int g_min = sh.norm_i[rmin_i], g_max = sh.norm_i[rmax_i];
    const int p_min = sh.norm_p[rmin_p], p_max = sh.norm_p[rmax_p];
    const int G = g_max - g_min + 1;
    const int PR = p_max - p_min + 1;
    if (G > SB_GMAX || PR > SB_GMAX || (unsigned long long)n * (unsigned long long)(G > 1 ? G - 1 : 1) >= 0xFFFFFFFFull) {
        for (int j = tid; j < n; j += SB_THREADS) mout[j] = 0;
        if (tid == 0) { status_out[inst] = 4; b_max_out[inst] = 0; }
        return;
    }
    // ---- binning tables: raw byte -> bin, last bin of an axis diverted to a side region --------------
    {
        const int t = tid;
        int vi = 0, vp = 0;
        if (t >= rmin_i && t <= rmax_i) {
            const int c = np_axis_bin(sh.norm_i[t], g_min, g_max, G);
            vi = c == G - 1 ? SB_OFF1 : c;
        }
        if (t >= rmin_p && t <= rmax_p) {
            const int r = np_axis_bin(sh.norm_p[t], p_min, p_max, G);
            vp = r == G - 1 ? SB_OFF2 : r;
        }
        sh.lut_i[t] = (unsigned short)vi;
        sh.lut_p[t] = (unsigned short)vp;
    }
    __syncthreads();

    // ---- pass B: anti-diagonal histograms in shared memory ------------------------------------------
    for (int idx = tid; idx < n_items; idx += SB_THREADS) {
        uint2 wi, wp;
        int u_lo = 0;
        if (idx < ncache) {
            wi = c_img[idx]; wp = c_prm[idx];
            if (!SMALL) {                                          // only the last group of a row is shifted
                const unsigned int rho = fast_div((unsigned)idx, m_ng, (unsigned)ng);
                const int g = idx - (int)rho * ng;
                u_lo = (g << 3) - min(g << 3, sx - 8);
            }
        } else {
            unsigned int ioff, poff;
            decode(idx, ioff, poff, u_lo);
            fetch(ioff, poff, wi, wp);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned int bi = ((u < 4 ? wi.x : wi.y) >> (8 * (u & 3))) & 0xFFu;
            const unsigned int bp = ((u < 4 ? wp.x : wp.y) >> (8 * (u & 3))) & 0xFFu;
            const unsigned int c = sh.lut_i[bi];
            unsigned int s = c + sh.lut_p[bp];
            const bool valid = SMALL ? (u < sx) : (u >= u_lo);
            s = valid ? s : (unsigned)SB_TRASH;
            atomicAdd(&sh.cnt[s], 1u);
            atomicAdd(&sh.sumc[s], c);
        }
    }
    __syncthreads();                                  // histograms complete
    // ---- scan over b (otsu.py:226-274, closed form) ------------------------------------------------
    // first moments of the samples that sit in the side regions (their bins are known from the region):
    unsigned long long side_c = 0ull, side_r = 0ull;
    for (int e = SB_OFF1 + tid; e < SB_TRASH; e += SB_THREADS) {
        const unsigned long long cn = sh.cnt[e];
        if (cn == 0ull) continue;
        if (e < SB_OFF2) { side_c += (unsigned long long)(G - 1) * cn; side_r += (unsigned long long)(e - SB_OFF1) * cn; }
        else if (e < SB_OFF1 + SB_OFF2) { side_c += sh.sumc[e]; side_r += (unsigned long long)(G - 1) * cn; }
        else { side_c += (unsigned long long)(G - 1) * cn; side_r += (unsigned long long)(G - 1) * cn; }
    }
    int b_max, found;
    otsu_scan_b<SB_THREADS>(sh.cnt, sh.sumc, 2 * G - 1, side_c, side_r, false, n, g_min, g_max, p_min, p_max, sh.scan, b_max, found);
    if (tid == 0) { b_max_out[inst] = b_max; status_out[inst] = found ? 0 : 1; }

    // ---- pass C: mask (otsu.py:276-282, closed form) -----------------------------------------------
    // background <=> I < min(b_max - g_min, g_max)  and  P < min(b_max - I, g_max + 1), I = norm_i[ri], P = norm_p[rp].
    // norm_p is monotone in rp, so per raw image level: background <=> rp < thr[ri] with
    // thr[ri] = #{rp in 0..255 : norm_p[rp] < min(b_max - I, g_max + 1)} (0 when I >= x_hi or no threshold was found;
    // a failed instance pastes nothing in the chain).  Two voxels per 32-bit subtract: lane = (0x8000 | thr) - rp - 1
    // keeps bit 15 exactly when rp < thr.
    {
        int t = 0;
        if (found) {
            const int I = sh.norm_i[tid];
            const int x_hi = min(b_max - g_min, g_max);
            if (I < x_hi) {
                const int T = min(b_max - I, g_max + 1);
                int lo = 0, hi = 256;                               // first rp with norm_p[rp] >= T
                while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)sh.norm_p[mid] < T) lo = mid + 1; else hi = mid; }
                t = lo;
            }
        }
        sh.thr[tid] = (unsigned short)(0x8000 | t);
    }
    __syncthreads();
    for (int idx = tid; idx < n_items; idx += SB_THREADS) {
        unsigned int ioff, poff; int u_lo;
        decode(idx, ioff, poff, u_lo);
        uint2 wi, wp;
        if (idx < ncache) { wi = c_img[idx]; wp = c_prm[idx]; }
        else fetch(ioff, poff, wi, wp);
        unsigned int m[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const unsigned int a = h ? wi.y : wi.x, b = h ? wp.y : wp.x;
            const unsigned int t01 = (unsigned)sh.thr[a & 0xFFu] | ((unsigned)sh.thr[(a >> 8) & 0xFFu] << 16);
            const unsigned int t23 = (unsigned)sh.thr[(a >> 16) & 0xFFu] | ((unsigned)sh.thr[a >> 24] << 16);
            const unsigned int d01 = t01 - __byte_perm(b, 0u, 0x4140) - 0x00010001u;      // halves: rp0, rp1
            const unsigned int d23 = t23 - __byte_perm(b, 0u, 0x4342) - 0x00010001u;      // halves: rp2, rp3
            m[h] = ~prmt_signs(d01, d23);                                                  // 255 = foreground
        }
        uint8_t* dst = mout + poff;
        if (!SMALL && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0)) {
            reinterpret_cast<unsigned int*>(dst)[0] = m[0];
            reinterpret_cast<unsigned int*>(dst)[1] = m[1];
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) if (!SMALL || u < sx) dst[u] = (uint8_t)(m[u >> 2] >> (8 * (u & 3)));
        }
    }
}

// grid = (slots, volumes): instance `slot` (visit order) of volume blockIdx.y
__global__ void __launch_bounds__(SB_THREADS, SB_MIN_CTAS)
soma_binarize_kernel(const uint8_t* __restrict__ volumes, const uint8_t* __restrict__ prm,
                     const int64_t* __restrict__ crop_off, int n_crops, int n_volumes,
                     int S, int H, int W, const int32_t* __restrict__ det_off, const int32_t* __restrict__ boxes,
                     const int32_t* __restrict__ order, const int32_t* __restrict__ n_valid,
                     uint8_t* __restrict__ mask, int32_t* __restrict__ b_max_out, int32_t* __restrict__ status_out) {
    __shared__ SomaShared sh;
    extern __shared__ __align__(16) unsigned char s_cache[];
    uint2* c_img = reinterpret_cast<uint2*>(s_cache);
    uint2* c_prm = c_img + SB_CACHE_ITEMS;

    const int slot = blockIdx.x, vol = blockIdx.y;
    const int base = det_off ? det_off[vol] : 0;
    const int n_here = det_off ? det_off[vol + 1] - base : n_crops;
    if (slot >= n_here || (n_valid && slot >= n_valid[vol])) return;
    const int inst = base + (order ? order[base + slot] : slot);

    const int64_t off = crop_off[inst];
    const int n = (int)(crop_off[inst + 1] - off);               // voxels of this crop (< 2^31)
    const int32_t* bb = boxes + 6 * (size_t)inst;
    const int bx1 = bb[0], by1 = bb[1], bz1 = bb[2];
    const int sx = bb[3] - bx1 + 1, sy = bb[4] - by1 + 1;
    if (n <= 0 || sx <= 0 || sy <= 0) {
        if (threadIdx.x == 0) { status_out[inst] = 2; b_max_out[inst] = 0; }
        return;
    }
    const int total_inst = det_off ? det_off[n_volumes] : n_crops;
    const size_t V = (size_t)S * H * W;                           // < 2^31 (checked by the launcher)
    const uint8_t* img_vol = volumes + (size_t)vol * V;
    const uint8_t* prm_crop = prm + off;
    const unsigned int img_base = ((unsigned)bz1 * (unsigned)H + (unsigned)by1) * (unsigned)W + (unsigned)bx1;
    // the funnel-shift loads may touch 3 bytes past a group and 3 before it (aligned word start): both buffers
    // must extend that far beyond this crop, which fails only for the last rows of the last volume / last crop
    const int n_rows = n / sx;
    const bool fast = (vol + 1 < n_volumes || (size_t)img_base + (size_t)(n_rows / sy - 1) * H * W + (size_t)(sy - 1) * W + sx + 3 <= V) &&
                      (inst + 1 < total_inst) && (vol > 0 || img_base >= 3u) && (off >= 3);
    uint8_t* mout = mask + off;
    // fast_div exactness needs numerator * divisor < 2^32
    if ((unsigned long long)n_rows * (unsigned long long)((sx + 7) >> 3) * (unsigned long long)((sx + 7) >> 3) >= 0xFFFFFFFFull ||
        (unsigned long long)n_rows * (unsigned long long)sy >= 0xFFFFFFFFull) {
        for (int j = threadIdx.x; j < n; j += SB_THREADS) mout[j] = 0;
        if (threadIdx.x == 0) { status_out[inst] = 4; b_max_out[inst] = 0; }
        return;
    }
    if (sx < 8) soma_binarize_cta<true, false>(sh, c_img, c_prm, img_vol, prm_crop, mout, n, sx, sy, H, W, img_base, inst, b_max_out, status_out);
    else if (fast) soma_binarize_cta<false, true>(sh, c_img, c_prm, img_vol, prm_crop, mout, n, sx, sy, H, W, img_base, inst, b_max_out, status_out);
    else soma_binarize_cta<false, false>(sh, c_img, c_prm, img_vol, prm_crop, mout, n, sx, sy, H, W, img_base, inst, b_max_out, status_out);
}

constexpr int SB_CACHE_BYTES = 2 * SB_CACHE_ITEMS * 8;

}  // namespace b200seg

using namespace b200seg;

extern "C" int b200seg_soma_binarize_dev(const uint8_t* volumes, int n_volumes, int S, int H, int W,
                                         const int32_t* det_off, int n_max, const int32_t* boxes,
                                         const uint8_t* prm, const int64_t* crop_off,
                                         const int32_t* order, const int32_t* n_valid, uint8_t* mask,
                                         int32_t* b_max, int32_t* status, b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(n_max >= 0 && n_volumes >= 0 && S > 0 && H > 0 && W > 0, "soma_binarize: bad sizes");
    if (n_max == 0 || n_volumes == 0) return 0;
    B200_CHECK_ARG(n_volumes == 1 || det_off, "soma_binarize: det_off is required for more than one volume");
    B200_CHECK_ARG(n_volumes <= 65535, "soma_binarize: too many volumes in one call");
    B200_CHECK_ARG((unsigned long long)S * H * W < (1ull << 31), "soma_binarize: volume too large (>= 2^31 voxels)");
    B200_CHECK_ARG(volumes && boxes && prm && crop_off && mask && b_max && status, "soma_binarize: null pointer");
    static OncePerDevice attr;
    int attr_dev;
    if (attr.needed(&attr_dev)) {
        B200_CUDA(cudaFuncSetAttribute(soma_binarize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_CACHE_BYTES));
        attr.mark(attr_dev);
    }
    dim3 grid(n_max, n_volumes);
    soma_binarize_kernel<<<grid, SB_THREADS, SB_CACHE_BYTES, stream>>>(volumes, prm, crop_off, n_max, n_volumes, S, H, W,
                                                                        det_off, boxes, order, n_valid, mask, b_max, status);
    B200_LAUNCH_CHECK("soma_binarize_kernel");
    return 0;
}
