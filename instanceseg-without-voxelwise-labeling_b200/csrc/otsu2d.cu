// otsu2d.cu -- batched per-instance 2D-Otsu binarization
// (replaces tools/otsu.py:199-284 `otsu_py_2d_fast`, k = -1, and the crop/normalise steps of
//  tools/binarization_soma.py:78-94).
//
// One CTA per instance crop; all crops of a volume (or of a batch of volumes) go in one launch.
// The reference's O(G) Python scan over b, each step summing histogram cells, collapses to a
// closed form: the background region of line y = -x + b is { (r,c) : r + c < b - 2*g_min,
// r <= G-2, c <= G-2 } (r = PRM bin, c = image bin; derived from otsu.py:232-235,251-253), so the
// criterion only needs ANTI-DIAGONAL sums of the joint histogram.  Per crop we therefore build,
// in shared memory with warp-aggregated atomics (__match_any_sync + redux), two integer histograms
// over s = r + c: the count and the sum of c.  Exact integer prefix sums over s then give p0 and
// both first moments for every b; the between-class criterion is evaluated in fp64 for all b in
// parallel and reduced with the reference's first-strictly-greater rule (otsu.py:247-250,271-274).
// The G x G joint histogram itself is never materialised (optional debug output only).
//
// Passes over the crop (L2-resident after the first touch): (0) raw max for the soma
// normalisation LUTs, (1) min/max, (2) diagonal histograms, (3) mask.  HBM sees the crop once
// and the mask once.
//
// numpy.histogram2d binning is reproduced exactly (fp64 linspace edges i*step+start with the last
// edge pinned, searchsorted-right, right edge inclusive; each axis over its own [min,max] split in
// G bins) through per-gray-level lookup tables.
#include "common.cuh"

namespace b200seg {

constexpr int OTSU_THREADS = 512;
constexpr int OTSU_GMAX = 2048;                 // gray range supported (reference: G*G fp64 histogram)
constexpr int OTSU_NDIAG = 2 * OTSU_GMAX - 1;

// bin of value v on an axis of G bins over [vmin,vmax]  (numpy histogramdd / linspace semantics)
__device__ __forceinline__ int np_axis_bin(int v, int vmin, int vmax, int G) {
    double lo = (double)vmin, hi = (double)vmax;
    if (vmin == vmax) { lo -= 0.5; hi += 0.5; }
    const double step = __ddiv_rn(__dsub_rn(hi, lo), (double)G);
    const double x = (double)v;
    int g = (int)floor(__ddiv_rn(__dsub_rn(x, lo), step));
    g = max(0, min(G - 1, g));
    // edge(i) = i*step + lo for i < G, edge(G) = hi
    while (g < G - 1 && __dadd_rn(__dmul_rn((double)(g + 1), step), lo) <= x) ++g;
    while (g > 0 && __dadd_rn(__dmul_rn((double)g, step), lo) > x) --g;
    return g;
}

__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

struct OtsuShared {
    unsigned int cnt[OTSU_NDIAG];     // count per anti-diagonal s = r + c (r,c <= G-2)
    unsigned int sumc[OTSU_NDIAG];    // sum of image bin c per anti-diagonal
    unsigned short lut_i[OTSU_GMAX];  // image gray level (v - g_min) -> bin
    unsigned short lut_p[OTSU_GMAX];  // prm level (v - p_min) -> bin
    unsigned short norm_i[256];       // soma normalisation LUTs (raw uint8 -> uint16)
    unsigned short norm_p[256];
    int red[4][OTSU_THREADS / 32];
    unsigned long long red64[3][OTSU_THREADS / 32];
    double best_var[OTSU_THREADS / 32];
    int best_b[OTSU_THREADS / 32];
    int bcast[8];
    unsigned long long tot[2];
};

// MODE 0: image/prm are uint16 sample arrays (crop i at crop_off[i]).
// MODE 1: soma fused: image is the raw uint8 volume, prm raw uint8 box crops.
template <int MODE>
__global__ void __launch_bounds__(OTSU_THREADS)
otsu2d_kernel(const void* __restrict__ image_, const void* __restrict__ prm_,
              const int64_t* __restrict__ crop_off, int n_crops,
              int S, int H, int W, const int32_t* __restrict__ boxes,
              const int32_t* __restrict__ order, const int32_t* __restrict__ n_valid,
              uint8_t* __restrict__ mask, int32_t* __restrict__ b_max_out, int32_t* __restrict__ g_info,
              int32_t* __restrict__ status_out, uint32_t* __restrict__ hist, const int64_t* __restrict__ hist_off) {
    __shared__ OtsuShared sh;
    const int slot = blockIdx.x;
    if (n_valid && slot >= *n_valid) return;
    const int inst = order ? order[slot] : slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = OTSU_THREADS / 32;

    const int64_t off = crop_off[inst];
    const long long n = (long long)(crop_off[inst + 1] - off);
    int bx1 = 0, by1 = 0, bz1 = 0, sx = 1, sy = 1;
    if (MODE == 1) {
        const int32_t* bb = boxes + 6 * inst;
        bx1 = bb[0]; by1 = bb[1]; bz1 = bb[2];
        sx = bb[3] - bb[0] + 1; sy = bb[4] - bb[1] + 1;
    }
    const uint16_t* img16 = (const uint16_t*)image_ + (MODE == 0 ? off : 0);
    const uint16_t* prm16 = (const uint16_t*)prm_ + (MODE == 0 ? off : 0);
    const uint8_t* vol8 = (const uint8_t*)image_;
    const uint8_t* prm8 = (const uint8_t*)prm_ + (MODE == 1 ? off : 0);
    uint8_t* mout = mask + off;

    if (n <= 0) {
        if (tid == 0) { status_out[inst] = 2; b_max_out[inst] = 0; }
        return;
    }

    auto raw_image = [&](long long j) -> int {
        if (MODE == 0) return (int)img16[j];
        const int x = (int)(j % sx);
        const long long t = j / sx;
        const int y = (int)(t % sy), z = (int)(t / sy);
        return (int)vol8[((size_t)(bz1 + z) * H + (by1 + y)) * W + (bx1 + x)];
    };
    auto raw_prm = [&](long long j) -> int { return MODE == 0 ? (int)prm16[j] : (int)prm8[j]; };

    // ---- pass 0 (soma): raw maxima -> normalisation LUTs (binarization_soma.py:85-91) ----------
    if (MODE == 1) {
        int mi = 0, mp = 0;
        for (long long j = tid; j < n; j += OTSU_THREADS) { mi = max(mi, raw_image(j)); mp = max(mp, raw_prm(j)); }
        mi = warp_max(mi); mp = warp_max(mp);
        if (lane == 0) { sh.red[0][warp] = mi; sh.red[1][warp] = mp; }
        __syncthreads();
        if (warp == 0) {
            mi = lane < NW ? sh.red[0][lane] : 0; mp = lane < NW ? sh.red[1][lane] : 0;
            mi = warp_max(mi); mp = warp_max(mp);
            if (lane == 0) { sh.bcast[0] = mi; sh.bcast[1] = mp; }
        }
        __syncthreads();
        const int gray_max = sh.bcast[0], prm_max = sh.bcast[1];
        if (prm_max == 0) {                               // no positive PRM voxel: instance skipped (:74-76)
            for (long long j = tid; j < n; j += OTSU_THREADS) mout[j] = 0;
            if (tid == 0) { status_out[inst] = 3; b_max_out[inst] = 0; }
            return;
        }
        if (tid < 256) {
            // np.clip(v / gray_max * 300, 0, 300) + 30 -> astype(uint16) (truncation)
            double f = gray_max > 0 ? __dmul_rn(__ddiv_rn((double)tid, (double)gray_max), 300.0) : 0.0;
            f = fmin(fmax(f, 0.0), 300.0);
            sh.norm_i[tid] = gray_max > 0 ? (unsigned short)(int)__dadd_rn(f, 30.0) : (unsigned short)0;
            // np.round(p / max * 300 + 30) -> uint16   (round half to even)
            const double p = __dadd_rn(__dmul_rn(__ddiv_rn((double)tid, (double)prm_max), 300.0), 30.0);
            sh.norm_p[tid] = (unsigned short)(int)rint(p);
        }
        __syncthreads();
    }
    auto val_image = [&](long long j) -> int { const int r = raw_image(j); return MODE == 1 ? (int)sh.norm_i[r] : r; };
    auto val_prm = [&](long long j) -> int { const int r = raw_prm(j); return MODE == 1 ? (int)sh.norm_p[r] : r; };

    // ---- pass 1: min / max of both attributes (otsu.py:201) ---------------------------------------
    int g_min, g_max, p_min, p_max;
    {
        int a = 0x7fffffff, b = -1, c = 0x7fffffff, d = -1;
        for (long long j = tid; j < n; j += OTSU_THREADS) {
            const int vi = val_image(j), vp = val_prm(j);
            a = min(a, vi); b = max(b, vi); c = min(c, vp); d = max(d, vp);
        }
        a = warp_min(a); b = warp_max(b); c = warp_min(c); d = warp_max(d);
        if (lane == 0) { sh.red[0][warp] = a; sh.red[1][warp] = b; sh.red[2][warp] = c; sh.red[3][warp] = d; }
        __syncthreads();
        if (warp == 0) {
            a = lane < NW ? sh.red[0][lane] : 0x7fffffff; b = lane < NW ? sh.red[1][lane] : -1;
            c = lane < NW ? sh.red[2][lane] : 0x7fffffff; d = lane < NW ? sh.red[3][lane] : -1;
            a = warp_min(a); b = warp_max(b); c = warp_min(c); d = warp_max(d);
            if (lane == 0) { sh.bcast[2] = a; sh.bcast[3] = b; sh.bcast[4] = c; sh.bcast[5] = d; }
        }
        __syncthreads();
        g_min = sh.bcast[2]; g_max = sh.bcast[3]; p_min = sh.bcast[4]; p_max = sh.bcast[5];
    }
    const int G = g_max - g_min + 1;
    const int PR = p_max - p_min + 1;
    if (tid == 0 && g_info) {
        g_info[inst * 4 + 0] = g_min; g_info[inst * 4 + 1] = g_max;
        g_info[inst * 4 + 2] = p_min; g_info[inst * 4 + 3] = p_max;
    }
    if (G > OTSU_GMAX || PR > OTSU_GMAX || (unsigned long long)n * (unsigned long long)(G > 1 ? G - 1 : 1) >= 0xFFFFFFFFull) {
        for (long long j = tid; j < n; j += OTSU_THREADS) mout[j] = 255;
        if (tid == 0) { status_out[inst] = 4; b_max_out[inst] = 0; }
        return;
    }

    // ---- binning LUTs + clear diagonal histograms ---------------------------------------------------
    for (int v = tid; v < G; v += OTSU_THREADS) sh.lut_i[v] = (unsigned short)np_axis_bin(g_min + v, g_min, g_max, G);
    for (int v = tid; v < PR; v += OTSU_THREADS) sh.lut_p[v] = (unsigned short)np_axis_bin(p_min + v, p_min, p_max, G);
    const int ndiag = 2 * G - 1;
    for (int s = tid; s < ndiag; s += OTSU_THREADS) { sh.cnt[s] = 0u; sh.sumc[s] = 0u; }
    __syncthreads();

    // ---- pass 2: anti-diagonal histograms, warp-aggregated shared-memory atomics -------------------
    unsigned long long tot_c = 0ull, tot_r = 0ull;
    uint32_t* hist_i = hist ? hist + hist_off[inst] : nullptr;
    const long long n_round = (n + OTSU_THREADS - 1) / OTSU_THREADS * OTSU_THREADS;
    for (long long j = tid; j < n_round; j += OTSU_THREADS) {
        const bool in = j < n;
        int c = 0, r = 0;
        if (in) {
            c = sh.lut_i[val_image(j) - g_min];
            r = sh.lut_p[val_prm(j) - p_min];
            tot_c += (unsigned)c; tot_r += (unsigned)r;
            if (hist_i) atomicAdd(&hist_i[(size_t)r * G + c], 1u);
        }
        const bool ok = in && c <= G - 2 && r <= G - 2;
        const unsigned key = ok ? (unsigned)(r + c) : (0x80000000u | (unsigned)lane);
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const unsigned csum = __reduce_add_sync(peers, (unsigned)c);
        if (ok && lane == (__ffs(peers) - 1)) {
            atomicAdd(&sh.cnt[r + c], (unsigned)__popc(peers));
            atomicAdd(&sh.sumc[r + c], csum);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        tot_c += __shfl_xor_sync(0xffffffffu, tot_c, o);
        tot_r += __shfl_xor_sync(0xffffffffu, tot_r, o);
    }
    if (lane == 0) { sh.red64[0][warp] = tot_c; sh.red64[1][warp] = tot_r; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long a = 0, b = 0;
        for (int w = 0; w < NW; ++w) { a += sh.red64[0][w]; b += sh.red64[1][w]; }
        sh.tot[0] = a; sh.tot[1] = b;
    }
    __syncthreads();

    // ---- scan over b (otsu.py:226-274, closed form) ------------------------------------------------
    // candidates b = b_dw + i, i in [0, nb); background = diagonals s <= i.
    const int b_dw = 2 * g_min + 1, b_up = 2 * g_max - 1;
    const int nb = max(1, b_up - b_dw);
    const double Nd = (double)n;
    // bin centres are affine in the bin index up to fp64 rounding: c1[c] = start + (c + 0.5) * step
    double lo1 = (double)g_min, hi1 = (double)g_max, lo2 = (double)p_min, hi2 = (double)p_max;
    if (g_min == g_max) { lo1 -= 0.5; hi1 += 0.5; }
    if (p_min == p_max) { lo2 -= 0.5; hi2 += 0.5; }
    const double step1 = (hi1 - lo1) / (double)G, step2 = (hi2 - lo2) / (double)G;
    const double ut0 = (lo1 * Nd + step1 * ((double)sh.tot[0] + 0.5 * Nd)) / Nd;
    const double ut1 = (lo2 * Nd + step2 * ((double)sh.tot[1] + 0.5 * Nd)) / Nd;

    const int chunk = (nb + OTSU_THREADS - 1) / OTSU_THREADS;
    const int i0 = min(nb, tid * chunk), i1 = min(nb, i0 + chunk);
    unsigned long long lp = 0, lc = 0, lr = 0;       // local sums of cnt, sum c, sum r over my chunk
    for (int i = i0; i < i1; ++i) {
        const unsigned long long cn = i < ndiag ? sh.cnt[i] : 0u, sc = i < ndiag ? sh.sumc[i] : 0u;
        lp += cn; lc += sc; lr += (unsigned long long)i * cn - sc;
    }
    // exclusive scan of the per-thread triples across the CTA
    unsigned long long ep = lp, ec = lc, er = lr;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long a = __shfl_up_sync(0xffffffffu, ep, o), b = __shfl_up_sync(0xffffffffu, ec, o),
                                 c = __shfl_up_sync(0xffffffffu, er, o);
        if (lane >= o) { ep += a; ec += b; er += c; }
    }
    if (lane == 31) { sh.red64[0][warp] = ep; sh.red64[1][warp] = ec; sh.red64[2][warp] = er; }
    __syncthreads();
    unsigned long long bp = 0, bc = 0, br = 0;
    for (int w = 0; w < warp; ++w) { bp += sh.red64[0][w]; bc += sh.red64[1][w]; br += sh.red64[2][w]; }
    unsigned long long P = bp + ep - lp, MC = bc + ec - lc, MR = br + er - lr;   // exclusive prefix
    double my_var = 0.0;
    int my_b = 0x7fffffff;
    for (int i = i0; i < i1; ++i) {
        const unsigned long long cn = i < ndiag ? sh.cnt[i] : 0u, sc = i < ndiag ? sh.sumc[i] : 0u;
        P += cn; MC += sc; MR += (unsigned long long)i * cn - sc;                // inclusive at i
        const double Pd = (double)P;
        const double p0 = Pd / Nd;
        const double u00 = (lo1 * Pd + step1 * ((double)MC + 0.5 * Pd)) / Nd;
        const double u01 = (lo2 * Pd + step2 * ((double)MR + 0.5 * Pd)) / Nd;
        const double p1 = 1.0 - p0;
        const double u10 = (ut0 - p0 * u00) / p1, u11 = (ut1 - p0 * u01) / p1;
        const double d0 = u00 - ut0, d1 = u01 - ut1, f0 = u10 - ut0, f1 = u11 - ut1;
        const double var_b = ((p0 * d0) * d0 + (p1 * f0) * f0) + ((p0 * d1) * d1 + (p1 * f1) * f1);
        if (var_b > my_var) { my_var = var_b; my_b = b_dw + i; }                 // first strictly greater
    }
    // arg-max: largest var, earliest b among equals
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, my_var, o);
        const int ob = __shfl_xor_sync(0xffffffffu, my_b, o);
        if (ov > my_var || (ov == my_var && ob < my_b)) { my_var = ov; my_b = ob; }
    }
    if (lane == 0) { sh.best_var[warp] = my_var; sh.best_b[warp] = my_b; }
    __syncthreads();
    if (tid == 0) {
        double bv = 0.0; int bb = 0x7fffffff;
        for (int w = 0; w < NW; ++w)
            if (sh.best_var[w] > bv || (sh.best_var[w] == bv && sh.best_b[w] < bb)) { bv = sh.best_var[w]; bb = sh.best_b[w]; }
        const int found = (bv > 0.0 && bb != 0x7fffffff);
        sh.bcast[6] = found ? bb : 0;
        sh.bcast[7] = found;
        b_max_out[inst] = found ? bb : 0;
        status_out[inst] = found ? 0 : 1;
    }
    __syncthreads();
    const int b_max = sh.bcast[6];
    const int found = sh.bcast[7];

    // ---- pass 3: mask (otsu.py:276-282, closed form) -----------------------------------------------
    // background <=> I < min(b_max - g_min, g_max)  and  P < min(b_max - I, g_max + 1)
    const int x_hi = min(b_max - g_min, g_max);
    for (long long j = tid; j < n; j += OTSU_THREADS) {
        uint8_t m = 255;
        if (found) {
            const int I = val_image(j), Pv = val_prm(j);
            if (I < x_hi && Pv < min(b_max - I, g_max + 1)) m = 0;
        }
        mout[j] = m;
    }
}

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
    otsu2d_kernel<0><<<n_crops, OTSU_THREADS, 0, stream>>>(image, prm, crop_off, n_crops, 0, 0, 0, nullptr,
                                                          nullptr, nullptr, mask, b_max, g_info, status, hist, hist_off);
    B200_LAUNCH_CHECK("otsu2d_kernel<0>");
    return 0;
}

extern "C" int b200seg_soma_binarize_dev(const uint8_t* volume, int S, int H, int W, const int32_t* boxes,
                                         const uint8_t* prm, const int64_t* crop_off, int n,
                                         const int32_t* order, const int32_t* n_valid, uint8_t* mask,
                                         int32_t* b_max, int32_t* status, b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(n >= 0 && S > 0 && H > 0 && W > 0, "soma_binarize: bad sizes");
    if (n == 0) return 0;
    B200_CHECK_ARG(volume && boxes && prm && crop_off && mask && b_max && status, "soma_binarize: null pointer");
    otsu2d_kernel<1><<<n, OTSU_THREADS, 0, stream>>>(volume, prm, crop_off, n, S, H, W, boxes, order, n_valid,
                                                    mask, b_max, nullptr, status, nullptr, nullptr);
    B200_LAUNCH_CHECK("otsu2d_kernel<1>");
    return 0;
}
