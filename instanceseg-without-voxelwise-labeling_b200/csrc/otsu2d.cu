// otsu2d.cu -- batched per-instance 2D-Otsu binarization
// (replaces tools/otsu.py:199-284 `otsu_py_2d_fast`, k = -1, and the crop/normalise steps of
//  tools/binarization_soma.py:78-94).
//
// One CTA per instance crop; all crops of a volume go in one launch.
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
// Data movement: pass A streams the crop ONCE from global memory (rows of the raw volume for the
// fused soma mode, with incremental (x,y,z) addressing and 4 independent loads in flight per
// thread), records the raw min/max and parks the samples in a shared-memory crop cache; the
// histogram and mask passes then run out of shared memory (samples beyond the cache capacity fall
// back to L2).  The caller's normalisation (binarization_soma.py:85-91) is monotone, so it folds
// into 256-entry lookup tables and the normalised min/max follow from the raw ones -- no extra pass.
// HBM sees the crop once and the mask once.
//
// numpy.histogram2d binning is reproduced exactly (fp64 linspace edges i*step+start with the last
// edge pinned, searchsorted-right, right edge inclusive; each axis over its own [min,max] split in
// G bins) through per-gray-level lookup tables.
#include "common.cuh"
#include <type_traits>

namespace b200seg {

constexpr int OTSU_THREADS = 512;
constexpr int OTSU_NW = OTSU_THREADS / 32;

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

template <int GMAX, int NLUT>
struct OtsuShared {
    unsigned int cnt[2 * GMAX - 1];   // count per anti-diagonal s = r + c (r,c <= G-2)
    unsigned int sumc[2 * GMAX - 1];  // sum of image bin c per anti-diagonal
    unsigned short bin_i[NLUT];       // MODE 0: gray level (v - g_min) -> bin; MODE 1: raw uint8 -> bin
    unsigned short bin_p[NLUT];
    unsigned short norm_i[256];       // MODE 1: soma normalisation (raw uint8 -> uint16 level)
    unsigned short norm_p[256];
    int red[4][OTSU_NW];
    unsigned long long red64[3][OTSU_NW];
    double best_var[OTSU_NW];
    int best_b[OTSU_NW];
    int bcast[8];
    unsigned long long tot[2];
};

// MODE 0: image/prm are uint16 sample arrays (crop i at crop_off[i]).
// MODE 1: soma fused: image is the raw uint8 volume, prm raw uint8 box crops.
template <int MODE, int GMAX>
__global__ void __launch_bounds__(OTSU_THREADS, MODE == 1 ? 3 : 2)
otsu2d_kernel(const void* __restrict__ image_, const void* __restrict__ prm_,
              const int64_t* __restrict__ crop_off, int n_crops,
              int S, int H, int W, const int32_t* __restrict__ det_off, const int32_t* __restrict__ boxes,
              const int32_t* __restrict__ order, const int32_t* __restrict__ n_valid,
              uint8_t* __restrict__ mask, int32_t* __restrict__ b_max_out, int32_t* __restrict__ g_info,
              int32_t* __restrict__ status_out, uint32_t* __restrict__ hist, const int64_t* __restrict__ hist_off,
              int cache_vox) {
    using E = typename std::conditional<MODE == 1, uint8_t, uint16_t>::type;
    constexpr int NLUT = MODE == 1 ? 256 : GMAX;
    __shared__ OtsuShared<GMAX, NLUT> sh;
    extern __shared__ __align__(16) unsigned char s_cache[];
    E* c_img = reinterpret_cast<E*>(s_cache);
    E* c_prm = c_img + cache_vox;

    // grid = (slots, volumes): instance `slot` (visit order) of volume blockIdx.y
    const int slot = blockIdx.x, vol = blockIdx.y;
    const int base = det_off ? det_off[vol] : 0;
    const int n_here = det_off ? det_off[vol + 1] - base : n_crops;
    if (slot >= n_here || (n_valid && slot >= n_valid[vol])) return;
    const int inst = base + (order ? order[base + slot] : slot);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int64_t off = crop_off[inst];
    const int n = (int)(crop_off[inst + 1] - off);               // samples of this crop (< 2^31)
    int bx1 = 0, by1 = 0, bz1 = 0, sx = 1, sy = 1;
    if (MODE == 1) {
        const int32_t* bb = boxes + 6 * inst;
        bx1 = bb[0]; by1 = bb[1]; bz1 = bb[2];
        sx = bb[3] - bb[0] + 1; sy = bb[4] - bb[1] + 1;
    }
    const E* g_img = reinterpret_cast<const E*>(image_) + (MODE == 0 ? (size_t)off : (size_t)vol * S * H * W);
    const E* g_prm = reinterpret_cast<const E*>(prm_) + off;
    uint8_t* mout = mask + off;
    const uint8_t fail_fill = MODE == 1 ? 0 : 255;      // chain mode: a failed instance pastes nothing

    if (n <= 0) {
        if (tid == 0) { status_out[inst] = 2; b_max_out[inst] = 0; }
        return;
    }
    const int ncache = min(n, cache_vox);

    // raw sample fetch straight from global memory (pass A, and the uncached tail of later passes)
    auto g_image_at = [&](int j) -> int {
        if (MODE == 0) return (int)g_img[j];
        const int x = j % sx;
        const int t = j / sx;
        const int y = t % sy, z = t / sy;
        return (int)g_img[((size_t)(bz1 + z) * H + (by1 + y)) * W + (bx1 + x)];
    };

    // ---- pass A: stream the crop once: raw min/max + fill the shared-memory cache ------------------
    int rmin_i = 0x7fffffff, rmax_i = -1, rmin_p = 0x7fffffff, rmax_p = -1;
    {
        // incremental (x,y,z) of sample j = tid + k*THREADS (no per-sample division)
        int x = 0, y = 0, z = 0;
        int dx = 0, dy = 0, dz = 0;
        if (MODE == 1) {
            x = tid % sx; const int t = tid / sx; y = t % sy; z = t / sy;
            dx = OTSU_THREADS % sx; const int u = OTSU_THREADS / sx; dy = u % sy; dz = u / sy;
        }
        constexpr int U = 4;
        for (int j0 = tid; j0 < n; j0 += U * OTSU_THREADS) {
            int vi[U], vp[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * OTSU_THREADS;
                vi[u] = -1; vp[u] = -1;
                if (j < n) {
                    if (MODE == 1) {
                        vi[u] = (int)g_img[((size_t)(bz1 + z) * H + (by1 + y)) * W + (bx1 + x)];
                        x += dx; if (x >= sx) { x -= sx; ++y; }
                        y += dy; if (y >= sy) { y -= sy; ++z; }
                        z += dz;
                    } else {
                        vi[u] = (int)g_img[j];
                    }
                    vp[u] = (int)g_prm[j];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * OTSU_THREADS;
                if (j < n) {
                    rmin_i = min(rmin_i, vi[u]); rmax_i = max(rmax_i, vi[u]);
                    rmin_p = min(rmin_p, vp[u]); rmax_p = max(rmax_p, vp[u]);
                    if (j < ncache) { c_img[j] = (E)vi[u]; c_prm[j] = (E)vp[u]; }
                }
            }
        }
        rmin_i = warp_min(rmin_i); rmax_i = warp_max(rmax_i); rmin_p = warp_min(rmin_p); rmax_p = warp_max(rmax_p);
        if (lane == 0) { sh.red[0][warp] = rmin_i; sh.red[1][warp] = rmax_i; sh.red[2][warp] = rmin_p; sh.red[3][warp] = rmax_p; }
        __syncthreads();
        if (warp == 0) {
            int a = lane < OTSU_NW ? sh.red[0][lane] : 0x7fffffff, b = lane < OTSU_NW ? sh.red[1][lane] : -1;
            int c = lane < OTSU_NW ? sh.red[2][lane] : 0x7fffffff, d = lane < OTSU_NW ? sh.red[3][lane] : -1;
            a = warp_min(a); b = warp_max(b); c = warp_min(c); d = warp_max(d);
            if (lane == 0) { sh.bcast[0] = a; sh.bcast[1] = b; sh.bcast[2] = c; sh.bcast[3] = d; }
        }
        __syncthreads();
        rmin_i = sh.bcast[0]; rmax_i = sh.bcast[1]; rmin_p = sh.bcast[2]; rmax_p = sh.bcast[3];
    }

    // ---- normalisation LUTs (soma, binarization_soma.py:85-91) and value range (otsu.py:201) -------
    int g_min, g_max, p_min, p_max;
    if (MODE == 1) {
        const int gray_max = rmax_i, prm_max = rmax_p;
        if (prm_max == 0) {                               // no positive PRM voxel: instance skipped (:74-76)
            for (int j = tid; j < n; j += OTSU_THREADS) mout[j] = 0;
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
        // both maps are monotone non-decreasing, so the normalised range is the image of the raw range
        g_min = sh.norm_i[rmin_i]; g_max = sh.norm_i[rmax_i];
        p_min = sh.norm_p[rmin_p]; p_max = sh.norm_p[rmax_p];
    } else {
        g_min = rmin_i; g_max = rmax_i; p_min = rmin_p; p_max = rmax_p;
    }
    const int G = g_max - g_min + 1;
    const int PR = p_max - p_min + 1;
    if (tid == 0 && g_info) {
        g_info[inst * 4 + 0] = g_min; g_info[inst * 4 + 1] = g_max;
        g_info[inst * 4 + 2] = p_min; g_info[inst * 4 + 3] = p_max;
    }
    if (G > GMAX || PR > GMAX || (unsigned long long)n * (unsigned long long)(G > 1 ? G - 1 : 1) >= 0xFFFFFFFFull) {
        for (int j = tid; j < n; j += OTSU_THREADS) mout[j] = fail_fill;
        if (tid == 0) { status_out[inst] = 4; b_max_out[inst] = 0; }
        return;
    }

    // ---- binning LUTs + clear diagonal histograms ---------------------------------------------------
    if (MODE == 1) {
        if (tid < 256) {                                  // raw level -> bin (only levels inside the raw range occur)
            const int t = tid;
            sh.bin_i[t] = (t >= rmin_i && t <= rmax_i) ? (unsigned short)np_axis_bin(sh.norm_i[t], g_min, g_max, G) : (unsigned short)0;
            sh.bin_p[t] = (t >= rmin_p && t <= rmax_p) ? (unsigned short)np_axis_bin(sh.norm_p[t], p_min, p_max, G) : (unsigned short)0;
        }
    } else {
        for (int v = tid; v < G; v += OTSU_THREADS) sh.bin_i[v] = (unsigned short)np_axis_bin(g_min + v, g_min, g_max, G);
        for (int v = tid; v < PR; v += OTSU_THREADS) sh.bin_p[v] = (unsigned short)np_axis_bin(p_min + v, p_min, p_max, G);
    }
    const int ndiag = 2 * G - 1;
    for (int s = tid; s < ndiag; s += OTSU_THREADS) { sh.cnt[s] = 0u; sh.sumc[s] = 0u; }
    __syncthreads();

    auto raw_i_at = [&](int j) -> int { return j < ncache ? (int)c_img[j] : g_image_at(j); };
    auto raw_p_at = [&](int j) -> int { return j < ncache ? (int)c_prm[j] : (int)g_prm[j]; };
    auto bin_of_i = [&](int raw) -> int { return MODE == 1 ? (int)sh.bin_i[raw] : (int)sh.bin_i[raw - g_min]; };
    auto bin_of_p = [&](int raw) -> int { return MODE == 1 ? (int)sh.bin_p[raw] : (int)sh.bin_p[raw - p_min]; };

    // ---- pass B: anti-diagonal histograms in shared memory ------------------------------------------
    // Warp aggregation: when all 32 lanes hit the same diagonal (background runs -- the contended case)
    // one lane adds the warp's count and redux-summed column index; otherwise lanes issue their own
    // shared-memory atomics (spread addresses).  (__match_any_sync-based grouping was measured 6x
    // slower here: its cost grows with the number of distinct keys in the warp.)
    unsigned int tot_c = 0u, tot_r = 0u;      // fit 32 bits: n * (G-1) < 2^32 is checked above
    uint32_t* hist_i = hist ? hist + hist_off[inst] : nullptr;
    auto add_sample = [&](bool in, int ri, int rp) {
        int c = 0, r = 0;
        if (in) {
            c = bin_of_i(ri);
            r = bin_of_p(rp);
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
        // cached part: each thread takes 4 consecutive samples (one 32-bit / 64-bit shared load per attribute)
        const int ngroups = (ncache + 3) >> 2;
        const int g_round = (ngroups + OTSU_THREADS - 1) / OTSU_THREADS * OTSU_THREADS;
        for (int g4 = tid; g4 < g_round; g4 += OTSU_THREADS) {
            int ri[4] = {0, 0, 0, 0}, rp[4] = {0, 0, 0, 0};
            const int j = g4 << 2;
            if (g4 < ngroups) {
                if (MODE == 1) {
                    const uint32_t wi = *reinterpret_cast<const uint32_t*>(c_img + j);
                    const uint32_t wp = *reinterpret_cast<const uint32_t*>(c_prm + j);
#pragma unroll
                    for (int u = 0; u < 4; ++u) { ri[u] = (wi >> (8 * u)) & 0xFF; rp[u] = (wp >> (8 * u)) & 0xFF; }
                } else {
                    const uint2 wi = *reinterpret_cast<const uint2*>(c_img + j);
                    const uint2 wp = *reinterpret_cast<const uint2*>(c_prm + j);
                    ri[0] = wi.x & 0xFFFF; ri[1] = wi.x >> 16; ri[2] = wi.y & 0xFFFF; ri[3] = wi.y >> 16;
                    rp[0] = wp.x & 0xFFFF; rp[1] = wp.x >> 16; rp[2] = wp.y & 0xFFFF; rp[3] = wp.y >> 16;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) add_sample(g4 < ngroups && j + u < ncache, ri[u], rp[u]);
        }
        // uncached tail (crop larger than the cache): straight from L2
        const int ntail = n - ncache;
        const int t_round = (ntail + OTSU_THREADS - 1) / OTSU_THREADS * OTSU_THREADS;
        for (int t = tid; t < t_round; t += OTSU_THREADS) {
            const bool in = t < ntail;
            const int j = ncache + t;
            add_sample(in, in ? g_image_at(j) : 0, in ? (int)g_prm[j] : 0);
        }
    }
    tot_c = __reduce_add_sync(0xffffffffu, tot_c);
    tot_r = __reduce_add_sync(0xffffffffu, tot_r);
    if (lane == 0) { sh.red64[0][warp] = tot_c; sh.red64[1][warp] = tot_r; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long a = 0, b = 0;
        for (int w = 0; w < OTSU_NW; ++w) { a += sh.red64[0][w]; b += sh.red64[1][w]; }
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
        for (int w = 0; w < OTSU_NW; ++w)
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

    // ---- pass C: mask (otsu.py:276-282, closed form) -----------------------------------------------
    // background <=> I < min(b_max - g_min, g_max)  and  P < min(b_max - I, g_max + 1)
    const int x_hi = min(b_max - g_min, g_max);
    auto mask_of = [&](int ri, int rp) -> uint32_t {
        if (!found) return (uint32_t)fail_fill;
        int I = ri, Pv = rp;
        if (MODE == 1) { I = sh.norm_i[ri]; Pv = sh.norm_p[rp]; }
        return (I < x_hi && Pv < min(b_max - I, g_max + 1)) ? 0u : 255u;
    };
    {
        const int ngroups = (ncache + 3) >> 2;
        const bool aligned = ((reinterpret_cast<uintptr_t>(mout)) & 3) == 0;
        for (int g4 = tid; g4 < ngroups; g4 += OTSU_THREADS) {
            const int j = g4 << 2;
            int ri[4], rp[4];
            if (MODE == 1) {
                const uint32_t wi = *reinterpret_cast<const uint32_t*>(c_img + j);
                const uint32_t wp = *reinterpret_cast<const uint32_t*>(c_prm + j);
#pragma unroll
                for (int u = 0; u < 4; ++u) { ri[u] = (wi >> (8 * u)) & 0xFF; rp[u] = (wp >> (8 * u)) & 0xFF; }
            } else {
                const uint2 wi = *reinterpret_cast<const uint2*>(c_img + j);
                const uint2 wp = *reinterpret_cast<const uint2*>(c_prm + j);
                ri[0] = wi.x & 0xFFFF; ri[1] = wi.x >> 16; ri[2] = wi.y & 0xFFFF; ri[3] = wi.y >> 16;
                rp[0] = wp.x & 0xFFFF; rp[1] = wp.x >> 16; rp[2] = wp.y & 0xFFFF; rp[3] = wp.y >> 16;
            }
            uint32_t m4 = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) m4 |= mask_of(ri[u], rp[u]) << (8 * u);
            if (aligned && j + 3 < ncache) {
                *reinterpret_cast<uint32_t*>(mout + j) = m4;
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) if (j + u < ncache) mout[j + u] = (uint8_t)(m4 >> (8 * u));
            }
        }
        for (int j = ncache + tid; j < n; j += OTSU_THREADS) mout[j] = (uint8_t)mask_of(g_image_at(j), (int)g_prm[j]);
    }
}

constexpr int OTSU_GMAX_GENERIC = 2048;     // gray range supported by the generic entry (reference: G*G fp64 histogram)
constexpr int OTSU_GMAX_SOMA = 512;         // soma levels live in [30,330]
constexpr int OTSU_CACHE_BYTES_GENERIC = 64 * 1024;
constexpr int OTSU_CACHE_BYTES_SOMA = 48 * 1024;

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
    auto kern = otsu2d_kernel<0, OTSU_GMAX_GENERIC>;
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, OTSU_CACHE_BYTES_GENERIC));
    kern<<<n_crops, OTSU_THREADS, OTSU_CACHE_BYTES_GENERIC, stream>>>(image, prm, crop_off, n_crops, 0, 0, 0, nullptr, nullptr,
                                                                      nullptr, nullptr, mask, b_max, g_info, status, hist,
                                                                      hist_off, OTSU_CACHE_BYTES_GENERIC / 4);
    B200_LAUNCH_CHECK("otsu2d_kernel<0>");
    return 0;
}

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
    B200_CHECK_ARG(volumes && boxes && prm && crop_off && mask && b_max && status, "soma_binarize: null pointer");
    auto kern = otsu2d_kernel<1, OTSU_GMAX_SOMA>;
    static bool attr_set = false;
    if (!attr_set) {
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, OTSU_CACHE_BYTES_SOMA));
        attr_set = true;
    }
    dim3 grid(n_max, n_volumes);
    kern<<<grid, OTSU_THREADS, OTSU_CACHE_BYTES_SOMA, stream>>>(volumes, prm, crop_off, n_max, S, H, W, det_off, boxes, order,
                                                                n_valid, mask, b_max, nullptr, status, nullptr, nullptr,
                                                                OTSU_CACHE_BYTES_SOMA / 2);
    B200_LAUNCH_CHECK("otsu2d_kernel<1>");
    return 0;
}
