// otsu_common.cuh -- pieces shared by the two 2D-Otsu kernels (otsu2d.cu: generic uint16 crops,
// soma_binarize.cu: fused crop + normalise + Otsu from the raw uint8 volume).
#pragma once
#include "common.cuh"

namespace b200seg {

constexpr int OTSU_THREADS = 512;
constexpr int OTSU_NW = OTSU_THREADS / 32;

// bin of value v on an axis of G bins over [vmin,vmax]  (numpy histogramdd / linspace semantics:
// fp64 edges edge(i) = i*step + lo for i < G, edge(G) = hi; searchsorted-right; right edge inclusive).
// The first guess comes from cheap fp32 arithmetic; the two fix-up loops compare against the exact
// fp64 edges, so the result does not depend on the quality of the guess.
__device__ __forceinline__ int np_axis_bin(int v, int vmin, int vmax, int G) {
    double lo = (double)vmin, hi = (double)vmax;
    if (vmin == vmax) { lo -= 0.5; hi += 0.5; }
    const double step = __ddiv_rn(__dsub_rn(hi, lo), (double)G);
    const double x = (double)v;
    int g = vmin == vmax ? 0 : (int)(((float)(v - vmin) * (float)G) / (float)(vmax - vmin));
    g = max(0, min(G - 1, g));
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

template <int THREADS>
struct OtsuScanShared {
    unsigned long long red64[5][THREADS / 32];
    double best_var[THREADS / 32];
    int best_b[THREADS / 32];
    int bcast[2];
};

// Scan over b of tools/otsu.py:226-274 in closed form.  Called by all THREADS threads of the CTA.
//   cnt[s], sumc[s], s < ndiag: number of samples / sum of image bins on anti-diagonal s = r + c
//   (only samples with r <= G-2 and c <= G-2);
//   part_c, part_r: this thread's share of the image / PRM bin sums over the samples NOT on those diagonals
//   (extra_is_total = false) or over ALL n samples (extra_is_total = true); reduced here.
// Candidates b = b_dw + i, i in [0, nb); the background of candidate i is the diagonals s <= i.  Exact integer
// prefix sums give P = #background, MC = sum c, MR = sum r for every candidate; the criterion is then evaluated
// in fp64 for all b in parallel with the reference's own operation structure (otsu.py:240-246,257-270:
// p0, un-normalised u0, u1 = (ut - p0*u0)/p1, var = p0*|u0-ut|^2 + p1*|u1-ut|^2).  The algebraically equal
// short form |a-ut|^2 * p0/(1-p0) was tried and rejected: it resolves near-ties differently from the reference
// (tests/test_gpu_parity.py::test_otsu_batch_vs_oracle_hist_threshold_mask).  The arg-max keeps the reference's
// first-strictly-greater rule (otsu.py:247-250,271-274); NaN (p1 == 0) never wins.
// Returns (uniformly) found and b_max.
template <int THREADS>
__device__ __forceinline__ void otsu_scan_b(const unsigned int* cnt, const unsigned int* sumc, int ndiag,
                                            unsigned long long part_c, unsigned long long part_r, bool extra_is_total,
                                            int n, int g_min, int g_max, int p_min, int p_max,
                                            OtsuScanShared<THREADS>& ss, int& b_max, int& found) {
    constexpr int NW = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = g_max - g_min + 1;
    const int b_dw = 2 * g_min + 1, b_up = 2 * g_max - 1;
    const int nb = max(1, b_up - b_dw);
    const int ntot = max(nb, ndiag);                   // the sums must cover every diagonal, candidates only i < nb
    const double Nd = (double)n;
    // bin centres are affine in the bin index up to fp64 rounding: c1[c] = start + (c + 0.5) * step
    double lo1 = (double)g_min, hi1 = (double)g_max, lo2 = (double)p_min, hi2 = (double)p_max;
    if (g_min == g_max) { lo1 -= 0.5; hi1 += 0.5; }
    if (p_min == p_max) { lo2 -= 0.5; hi2 += 0.5; }
    const double step1 = (hi1 - lo1) / (double)G, step2 = (hi2 - lo2) / (double)G;

    const int chunk = (ntot + THREADS - 1) / THREADS;
    const int i0 = min(ntot, tid * chunk), i1 = min(ntot, i0 + chunk);
    __syncthreads();                                  // the histograms are complete; ss is free
    unsigned long long lp = 0, lc = 0, lr = 0;       // local sums of cnt, sum c, sum r over my chunk
    for (int i = i0; i < i1; ++i) {
        const unsigned long long cn = i < ndiag ? cnt[i] : 0u, sc = i < ndiag ? sumc[i] : 0u;
        lp += cn; lc += sc; lr += (unsigned long long)i * cn - sc;
    }
    // inclusive warp scan of the triples + warp sums of the extra parts
    unsigned long long ep = lp, ec = lc, er = lr;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long a = __shfl_up_sync(0xffffffffu, ep, o), b = __shfl_up_sync(0xffffffffu, ec, o),
                                 c = __shfl_up_sync(0xffffffffu, er, o);
        if (lane >= o) { ep += a; ec += b; er += c; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { part_c += __shfl_xor_sync(0xffffffffu, part_c, o); part_r += __shfl_xor_sync(0xffffffffu, part_r, o); }
    if (lane == 31) {
        ss.red64[0][warp] = ep; ss.red64[1][warp] = ec; ss.red64[2][warp] = er;
        ss.red64[3][warp] = part_c; ss.red64[4][warp] = part_r;
    }
    __syncthreads();
    unsigned long long bp = 0, bc = 0, br = 0, tot_c = 0, tot_r = 0, all_c = 0, all_r = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const unsigned long long a = ss.red64[0][w], b = ss.red64[1][w], c = ss.red64[2][w];
        if (w < warp) { bp += a; bc += b; br += c; }
        all_c += b; all_r += c;
        tot_c += ss.red64[3][w]; tot_r += ss.red64[4][w];
    }
    if (!extra_is_total) { tot_c += all_c; tot_r += all_r; }
    const double ut0 = (lo1 * Nd + step1 * ((double)tot_c + 0.5 * Nd)) / Nd;
    const double ut1 = (lo2 * Nd + step2 * ((double)tot_r + 0.5 * Nd)) / Nd;
    unsigned long long P = bp + ep - lp, MC = bc + ec - lc, MR = br + er - lr;   // exclusive prefix
    double my_var = 0.0;
    int my_b = 0x7fffffff;
    const int e1 = min(i1, nb);
    for (int i = i0; i < e1; ++i) {
        const unsigned long long cn = i < ndiag ? cnt[i] : 0u, sc = i < ndiag ? sumc[i] : 0u;
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
    if (lane == 0) { ss.best_var[warp] = my_var; ss.best_b[warp] = my_b; }
    __syncthreads();
    if (tid == 0) {
        double bv = 0.0; int bb = 0x7fffffff;
        for (int w = 0; w < NW; ++w)
            if (ss.best_var[w] > bv || (ss.best_var[w] == bv && ss.best_b[w] < bb)) { bv = ss.best_var[w]; bb = ss.best_b[w]; }
        const int f = (bv > 0.0 && bb != 0x7fffffff);
        ss.bcast[0] = f ? bb : 0;
        ss.bcast[1] = f;
    }
    __syncthreads();
    b_max = ss.bcast[0];
    found = ss.bcast[1];
}

}  // namespace b200seg
