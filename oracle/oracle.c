/*
 * oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker / reported CPU baseline -- never as the product path.
 *
 * Parity pinning: every function here is checked (tests/test_oracle_vs_reference.py,
 * tests/golden/) against the reference's own code run in the authoring container:
 *   - oracle_nms_3d / oracle_bbox_overlaps_3d against oracle/_ref/cython_{nms,bbox}_3d (built from
 *     the reference .pyx by oracle/build_ref.py),
 *   - oracle_otsu_2d_fast against tools/otsu.py:otsu_py_2d_fast (golden vectors made by
 *     tests/golden/make_golden.py which imports the reference),
 *   - oracle_peak_stimulation against lib/prm/peak_stimulation_3d.py (same script),
 *   - oracle_roialign3d_{fwd,bwd} against oracle/_ref/libref_roialign3d.so on the GPU box
 *     (the reference has no CPU RoIAlign; functions/roi_align_3d.py:31-32 raises).
 *
 * Build: gcc -O2 -ffp-contract=off (baseline x86-64: no FMA, fp32 expressions evaluated in fp32,
 * exactly like the Cython-generated C of the reference).
 *
 * Citations are path:line relative to the reference checkout.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * 3D NMS -- lib/utils/cython_nms_3d.pyx:39-96 (nms_3d), :102-159 (nms_3d_volume)
 * dets [n,7] = x1,y1,z1,x2,y2,z2,score (fp32, C-contiguous).
 * Tie rule: the reference visits `scores.argsort()[::-1]` (cython_nms_3d.pyx:49).  numpy's default argsort is an
 * introsort (insertion sort below 17 elements, where it IS stable; unstable and build dependent above), so equal keys come
 * out in ascending index order and the reversal puts the HIGHER index first.  That is the rule here: key descending, original
 * index DESCENDING -- identical to the reference wherever numpy's sort is stable, documented deviation elsewhere.  NaN keys
 * sort first (numpy sorts NaN last ascending, then the reference reverses).
 * keep_out receives the kept ORIGINAL indices in ascending order (np.where(suppressed==0), :96).
 * ------------------------------------------------------------------------------------------ */
typedef struct { float key; int idx; } sort_item;

static int cmp_desc(const void* pa, const void* pb) {
    const sort_item* a = (const sort_item*)pa;
    const sort_item* b = (const sort_item*)pb;
    int an = isnan(a->key), bn = isnan(b->key);
    if (an != bn) return an ? -1 : 1;
    if (!an) {
        if (a->key > b->key) return -1;
        if (a->key < b->key) return 1;
    }
    return (a->idx < b->idx) - (a->idx > b->idx);      /* equal keys: higher original index first */
}

static inline float f_max(float a, float b) { return a >= b ? a : b; }   /* cython_nms_3d.pyx:30-31 */
static inline float f_min(float a, float b) { return a <= b ? a : b; }   /* cython_nms_3d.pyx:33-34 */

long oracle_nms_3d(const float* dets, long n, float thresh, int by_volume, int64_t* keep_out) {
    if (n <= 0) return 0;
    float* vol = (float*)malloc(sizeof(float) * n);
    sort_item* order = (sort_item*)malloc(sizeof(sort_item) * n);
    unsigned char* sup = (unsigned char*)calloc(n, 1);
    for (long i = 0; i < n; ++i) {
        const float* d = dets + 7 * i;
        /* numpy fp32: (x2 - x1 + 1) * (y2 - y1 + 1) * (z2 - z1 + 1)   (:48) */
        float a = (d[3] - d[0]) + 1.0f;
        float b = (d[4] - d[1]) + 1.0f;
        float c = (d[5] - d[2]) + 1.0f;
        float ab = a * b;
        vol[i] = ab * c;
        order[i].key = by_volume ? vol[i] : d[6];
        order[i].idx = (int)i;
    }
    qsort(order, n, sizeof(sort_item), cmp_desc);
    for (long _i = 0; _i < n; ++_i) {
        int i = order[_i].idx;
        if (sup[i]) continue;
        const float* di = dets + 7 * i;
        float ix1 = di[0], iy1 = di[1], iz1 = di[2], ix2 = di[3], iy2 = di[4], iz2 = di[5];
        float ivol = vol[i];
        for (long _j = _i + 1; _j < n; ++_j) {
            int j = order[_j].idx;
            if (sup[j]) continue;
            const float* dj = dets + 7 * j;
            float xx1 = f_max(ix1, dj[0]);
            float yy1 = f_max(iy1, dj[1]);
            float zz1 = f_max(iz1, dj[2]);
            float xx2 = f_min(ix2, dj[3]);
            float yy2 = f_min(iy2, dj[4]);
            float zz2 = f_min(iz2, dj[5]);
            /* (xx2 - xx1 + 1): fp32 subtract, fp64 "+ 1.0", rounded to fp32 == fp32 add (:88-90) */
            float w = f_max(0.0f, (float)((double)(xx2 - xx1) + 1.0));
            float h = f_max(0.0f, (float)((double)(yy2 - yy1) + 1.0));
            float s = f_max(0.0f, (float)((double)(zz2 - zz1) + 1.0));
            float wh = w * h;
            float inter = wh * s;
            float uni = ivol + vol[j];
            uni = uni - inter;
            float ovr = inter / uni;
            if (ovr >= thresh) sup[j] = 1;
        }
    }
    long m = 0;
    for (long i = 0; i < n; ++i) if (!sup[i]) keep_out[m++] = i;
    free(vol); free(order); free(sup);
    return m;
}

/* descending order used by the binarization scripts (binarization_soma.py:60: argsort()[::-1]),
 * same tie rule as above.  order_out[n] = original indices in visit order. */
void oracle_argsort_desc(const float* keys, long n, int64_t* order_out) {
    sort_item* order = (sort_item*)malloc(sizeof(sort_item) * (n > 0 ? n : 1));
    for (long i = 0; i < n; ++i) { order[i].key = keys[i]; order[i].idx = (int)i; }
    qsort(order, n, sizeof(sort_item), cmp_desc);
    for (long i = 0; i < n; ++i) order_out[i] = order[i].idx;
    free(order);
}

/* ------------------------------------------------------------------------------------------
 * 3D box IoU -- lib/utils/cython_bbox_3d.pyx:32-80.  Mixed fp32/fp64 exactly as the generated C:
 *   side   = f32(hi - lo) + 1.0           (fp64)
 *   box_volume = f32(side*side*side)      (query box, :52-56)
 *   iw,ih,iss  = f32(f32(min - max) + 1.0)
 *   uv (C double) = side_x*side_y*side_z (fp64, boxes[n]) + (double)box_volume - (double)f32(f32(iw*ih)*iss)
 *   overlaps[n,k] = f32( (double)f32(f32(iw*ih)*iss) / uv )
 * ------------------------------------------------------------------------------------------ */
void oracle_bbox_overlaps_3d(const float* boxes, long N, const float* query, long K, float* out) {
    memset(out, 0, sizeof(float) * (size_t)N * (size_t)K);
    for (long k = 0; k < K; ++k) {
        const float* q = query + 6 * k;
        float box_volume = (float)((((double)(q[3] - q[0]) + 1.0) * ((double)(q[4] - q[1]) + 1.0)) *
                                   ((double)(q[5] - q[2]) + 1.0));
        for (long n = 0; n < N; ++n) {
            const float* b = boxes + 6 * n;
            float mn = q[3] < b[3] ? q[3] : b[3];
            float mx = q[0] > b[0] ? q[0] : b[0];
            float iw = (float)((double)(mn - mx) + 1.0);
            if (iw > 0) {
                mn = q[4] < b[4] ? q[4] : b[4];
                mx = q[1] > b[1] ? q[1] : b[1];
                float ih = (float)((double)(mn - mx) + 1.0);
                if (ih > 0) {
                    mn = q[5] < b[5] ? q[5] : b[5];
                    mx = q[2] > b[2] ? q[2] : b[2];
                    float iss = (float)((double)(mn - mx) + 1.0);
                    if (iss > 0) {
                        float iwh = iw * ih;
                        float inter = iwh * iss;
                        double uv = ((((double)(b[3] - b[0]) + 1.0) * ((double)(b[4] - b[1]) + 1.0)) *
                                     ((double)(b[5] - b[2]) + 1.0) + (double)box_volume) - (double)inter;
                        out[n * K + k] = (float)((double)inter / uv);
                    }
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * RoIAlign3D -- lib/modeling/roi_xfrom/roi_align_3d/src/roi_align_kernel_3d.cu
 * forward :16-151, backward :180-338.  fp32 arithmetic, sequential sums.
 * rois [R,7] = batch, x1,y1,z1,x2,y2,z2.  Pooled sizes (Ps,Ph,Pw).
 * Forward flat output index decomposes ps fastest, then pw, then ph (:87-91): the reference
 * writes (H,W,S) order into a tensor shaped [R,C,Ps,Ph,Pw]; we reproduce that quirk as is.
 * Backward reads top_diff[(n*C+c)*Ps*Ph*Pw + ps*Ph*Pw + ph*Pw + pw] (:272-275) and uses the
 * z guard "z < -0.1" (:187).  The adds are performed in output-index order (the reference uses
 * atomicAdd, i.e. an unspecified order).
 * ------------------------------------------------------------------------------------------ */
/* Sample coordinate  start + p*bin + (i+.5f)*bin/grid  (.cu:130-138).  The reference is CUDA built
 * with nvcc's default -fmad=true: the sm_100a PTX of the unmodified .cu contracts start + p*bin into
 * ONE fma and keeps the rest as mul, div.rn, add (checked with `nvcc -ptx`).  The coordinate is
 * evaluated the same way here, because on white-noise features a 1-ulp coordinate difference moves
 * the interpolated value by ~1e-5 -- more than the parity tolerance. */
#define SAMPLE_COORD(start, p, bin, i, g) (fmaf((float)(p), (bin), (start)) + (((i) + .5f) * (bin)) / (float)(g))

static float trilinear(const float* data, int S, int H, int W, float z, float y, float x) {
    if (z < -1.0 || z > S || y < -1.0 || y > H || x < -1.0 || x > W) return 0;
    if (z <= 0) z = 0;
    if (y <= 0) y = 0;
    if (x <= 0) x = 0;
    int zl = (int)z, yl = (int)y, xl = (int)x, zh, yh, xh;
    if (zl >= S - 1) { zh = zl = S - 1; z = (float)zl; } else zh = zl + 1;
    if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
    if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
    float lz = z - zl, ly = y - yl, lx = x - xl;
    float hz = (float)(1. - lz), hy = (float)(1. - ly), hx = (float)(1. - lx);
    float v1 = data[zl * (W * H) + yl * W + xl], v2 = data[zl * (W * H) + yl * W + xh];
    float v3 = data[zl * (W * H) + yh * W + xl], v4 = data[zl * (W * H) + yh * W + xh];
    float v5 = data[zh * (W * H) + yl * W + xl], v6 = data[zh * (W * H) + yl * W + xh];
    float v7 = data[zh * (W * H) + yh * W + xl], v8 = data[zh * (W * H) + yh * W + xh];
    float w1 = hz * hy * hx, w2 = hz * hy * lx, w3 = hz * ly * hx, w4 = hz * ly * lx;
    float w5 = lz * hy * hx, w6 = lz * hy * lx, w7 = lz * ly * hx, w8 = lz * ly * lx;
    return (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4 + w5 * v5 + w6 * v6 + w7 * v7 + w8 * v8);
}

void oracle_roialign3d_fwd(const float* feat, int B, int C, int S, int H, int W,
                           const float* rois, int R, int Ps, int Ph, int Pw,
                           float scale, int sr, float* out) {
    (void)B;
    long total = (long)R * C * Ps * Ph * Pw;
    for (long index = 0; index < total; ++index) {
        int ps = index % Ps;
        int pw = (index / Ps) % Pw;
        int ph = (index / Ps / Pw) % Ph;
        int c = (index / Ps / Pw / Ph) % C;
        int n = index / Ps / Pw / Ph / C;
        const float* r = rois + n * 7;
        int b = (int)r[0];
        float sw = r[1] * scale, sh = r[2] * scale, ss = r[3] * scale;
        float ew = r[4] * scale, eh = r[5] * scale, es = r[6] * scale;
        float rs = fmaxf(es - ss, 1.f), rw = fmaxf(ew - sw, 1.f), rh = fmaxf(eh - sh, 1.f);
        float bs = rs / Ps, bh = rh / Ph, bw = rw / Pw;
        const float* data = feat + ((long)b * C + c) * S * H * W;
        int gs = sr > 0 ? sr : (int)ceil(rs / Ps);
        int gh = sr > 0 ? sr : (int)ceil(rh / Ph);
        int gw = sr > 0 ? sr : (int)ceil(rw / Pw);
        const float count = (float)(gs * gh * gw);
        float acc = 0.f;
        for (int iz = 0; iz < gs; ++iz) {
            const float z = SAMPLE_COORD(ss, ps, bs, iz, gs);
            for (int iy = 0; iy < gh; ++iy) {
                const float y = SAMPLE_COORD(sh, ph, bh, iy, gh);
                for (int ix = 0; ix < gw; ++ix) {
                    const float x = SAMPLE_COORD(sw, pw, bw, ix, gw);
                    acc += trilinear(data, S, H, W, z, y, x);
                }
            }
        }
        out[index] = acc / count;
    }
}

void oracle_roialign3d_bwd(const float* top, int B, int C, int S, int H, int W,
                           const float* rois, int R, int Ps, int Ph, int Pw,
                           float scale, int sr, float* grad_in /* [B,C,S,H,W], zero-filled here */) {
    memset(grad_in, 0, sizeof(float) * (size_t)B * C * S * H * W);
    long total = (long)R * C * Ps * Ph * Pw;
    for (long index = 0; index < total; ++index) {
        int ps = index % Ps;
        int pw = (index / Ps) % Pw;
        int ph = (index / Ps / Pw) % Ph;
        int c = (index / Ps / Pw / Ph) % C;
        int n = index / Ps / Pw / Ph / C;
        const float* r = rois + n * 7;
        int b = (int)r[0];
        float sw = r[1] * scale, sh = r[2] * scale, ss = r[3] * scale;
        float ew = r[4] * scale, eh = r[5] * scale, es = r[6] * scale;
        float rw = fmaxf(ew - sw, 1.f), rh = fmaxf(eh - sh, 1.f), rs = fmaxf(es - ss, 1.f);
        float bh = rh / Ph, bw = rw / Pw, bs = rs / Ps;
        float* g = grad_in + ((long)b * C + c) * S * H * W;
        const float t = top[((long)n * C + c) * Ps * Ph * Pw + ps * Ph * Pw + ph * Pw + pw];
        int gs = sr > 0 ? sr : (int)ceil(rs / Ps);
        int gh = sr > 0 ? sr : (int)ceil(rh / Ph);
        int gw = sr > 0 ? sr : (int)ceil(rw / Pw);
        const float count = (float)(gs * gh * gw);
        for (int iz = 0; iz < gs; ++iz) {
            float z = SAMPLE_COORD(ss, ps, bs, iz, gs);
            for (int iy = 0; iy < gh; ++iy) {
                float y = SAMPLE_COORD(sh, ph, bh, iy, gh);
                for (int ix = 0; ix < gw; ++ix) {
                    float x = SAMPLE_COORD(sw, pw, bw, ix, gw);
                    float zz = z, yy = y, xx = x;
                    if (zz < -0.1 || zz > S || yy < -1.0 || yy > H || xx < -1.0 || xx > W) continue;
                    if (zz <= 0) zz = 0;
                    if (yy <= 0) yy = 0;
                    if (xx <= 0) xx = 0;
                    int zl = (int)zz, yl = (int)yy, xl = (int)xx, zh, yh, xh;
                    if (zl >= S - 1) { zh = zl = S - 1; zz = (float)zl; } else zh = zl + 1;
                    if (yl >= H - 1) { yh = yl = H - 1; yy = (float)yl; } else yh = yl + 1;
                    if (xl >= W - 1) { xh = xl = W - 1; xx = (float)xl; } else xh = xl + 1;
                    float lz = zz - zl, ly = yy - yl, lx = xx - xl;
                    float hz = (float)(1. - lz), hy = (float)(1. - ly), hx = (float)(1. - lx);
                    float w1 = hz * hy * hx, w2 = hz * hy * lx, w3 = hz * ly * hx, w4 = hz * ly * lx;
                    float w5 = lz * hy * hx, w6 = lz * hy * lx, w7 = lz * ly * hx, w8 = lz * ly * lx;
                    g[zl * H * W + yl * W + xl] += t * w1 / count;
                    g[zl * H * W + yl * W + xh] += t * w2 / count;
                    g[zl * H * W + yh * W + xl] += t * w3 / count;
                    g[zl * H * W + yh * W + xh] += t * w4 / count;
                    g[zh * H * W + yl * W + xl] += t * w5 / count;
                    g[zh * H * W + yl * W + xh] += t * w6 / count;
                    g[zh * H * W + yh * W + xl] += t * w7 / count;
                    g[zh * H * W + yh * W + xh] += t * w8 / count;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * PRM peak stimulation -- lib/prm/peak_stimulation_3d.py:9-41 with the median filter of
 * lib/prm/peak_response_mapping_3d.py:45-49.
 * ATen max_pool3d semantics (window scanned z,y,x; update when val > max || isnan(val); initial
 * index = window start, initial max = -inf) on the -inf padded volume; a voxel is a peak when the
 * arg-max of its own window is itself (:18,25).
 * filter_mode: 0 none, 1 lower median per (b,c) (torch.median), 2 per-(b,c) thresholds given in
 * thr_in[B*A].  peaks_out [cap,5] int64 rows (b,c,z,y,x) in lexicographic order (torch.nonzero).
 * agg_out[B*A] (may be NULL) = sum(input*peak)/sum(peak) accumulated in fp32 raster order.
 * thr_out[B*A] (may be NULL) receives the thresholds used.  Returns the number of peaks (which may
 * exceed cap; only the first cap rows are written).
 * ------------------------------------------------------------------------------------------ */
static int cmp_float_asc(const void* a, const void* b) {
    float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

long oracle_peak_stimulation(const float* in, int B, int A, int S, int H, int W, int win,
                             int filter_mode, const float* thr_in, int64_t* peaks_out, long cap,
                             float* agg_out, float* thr_out) {
    const int off = (win - 1) / 2;
    const long V = (long)S * H * W;
    long npk = 0;
    float* tmp = filter_mode == 1 ? (float*)malloc(sizeof(float) * V) : NULL;
    for (int b = 0; b < B; ++b) for (int a = 0; a < A; ++a) {
        const float* v = in + ((long)b * A + a) * V;
        float thr = 0.f;
        if (filter_mode == 1) {
            int has_nan = 0;
            for (long i = 0; i < V; ++i) { tmp[i] = v[i]; has_nan |= isnan(v[i]); }
            if (has_nan) thr = NAN;
            else { qsort(tmp, V, sizeof(float), cmp_float_asc); thr = tmp[(V - 1) / 2]; }
        } else if (filter_mode == 2) thr = thr_in[b * A + a];
        if (thr_out) thr_out[b * A + a] = thr;
        float num = 0.f, den = 0.f;
        for (int z = 0; z < S; ++z) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
            /* window in padded coords: [z, z+win) etc.; own padded coord (z+off, y+off, x+off) */
            float best = -INFINITY;
            int bz = z, by = y, bx = x;                      /* window start (padded coords) */
            for (int dz = 0; dz < win; ++dz) for (int dy = 0; dy < win; ++dy) for (int dx = 0; dx < win; ++dx) {
                int pz = z + dz - off, py = y + dy - off, px = x + dx - off;   /* unpadded coords */
                float val = (pz < 0 || pz >= S || py < 0 || py >= H || px < 0 || px >= W)
                                ? -INFINITY : v[((long)pz * H + py) * W + px];
                if (val > best || isnan(val)) { best = val; bz = z + dz; by = y + dy; bx = x + dx; }
            }
            int is_peak = (bz == z + off && by == y + off && bx == x + off);
            float c = v[((long)z * H + y) * W + x];
            if (is_peak && filter_mode != 0) is_peak = (c >= thr);
            if (is_peak) {
                if (npk < cap) {
                    int64_t* p = peaks_out + 5 * npk;
                    p[0] = b; p[1] = a; p[2] = z; p[3] = y; p[4] = x;
                }
                ++npk;
                num += c; den += 1.f;
            }
        }
        if (agg_out) agg_out[b * A + a] = num / den;
    }
    free(tmp);
    return npk;
}

/* ------------------------------------------------------------------------------------------
 * 2D Otsu -- tools/otsu.py:199-284 (otsu_py_2d_fast, k = -1 only).
 * image/prm: n integer samples (uint16 here; the callers pass uint16, binarization_soma.py:85-91).
 * Follows the reference structure: numpy histogram2d binning (each axis over its own min..max,
 * G = g_max-g_min+1 bins, fp64 linspace edges, searchsorted-right, last edge inclusive), hist.T,
 * prob, first moments, then the incremental oblique-line scan over b with numpy's pairwise
 * summation, first-strictly-greater arg-max starting at 0, and the gray-level mask loop.
 * Outputs: mask (0/255), *b_max, optional hist (G*G doubles, [prm_bin][img_bin] = hist.T).
 * Returns 0, or 1 when no b wins (the reference then raises NameError on k_max, :277),
 * or -1 when G is too large for this restatement.
 * ------------------------------------------------------------------------------------------ */
static double pairwise_sum(const double* a, long n, long stride) {
    if (n < 8) {
        double res = 0.;
        for (long i = 0; i < n; ++i) res += a[i * stride];
        return res;
    } else if (n <= 128) {
        double r[8];
        long i;
        for (i = 0; i < 8; ++i) r[i] = a[i * stride];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[(i + j) * stride];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i * stride];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum(a, n2, stride) + pairwise_sum(a + n2 * stride, n - n2, stride);
    }
}

static void np_linspace(double start, double stop, long num, double* y) {   /* endpoint=True */
    long div = num - 1;
    double delta = stop - start;
    if (div > 0) {
        double step = delta / (double)div;
        if (step == 0) for (long i = 0; i < num; ++i) y[i] = ((double)i / (double)div) * delta + start;
        else for (long i = 0; i < num; ++i) y[i] = (double)i * step + start;
    } else {
        for (long i = 0; i < num; ++i) y[i] = (double)i * delta + start;
    }
    if (num > 1) y[num - 1] = stop;
}

static long searchsorted_right(const double* e, long m, double v) {
    long lo = 0, hi = m;
    while (lo < hi) { long mid = (lo + hi) >> 1; if (e[mid] <= v) lo = mid + 1; else hi = mid; }
    return lo;
}

/* bin index of value v on an axis with G bins over [vmin,vmax] (numpy histogramdd semantics) */
static void axis_edges(double vmin, double vmax, long G, double* edges) {
    if (vmin == vmax) { vmin -= 0.5; vmax += 0.5; }
    np_linspace(vmin, vmax, G + 1, edges);
}
static long axis_bin(const double* edges, long G, double v) {
    long idx = searchsorted_right(edges, G + 1, v);
    if (v == edges[G]) idx -= 1;
    return idx - 1;    /* core bin; inputs lie inside [min,max] so 0 <= bin < G */
}

int oracle_otsu_2d_fast(const uint16_t* image, const uint16_t* prm, long n,
                        uint8_t* mask_out, int* b_max_out, double* hist_out, int* G_out) {
    long g_min = 65536, g_max = -1, p_min = 65536, p_max = -1;
    for (long i = 0; i < n; ++i) {
        if (image[i] < g_min) g_min = image[i];
        if (image[i] > g_max) g_max = image[i];
        if (prm[i] < p_min) p_min = prm[i];
        if (prm[i] > p_max) p_max = prm[i];
    }
    if (n <= 0) return -1;
    const long G = g_max - g_min + 1;
    if (G_out) *G_out = (int)G;
    if (G < 1 || G > 4096) return -1;
    double* e1 = (double*)malloc(sizeof(double) * (G + 1));
    double* e2 = (double*)malloc(sizeof(double) * (G + 1));
    axis_edges((double)g_min, (double)g_max, G, e1);
    axis_edges((double)p_min, (double)p_max, G, e2);
    double* c1 = (double*)malloc(sizeof(double) * G);
    double* c2 = (double*)malloc(sizeof(double) * G);
    for (long i = 0; i < G; ++i) { c1[i] = (e1[i] + e1[i + 1]) / 2; c2[i] = (e2[i] + e2[i + 1]) / 2; }
    /* hist[r][c]: r = prm bin, c = image bin (hist.T, :206) */
    double* hist = (double*)calloc((size_t)G * G, sizeof(double));
    long* lut1 = (long*)malloc(sizeof(long) * G);
    long np_ = p_max - p_min + 1;
    long* lut2 = (long*)malloc(sizeof(long) * np_);
    for (long v = 0; v < G; ++v) lut1[v] = axis_bin(e1, G, (double)(g_min + v));
    for (long v = 0; v < np_; ++v) lut2[v] = axis_bin(e2, G, (double)(p_min + v));
    for (long i = 0; i < n; ++i) hist[lut2[prm[i] - p_min] * G + lut1[image[i] - g_min]] += 1.0;
    if (hist_out) memcpy(hist_out, hist, sizeof(double) * (size_t)G * G);
    double total = pairwise_sum(hist, G * G, 1);
    double* prob = (double*)malloc(sizeof(double) * (size_t)G * G);
    double* ui = (double*)malloc(sizeof(double) * (size_t)G * G);
    double* up = (double*)malloc(sizeof(double) * (size_t)G * G);
    for (long r = 0; r < G; ++r) for (long c = 0; c < G; ++c) {
        double p = hist[r * G + c] / total;
        prob[r * G + c] = p;
        ui[r * G + c] = c1[c] * p;
        up[r * G + c] = c2[r] * p;
    }
    const double ut0 = pairwise_sum(ui, G * G, 1), ut1 = pairwise_sum(up, G * G, 1);
    double var_b_max = 0;
    long b_max = 0;
    int found = 0;
    const long b_dw = 2 * g_min + 1, b_up = 2 * g_max - 1;       /* int((1-k)*g_min+1), int((1-k)*g_max-1) */
    const long nx = G - 1;                                        /* x_range = g_min .. g_max-1 */
    long* yr = (long*)malloc(sizeof(long) * (nx > 0 ? nx : 1));
    double* gat = (double*)malloc(sizeof(double) * 3 * (nx > 0 ? nx : 1));
    double p0 = 0, u00 = 0, u01 = 0;
    long b = b_dw;
    for (long j = 0; j < nx; ++j) {
        long line = b - (g_min + j);
        long y = (line < g_max ? line : g_max) - g_min;
        yr[j] = y > 0 ? y : 0;
    }
    for (long j = 0; j < nx; ++j) if (yr[j] > 0) {
        p0 += pairwise_sum(prob + j, yr[j], G);
        u00 += pairwise_sum(ui + j, yr[j], G);
        u01 += pairwise_sum(up + j, yr[j], G);
    }
    for (;;) {
        double p1 = 1. - p0;
        double u10 = (ut0 - p0 * u00) / p1, u11 = (ut1 - p0 * u01) / p1;
        double d0 = u00 - ut0, d1 = u01 - ut1, f0 = u10 - ut0, f1 = u11 - ut1;
        double var_b = ((p0 * d0) * d0 + (p1 * f0) * f0) + ((p0 * d1) * d1 + (p1 * f1) * f1);
        if (var_b > var_b_max) { var_b_max = var_b; b_max = b; found = 1; }
        ++b;
        if (b >= b_up) break;
        /* incremental update (:251-268) */
        long m = 0;
        for (long j = 0; j < nx; ++j) {
            long line = b - (g_min + j);
            long y = (line < g_max ? line : g_max) - g_min;
            long ynew = y > 0 ? y : 0;
            if (ynew - yr[j] > 0 && yr[j] > 0) {            /* ovlp: one more cell in row yr[j] */
                gat[m] = prob[yr[j] * G + j];
                gat[nx + m] = ui[yr[j] * G + j];
                gat[2 * nx + m] = up[yr[j] * G + j];
                ++m;
            }
        }
        p0 += pairwise_sum(gat, m, 1);
        u00 += pairwise_sum(gat + nx, m, 1);
        u01 += pairwise_sum(gat + 2 * nx, m, 1);
        for (long j = 0; j < nx; ++j) {
            long line = b - (g_min + j);
            long y = (line < g_max ? line : g_max) - g_min;
            long ynew = y > 0 ? y : 0;
            if (ynew - yr[j] > 0 && !(yr[j] > 0)) {         /* bg_add: a whole new column prefix */
                p0 += pairwise_sum(prob + j, ynew, G);
                u00 += pairwise_sum(ui + j, ynew, G);
                u01 += pairwise_sum(up + j, ynew, G);
            }
            yr[j] = ynew;
        }
    }
    int status = found ? 0 : 1;
    if (b_max_out) *b_max_out = (int)b_max;
    if (mask_out) {
        memset(mask_out, 255, n);
        if (found) {
            long x_g_min = b_max - g_min;                    /* int((g_min-b_max)/k_max), k=-1 */
            long x_hi = x_g_min < g_max ? x_g_min : g_max;
            for (long i = 0; i < n; ++i) {
                long I = image[i];
                if (I >= g_min && I < x_hi) {
                    long y0 = b_max - I;
                    if (y0 > g_max + 1) y0 = g_max + 1;
                    if ((long)prm[i] < y0) mask_out[i] = 0;
                }
            }
        }
    }
    free(e1); free(e2); free(c1); free(c2); free(hist); free(lut1); free(lut2);
    free(prob); free(ui); free(up); free(yr); free(gat);
    return status;
}

/* ------------------------------------------------------------------------------------------
 * Label paste-back -- tools/binarization_soma.py:66-104 (the "write where still 0" rule,
 * :100-102, and the survivor test `mask_id in np.unique(seg)`, :103).
 * seg [S,H,W] uint16 (updated in place).  Instances are given in VISIT order; instance i has
 * label ids[i], box (x1,y1,z1,x2,y2,z2) inclusive integer voxel coords inside the volume and a
 * uint8 mask crop (non-zero = foreground) of shape [z2-z1+1, y2-y1+1, x2-x1+1] at mask_off[i].
 * survive_out[i] = 1 when label ids[i] is present in the volume after its own paste
 * (later instances never overwrite, so this equals "present at the end").
 * ------------------------------------------------------------------------------------------ */
void oracle_paste_labels(uint16_t* seg, int S, int H, int W, int n_inst, const int32_t* boxes,
                         const uint16_t* ids, const uint8_t* masks, const int64_t* mask_off,
                         uint8_t* survive_out) {
    (void)S;
    for (int i = 0; i < n_inst; ++i) {
        const int32_t* b = boxes + 6 * i;
        int sx = b[3] - b[0] + 1, sy = b[4] - b[1] + 1, sz = b[5] - b[2] + 1;
        const uint8_t* m = masks + mask_off[i];
        int any = 0;
        for (int z = 0; z < sz; ++z) for (int y = 0; y < sy; ++y) for (int x = 0; x < sx; ++x) {
            uint16_t* p = seg + ((long)(b[2] + z) * H + (b[1] + y)) * W + (b[0] + x);
            if (*p == 0 && m[((long)z * sy + y) * sx + x]) { *p = ids[i]; any = 1; }
        }
        if (survive_out) survive_out[i] = (uint8_t)any;
    }
}
