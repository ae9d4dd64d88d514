// roialign3d.cu -- RoIAlign3D forward / backward for sm_100a
// (replaces lib/modeling/roi_xfrom/roi_align_3d/src/roi_align_kernel_3d.cu:81-151 and :238-338).
//
// The reference spends 8 scattered taps x sr^3 samples per output element (64 taps at sr=2) with
// adjacent threads striding in z, and the backward issues 64 global atomicAdds per element.
// Trilinear sampling is separable: the sr samples of a bin along one axis collapse into a small
// dense matrix W_axis[f][p] (f = footprint voxel, p = bin).  Then
//     out[ps,ph,pw] = (1/count) * sum_z Wz[z][ps] sum_y Wy[y][ph] sum_x Wx[x][pw] feat[z,y,x]
// which is three tiny tensor contractions per (roi, channel) instead of 64 taps per output.
//
// forward:  CTA = (roi, 32-channel chunk).  Axis tables are built once per CTA in shared memory;
//           each thread produces a whole P-vector per pass (x-pass reads the feature rows straight
//           from L2, y/z passes stream shared-memory intermediates), weights are broadcast 128-bit
//           shared loads.  The z-pass thread owns (c, ph, pw) and writes its P results contiguously:
//           exactly the reference's (H,W,S) bin order, so output stores of a CTA are one contiguous,
//           fully coalesced span.  Nothing is pre-zeroed.
// backward: deterministic and atomics-free ("owner computes").  CTA = (batch, channel chunk,
//           8x8x8 feature tile); it walks the RoIs of its batch in index order, runs the adjoint
//           three-pass contraction restricted to the tile, accumulates in REGISTERS (each thread owns
//           fixed feature rows) and finally writes every grad_in element exactly once (zero fill
//           folded in).  The reference's quirks are kept in layout 0: grad_out is read in (S,H,W)
//           order and the z guard is -0.1 (roi_align_kernel_3d.cu:187,272-275).
#include "common.cuh"

namespace b200seg {

constexpr int RA_THREADS = 256;
constexpr int RA_FMAX = 16;        // max footprint voxels per axis handled by the separable forward path
constexpr int RA_CC = 32;          // channels per forward CTA
constexpr int RA_FWD_SMEM_FLOATS = 16384;   // 64 KB of intermediates

struct AxisP {
    float start, bin;
    int g;          // samples per bin along this axis
    int dim;
    double guard;   // samples with coordinate < guard (or > dim) contribute nothing (-1.0, or -0.1 for bwd z)
};

struct Tap { int low, high; float l, h; bool valid; };

// one sample of roi_align_kernel_3d.cu:130-138 + :19-58 (forward) / :187-224 (backward)
__device__ __forceinline__ Tap axis_sample(const AxisP& a, int p, int i) {
    Tap t;
    // fma(p, bin, start) + ((i+.5)*bin)/g : the contraction nvcc applies to the reference kernel
    float c = __fadd_rn(__fmaf_rn((float)p, a.bin, a.start), __fdiv_rn(__fmul_rn(i + .5f, a.bin), (float)a.g));
    t.valid = !((double)c < a.guard || c > (float)a.dim);
    if (c <= 0) c = 0;
    int low = (int)c;
    if (low >= a.dim - 1) { t.high = t.low = a.dim - 1; c = (float)t.low; }
    else { t.low = low; t.high = low + 1; }
    t.l = c - t.low;
    t.h = 1.f - t.l;
    return t;
}

__device__ __forceinline__ void roi_axes(const float* __restrict__ roi, float scale, int sr,
                                         int Ps, int Ph, int Pw, int S, int H, int W, double zguard,
                                         AxisP& az, AxisP& ay, AxisP& ax, int& batch, float& count) {
    batch = (int)roi[0];
    const float sw = roi[1] * scale, sh = roi[2] * scale, ss = roi[3] * scale;
    const float ew = roi[4] * scale, eh = roi[5] * scale, es = roi[6] * scale;
    const float rs = fmaxf(es - ss, 1.f), rw = fmaxf(ew - sw, 1.f), rh = fmaxf(eh - sh, 1.f);
    az.start = ss; az.bin = rs / Ps; az.g = sr > 0 ? sr : (int)ceilf(rs / Ps); az.dim = S; az.guard = zguard;
    ay.start = sh; ay.bin = rh / Ph; ay.g = sr > 0 ? sr : (int)ceilf(rh / Ph); ay.dim = H; ay.guard = -1.0;
    ax.start = sw; ax.bin = rw / Pw; ax.g = sr > 0 ? sr : (int)ceilf(rw / Pw); ax.dim = W; ax.guard = -1.0;
    count = (float)(az.g * ay.g * ax.g);
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
struct __align__(16) FwdShared {
    float w[3][RA_FMAX][16];     // [axis z,y,x][footprint voxel][bin]
    int lo[3], hi[3];
};

template <typename T, int PT>
__global__ void __launch_bounds__(RA_THREADS)
roialign3d_fwd_kernel(const T* __restrict__ feat, const float* __restrict__ rois, T* __restrict__ out,
                      int C, int S, int H, int W, int Ps, int Ph, int Pw, float scale, int sr, int layout) {
    extern __shared__ __align__(16) float s_buf[];
    __shared__ FwdShared sh;
    const int r = blockIdx.x;
    const int c0 = blockIdx.y * RA_CC;
    const int nc = min(RA_CC, C - c0);
    const int tid = threadIdx.x;

    AxisP ax[3];
    int batch; float count;
    roi_axes(rois + (size_t)r * 7, scale, sr, Ps, Ph, Pw, S, H, W, -1.0, ax[0], ax[1], ax[2], batch, count);
    const int Pa[3] = {Ps, Ph, Pw};

    // ---- axis tables: footprint range, then dense W[f][p] ------------------------------------------
    if (tid < 3) { sh.lo[tid] = 0x7fffffff; sh.hi[tid] = -1; }
    for (int i = tid; i < 3 * RA_FMAX * 16; i += RA_THREADS) (&sh.w[0][0][0])[i] = 0.f;
    __syncthreads();
    if (tid < 3 * 16) {
        const int a = tid >> 4, p = tid & 15;
        if (p < Pa[a]) {
            int lo = 0x7fffffff, hi = -1;
            for (int i = 0; i < ax[a].g; ++i) {
                const Tap t = axis_sample(ax[a], p, i);
                if (t.valid) { lo = min(lo, t.low); hi = max(hi, t.high); }
            }
            if (hi >= 0) { atomicMin(&sh.lo[a], lo); atomicMax(&sh.hi[a], hi); }
        }
    }
    __syncthreads();
    const int zlo = sh.lo[0], ylo = sh.lo[1], xlo = sh.lo[2];
    const int Fz = sh.hi[0] - zlo + 1, Fy = sh.hi[1] - ylo + 1, Fx = sh.hi[2] - xlo + 1;
    const bool empty = sh.hi[0] < 0 || sh.hi[1] < 0 || sh.hi[2] < 0;
    const size_t P3 = (size_t)Ps * Ph * Pw;
    T* out_r = out + ((size_t)r * C + c0) * P3;
    if (empty) {                                        // every sample falls outside: zeros
        for (size_t i = tid; i < (size_t)nc * P3; i += RA_THREADS) out_r[i] = from_f<T>(0.f);
        return;
    }
    const T* feat_b = feat + ((size_t)batch * C + c0) * S * H * W;
    const bool separable = Fz <= RA_FMAX && Fy <= RA_FMAX && Fx <= RA_FMAX &&
                           (Fz * Fy * PT + Fz * PT * PT) <= RA_FWD_SMEM_FLOATS;
    if (!separable) {
        // direct evaluation (reference arithmetic), coalesced over the output span of this CTA
        for (size_t idx = tid; idx < (size_t)nc * P3; idx += RA_THREADS) {
            const int c = (int)(idx / P3);
            const int e = (int)(idx % P3);
            int ps, ph, pw;
            if (layout == 0) { ps = e % Ps; pw = (e / Ps) % Pw; ph = e / (Ps * Pw); }
            else { pw = e % Pw; ph = (e / Pw) % Ph; ps = e / (Pw * Ph); }
            const T* data = feat_b + (size_t)c * S * H * W;
            float acc = 0.f;
            for (int iz = 0; iz < ax[0].g; ++iz) {
                const Tap tz = axis_sample(ax[0], ps, iz);
                for (int iy = 0; iy < ax[1].g; ++iy) {
                    const Tap ty = axis_sample(ax[1], ph, iy);
                    for (int ix = 0; ix < ax[2].g; ++ix) {
                        const Tap tx = axis_sample(ax[2], pw, ix);
                        if (!(tz.valid && ty.valid && tx.valid)) continue;
                        const T* p0 = data + ((size_t)tz.low * H + ty.low) * W;
                        const T* p1 = data + ((size_t)tz.low * H + ty.high) * W;
                        const T* p2 = data + ((size_t)tz.high * H + ty.low) * W;
                        const T* p3 = data + ((size_t)tz.high * H + ty.high) * W;
                        acc += tz.h * ty.h * tx.h * to_f(p0[tx.low]) + tz.h * ty.h * tx.l * to_f(p0[tx.high]) +
                               tz.h * ty.l * tx.h * to_f(p1[tx.low]) + tz.h * ty.l * tx.l * to_f(p1[tx.high]) +
                               tz.l * ty.h * tx.h * to_f(p2[tx.low]) + tz.l * ty.h * tx.l * to_f(p2[tx.high]) +
                               tz.l * ty.l * tx.h * to_f(p3[tx.low]) + tz.l * ty.l * tx.l * to_f(p3[tx.high]);
                    }
                }
            }
            out_r[idx] = from_f<T>(acc / count);
        }
        return;
    }
    if (tid < 3 * 16) {
        const int a = tid >> 4, p = tid & 15;
        if (p < Pa[a]) {
            const int lo = sh.lo[a];
            for (int i = 0; i < ax[a].g; ++i) {
                const Tap t = axis_sample(ax[a], p, i);
                if (t.valid) { sh.w[a][t.low - lo][p] += t.h; sh.w[a][t.high - lo][p] += t.l; }
            }
        }
    }
    __syncthreads();

    const int per_c = Fz * Fy * PT + Fz * PT * PT;
    const int cc_fit = min(nc, RA_FWD_SMEM_FLOATS / per_c);
    for (int cs = 0; cs < nc; cs += cc_fit) {
        const int cc = min(cc_fit, nc - cs);
        float* T1 = s_buf;                              // [cc][Fz][Fy][PT]
        float* T2 = s_buf + (size_t)cc * Fz * Fy * PT;  // [cc][Fz][PT(ph)][PT(pw)]
        // ---- pass X: thread = (c, z, y) row; PT-vector over pw ------------------------------------
        for (int item = tid; item < cc * Fz * Fy; item += RA_THREADS) {
            const int y = item % Fy, z = (item / Fy) % Fz, c = item / (Fy * Fz);
            const T* row = feat_b + (((size_t)(cs + c) * S + (zlo + z)) * H + (ylo + y)) * W + xlo;
            float acc[PT];
#pragma unroll
            for (int p = 0; p < PT; ++p) acc[p] = 0.f;
            for (int x = 0; x < Fx; ++x) {
                const float v = to_f(row[x]);
                const float4* wv = reinterpret_cast<const float4*>(&sh.w[2][x][0]);
#pragma unroll
                for (int q = 0; q < PT / 4; ++q) {
                    const float4 w4 = wv[q];
                    acc[4 * q + 0] += w4.x * v; acc[4 * q + 1] += w4.y * v;
                    acc[4 * q + 2] += w4.z * v; acc[4 * q + 3] += w4.w * v;
                }
            }
            float4* dst = reinterpret_cast<float4*>(T1 + (size_t)item * PT);
#pragma unroll
            for (int q = 0; q < PT / 4; ++q) dst[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        }
        __syncthreads();
        // ---- pass Y: thread = (c, z, pw); PT-vector over ph ----------------------------------------
        for (int item = tid; item < cc * Fz * PT; item += RA_THREADS) {
            const int pw = item % PT, cz = item / PT;
            if (pw >= Pw) continue;
            float acc[PT];
#pragma unroll
            for (int p = 0; p < PT; ++p) acc[p] = 0.f;
            const float* src = T1 + (size_t)cz * Fy * PT + pw;
            for (int y = 0; y < Fy; ++y) {
                const float v = src[y * PT];
                const float4* wv = reinterpret_cast<const float4*>(&sh.w[1][y][0]);
#pragma unroll
                for (int q = 0; q < PT / 4; ++q) {
                    const float4 w4 = wv[q];
                    acc[4 * q + 0] += w4.x * v; acc[4 * q + 1] += w4.y * v;
                    acc[4 * q + 2] += w4.z * v; acc[4 * q + 3] += w4.w * v;
                }
            }
            float* dst = T2 + (size_t)cz * PT * PT + pw;
#pragma unroll
            for (int p = 0; p < PT; ++p) dst[p * PT] = acc[p];
        }
        __syncthreads();
        // ---- pass Z: thread = (c, ph, pw); PT-vector over ps, written contiguously -----------------
        for (int item = tid; item < cc * Ph * Pw; item += RA_THREADS) {
            const int pw = item % Pw, ph = (item / Pw) % Ph, c = item / (Pw * Ph);
            float acc[PT];
#pragma unroll
            for (int p = 0; p < PT; ++p) acc[p] = 0.f;
            const float* src = T2 + (size_t)c * Fz * PT * PT + ph * PT + pw;
            for (int z = 0; z < Fz; ++z) {
                const float v = src[(size_t)z * PT * PT];
                const float4* wv = reinterpret_cast<const float4*>(&sh.w[0][z][0]);
#pragma unroll
                for (int q = 0; q < PT / 4; ++q) {
                    const float4 w4 = wv[q];
                    acc[4 * q + 0] += w4.x * v; acc[4 * q + 1] += w4.y * v;
                    acc[4 * q + 2] += w4.z * v; acc[4 * q + 3] += w4.w * v;
                }
            }
            T* o = out_r + (size_t)(cs + c) * P3;
            if (layout == 0) {
                T* o2 = o + ((size_t)ph * Pw + pw) * Ps;
#pragma unroll
                for (int p = 0; p < PT; ++p) if (p < Ps) o2[p] = from_f<T>(acc[p] / count);
            } else {
#pragma unroll
                for (int p = 0; p < PT; ++p) if (p < Ps) o[((size_t)p * Ph + ph) * Pw + pw] = from_f<T>(acc[p] / count);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
constexpr int RB_T = 8;            // feature tile edge (z,y,x)

struct RoiMeta {                   // one per RoI, built by roialign3d_prep_kernel
    float start[3], bin[3];
    int g[3];
    int lo[3], hi[3];              // footprint (inclusive), hi < lo when no valid sample on that axis
    int batch;
    float count;
    int pad[3];
};

__global__ void roialign3d_prep_kernel(const float* __restrict__ rois, int R, RoiMeta* __restrict__ meta,
                                       int S, int H, int W, int Ps, int Ph, int Pw, float scale, int sr, double zguard) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    AxisP ax[3];
    int batch; float count;
    roi_axes(rois + (size_t)r * 7, scale, sr, Ps, Ph, Pw, S, H, W, zguard, ax[0], ax[1], ax[2], batch, count);
    const int Pa[3] = {Ps, Ph, Pw};
    RoiMeta m;
    for (int a = 0; a < 3; ++a) {
        int lo = 0x7fffffff, hi = -1;
        for (int p = 0; p < Pa[a]; ++p)
            for (int i = 0; i < ax[a].g; ++i) {
                const Tap t = axis_sample(ax[a], p, i);
                if (t.valid) { lo = min(lo, t.low); hi = max(hi, t.high); }
            }
        m.start[a] = ax[a].start; m.bin[a] = ax[a].bin; m.g[a] = ax[a].g; m.lo[a] = lo; m.hi[a] = hi;
    }
    m.batch = batch; m.count = count; m.pad[0] = m.pad[1] = m.pad[2] = 0;
    meta[r] = m;
}

template <int PT> struct BwdCfg { static constexpr int CC = PT == 8 ? 16 : 4; };

template <typename T, int PT>
__global__ void __launch_bounds__(RA_THREADS)
roialign3d_bwd_kernel(const T* __restrict__ gout, const RoiMeta* __restrict__ meta, T* __restrict__ gin,
                      int C, int S, int H, int W, int R, int Ps, int Ph, int Pw, int layout, double zguard,
                      int tiles_x, int tiles_y) {
    constexpr int CC = BwdCfg<PT>::CC;
    constexpr int ROWS = CC * RB_T * RB_T / RA_THREADS;        // feature rows (8 voxels) owned per thread
    static_assert(CC * RB_T * RB_T % RA_THREADS == 0, "row ownership must be exact");
    extern __shared__ __align__(16) float s_dynb[];
    __shared__ __align__(16) float s_w[3][RB_T][16];            // tile-local adjoint tables W[t][p]
    float* s_T2 = s_dynb;                                       // [c][tz][ph][pw]   CC*RB_T*PT*PT
    float* s_T1 = s_dynb + CC * RB_T * PT * PT;                 // [c][tz][ty][pw]   CC*RB_T*RB_T*PT

    const int tile = blockIdx.x;
    const int tx0 = (tile % tiles_x) * RB_T, ty0 = ((tile / tiles_x) % tiles_y) * RB_T, tz0 = (tile / (tiles_x * tiles_y)) * RB_T;
    const int c0 = blockIdx.y * CC;
    const int nc = min(CC, C - c0);
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int t0[3] = {tz0, ty0, tx0};
    const int dims[3] = {S, H, W};
    const int Pa[3] = {Ps, Ph, Pw};
    const size_t P3 = (size_t)Ps * Ph * Pw;

    float acc[ROWS][RB_T];
#pragma unroll
    for (int i = 0; i < ROWS; ++i)
#pragma unroll
        for (int x = 0; x < RB_T; ++x) acc[i][x] = 0.f;

    for (int r = 0; r < R; ++r) {
        const RoiMeta m = meta[r];                              // uniform broadcast load
        if (m.batch != b) continue;
        bool hit = true;
        int rlo[3], rhi[3];                                     // tile-local index ranges touched by this RoI
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            rlo[a] = max(m.lo[a], t0[a]) - t0[a];
            rhi[a] = min(m.hi[a], min(t0[a] + RB_T, dims[a]) - 1) - t0[a];
            hit = hit && (m.hi[a] >= m.lo[a]) && rhi[a] >= rlo[a];
        }
        if (!hit) continue;                                     // CTA-uniform
        __syncthreads();                                        // previous RoI done with s_w / s_T*
        for (int i = tid; i < 3 * RB_T * 16; i += RA_THREADS) (&s_w[0][0][0])[i] = 0.f;
        __syncthreads();
        if (tid < 3 * 16) {
            const int a = tid >> 4, p = tid & 15;
            if (p < Pa[a]) {
                AxisP ap;
                ap.start = m.start[a]; ap.bin = m.bin[a]; ap.g = m.g[a]; ap.dim = dims[a];
                ap.guard = a == 0 ? zguard : -1.0;
                for (int i = 0; i < ap.g; ++i) {
                    const Tap t = axis_sample(ap, p, i);
                    if (!t.valid) continue;
                    const int l = t.low - t0[a], h = t.high - t0[a];
                    if (l >= 0 && l < RB_T) s_w[a][l][p] += t.h;
                    if (h >= 0 && h < RB_T) s_w[a][h][p] += t.l;
                }
            }
        }
        __syncthreads();
        const int nz = rhi[0] - rlo[0] + 1, ny = rhi[1] - rlo[1] + 1;
        const T* g_r = gout + ((size_t)r * C + c0) * P3;
        // ---- pass Z^T: thread = (c, ph, pw): T2[c][tz][ph][pw] = sum_ps Wz[tz][ps] * G[c][ps][ph][pw]
        for (int item = tid; item < nc * Ph * Pw; item += RA_THREADS) {
            const int pw = item % Pw, ph = (item / Pw) % Ph, c = item / (Pw * Ph);
            float g[PT];
            const T* gc = g_r + (size_t)c * P3;
#pragma unroll
            for (int p = 0; p < PT; ++p) {
                g[p] = 0.f;
                // grad_out is read in (S,H,W) order in both layouts: that is what the reference does
                // (.cu:272-275) and it is also the exact adjoint of the layout-1 forward.
                if (p < Ps) g[p] = to_f(gc[((size_t)p * Ph + ph) * Pw + pw]);
            }
            for (int z = 0; z < nz; ++z) {
                const float4* wv = reinterpret_cast<const float4*>(&s_w[0][rlo[0] + z][0]);
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < PT / 4; ++q) {
                    const float4 w4 = wv[q];
                    s += w4.x * g[4 * q] + w4.y * g[4 * q + 1] + w4.z * g[4 * q + 2] + w4.w * g[4 * q + 3];
                }
                s_T2[((c * RB_T + z) * PT + ph) * PT + pw] = s;
            }
        }
        __syncthreads();
        // ---- pass Y^T: thread = (c, tz, pw): T1[c][tz][ty][pw] = sum_ph Wy[ty][ph] * T2[c][tz][ph][pw]
        for (int item = tid; item < nc * nz * Pw; item += RA_THREADS) {
            const int pw = item % Pw, z = (item / Pw) % nz, c = item / (Pw * nz);
            float g[PT];
#pragma unroll
            for (int p = 0; p < PT; ++p) g[p] = p < Ph ? s_T2[((c * RB_T + z) * PT + p) * PT + pw] : 0.f;
            for (int y = 0; y < ny; ++y) {
                const float4* wv = reinterpret_cast<const float4*>(&s_w[1][rlo[1] + y][0]);
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < PT / 4; ++q) {
                    const float4 w4 = wv[q];
                    s += w4.x * g[4 * q] + w4.y * g[4 * q + 1] + w4.z * g[4 * q + 2] + w4.w * g[4 * q + 3];
                }
                s_T1[((c * RB_T + z) * RB_T + y) * PT + pw] = s;
            }
        }
        __syncthreads();
        // ---- pass X^T: thread owns fixed feature rows (c, tz, ty); accumulate in registers ----------
        const float count = m.count;
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            const int row = tid + i * RA_THREADS;               // row = (c * RB_T + tz) * RB_T + ty
            const int ty = row % RB_T, tz = (row / RB_T) % RB_T, c = row / (RB_T * RB_T);
            if (c < nc && tz >= rlo[0] && tz <= rhi[0] && ty >= rlo[1] && ty <= rhi[1]) {
                const float* src = s_T1 + ((c * RB_T + (tz - rlo[0])) * RB_T + (ty - rlo[1])) * PT;
                float g[PT];
#pragma unroll
                for (int q = 0; q < PT / 4; ++q) {
                    const float4 v = reinterpret_cast<const float4*>(src)[q];
                    g[4 * q] = v.x; g[4 * q + 1] = v.y; g[4 * q + 2] = v.z; g[4 * q + 3] = v.w;
                }
#pragma unroll
                for (int p = 0; p < PT; ++p) if (p >= Pw) g[p] = 0.f;     // columns >= Pw of s_T1 are never written
#pragma unroll
                for (int x = 0; x < RB_T; ++x) {
                    const float4* wv = reinterpret_cast<const float4*>(&s_w[2][x][0]);
                    float s = 0.f;
#pragma unroll
                    for (int q = 0; q < PT / 4; ++q) {
                        const float4 w4 = wv[q];
                        s += w4.x * g[4 * q] + w4.y * g[4 * q + 1] + w4.z * g[4 * q + 2] + w4.w * g[4 * q + 3];
                    }
                    acc[i][x] += s / count;
                }
            }
        }
    }
    // ---- every grad_in element of the tile is written exactly once --------------------------------
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
        const int row = tid + i * RA_THREADS;
        const int ty = row % RB_T, tz = (row / RB_T) % RB_T, c = row / (RB_T * RB_T);
        const int z = tz0 + tz, y = ty0 + ty;
        if (c < nc && z < S && y < H) {
            T* dst = gin + ((((size_t)b * C + c0 + c) * S + z) * H + y) * W + tx0;
#pragma unroll
            for (int x = 0; x < RB_T; ++x) if (tx0 + x < W) dst[x] = from_f<T>(acc[i][x]);
        }
    }
}

}  // namespace b200seg

using namespace b200seg;

extern "C" size_t b200seg_roialign3d_workspace_bytes(int R) {
    return align_up((size_t)(R > 0 ? R : 1) * sizeof(RoiMeta), 256) + 256;
}

template <typename T>
static int launch_fwd(const void* features, const float* rois, void* output, int C, int S, int H, int W, int R,
                      int Ps, int Ph, int Pw, float scale, int sr, int layout, cudaStream_t stream) {
    const int pmax = Ps > Ph ? (Ps > Pw ? Ps : Pw) : (Ph > Pw ? Ph : Pw);
    dim3 grid(R, (C + RA_CC - 1) / RA_CC);
    const size_t smem = RA_FWD_SMEM_FLOATS * sizeof(float);
    if (pmax <= 8) {
        B200_CUDA(cudaFuncSetAttribute(roialign3d_fwd_kernel<T, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        roialign3d_fwd_kernel<T, 8><<<grid, RA_THREADS, smem, stream>>>((const T*)features, rois, (T*)output, C, S, H, W,
                                                                        Ps, Ph, Pw, scale, sr, layout);
    } else {
        B200_CUDA(cudaFuncSetAttribute(roialign3d_fwd_kernel<T, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        roialign3d_fwd_kernel<T, 16><<<grid, RA_THREADS, smem, stream>>>((const T*)features, rois, (T*)output, C, S, H, W,
                                                                         Ps, Ph, Pw, scale, sr, layout);
    }
    B200_LAUNCH_CHECK("roialign3d_fwd_kernel");
    return 0;
}

extern "C" int b200seg_roialign3d_fwd_dev(const void* features, int dtype, const float* rois, void* output,
                                          int B, int C, int S, int H, int W, int R, int Ps, int Ph, int Pw,
                                          float spatial_scale, int sampling_ratio, int layout,
                                          b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(B > 0 && C > 0 && S > 0 && H > 0 && W > 0 && R >= 0, "roialign3d_fwd: bad sizes");
    B200_CHECK_ARG(Ps > 0 && Ph > 0 && Pw > 0 && Ps <= 16 && Ph <= 16 && Pw <= 16,
                   "roialign3d_fwd: pooled size must be in 1..16 (got %d,%d,%d)", Ps, Ph, Pw);
    B200_CHECK_ARG(layout == 0 || layout == 1, "roialign3d_fwd: bad layout");
    B200_CHECK_ARG(dtype == B200SEG_F32 || dtype == B200SEG_BF16, "roialign3d_fwd: bad dtype");
    if (R == 0) return 0;
    B200_CHECK_ARG(features && rois && output, "roialign3d_fwd: null pointer");
    B200_CHECK_ARG((C + RA_CC - 1) / RA_CC <= 65535, "roialign3d_fwd: too many channels");
    if (dtype == B200SEG_F32)
        return launch_fwd<float>(features, rois, output, C, S, H, W, R, Ps, Ph, Pw, spatial_scale, sampling_ratio, layout, stream);
    return launch_fwd<__nv_bfloat16>(features, rois, output, C, S, H, W, R, Ps, Ph, Pw, spatial_scale, sampling_ratio, layout, stream);
}

template <typename T>
static int launch_bwd(const void* grad_out, const RoiMeta* meta, void* grad_in, int B, int C, int S, int H, int W, int R,
                      int Ps, int Ph, int Pw, int layout, double zguard, cudaStream_t stream) {
    const int pmax = Ps > Ph ? (Ps > Pw ? Ps : Pw) : (Ph > Pw ? Ph : Pw);
    const int tiles_x = (W + RB_T - 1) / RB_T, tiles_y = (H + RB_T - 1) / RB_T, tiles_z = (S + RB_T - 1) / RB_T;
    if (pmax <= 8) {
        dim3 grid(tiles_x * tiles_y * tiles_z, (C + BwdCfg<8>::CC - 1) / BwdCfg<8>::CC, B);
        const size_t smem = (size_t)BwdCfg<8>::CC * RB_T * (8 * 8 + RB_T * 8) * sizeof(float);
        B200_CUDA(cudaFuncSetAttribute(roialign3d_bwd_kernel<T, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        roialign3d_bwd_kernel<T, 8><<<grid, RA_THREADS, smem, stream>>>((const T*)grad_out, meta, (T*)grad_in, C, S, H, W, R,
                                                                     Ps, Ph, Pw, layout, zguard, tiles_x, tiles_y);
    } else {
        dim3 grid(tiles_x * tiles_y * tiles_z, (C + BwdCfg<16>::CC - 1) / BwdCfg<16>::CC, B);
        const size_t smem = (size_t)BwdCfg<16>::CC * RB_T * (16 * 16 + RB_T * 16) * sizeof(float);
        B200_CUDA(cudaFuncSetAttribute(roialign3d_bwd_kernel<T, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        roialign3d_bwd_kernel<T, 16><<<grid, RA_THREADS, smem, stream>>>((const T*)grad_out, meta, (T*)grad_in, C, S, H, W, R,
                                                                      Ps, Ph, Pw, layout, zguard, tiles_x, tiles_y);
    }
    B200_LAUNCH_CHECK("roialign3d_bwd_kernel");
    return 0;
}

extern "C" int b200seg_roialign3d_bwd_dev(const void* grad_out, int dtype, const float* rois, void* grad_in,
                                          int B, int C, int S, int H, int W, int R, int Ps, int Ph, int Pw,
                                          float spatial_scale, int sampling_ratio, int layout,
                                          void* workspace, size_t workspace_bytes, b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(B > 0 && C > 0 && S > 0 && H > 0 && W > 0 && R >= 0, "roialign3d_bwd: bad sizes");
    B200_CHECK_ARG(Ps > 0 && Ph > 0 && Pw > 0 && Ps <= 16 && Ph <= 16 && Pw <= 16,
                   "roialign3d_bwd: pooled size must be in 1..16 (got %d,%d,%d)", Ps, Ph, Pw);
    B200_CHECK_ARG(layout == 0 || layout == 1, "roialign3d_bwd: bad layout");
    B200_CHECK_ARG(dtype == B200SEG_F32 || dtype == B200SEG_BF16, "roialign3d_bwd: bad dtype");
    B200_CHECK_ARG(grad_in && (R == 0 || (grad_out && rois && workspace)), "roialign3d_bwd: null pointer");
    B200_CHECK_ARG(B <= 65535, "roialign3d_bwd: batch too large");
    if (workspace_bytes < b200seg_roialign3d_workspace_bytes(R)) {
        set_error("roialign3d_bwd: workspace too small");
        return B200SEG_EWORKSPACE;
    }
    RoiMeta* meta = (RoiMeta*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    // layout 0 keeps the reference's backward z guard (-0.1); layout 1 is the exact adjoint (-1.0)
    const double zguard = layout == 0 ? -0.1 : -1.0;
    if (R > 0) {
        roialign3d_prep_kernel<<<(R + 127) / 128, 128, 0, stream>>>(rois, R, meta, S, H, W, Ps, Ph, Pw, spatial_scale,
                                                                   sampling_ratio, zguard);
        B200_LAUNCH_CHECK("roialign3d_prep_kernel");
    }
    if (dtype == B200SEG_F32)
        return launch_bwd<float>(grad_out, meta, grad_in, B, C, S, H, W, R, Ps, Ph, Pw, layout, zguard, stream);
    return launch_bwd<__nv_bfloat16>(grad_out, meta, grad_in, B, C, S, H, W, R, Ps, Ph, Pw, layout, zguard, stream);
}
