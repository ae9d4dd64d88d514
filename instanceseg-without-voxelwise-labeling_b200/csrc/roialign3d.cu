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
template <int PT> struct FwdCfg { static constexpr int SMEM_FLOATS = PT == 8 ? 10240 : 24576; };   // 40 KB / 96 KB of per-warp buffers

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

// Blackwell packed fp32: two independent round-to-nearest FMAs per instruction (FFMA2).  A 128-bit shared-memory
// load of four weights lands in two aligned register pairs, so the packed operands need no shuffling.
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ float lo2(unsigned long long v) { return __uint_as_float((unsigned int)v); }
__device__ __forceinline__ float hi2(unsigned long long v) { return __uint_as_float((unsigned int)(v >> 32)); }

// acc[0..PT) += w[0..PT) * v  with w a 16-byte aligned shared-memory row, acc kept as PT/2 packed pairs
template <int PT>
__device__ __forceinline__ void axpy_row(unsigned long long (&acc)[PT / 2], const float* __restrict__ wrow, float v) {
    const unsigned long long vv = pack2(v, v);
#pragma unroll
    for (int q = 0; q < PT / 4; ++q) {
        const ulonglong2 w = reinterpret_cast<const ulonglong2*>(wrow)[q];
        acc[2 * q] = fma2(w.x, vv, acc[2 * q]);
        acc[2 * q + 1] = fma2(w.y, vv, acc[2 * q + 1]);
    }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
struct __align__(16) FwdShared {
    float w[3][RA_FMAX][16];     // [axis z,y,x][footprint voxel][bin]
    int row_off[RA_FMAX * RA_FMAX];   // offset of footprint row (z, y) inside one channel volume
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
    // per-warp buffers of the separable path (floats): A = stage [rows][RSx] + T1 [rows][PT] (reused as the output
    // staging area [P3]), B = T2 [Fz][PT][PT]
    const int rows = Fz * Fy;
    const int RSx = Fx | 1;                               // odd row stride: conflict-free column walks
    const int sizeA = max(rows * (RSx + PT), (int)((P3 + 3) & ~(size_t)3));
    const int need = sizeA + Fz * PT * PT;
    const bool separable = Fz <= RA_FMAX && Fy <= RA_FMAX && Fx <= RA_FMAX && need <= FwdCfg<PT>::SMEM_FLOATS;
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
    const int icount = ax[0].g * ax[1].g * ax[2].g;
    const bool pow2 = (icount & (icount - 1)) == 0;       // sr = 2: count = 8 -> scaling by 1/count is exact
    __syncthreads();
    if (pow2) for (int i = tid; i < RA_FMAX * 16; i += RA_THREADS) (&sh.w[0][0][0])[i] *= 1.0f / count;
    // offset of every footprint row (z, y) inside one channel volume
    for (int row = tid; row < rows; row += RA_THREADS) {
        const int z = row / Fy, y = row - z * Fy;
        sh.row_off[row] = ((zlo + z) * H + (ylo + y)) * W + xlo;
    }
    __syncthreads();

    // ---- from here on every warp works on its own channels; no block barrier ------------------------
    const int lane = tid & 31, warp = tid >> 5;
    const int n_active = min(RA_THREADS / 32, FwdCfg<PT>::SMEM_FLOATS / need);
    if (warp >= n_active) return;
    float* bufA = s_buf + (size_t)warp * need;            // stage [rows][RSx], then T1 [rows][PT] behind it; later out [P3]
    float* sT1 = bufA + rows * RSx;
    float* sT2 = bufA + sizeA;                            // [Fz][PT(ph)][PT(pw)]
    const int XS = Fx <= 8 ? 8 : 16;                      // lanes per footprint row while staging
    const int lx = lane & (XS - 1), lr = lane / XS, rpi = 32 / XS;
    constexpr int ZP_ITERS = (PT * PT + 31) / 32;
    int zp_src[ZP_ITERS], zp_dst[ZP_ITERS];               // pass Z items of this lane: (ph, pw) = item / Pw, item % Pw
#pragma unroll
    for (int k = 0; k < ZP_ITERS; ++k) {
        const int item = lane + 32 * k;
        const int ph = item / Pw, pw = item - ph * Pw;
        zp_src[k] = item < Ph * Pw ? ph * PT + pw : -1;
        zp_dst[k] = layout == 0 ? item * Ps : item;       // layout 0: (H,W,S) order, bins of one (ph,pw) contiguous
    }
    const int ps_stride = layout == 0 ? 1 : Ph * Pw;
    const size_t SHW = (size_t)S * H * W;

    for (int c = warp; c < nc; c += n_active) {
        const T* fc = feat_b + (size_t)c * SHW;
        // ---- stage the footprint: XS lanes per row, coalesced within a row ---------------------------
        __syncwarp();
        if (lx < Fx) {
            const T* fcl = fc + lx;
            float* dst = bufA + lr * RSx + lx;
            const int dstep = rpi * RSx;
            int row = lr;
            for (; row + 3 * rpi < rows; row += 4 * rpi, dst += 4 * dstep) {      // 4 independent loads in flight
                const int o0 = sh.row_off[row], o1 = sh.row_off[row + rpi], o2 = sh.row_off[row + 2 * rpi], o3 = sh.row_off[row + 3 * rpi];
                const T v0 = fcl[o0], v1 = fcl[o1], v2 = fcl[o2], v3 = fcl[o3];
                dst[0] = to_f(v0); dst[dstep] = to_f(v1); dst[2 * dstep] = to_f(v2); dst[3 * dstep] = to_f(v3);
            }
            for (; row < rows; row += rpi, dst += dstep) *dst = to_f(fcl[sh.row_off[row]]);
        }
        __syncwarp();
        // ---- pass X: lane = footprint row; PT-vector over pw -------------------------------------------
        for (int row = lane; row < rows; row += 32) {
            unsigned long long acc[PT / 2];
#pragma unroll
            for (int p = 0; p < PT / 2; ++p) acc[p] = 0ull;
            const float* src = bufA + row * RSx;
            for (int x = 0; x < Fx; ++x) axpy_row<PT>(acc, &sh.w[2][x][0], src[x]);
#pragma unroll
            for (int p = 0; p < PT / 2; ++p) { sT1[row * PT + 2 * p] = lo2(acc[p]); sT1[row * PT + 2 * p + 1] = hi2(acc[p]); }   // sT1 is only 4-byte aligned
        }
        __syncwarp();
        // ---- pass Y: lane = (z, pw); PT-vector over ph -------------------------------------------------
        {
            const int pw = lane % PT;
            for (int z = lane / PT; z < Fz; z += 32 / PT) {
                unsigned long long acc[PT / 2];
#pragma unroll
                for (int p = 0; p < PT / 2; ++p) acc[p] = 0ull;
                const float* src = sT1 + z * Fy * PT + pw;
                for (int y = 0; y < Fy; ++y) axpy_row<PT>(acc, &sh.w[1][y][0], src[y * PT]);
                float* dst = sT2 + z * PT * PT + pw;
#pragma unroll
                for (int p = 0; p < PT / 2; ++p) { dst[(2 * p) * PT] = lo2(acc[p]); dst[(2 * p + 1) * PT] = hi2(acc[p]); }
            }
        }
        __syncwarp();
        // ---- pass Z: lane = (ph, pw); PT-vector over ps, staged in bufA in the output order -----------
#pragma unroll
        for (int k = 0; k < ZP_ITERS; ++k) {
            if (zp_src[k] < 0) continue;
            unsigned long long acc2[PT / 2];
#pragma unroll
            for (int p = 0; p < PT / 2; ++p) acc2[p] = 0ull;
            const float* src = sT2 + zp_src[k];
            for (int z = 0; z < Fz; ++z) axpy_row<PT>(acc2, &sh.w[0][z][0], src[z * PT * PT]);
            float acc[PT];
#pragma unroll
            for (int p = 0; p < PT / 2; ++p) { acc[2 * p] = lo2(acc2[p]); acc[2 * p + 1] = hi2(acc2[p]); }
            float* dst = bufA + zp_dst[k];
            if (pow2) {                                    // 1/count already folded into the z table
#pragma unroll
                for (int p = 0; p < PT; ++p) if (p < Ps) dst[p * ps_stride] = acc[p];
            } else {
#pragma unroll
                for (int p = 0; p < PT; ++p) if (p < Ps) dst[p * ps_stride] = acc[p] / count;
            }
        }
        __syncwarp();
        // ---- coalesced copy-out of the P3 bins of this (roi, channel) ----------------------------------
        T* o = out_r + (size_t)c * P3;
        {
            const int n3 = (int)P3;
            int i = lane;
            for (; i + 96 < n3; i += 128) {
                const float v0 = bufA[i], v1 = bufA[i + 32], v2 = bufA[i + 64], v3 = bufA[i + 96];
                o[i] = from_f<T>(v0); o[i + 32] = from_f<T>(v1); o[i + 64] = from_f<T>(v2); o[i + 96] = from_f<T>(v3);
            }
            for (; i < n3; i += 32) o[i] = from_f<T>(bufA[i]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// "Owner computes": a CTA owns an 8x8x8 feature tile of one batch element and 8 channels (one channel per
// warp).  roialign3d_prep_kernel turns every RoI into (a) its footprint box and (b) dense per-axis adjoint
// tables W[axis][voxel coordinate][bin] over the whole feature extent, so a tile's slice of a table is one
// contiguous 8 x PT block.  The CTA compacts the RoIs that hit its tile into an ordered list once; then every
// warp walks that list on its own (no block barrier inside the loop): three small contractions per
// (RoI, channel) restricted to the tile,
//     T2[z][ph][pw] = sum_ps Wz[z][ps] G[ps][ph][pw],   T1[z][y][pw] = sum_ph Wy[y][ph] T2[z][ph][pw],
//     acc[z][y][x] += (sum_pw Wx[x][pw] T1[z][y][pw]) / count,
// with the tile gradient held in registers (a lane owns two (z,y) rows of 8 voxels).  RoIs are visited in index
// order, so the result is deterministic, and every grad_in element is written exactly once (zero fill folded in).
constexpr int RB_T = 8;            // feature tile edge (z,y,x)
constexpr int RB_WARPS = 8;        // channels per CTA
constexpr int RB_LIST = 1024;      // RoIs compacted per round

struct __align__(16) RoiBox {      // one per RoI, built by roialign3d_prep_kernel
    int lo[3], hi[3];              // footprint (inclusive), hi < lo when no valid sample on that axis
    int batch;
    float inv_count;
};

// grid = R, block = 64.  tables [R][S+H+W][PT] are cleared by the launcher beforehand.
template <int PT>
__global__ void __launch_bounds__(64)
roialign3d_prep_kernel(const float* __restrict__ rois, int R, RoiBox* __restrict__ boxes, float* __restrict__ tables,
                       int S, int H, int W, int Ps, int Ph, int Pw, float scale, int sr, double zguard) {
    const int r = blockIdx.x;
    __shared__ int s_lo[3], s_hi[3];
    AxisP ax[3];
    int batch; float count;
    roi_axes(rois + (size_t)r * 7, scale, sr, Ps, Ph, Pw, S, H, W, zguard, ax[0], ax[1], ax[2], batch, count);
    const int Pa[3] = {Ps, Ph, Pw};
    const int seg[3] = {0, S, S + H};
    if (threadIdx.x < 3) { s_lo[threadIdx.x] = 0x7fffffff; s_hi[threadIdx.x] = -1; }
    __syncthreads();
    if (threadIdx.x < 3 * PT) {
        const int a = threadIdx.x / PT, p = threadIdx.x % PT;
        if (p < Pa[a]) {
            float* col = tables + ((size_t)r * (S + H + W) + seg[a]) * PT + p;     // column p of axis a: stride PT
            int lo = 0x7fffffff, hi = -1;
            for (int i = 0; i < ax[a].g; ++i) {                                    // one thread per column: no races
                const Tap t = axis_sample(ax[a], p, i);
                if (!t.valid) continue;
                col[(size_t)t.low * PT] += t.h;
                col[(size_t)t.high * PT] += t.l;
                lo = min(lo, t.low); hi = max(hi, t.high);
            }
            if (hi >= 0) { atomicMin(&s_lo[a], lo); atomicMax(&s_hi[a], hi); }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        RoiBox b;
        for (int a = 0; a < 3; ++a) { b.lo[a] = s_lo[a]; b.hi[a] = s_hi[a]; }
        b.batch = batch; b.inv_count = 1.0f / count;
        boxes[r] = b;
    }
}

// dot product of PT table weights (shared memory, 16-byte aligned row) with PT register values
template <int PT>
__device__ __forceinline__ float dot_row(const float* __restrict__ wrow, const float (&g)[PT]) {
    float s0 = 0.f, s1 = 0.f;                                  // two chains: halves the dependent-FMA latency
#pragma unroll                                                  // (packed FFMA2 was measured slower here: the register
    for (int q = 0; q < PT / 4; ++q) {                          //  pairs of g cost more moves than the FMAs they save)
        const float4 w4 = reinterpret_cast<const float4*>(wrow)[q];
        s0 = fmaf(w4.x, g[4 * q], s0); s1 = fmaf(w4.y, g[4 * q + 1], s1);
        s0 = fmaf(w4.z, g[4 * q + 2], s0); s1 = fmaf(w4.w, g[4 * q + 3], s1);
    }
    return s0 + s1;
}

// PC > 0: cubic pooled size known at compile time (Ps == Ph == Pw == PC): constant strides, no divisions.
template <typename T, int PT, int PC>
__global__ void __launch_bounds__(RB_WARPS * 32)
roialign3d_bwd_kernel(const T* __restrict__ gout, const RoiBox* __restrict__ boxes, const float* __restrict__ tables,
                      T* __restrict__ gin, int C, int S, int H, int W, int R, int Ps_, int Ph_, int Pw_,
                      int tiles_x, int tiles_y) {
    constexpr int WARP_FLOATS = 3 * RB_T * PT + RB_T * PT * PT + RB_T * RB_T * PT;
    constexpr int ZT_ITERS = (PT * PT + 31) / 32;                 // (ph,pw) items per lane in pass Z^T
    constexpr int ZPI = 32 / PT;                                  // z slices per iteration in pass Y^T (lane = z * PT + pw)
    extern __shared__ __align__(16) float s_dynb[];
    __shared__ int s_list[RB_LIST];
    __shared__ int s_wcnt[RB_WARPS];

    const int Ps = PC > 0 ? PC : Ps_, Ph = PC > 0 ? PC : Ph_, Pw = PC > 0 ? PC : Pw_;
    const int PP = Ph * Pw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* sW = s_dynb + (size_t)warp * WARP_FLOATS;              // [3][RB_T][PT]   tile slice of the adjoint tables
    float* sT2 = sW + 3 * RB_T * PT;                              // [RB_T][PT(ph)][PT(pw)]
    float* sT1 = sT2 + RB_T * PT * PT;                            // [RB_T][RB_T][PT(pw)]

    const int tile = blockIdx.x;
    const int t0[3] = {(tile / (tiles_x * tiles_y)) * RB_T, ((tile / tiles_x) % tiles_y) * RB_T, (tile % tiles_x) * RB_T};
    const int dims[3] = {S, H, W};
    const int seg[3] = {0, S, S + H};
    const int c = blockIdx.y * RB_WARPS + warp;
    const bool c_ok = c < C;
    const int b = blockIdx.z;
    const size_t P3 = (size_t)Ps * PP;
    const int t1[3] = {min(t0[0] + RB_T, S) - 1, min(t0[1] + RB_T, H) - 1, min(t0[2] + RB_T, W) - 1};   // last voxel of the tile

    // pass Z^T items of this lane: item = lane + 32 k -> (ph, pw), fixed for the whole kernel
    int zt_off[ZT_ITERS];                                         // offset of (ph, pw) inside a [PT][PT] plane of sT2
#pragma unroll
    for (int k = 0; k < ZT_ITERS; ++k) {
        const int item = lane + 32 * k;
        const int ph = item / Pw, pw = item - ph * Pw;
        zt_off[k] = item < PP ? ph * PT + pw : -1;
    }
    // lane owns rows (z = lane / 8 + 4 i, y = lane % 8), i = 0, 1
    const int my_y = lane & 7, my_z = lane >> 3;
    float acc[2][RB_T];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int x = 0; x < RB_T; ++x) acc[i][x] = 0.f;

    for (int r0 = 0; r0 < R; r0 += RB_LIST) {
        // ---- ordered list of the RoIs of this round that hit the tile ------------------------------
        __syncthreads();                                          // every warp is done with the previous list
        const int rn = min(RB_LIST, R - r0);
        int n_list = 0;
        for (int q0 = 0; q0 < rn; q0 += RB_WARPS * 32) {
            const int q = q0 + tid;
            bool hit = false;
            if (q < rn) {
                const RoiBox bx = boxes[r0 + q];
                hit = bx.batch == b;
#pragma unroll
                for (int a = 0; a < 3; ++a) hit = hit && bx.hi[a] >= bx.lo[a] && bx.hi[a] >= t0[a] && bx.lo[a] <= t1[a];
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_wcnt[warp] = __popc(m);
            __syncthreads();
            int before = n_list;
            for (int w = 0; w < warp; ++w) before += s_wcnt[w];
            if (hit) s_list[before + __popc(m & ((1u << lane) - 1u))] = r0 + q;
            for (int w = 0; w < RB_WARPS; ++w) n_list += s_wcnt[w];
            __syncthreads();
        }
        if (!c_ok) continue;                                      // (still takes part in the barriers above)

        for (int k = 0; k < n_list; ++k) {
            const int r = s_list[k];
            const RoiBox bx = boxes[r];                           // uniform address: one broadcast transaction
            int rlo[3], rhi[3];                                   // tile-local index ranges touched by this RoI
#pragma unroll
            for (int a = 0; a < 3; ++a) { rlo[a] = max(bx.lo[a], t0[a]) - t0[a]; rhi[a] = min(bx.hi[a], t1[a]) - t0[a]; }
            // ---- tile slice of the three tables: [RB_T][PT] contiguous per axis ------------------------
            const float* tab = tables + (size_t)r * (S + H + W) * PT;
            __syncwarp();
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float2* src = reinterpret_cast<const float2*>(tab + (size_t)(seg[a] + t0[a]) * PT);
#pragma unroll
                for (int e = 0; e < RB_T * PT / 64; ++e) {
                    const int f2 = lane + 32 * e;                 // float2 index inside the [RB_T][PT] slice
                    const int t = (2 * f2) / PT;
                    const float2 v = (t0[a] + t < dims[a]) ? __ldg(src + f2) : make_float2(0.f, 0.f);
                    reinterpret_cast<float2*>(sW + a * RB_T * PT)[f2] = v;
                }
            }
            __syncwarp();
            const int nz = rhi[0] - rlo[0] + 1, ny = rhi[1] - rlo[1] + 1;
            const T* gc = gout + ((size_t)r * C + c) * P3 + lane;
            // ---- pass Z^T: lane = (ph, pw) ---------------------------------------------------------------
            // grad_out is read in (S,H,W) order in both layouts: that is what the reference does (.cu:272-275)
            // and it is also the exact adjoint of the layout-1 forward.
#pragma unroll
            for (int it = 0; it < ZT_ITERS; ++it) {
                if (zt_off[it] < 0) continue;
                float g[PT];
#pragma unroll
                for (int p = 0; p < PT; ++p) g[p] = p < Ps ? to_f(gc[p * PP + 32 * it]) : 0.f;
                float* dst = sT2 + zt_off[it];
#pragma unroll 2
                for (int z = 0; z < nz; ++z) dst[z * PT * PT] = dot_row<PT>(sW + (rlo[0] + z) * PT, g);
            }
            __syncwarp();
            // ---- pass Y^T: lane = (z, pw) ----------------------------------------------------------------
            {
                const int pw = lane % PT;
#pragma unroll
                for (int it = 0; it < RB_T / ZPI; ++it) {
                    const int z = lane / PT + it * ZPI;
                    if (z < nz && pw < Pw) {
                        float g[PT];
#pragma unroll
                        for (int p = 0; p < PT; ++p) g[p] = p < Ph ? sT2[(z * PT + p) * PT + pw] : 0.f;
                        float* dst = sT1 + z * RB_T * PT + pw;
#pragma unroll 2
                        for (int y = 0; y < ny; ++y) dst[y * PT] = dot_row<PT>(sW + (RB_T + rlo[1] + y) * PT, g);
                    }
                }
            }
            __syncwarp();
            // ---- pass X^T: lane owns fixed feature rows; accumulate in registers ------------------------
            const float inv_count = bx.inv_count;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int tz = my_z + 4 * i;
                if (tz >= rlo[0] && tz <= rhi[0] && my_y >= rlo[1] && my_y <= rhi[1]) {
                    const float* src = sT1 + ((tz - rlo[0]) * RB_T + (my_y - rlo[1])) * PT;
                    float g[PT];
#pragma unroll
                    for (int q = 0; q < PT / 4; ++q) {
                        const float4 v = reinterpret_cast<const float4*>(src)[q];
                        g[4 * q] = v.x; g[4 * q + 1] = v.y; g[4 * q + 2] = v.z; g[4 * q + 3] = v.w;
                    }
#pragma unroll
                    for (int p = 0; p < PT; ++p) if (p >= Pw) g[p] = 0.f;     // columns >= Pw of sT1 are never written
                    // rows of the Wx slice outside the footprint are zero, so all 8 voxels can be updated blindly
#pragma unroll
                    for (int x = 0; x < RB_T; ++x) acc[i][x] = fmaf(dot_row<PT>(sW + (2 * RB_T + x) * PT, g), inv_count, acc[i][x]);
                }
            }
        }
    }
    // ---- every grad_in element of the tile is written exactly once --------------------------------
    if (c_ok) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int z = t0[0] + my_z + 4 * i, y = t0[1] + my_y;
            if (z < S && y < H) {
                T* dst = gin + ((((size_t)b * C + c) * S + z) * H + y) * W + t0[2];
#pragma unroll
                for (int x = 0; x < RB_T; ++x) if (t0[2] + x < W) dst[x] = from_f<T>(acc[i][x]);
            }
        }
    }
}

}  // namespace b200seg

using namespace b200seg;

static size_t bwd_table_floats(int R, int S, int H, int W, int PT) { return (size_t)R * (size_t)(S + H + W) * PT; }

extern "C" size_t b200seg_roialign3d_workspace_bytes(int R, int S, int H, int W, int P_max) {
    const int PT = P_max <= 8 ? 8 : 16;
    const size_t r = R > 0 ? R : 1;
    if (S <= 0 || H <= 0 || W <= 0) return 256;
    return align_up(r * sizeof(RoiBox), 256) + align_up(bwd_table_floats((int)r, S, H, W, PT) * sizeof(float), 256) + 256;
}

template <typename T>
static int launch_fwd(const void* features, const float* rois, void* output, int C, int S, int H, int W, int R,
                      int Ps, int Ph, int Pw, float scale, int sr, int layout, cudaStream_t stream) {
    const int pmax = Ps > Ph ? (Ps > Pw ? Ps : Pw) : (Ph > Pw ? Ph : Pw);
    dim3 grid(R, (C + RA_CC - 1) / RA_CC);
    if (pmax <= 8) {
        const size_t smem = FwdCfg<8>::SMEM_FLOATS * sizeof(float);
        B200_CUDA(cudaFuncSetAttribute(roialign3d_fwd_kernel<T, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        roialign3d_fwd_kernel<T, 8><<<grid, RA_THREADS, smem, stream>>>((const T*)features, rois, (T*)output, C, S, H, W,
                                                                        Ps, Ph, Pw, scale, sr, layout);
    } else {
        const size_t smem = FwdCfg<16>::SMEM_FLOATS * sizeof(float);
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

template <typename T, int PT>
static int launch_bwd_pt(const void* grad_out, const float* rois, void* grad_in, int B, int C, int S, int H, int W, int R,
                         int Ps, int Ph, int Pw, float scale, int sr, double zguard, RoiBox* boxes, float* tables,
                         cudaStream_t stream) {
    if (R > 0) {
        B200_CUDA(cudaMemsetAsync(tables, 0, bwd_table_floats(R, S, H, W, PT) * sizeof(float), stream));
        roialign3d_prep_kernel<PT><<<R, 64, 0, stream>>>(rois, R, boxes, tables, S, H, W, Ps, Ph, Pw, scale, sr, zguard);
        B200_LAUNCH_CHECK("roialign3d_prep_kernel");
    }
    const int tiles_x = (W + RB_T - 1) / RB_T, tiles_y = (H + RB_T - 1) / RB_T, tiles_z = (S + RB_T - 1) / RB_T;
    dim3 grid(tiles_x * tiles_y * tiles_z, (C + RB_WARPS - 1) / RB_WARPS, B);
    const size_t smem = (size_t)RB_WARPS * (3 * RB_T * PT + RB_T * PT * PT + RB_T * RB_T * PT) * sizeof(float);
    constexpr int PC = PT == 8 ? 7 : 14;                           // the pooled sizes the model uses (box head 7, mask head 14)
    if (Ps == PC && Ph == PC && Pw == PC) {
        B200_CUDA(cudaFuncSetAttribute(roialign3d_bwd_kernel<T, PT, PC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        roialign3d_bwd_kernel<T, PT, PC><<<grid, RB_WARPS * 32, smem, stream>>>((const T*)grad_out, boxes, tables, (T*)grad_in, C, S, H, W, R,
                                                                               Ps, Ph, Pw, tiles_x, tiles_y);
    } else {
        B200_CUDA(cudaFuncSetAttribute(roialign3d_bwd_kernel<T, PT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        roialign3d_bwd_kernel<T, PT, 0><<<grid, RB_WARPS * 32, smem, stream>>>((const T*)grad_out, boxes, tables, (T*)grad_in, C, S, H, W, R,
                                                                              Ps, Ph, Pw, tiles_x, tiles_y);
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
    B200_CHECK_ARG(B <= 65535 && (C + RB_WARPS - 1) / RB_WARPS <= 65535, "roialign3d_bwd: batch / channel count too large");
    const int pmax = Ps > Ph ? (Ps > Pw ? Ps : Pw) : (Ph > Pw ? Ph : Pw);
    if (workspace_bytes < b200seg_roialign3d_workspace_bytes(R, S, H, W, pmax)) {
        set_error("roialign3d_bwd: workspace too small");
        return B200SEG_EWORKSPACE;
    }
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    RoiBox* boxes = (RoiBox*)ws;
    float* tables = (float*)(ws + align_up((size_t)(R > 0 ? R : 1) * sizeof(RoiBox), 256));
    // layout 0 keeps the reference's backward z guard (-0.1); layout 1 is the exact adjoint (-1.0)
    const double zguard = layout == 0 ? -0.1 : -1.0;
    if (pmax <= 8) {
        if (dtype == B200SEG_F32)
            return launch_bwd_pt<float, 8>(grad_out, rois, grad_in, B, C, S, H, W, R, Ps, Ph, Pw, spatial_scale, sampling_ratio, zguard, boxes, tables, stream);
        return launch_bwd_pt<__nv_bfloat16, 8>(grad_out, rois, grad_in, B, C, S, H, W, R, Ps, Ph, Pw, spatial_scale, sampling_ratio, zguard, boxes, tables, stream);
    }
    if (dtype == B200SEG_F32)
        return launch_bwd_pt<float, 16>(grad_out, rois, grad_in, B, C, S, H, W, R, Ps, Ph, Pw, spatial_scale, sampling_ratio, zguard, boxes, tables, stream);
    return launch_bwd_pt<__nv_bfloat16, 16>(grad_out, rois, grad_in, B, C, S, H, W, R, Ps, Ph, Pw, spatial_scale, sampling_ratio, zguard, boxes, tables, stream);
}
