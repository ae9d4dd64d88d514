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
    // -1.0 is a float: (double)c < -1.0 <=> c < -1.0f, which keeps the forward's samples off the fp64 pipe
    t.valid = a.guard == -1.0 ? !(c < -1.0f || c > (float)a.dim) : !((double)c < a.guard || c > (float)a.dim);
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

// Table building shared by both forward kernels (block level, RA_THREADS threads).  On return sh.lo / sh.hi hold the
// footprint, and -- unless the footprint is empty or `tables` is false -- sh.w the dense axis tables with 1/count
// folded into the z table when count is a power of two.  Returns true when every sample falls outside the volume.
__device__ __forceinline__ bool fwd_ranges(FwdShared& sh, const AxisP (&ax)[3], const int (&Pa)[3]) {
    const int tid = threadIdx.x;
    if (tid < 3) { sh.lo[tid] = 0x7fffffff; sh.hi[tid] = -1; }
    for (int i = tid; i < 3 * RA_FMAX * 16; i += (int)blockDim.x) (&sh.w[0][0][0])[i] = 0.f;
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
    return sh.hi[0] < 0 || sh.hi[1] < 0 || sh.hi[2] < 0;
}

// footprints of at most RA_FMAX voxels per axis only.  Ends with a block barrier.
__device__ __forceinline__ void fwd_tables(FwdShared& sh, const AxisP (&ax)[3], const int (&Pa)[3], float count, bool pow2) {
    const int tid = threadIdx.x;
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
    if (pow2) for (int i = tid; i < RA_FMAX * 16; i += (int)blockDim.x) (&sh.w[0][0][0])[i] *= 1.0f / count;
    __syncthreads();
}

// direct evaluation (reference arithmetic) of nc channels of one RoI, coalesced over the output span
template <typename T>
__device__ __forceinline__ void fwd_direct(const T* __restrict__ feat_b, T* __restrict__ out_r, const AxisP (&ax)[3], int nc,
                                           int S, int H, int W, int Ps, int Ph, int Pw, int layout, float count) {
    const size_t P3 = (size_t)Ps * Ph * Pw;
    for (size_t idx = threadIdx.x; idx < (size_t)nc * P3; idx += blockDim.x) {
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
}

// Generic separable path (footprints up to RA_FMAX voxels per axis, pooled sizes up to 16): nc (<= RA_CC) channels of one
// RoI, every warp on its own channels, no block barrier inside.  Needs the tables of fwd_tables and, behind them,
// sh.row_off (filled by the caller + one block barrier).
template <typename T, int PT>
__device__ __forceinline__ void fwd_generic_warps(float* s_buf, const FwdShared& sh, const T* __restrict__ feat_b,
                                                  T* __restrict__ out_r, int nc, int S, int H, int W, int Ps, int Ph, int Pw,
                                                  int layout, int Fz, int Fy, int Fx, int need, int sizeA, float count, bool pow2) {
    const int tid = threadIdx.x;
    const size_t P3 = (size_t)Ps * Ph * Pw;
    const int rows = Fz * Fy;
    const int RSx = Fx | 1;                               // odd row stride: conflict-free column walks
    const int lane = tid & 31, warp = tid >> 5;
    const int n_active = min((int)blockDim.x / 32, FwdCfg<PT>::SMEM_FLOATS / need);
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

// Generic forward kernel: pooled sizes 9..16 (the mask head's 14^3) -- and the body the fast kernel below falls back to.
// cpb = channels per CTA (a multiple of RA_CC).
template <typename T, int PT>
__device__ __forceinline__ void fwd_generic_cta(float* s_buf, FwdShared& sh, const T* __restrict__ feat_b, T* __restrict__ out_r,
                                                const AxisP (&ax)[3], const int (&Pa)[3], int ncta, int S, int H, int W,
                                                int layout, float count) {
    const int tid = threadIdx.x;
    const int Ps = Pa[0], Ph = Pa[1], Pw = Pa[2];
    const size_t P3 = (size_t)Ps * Ph * Pw;
    const int zlo = sh.lo[0], ylo = sh.lo[1], xlo = sh.lo[2];
    const int Fz = sh.hi[0] - zlo + 1, Fy = sh.hi[1] - ylo + 1, Fx = sh.hi[2] - xlo + 1;
    // per-warp buffers of the separable path (floats): A = stage [rows][RSx] + T1 [rows][PT] (reused as the output
    // staging area [P3]), B = T2 [Fz][PT][PT]
    const int rows = Fz * Fy;
    const int RSx = Fx | 1;
    const int sizeA = max(rows * (RSx + PT), (int)((P3 + 3) & ~(size_t)3));
    const int need = sizeA + Fz * PT * PT;
    const bool separable = Fz <= RA_FMAX && Fy <= RA_FMAX && Fx <= RA_FMAX && need <= FwdCfg<PT>::SMEM_FLOATS;
    if (!separable) {
        fwd_direct<T>(feat_b, out_r, ax, ncta, S, H, W, Ps, Ph, Pw, layout, count);
        return;
    }
    const int icount = ax[0].g * ax[1].g * ax[2].g;
    const bool pow2 = (icount & (icount - 1)) == 0;       // sr = 2: count = 8 -> scaling by 1/count is exact
    fwd_tables(sh, ax, Pa, count, pow2);
    for (int row = tid; row < rows; row += (int)blockDim.x) {   // offset of every footprint row (z, y) inside one channel volume
        const int z = row / Fy, y = row - z * Fy;
        sh.row_off[row] = ((zlo + z) * H + (ylo + y)) * W + xlo;
    }
    __syncthreads();
    const size_t SHW = (size_t)S * H * W;
    for (int c0 = 0; c0 < ncta; c0 += RA_CC)
        fwd_generic_warps<T, PT>(s_buf, sh, feat_b + (size_t)c0 * SHW, out_r + (size_t)c0 * P3, min(RA_CC, ncta - c0), S, H, W,
                                 Ps, Ph, Pw, layout, Fz, Fy, Fx, need, sizeA, count, pow2);
}

template <typename T, int PT>
__global__ void __launch_bounds__(RA_THREADS)
roialign3d_fwd_kernel(const T* __restrict__ feat, const float* __restrict__ rois, T* __restrict__ out,
                      int C, int S, int H, int W, int Ps, int Ph, int Pw, float scale, int sr, int layout) {
    extern __shared__ __align__(16) float s_buf[];
    __shared__ FwdShared sh;
    const int r = blockIdx.x;
    const int c0 = blockIdx.y * RA_CC;
    const int nc = min(RA_CC, C - c0);
    AxisP ax[3];
    int batch; float count;
    roi_axes(rois + (size_t)r * 7, scale, sr, Ps, Ph, Pw, S, H, W, -1.0, ax[0], ax[1], ax[2], batch, count);
    const int Pa[3] = {Ps, Ph, Pw};
    const bool empty = fwd_ranges(sh, ax, Pa);
    const size_t P3 = (size_t)Ps * Ph * Pw;
    T* out_r = out + ((size_t)r * C + c0) * P3;
    if (empty) {                                        // every sample falls outside: zeros
        for (size_t i = threadIdx.x; i < (size_t)nc * P3; i += RA_THREADS) out_r[i] = from_f<T>(0.f);
        return;
    }
    const T* feat_b = feat + ((size_t)batch * C + c0) * S * H * W;
    fwd_generic_cta<T, PT>(s_buf, sh, feat_b, out_r, ax, Pa, nc, S, H, W, layout, count);
}

// ------------------------------------------------------------------------------------------------
// forward, fast path: pooled sizes <= 8 (the box head's 7^3) and footprints of <= 8 voxels per axis
// ------------------------------------------------------------------------------------------------
// The generic kernel above spends most of its issue slots on shared-memory traffic (one weight-row load per feature
// value) and on half-empty warps (25..49 items over 32 lanes).  Here the lanes of a warp are CHANNELS, so the axis
// weights are warp-uniform scalars that live in REGISTERS and every lane runs the same dense little contraction with
// all indices static (measured on B200: a shared-memory load costs the same whether its 32 lanes read 32 or 256 distinct
// bytes, and a 128-bit load costs 4.3x a 32-bit one -- profiles/exp/lds_patterns.cu -- so broadcast weight loads are as
// expensive as data loads and have to go):
//   * a warp task = (RoI, 16 channels).  lane = (channel pair cp = lane & 7, pw pair g = lane >> 3); a staged voxel holds
//     its 16 channels contiguously, so one 64-bit shared load yields the packed operand (ch 2cp, ch 2cp+1) of an FFMA2 whose
//     other operand is a scalar weight (FFMA2's .F32 broadcast form): two FMAs per issue slot, no register shuffling.
//   * phase A walks the footprint plane by plane along the axis that is contracted LAST (y in the reference's (H,W,S) bin
//     order, z in layout 1).  Planes arrive through cp.async into a double buffer; per
//     footprint row the x contraction feeds a dense scatter into the bins of the second axis (U[p2][pw pair], 16 packed
//     accumulators); the finished plane goes to shared memory as columns [channel][pw, p2].
//   * phase B: lane = column (4 per lane, packed in pairs); the last contraction runs down the planes and every store
//     instruction writes 32 consecutive output elements -- in both bin orders the column index IS the fast part of the
//     output index, which is why the pass order depends on the layout.
// A CTA owns one RoI (tables built once by warp 0: lane = (axis, bin), ranges by shuffles, every lane accumulates its own
// table column in registers -- one block barrier) and up to `cpb` = 96 channels = six 16-channel tasks.  A task belongs to a
// TEAM of two warps: they take alternate planes in phase A (each with its own staging double buffer) and alternate blocks of
// 128 columns in phase B, with named barriers in between -- twice the issue streams on the same column buffer.  This was the
// one change that moved the kernel (0.121 -> 0.108 ms): it is bound by the latency of a task's dependent phases and by how
// many tasks fit the shared memory of an SM, not by issue slots or bandwidth.  As many teams as fit the pool
// work at a time, the other warps retire.  RoIs that do not qualify take the generic body above inside the same launch.
constexpr int RF_THREADS = 192;         // 6 warps, 2 CTAs per SM: 168 registers per thread
constexpr int RF_TEAM = 2;              // warps per task (3 measured: same time)
constexpr int RF_POOL_FLOATS = 25600;   // 100 KB: two CTAs per SM
constexpr int RF_G = 16;                // channels per warp task
constexpr int RF_CHS = 20;              // floats per staged voxel (16 channels + pad: staging stores are conflict-free)
constexpr int RF_F = 8;                 // max footprint voxels per axis

struct __align__(16) FastShared {
    float w2[RF_F][8];                  // second axis: [footprint voxel][bin]
    float w3[RF_F][8];                  // last axis
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// acc += (w, w) * v   (scalar weight, packed pair of channels / columns)
__device__ __forceinline__ unsigned long long fma2s(float w, unsigned long long v, unsigned long long acc) {
    return fma2(pack2(w, w), v, acc);
}

// rows of one staged plane: x contraction + dense scatter into the bins of the second axis (processing rows in pairs
// inside one basic block -- four x chains instead of two -- was measured: no gain, the kernel is not FFMA2-latency bound)
template <int FX>
__device__ __forceinline__ void fast_plane_rows(const float* __restrict__ stage, int F2, const float (&wx)[RF_F][2],
                                                const float (&w2)[RF_F][8], unsigned long long (&U)[8][2]) {
#pragma unroll
    for (int f2 = 0; f2 < RF_F; ++f2) {
        if (f2 < F2) {                                                     // warp-uniform
            const float* rowp = stage + f2 * FX * RF_CHS;
            unsigned long long t0 = 0ull, t1 = 0ull;
#pragma unroll
            for (int x = 0; x < FX; ++x) {
                const unsigned long long v = *reinterpret_cast<const unsigned long long*>(rowp + x * RF_CHS);
                t0 = fma2s(wx[x][0], v, t0);
                t1 = fma2s(wx[x][1], v, t1);
            }
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                U[p][0] = fma2s(w2[f2][p], t0, U[p][0]);
                U[p][1] = fma2s(w2[f2][p], t1, U[p][1]);
            }
        }
    }
}

// phase B for one warp task: columns [0, NQ) of the column buffer, F3N planes deep
template <typename T, int F3N>
__device__ __forceinline__ void fast_columns(const float* __restrict__ ubuf, const FastShared& fs, T* __restrict__ og,
                                             int NQ, int NC, int UST, int P3, int P3n, int lane, int half) {
    float w3[F3N][8];
#pragma unroll
    for (int f = 0; f < F3N; ++f) {
        const float4 a = *reinterpret_cast<const float4*>(&fs.w3[f][0]), b = *reinterpret_cast<const float4*>(&fs.w3[f][4]);
        w3[f][0] = a.x; w3[f][1] = a.y; w3[f][2] = a.z; w3[f][3] = a.w;
        w3[f][4] = b.x; w3[f][5] = b.y; w3[f][6] = b.z; w3[f][7] = b.w;
    }
    // the two warps of a pair take alternate blocks of 128 columns
    for (int q0 = half * 128; q0 < NQ; q0 += RF_TEAM * 128) {   // 4 columns per lane
        unsigned long long acc[8][2];
#pragma unroll
        for (int p = 0; p < 8; ++p) { acc[p][0] = 0ull; acc[p][1] = 0ull; }
        // columns beyond the buffer are clamped (their results are never stored)
        const int o0 = min(q0 + lane, UST - 1), o1 = min(q0 + 32 + lane, UST - 1);
        const int o2 = min(q0 + 64 + lane, UST - 1), o3 = min(q0 + 96 + lane, UST - 1);
#pragma unroll
        for (int f3 = 0; f3 < F3N; ++f3) {
            const float* up = ubuf + f3 * UST;
            const unsigned long long v01 = pack2(up[o0], up[o1]);
            const unsigned long long v23 = pack2(up[o2], up[o3]);
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                acc[p][0] = fma2s(w3[f3][p], v01, acc[p][0]);
                acc[p][1] = fma2s(w3[f3][p], v23, acc[p][1]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int q = q0 + 32 * i + lane;
            if (q < NQ) {
                const int ch = q / NC;                     // NC is a compile-time constant for the model's pooled sizes
                T* o = og + (ch * P3 + (q - ch * NC));
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    if (p < P3n) {
                        const float v = (i & 1) ? hi2(acc[p][i >> 1]) : lo2(acc[p][i >> 1]);
                        o[p * NC] = from_f<T>(v);
                    }
                }
            }
        }
    }
}

// PC > 0: cubic pooled size known at compile time (Ps == Ph == Pw == PC): all column strides and loop bounds are constants.
// L0: the reference's (H,W,S) bin order (layout 0); otherwise (S,H,W) (layout 1).
template <typename T, int PC, bool L0>
__global__ void __launch_bounds__(RF_THREADS, 2)
roialign3d_fwd_fast_kernel(const T* __restrict__ feat, const float* __restrict__ rois, T* __restrict__ out,
                           int C, int S, int H, int W, int Ps_, int Ph_, int Pw_, float scale, int sr, int cpb) {
    extern __shared__ __align__(16) float s_buf[];
    __shared__ FwdShared sh;
    __shared__ FastShared fs;
    const int Ps = PC > 0 ? PC : Ps_, Ph = PC > 0 ? PC : Ph_, Pw = PC > 0 ? PC : Pw_;
    const int r = blockIdx.x;
    const int c_begin = blockIdx.y * cpb;
    const int ncta = min(cpb, C - c_begin);
    const int tid = threadIdx.x;
    constexpr int layout = L0 ? 0 : 1;

    AxisP ax[3];
    int batch; float count;
    roi_axes(rois + (size_t)r * 7, scale, sr, Ps, Ph, Pw, S, H, W, -1.0, ax[0], ax[1], ax[2], batch, count);
    const int Pa[3] = {Ps, Ph, Pw};
    const int P3 = Ps * Ph * Pw;
    T* out_r = out + ((size_t)r * C + c_begin) * P3;
    const size_t SHW = (size_t)S * H * W;
    const T* feat_b = feat + ((size_t)batch * C + c_begin) * SHW;
    constexpr int a2 = L0 ? 0 : 1, a3 = L0 ? 1 : 0;        // axis roles: x first, a2 scattered in registers, a3 last
    const int lane = tid & 31, warp = tid >> 5;

    // ---- footprint and axis tables: warp 0 alone, lane = (axis, bin); ranges by shuffles, every lane accumulates its own
    //      table column in registers (same order per element as fwd_tables) and writes it once: ONE block barrier instead of
    //      the five of fwd_ranges + fwd_tables, no zero fill, no shared-memory atomics ------------------------------------
    if (warp == 0) {
        const int ta = min(lane >> 3, 2), tp = lane & 7;
        const AxisP mine = ta == 0 ? ax[0] : (ta == 1 ? ax[1] : ax[2]);
        const int myP = ta == 0 ? Ps : (ta == 1 ? Ph : Pw);
        const bool t_on = lane < 24 && tp < myP;
        int lo = 0x7fffffff, hi = -1;
        if (t_on) {
            for (int i = 0; i < mine.g; ++i) {
                const Tap t = axis_sample(mine, tp, i);
                if (t.valid) { lo = min(lo, t.low); hi = max(hi, t.high); }
            }
        }
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
        }
        float wcol[RF_F];                                  // column tp of this lane's axis table: W[f][tp], f = footprint voxel
#pragma unroll
        for (int f = 0; f < RF_F; ++f) wcol[f] = 0.f;
        if (t_on) {
            for (int i = 0; i < mine.g; ++i) {
                const Tap t = axis_sample(mine, tp, i);
                if (t.valid) {
                    const int il = t.low - lo, ih = t.high - lo;
#pragma unroll
                    for (int f = 0; f < RF_F; ++f) { if (il == f) wcol[f] += t.h; if (ih == f) wcol[f] += t.l; }
                }
            }
        }
        const float zs = ta == 0 ? 1.0f / count : 1.0f;    // only used when count is a power of two (fast path): exact
        if (lane < 24) {
            float* dst = ta == 2 ? &sh.w[2][0][tp] : (ta == a2 ? &fs.w2[0][tp] : &fs.w3[0][tp]);
            const int fstride = ta == 2 ? 16 : 8;
#pragma unroll
            for (int f = 0; f < RF_F; ++f) dst[f * fstride] = wcol[f] * zs;
            if (tp == 0) { sh.lo[ta] = lo; sh.hi[ta] = hi; }
        }
    }
    __syncthreads();
    const int zlo = sh.lo[0], ylo = sh.lo[1], xlo = sh.lo[2];
    if (sh.hi[0] < 0 || sh.hi[1] < 0 || sh.hi[2] < 0) {   // every sample falls outside: zeros
        for (size_t i = tid; i < (size_t)ncta * P3; i += RF_THREADS) out_r[i] = from_f<T>(0.f);
        return;
    }
    const int Fz = sh.hi[0] - zlo + 1, Fy = sh.hi[1] - ylo + 1, Fx = sh.hi[2] - xlo + 1;
    const int F2 = L0 ? Fz : Fy, F3 = L0 ? Fy : Fz;
    const int P2 = Pa[a2], P3n = Pa[a3];
    const int NC = Pw * P2;                                // columns per channel
    const int UST = RF_G * NC;                             // floats per plane of the column buffer
    const int stage_f = F2 * Fx * RF_CHS;
    const int need = 2 * RF_TEAM * stage_f + F3 * UST;     // a double-buffered staging area per warp of the team
    const int icount = ax[0].g * ax[1].g * ax[2].g;
    const bool pow2 = (icount & (icount - 1)) == 0;        // sr = 2: count = 8 -> 1/count folds into the z table exactly
    const bool fast = pow2 && Fz <= RF_F && Fy <= RF_F && Fx <= RF_F && need <= RF_POOL_FLOATS;
    if (!fast) {                                           // CTA-uniform: the generic body rebuilds its own (larger) tables
        __syncthreads();
        fwd_ranges(sh, ax, Pa);
        fwd_generic_cta<T, 8>(s_buf, sh, feat_b, out_r, ax, Pa, ncta, S, H, W, layout, count);
        return;
    }

    // a task (RoI, 16 channels) belongs to a PAIR of warps: alternate planes in phase A, alternate column blocks in phase B,
    // named barriers in between -- twice the issue streams on the same shared memory
    const int pair = warp / RF_TEAM, half = warp - pair * RF_TEAM;
    const int n_active = min(RF_THREADS / (32 * RF_TEAM), RF_POOL_FLOATS / need);
    if (pair >= n_active) return;
    float* stage = s_buf + (size_t)pair * need + half * 2 * stage_f;
    float* ubuf = s_buf + (size_t)pair * need + 2 * RF_TEAM * stage_f;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, %1;" :: "r"(1 + pair), "n"(32 * RF_TEAM) : "memory"); };
    const int cp = lane & 7, g = lane >> 3;
    float wx[RF_F][2];
#pragma unroll
    for (int x = 0; x < RF_F; ++x) { wx[x][0] = sh.w[2][x][2 * g]; wx[x][1] = sh.w[2][x][2 * g + 1]; }
    float w2[RF_F][8];
#pragma unroll
    for (int f = 0; f < RF_F; ++f) {
        const float4 a = *reinterpret_cast<const float4*>(&fs.w2[f][0]), b = *reinterpret_cast<const float4*>(&fs.w2[f][4]);
        w2[f][0] = a.x; w2[f][1] = a.y; w2[f][2] = a.z; w2[f][3] = a.w;
        w2[f][4] = b.x; w2[f][5] = b.y; w2[f][6] = b.z; w2[f][7] = b.w;
    }
    // staging geometry: lane = (x = lane & 7, channel slot li = lane >> 3): channels li, li+4, li+8, li+12 of every row
    const int lx = lane & 7, li = lane >> 3;
    const int row_stride = L0 ? H * W : W;                 // global step between rows of a plane (axis 2)
    const int plane_stride = L0 ? W : H * W;               // global step between planes (axis 3)
    const int org = (zlo * H + ylo) * W + xlo + lx;
    const unsigned stage_s = (unsigned)__cvta_generic_to_shared(stage) + (unsigned)(lx * RF_CHS + li) * 4u;
    const unsigned row_bytes = (unsigned)(Fx * RF_CHS) * 4u;
    const int groups = (ncta + RF_G - 1) / RF_G;

    for (int grp = pair; grp < groups; grp += n_active) {
        const int ncg = min(RF_G, ncta - grp * RF_G);
        const T* fg = feat_b + ((size_t)grp * RF_G + li) * SHW + org;
        // planes arrive through cp.async (4-byte copies: rows start at arbitrary x) into a double buffer
        auto issue_plane = [&](int k, int it_) {
            if (lx < Fx) {
                const T* s = fg + (size_t)k * plane_stride;
                unsigned d = stage_s + (unsigned)((it_ & 1) * stage_f) * 4u;
                if (ncg == RF_G) {
#pragma unroll 2
                    for (int f2 = 0; f2 < F2; ++f2, s += row_stride, d += row_bytes) {
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            if constexpr (sizeof(T) == 4) {
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d + 16u * c4), "l"(s + (size_t)(4 * c4) * SHW) : "memory");
                            } else {
                                const float v = to_f(s[(size_t)(4 * c4) * SHW]);
                                asm volatile("st.shared.f32 [%0], %1;" :: "r"(d + 16u * c4), "f"(v) : "memory");
                            }
                        }
                    }
                } else {
                    for (int f2 = 0; f2 < F2; ++f2, s += row_stride, d += row_bytes) {
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            if (li + 4 * c4 < ncg) {
                                const float v = to_f(s[(size_t)(4 * c4) * SHW]);
                                asm volatile("st.shared.f32 [%0], %1;" :: "r"(d + 16u * c4), "f"(v) : "memory");
                            }
                        }
                    }
                }
            }
            cp_async_commit();
        };
        __syncwarp();
        if (half < F3) issue_plane(half, 0);
        for (int k = half, it = 0; k < F3; k += RF_TEAM, ++it) {
            if (k + RF_TEAM < F3) { issue_plane(k + RF_TEAM, it + 1); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            __syncwarp();
            unsigned long long U[8][2];
#pragma unroll
            for (int p = 0; p < 8; ++p) { U[p][0] = 0ull; U[p][1] = 0ull; }
            const float* st = stage + (it & 1) * stage_f + 2 * cp;
            switch (Fx) {
                case 1: fast_plane_rows<1>(st, F2, wx, w2, U); break;
                case 2: fast_plane_rows<2>(st, F2, wx, w2, U); break;
                case 3: fast_plane_rows<3>(st, F2, wx, w2, U); break;
                case 4: fast_plane_rows<4>(st, F2, wx, w2, U); break;
                case 5: fast_plane_rows<5>(st, F2, wx, w2, U); break;
                case 6: fast_plane_rows<6>(st, F2, wx, w2, U); break;
                case 7: fast_plane_rows<7>(st, F2, wx, w2, U); break;
                default: fast_plane_rows<8>(st, F2, wx, w2, U); break;
            }
            // finished plane -> columns [channel][L0 ? pw * P2 + p2 : p2 * Pw + pw]
            float* up = ubuf + k * UST + (2 * cp) * NC;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int pw = 2 * g + j;
                if (pw < Pw) {
                    float* upc = up + (L0 ? pw * P2 : pw);
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        if (p < P2) {
                            const int col = L0 ? p : p * Pw;
                            upc[col] = lo2(U[p][j]);
                            upc[NC + col] = hi2(U[p][j]);
                        }
                    }
                }
            }
            __syncwarp();
        }
        pair_sync();                                       // all planes of the task are in the column buffer
        // ---- phase B: last contraction down the planes, lane = 4 columns, coalesced stores ---------------------
        T* og = out_r + (size_t)grp * RF_G * P3;
        const int NQ = ncg * NC;
        switch (F3) {
            case 1: fast_columns<T, 1>(ubuf, fs, og, NQ, NC, UST, P3, P3n, lane, half); break;
            case 2: fast_columns<T, 2>(ubuf, fs, og, NQ, NC, UST, P3, P3n, lane, half); break;
            case 3: fast_columns<T, 3>(ubuf, fs, og, NQ, NC, UST, P3, P3n, lane, half); break;
            case 4: fast_columns<T, 4>(ubuf, fs, og, NQ, NC, UST, P3, P3n, lane, half); break;
            case 5: fast_columns<T, 5>(ubuf, fs, og, NQ, NC, UST, P3, P3n, lane, half); break;
            case 6: fast_columns<T, 6>(ubuf, fs, og, NQ, NC, UST, P3, P3n, lane, half); break;
            case 7: fast_columns<T, 7>(ubuf, fs, og, NQ, NC, UST, P3, P3n, lane, half); break;
            default: fast_columns<T, 8>(ubuf, fs, og, NQ, NC, UST, P3, P3n, lane, half); break;
        }
        pair_sync();                                       // the column buffer is free again
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

// ------------------------------------------------------------------------------------------------
// backward, fast path (round 2): pooled sizes <= 8, footprints <= 8 voxels per axis, power-of-two sample counts
// ------------------------------------------------------------------------------------------------
// Two launches, still deterministic and atomics-free:
//   1. roialign3d_bwd_roi_kernel -- the forward's fast kernel run backwards, per RoI: lanes are channels, a task = (RoI, 16
//      channels) worked by a team of two warps.  Phase B^T: lane = 4 columns (ph, pw) of grad_out (coalesced loads, the
//      reference reads grad_out in (S,H,W) order in both layouts), contraction with the z table into Fz planes of the column
//      buffer.  Phase A^T per plane: lane = (channel pair, pw pair) reads its columns, per footprint row the y table gives
//      t[pw], the x table the lane's partial of every footprint voxel (its two pw bins only); the partials of the four pw
//      pairs go to shared memory and are summed while the plane is written out -- the FOOTPRINT GRADIENT of the RoI,
//      dF[r][channel group][z][y][x][16 channels], 64-byte runs, into the caller's workspace (a fixed slot of 8^3 voxels
//      per (RoI, channel group)).
//   2. roialign3d_bwd_gather_kernel -- owner computes: a CTA owns an 8^3 feature tile x 16 channels, lists the RoIs that hit
//      the tile in index order and adds their footprint gradients voxel by voxel into a shared-memory tile (one RoI after
//      the other: fixed summation order), then writes every grad_in element exactly once, x fastest.  RoIs that do not
//      qualify for the fast path are evaluated directly from the per-axis adjoint tables at this point (rare).
// Against the round-1 kernel (every tile re-ran the three contractions of every RoI it touches, 2.5 visits per RoI, one
// channel per warp) the contraction work is done once per RoI with full lanes; the price is the footprint-gradient round
// trip through HBM (about 0.6 x the size of grad_out on the BASELINE shapes).
constexpr int RG_SLOT = RF_F * RF_F * RF_F * RF_G;     // floats per (RoI, channel group) slot of the footprint gradient

struct __align__(16) BwdRoi {      // one per RoI, written by the first channel slab's CTA of roialign3d_bwd_roi_kernel
    int lo[3], hi[3];              // footprint (inclusive), hi < lo when no valid sample on that axis
    int batch;
    int fast;                      // 1: footprint gradient in the workspace; 0: direct evaluation from the tables
};

__device__ __forceinline__ unsigned long long mul2s(float w, unsigned long long v) {
    unsigned long long d;
    const unsigned long long ww = pack2(w, w);
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(ww), "l"(v));
    return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long shfl_xor2(unsigned long long v, int m) {
    const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m), hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
    return (unsigned long long)lo | ((unsigned long long)hi << 32);
}

// Column buffer of the backward: plane stride UB (a multiple of 128 columns, so that phase B^T stores whole blocks without
// guards); inside a plane channel PAIRS are interleaved -- element (channel ch, column col) sits at
// ((ch >> 1) * NC + col) * 2 + (ch & 1) -- so that phase A^T fetches its packed operand with one 64-bit load.

// phase B^T for one warp of a team: columns [0, NQ) in blocks of 128, alternate blocks per warp.  PCC > 0: Ps known.
template <typename T, int F3N, int PCC>
__device__ __forceinline__ void bwd_fast_columns(float* __restrict__ ubuf, const FastShared& fs, const T* __restrict__ gg,
                                                 int NQ, int NC, int UB, int P3, int P3n, int lane, int half) {
    float w3[F3N][8];
#pragma unroll
    for (int f = 0; f < F3N; ++f) {
        const float4 a = *reinterpret_cast<const float4*>(&fs.w3[f][0]), b = *reinterpret_cast<const float4*>(&fs.w3[f][4]);
        w3[f][0] = a.x; w3[f][1] = a.y; w3[f][2] = a.z; w3[f][3] = a.w;
        w3[f][4] = b.x; w3[f][5] = b.y; w3[f][6] = b.z; w3[f][7] = b.w;
    }
    constexpr int NP = PCC > 0 ? PCC : 8;
    for (int q0 = half * 128; q0 < NQ; q0 += RF_TEAM * 128) {
        unsigned long long g01[NP], g23[NP];
        int dsto[4];
        {
            float gv[NP][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int q = min(q0 + 32 * i + lane, NQ - 1);     // clamped: columns past the end repeat the last one (never read back)
                const int ch = q / NC, col = q - ch * NC;
                const T* o = gg + (ch * P3 + col);
                dsto[i] = ((ch >> 1) * NC + col) * 2 + (ch & 1);
#pragma unroll
                for (int p = 0; p < NP; ++p) gv[p][i] = (PCC > 0 || p < P3n) ? to_f(o[p * NC]) : 0.f;
            }
#pragma unroll
            for (int p = 0; p < NP; ++p) { g01[p] = pack2(gv[p][0], gv[p][1]); g23[p] = pack2(gv[p][2], gv[p][3]); }
        }
        const bool last = q0 + 128 > NQ;                           // warp-uniform: only the last block has columns to skip
#pragma unroll
        for (int f3 = 0; f3 < F3N; ++f3) {
            unsigned long long a01 = mul2s(w3[f3][0], g01[0]), a23 = mul2s(w3[f3][0], g23[0]);
#pragma unroll
            for (int p = 1; p < NP; ++p) {
                a01 = fma2s(w3[f3][p], g01[p], a01);
                a23 = fma2s(w3[f3][p], g23[p], a23);
            }
            float* up = ubuf + f3 * UB;
            if (!last) {
                up[dsto[0]] = lo2(a01); up[dsto[1]] = hi2(a01); up[dsto[2]] = lo2(a23); up[dsto[3]] = hi2(a23);
            } else {
                if (q0 + lane < NQ) up[dsto[0]] = lo2(a01);
                if (q0 + 32 + lane < NQ) up[dsto[1]] = hi2(a01);
                if (q0 + 64 + lane < NQ) up[dsto[2]] = lo2(a23);
                if (q0 + 96 + lane < NQ) up[dsto[3]] = hi2(a23);
            }
        }
    }
}

// phase A^T rows of one plane: t[pw] from the columns, then this lane's partial (its two pw bins) of every voxel of the row;
// the partials of the four pw pairs (lanes 8 apart) are summed by a butterfly reduce-scatter over groups of four voxels, so
// that lane (cp, g) ends up with voxel 4 m + g of the row for channels (2cp, 2cp+1): one coalesced 64-bit store per group.
template <int FX>
__device__ __forceinline__ void bwd_fast_plane_rows(float* __restrict__ dplane, int F2, int g, const float (&wx)[RF_F][2],
                                                    const float (&w2)[RF_F][8], const unsigned long long (&U)[8][2], bool st_ok) {
#pragma unroll
    for (int f2 = 0; f2 < RF_F; ++f2) {
        if (f2 < F2) {                                                     // warp-uniform
            unsigned long long t0 = mul2s(w2[f2][0], U[0][0]), t1 = mul2s(w2[f2][0], U[0][1]);
#pragma unroll
            for (int p = 1; p < 8; ++p) {
                t0 = fma2s(w2[f2][p], U[p][0], t0);
                t1 = fma2s(w2[f2][p], U[p][1], t1);
            }
            float* rowp = dplane + f2 * FX * RF_G;
#pragma unroll
            for (int m = 0; m < (FX + 3) / 4; ++m) {
                unsigned long long d[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int x = 4 * m + e;
                    d[e] = x < FX ? fma2s(wx[x][1], t1, mul2s(wx[x][0], t0)) : 0ull;
                }
                // step 1 (lanes 16 apart): g < 2 keeps voxels 0,1, g >= 2 keeps voxels 2,3
                const bool up2 = (g & 2) != 0;
                const unsigned long long s0 = up2 ? d[0] : d[2], s1 = up2 ? d[1] : d[3];       // what the partner wants
                const unsigned long long k0 = up2 ? d[2] : d[0], k1 = up2 ? d[3] : d[1];
                const unsigned long long e0 = add2(k0, shfl_xor2(s0, 16)), e1 = add2(k1, shfl_xor2(s1, 16));
                // step 2 (lanes 8 apart): even g keeps the first of its two voxels, odd g the second
                const bool up1 = (g & 1) != 0;
                const unsigned long long s = up1 ? e0 : e1, k = up1 ? e1 : e0;
                const unsigned long long tot = add2(k, shfl_xor2(s, 8));
                const int x = 4 * m + g;
                if (x < FX && st_ok) *reinterpret_cast<unsigned long long*>(rowp + x * RF_G) = tot;
            }
        }
    }
}

// grid (R, channel slabs).  boxes / flags for the gather pass are written by the CTAs of slab 0.
template <typename T, int PC>
__global__ void __launch_bounds__(RF_THREADS, 2)
roialign3d_bwd_roi_kernel(const T* __restrict__ gout, const float* __restrict__ rois, BwdRoi* __restrict__ broi,
                          float* __restrict__ tables, float* __restrict__ dF,
                          int C, int S, int H, int W, int Ps_, int Ph_, int Pw_, float scale, int sr, double zguard, int cpb) {
    extern __shared__ __align__(16) float s_buf[];
    __shared__ FwdShared sh;
    __shared__ FastShared fs;
    const int Ps = PC > 0 ? PC : Ps_, Ph = PC > 0 ? PC : Ph_, Pw = PC > 0 ? PC : Pw_;
    const int r = blockIdx.x;
    const int c_begin = blockIdx.y * cpb;
    const int ncta = min(cpb, C - c_begin);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    AxisP ax[3];
    int batch; float count;
    roi_axes(rois + (size_t)r * 7, scale, sr, Ps, Ph, Pw, S, H, W, zguard, ax[0], ax[1], ax[2], batch, count);
    const int P3 = Ps * Ph * Pw;
    // axis roles (grad_out is read in (S,H,W) order): x first, y scattered in registers (a2 = 1), z last (a3 = 0)
    if (warp == 0) {
        const int ta = min(lane >> 3, 2), tp = lane & 7;
        const AxisP mine = ta == 0 ? ax[0] : (ta == 1 ? ax[1] : ax[2]);
        const int myP = ta == 0 ? Ps : (ta == 1 ? Ph : Pw);
        const bool t_on = lane < 24 && tp < myP;
        int lo = 0x7fffffff, hi = -1;
        if (t_on) {
            for (int i = 0; i < mine.g; ++i) {
                const Tap t = axis_sample(mine, tp, i);
                if (t.valid) { lo = min(lo, t.low); hi = max(hi, t.high); }
            }
        }
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
        }
        float wcol[RF_F];
#pragma unroll
        for (int f = 0; f < RF_F; ++f) wcol[f] = 0.f;
        if (t_on) {
            for (int i = 0; i < mine.g; ++i) {
                const Tap t = axis_sample(mine, tp, i);
                if (t.valid) {
                    const int il = t.low - lo, ih = t.high - lo;
#pragma unroll
                    for (int f = 0; f < RF_F; ++f) { if (il == f) wcol[f] += t.h; if (ih == f) wcol[f] += t.l; }
                }
            }
        }
        const float zs = ta == 0 ? 1.0f / count : 1.0f;    // only used when count is a power of two: exact
        if (lane < 24) {
            float* dst = ta == 2 ? &sh.w[2][0][tp] : (ta == 1 ? &fs.w2[0][tp] : &fs.w3[0][tp]);
            const int fstride = ta == 2 ? 16 : 8;
#pragma unroll
            for (int f = 0; f < RF_F; ++f) dst[f * fstride] = wcol[f] * zs;
            if (tp == 0) { sh.lo[ta] = lo; sh.hi[ta] = hi; }
        }
    }
    __syncthreads();
    const int zlo = sh.lo[0], ylo = sh.lo[1], xlo = sh.lo[2];
    const bool empty = sh.hi[0] < 0 || sh.hi[1] < 0 || sh.hi[2] < 0;
    const int Fz = sh.hi[0] - zlo + 1, Fy = sh.hi[1] - ylo + 1, Fx = sh.hi[2] - xlo + 1;
    const int F2 = Fy, F3 = Fz;
    const int P2 = Ph, P3n = Ps;
    const int NC = Ph * Pw;                                // columns per channel
    const int UB = (RF_G * NC + 127) & ~127;               // floats per plane of the column buffer
    const int need = F3 * UB;
    const int icount = ax[0].g * ax[1].g * ax[2].g;
    const bool pow2 = (icount & (icount - 1)) == 0;
    const bool fast = !empty && pow2 && Fz <= RF_F && Fy <= RF_F && Fx <= RF_F && need <= RF_POOL_FLOATS;
    if (blockIdx.y == 0) {
        if (tid == 0) {
            BwdRoi b;
            for (int a = 0; a < 3; ++a) { b.lo[a] = sh.lo[a]; b.hi[a] = empty ? sh.lo[a] - 1 : sh.hi[a]; }
            b.batch = batch; b.fast = fast ? 1 : 0;
            broi[r] = b;
        }
        if (!fast && !empty) {
            // direct evaluation in the gather pass needs the adjoint tables over the whole extent: W[axis][voxel][bin], 8 bins
            float* tab = tables + (size_t)r * (S + H + W) * 8;
            for (int i = tid; i < (S + H + W) * 8; i += RF_THREADS) tab[i] = 0.f;
            __syncthreads();
            if (tid < 24) {
                const int a = tid >> 3, p = tid & 7;
                const int Pa = a == 0 ? Ps : (a == 1 ? Ph : Pw);
                const int seg = a == 0 ? 0 : (a == 1 ? S : S + H);
                const AxisP mine = a == 0 ? ax[0] : (a == 1 ? ax[1] : ax[2]);
                if (p < Pa) {
                    float* col = tab + (size_t)seg * 8 + p;
                    for (int i = 0; i < mine.g; ++i) {                     // one thread per column: no races
                        const Tap t = axis_sample(mine, p, i);
                        if (!t.valid) continue;
                        col[(size_t)t.low * 8] += t.h;
                        col[(size_t)t.high * 8] += t.l;
                    }
                }
            }
        }
    }
    if (!fast) return;

    const int pair = warp / RF_TEAM, half = warp - pair * RF_TEAM;
    const int n_active = min(RF_THREADS / (32 * RF_TEAM), RF_POOL_FLOATS / need);
    if (pair >= n_active) return;
    float* ubuf = s_buf + (size_t)pair * need;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, %1;" :: "r"(1 + pair), "n"(32 * RF_TEAM) : "memory"); };
    const int cp = lane & 7, g = lane >> 3;
    float wx[RF_F][2];
#pragma unroll
    for (int x = 0; x < RF_F; ++x) { wx[x][0] = sh.w[2][x][2 * g]; wx[x][1] = sh.w[2][x][2 * g + 1]; }
    float w2[RF_F][8];
#pragma unroll
    for (int f = 0; f < RF_F; ++f) {
        const float4 a = *reinterpret_cast<const float4*>(&fs.w2[f][0]), b = *reinterpret_cast<const float4*>(&fs.w2[f][4]);
        w2[f][0] = a.x; w2[f][1] = a.y; w2[f][2] = a.z; w2[f][3] = a.w;
        w2[f][4] = b.x; w2[f][5] = b.y; w2[f][6] = b.z; w2[f][7] = b.w;
    }
    const int groups = (ncta + RF_G - 1) / RF_G;
    const int NG = (C + RF_G - 1) / RF_G;
    const int plane_vox = F2 * Fx;
    // this lane's columns in a plane: channel pair cp, pw = 2g + j (the bin 7 of a 7-bin axis carries zero weights: its index
    // is clamped to a valid column instead of guarded), all ph bins
    int ucol[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) ucol[j] = (cp * NC + min(2 * g + j, Pw - 1)) * 2;

    for (int grp = pair; grp < groups; grp += n_active) {
        const int ncg = min(RF_G, ncta - grp * RF_G);
        const T* gg = gout + ((size_t)r * C + c_begin + grp * RF_G) * P3;
        const int NQ = ncg * NC;
        if (ncg < RF_G) {                                  // tail group: absent channels must read zeros, not stale shared memory
            for (int i = lane + 32 * half; i < F3 * UB; i += 32 * RF_TEAM) ubuf[i] = 0.f;
            pair_sync();
        }
        switch (F3) {
            case 1: bwd_fast_columns<T, 1, PC>(ubuf, fs, gg, NQ, NC, UB, P3, P3n, lane, half); break;
            case 2: bwd_fast_columns<T, 2, PC>(ubuf, fs, gg, NQ, NC, UB, P3, P3n, lane, half); break;
            case 3: bwd_fast_columns<T, 3, PC>(ubuf, fs, gg, NQ, NC, UB, P3, P3n, lane, half); break;
            case 4: bwd_fast_columns<T, 4, PC>(ubuf, fs, gg, NQ, NC, UB, P3, P3n, lane, half); break;
            case 5: bwd_fast_columns<T, 5, PC>(ubuf, fs, gg, NQ, NC, UB, P3, P3n, lane, half); break;
            case 6: bwd_fast_columns<T, 6, PC>(ubuf, fs, gg, NQ, NC, UB, P3, P3n, lane, half); break;
            case 7: bwd_fast_columns<T, 7, PC>(ubuf, fs, gg, NQ, NC, UB, P3, P3n, lane, half); break;
            default: bwd_fast_columns<T, 8, PC>(ubuf, fs, gg, NQ, NC, UB, P3, P3n, lane, half); break;
        }
        pair_sync();                                       // every plane of the column buffer is complete
        float* slot = dF + ((size_t)r * NG + (c_begin / RF_G + grp)) * RG_SLOT + 2 * cp;
        const bool st_ok = 2 * cp < ncg;                   // (channel counts are even in practice; an odd tail writes one spare float inside the slot)
        for (int k = half; k < F3; k += RF_TEAM) {
            unsigned long long U[8][2];
            const float* up = ubuf + k * UB;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int p = 0; p < 8; ++p)
                    U[p][j] = *reinterpret_cast<const unsigned long long*>(up + ucol[j] + min(p, P2 - 1) * (2 * Pw));
            float* dplane = slot + (size_t)k * plane_vox * RF_G;
            switch (Fx) {
                case 1: bwd_fast_plane_rows<1>(dplane, F2, g, wx, w2, U, st_ok); break;
                case 2: bwd_fast_plane_rows<2>(dplane, F2, g, wx, w2, U, st_ok); break;
                case 3: bwd_fast_plane_rows<3>(dplane, F2, g, wx, w2, U, st_ok); break;
                case 4: bwd_fast_plane_rows<4>(dplane, F2, g, wx, w2, U, st_ok); break;
                case 5: bwd_fast_plane_rows<5>(dplane, F2, g, wx, w2, U, st_ok); break;
                case 6: bwd_fast_plane_rows<6>(dplane, F2, g, wx, w2, U, st_ok); break;
                case 7: bwd_fast_plane_rows<7>(dplane, F2, g, wx, w2, U, st_ok); break;
                default: bwd_fast_plane_rows<8>(dplane, F2, g, wx, w2, U, st_ok); break;
            }
        }
        pair_sync();                                       // the column buffer is free again
    }
}

// grid (tiles of one batch element, channel groups, B); block 256 = 16 row slots x 16 channels.
// A thread owns, for its channel, the tile rows (z, y) with z mod 4 == sz and y mod 4 == sy -- RGQ rows of 8 voxels, accumulated
// in REGISTERS over the RoIs of the tile's list in index order (no barrier inside the list walk, fixed summation order),
// and finally written as four runs of 8 consecutive x (two 128-bit stores per row for fp32).
constexpr int RGT = 8;                                     // tile edge in y and x
constexpr int RGZ = 4;                                     // tile depth (z): 4 -> two rows per thread, 8 -> four (measured: 4 is faster, twice the CTAs at half the registers)
constexpr int RGQ = RGZ / 2;                               // rows per thread
constexpr int RG_LIST = 256;                               // RoIs listed per round
template <typename T, int PC>
__global__ void __launch_bounds__(256)
roialign3d_bwd_gather_kernel(const T* __restrict__ gout, const BwdRoi* __restrict__ broi, const float* __restrict__ tables,
                             const float* __restrict__ rois, const float* __restrict__ dF, T* __restrict__ gin,
                             int C, int S, int H, int W, int R, int Ps_, int Ph_, int Pw_, float scale, int sr,
                             int tiles_x, int tiles_y) {
    __shared__ int s_list[RG_LIST];
    __shared__ BwdRoi s_box[RG_LIST];
    __shared__ int s_wcnt[8];
    const int Ps = PC > 0 ? PC : Ps_, Ph = PC > 0 ? PC : Ph_, Pw = PC > 0 ? PC : Pw_;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, cg = blockIdx.y, b = blockIdx.z;
    const int t0[3] = {(tile / (tiles_x * tiles_y)) * RGZ, ((tile / tiles_x) % tiles_y) * RGT, (tile % tiles_x) * RGT};
    const int t1[3] = {min(t0[0] + RGZ, S) - 1, min(t0[1] + RGT, H) - 1, min(t0[2] + RGT, W) - 1};
    const int NG = (C + RF_G - 1) / RF_G;
    const int ncg = min(RF_G, C - cg * RF_G);
    const int c = tid & 15, slot = tid >> 4;               // channel, row slot
    const int sz = slot >> 2, sy = slot & 3;
    const int P3 = Ps * Ph * Pw;
    const bool c_ok = c < ncg;
    float acc[RGQ][RGT];                                     // rows (sz + 4 i, sy + 4 j), i, j in {0, 1}: index 2 i + j
#pragma unroll
    for (int q = 0; q < RGQ; ++q)
#pragma unroll
        for (int x = 0; x < RGT; ++x) acc[q][x] = 0.f;

    for (int r0 = 0; r0 < R; r0 += RG_LIST) {
        __syncthreads();                                   // everybody is done with the previous list
        const int rn = min(RG_LIST, R - r0);
        bool hit = false;
        BwdRoi mine;
        if (tid < rn) {
            mine = broi[r0 + tid];
            hit = mine.batch == b;
#pragma unroll
            for (int a = 0; a < 3; ++a) hit = hit && mine.hi[a] >= mine.lo[a] && mine.hi[a] >= t0[a] && mine.lo[a] <= t1[a];
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_wcnt[warp] = __popc(m);
        __syncthreads();
        int before = 0, n_list = 0;
        for (int w = 0; w < 8; ++w) { if (w < warp) before += s_wcnt[w]; n_list += s_wcnt[w]; }
        if (hit) { const int pos = before + __popc(m & ((1u << lane) - 1u)); s_list[pos] = r0 + tid; s_box[pos] = mine; }
        __syncthreads();
        for (int k = 0; k < n_list; ++k) {                 // RoIs in index order: fixed summation order per voxel
            const int r = s_list[k];
            const BwdRoi bx = s_box[k];
            const int Fy = bx.hi[1] - bx.lo[1] + 1, Fx = bx.hi[2] - bx.lo[2] + 1;
            const int xa = max(bx.lo[2], t0[2]), xb = min(bx.hi[2], t1[2]);      // x range inside the tile (not empty: the RoI hits the tile)
            if (bx.fast) {
                // branch-free: all (up to 32) loads of this RoI are issued before the first add, one L2 latency per RoI
                const float* base = dF + ((size_t)r * NG + cg) * RG_SLOT + c;
                float v[RGQ][RGT];
#pragma unroll
                for (int q = 0; q < RGQ; ++q) {
                    const int gz = t0[0] + sz + 4 * (q >> 1), gy = t0[1] + sy + 4 * (q & 1);
                    const bool rok = c_ok && gz >= bx.lo[0] && gz <= bx.hi[0] && gy >= bx.lo[1] && gy <= bx.hi[1];
                    const int roff = rok ? (((gz - bx.lo[0]) * Fy + (gy - bx.lo[1])) * Fx - bx.lo[2] + t0[2]) * RF_G : 0;
#pragma unroll
                    for (int x = 0; x < RGT; ++x) v[q][x] = 0.f;
                    if (__any_sync(0xffffffffu, rok)) {            // warp-uniform: neither of the warp's two rows lies inside this RoI
#pragma unroll
                        for (int x = 0; x < RGT; ++x) {
                            const int gx = t0[2] + x;
                            if (rok && gx >= xa && gx <= xb) v[q][x] = __ldg(base + roff + x * RF_G);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < RGQ; ++q)
#pragma unroll
                    for (int x = 0; x < RGT; ++x) acc[q][x] += v[q][x];
            } else if (c_ok) {
                // direct evaluation from the per-axis adjoint tables (RoIs with large footprints or odd sample counts)
                const float* tab = tables + (size_t)r * (S + H + W) * 8;
                const T* gc = gout + ((size_t)r * C + cg * RF_G + c) * P3;
                const float* roi = rois + (size_t)r * 7;
                const float rs = fmaxf(roi[6] * scale - roi[3] * scale, 1.f), rw = fmaxf(roi[4] * scale - roi[1] * scale, 1.f),
                            rh = fmaxf(roi[5] * scale - roi[2] * scale, 1.f);
                const int gz_ = sr > 0 ? sr : (int)ceilf(rs / Ps), gy_ = sr > 0 ? sr : (int)ceilf(rh / Ph), gx_ = sr > 0 ? sr : (int)ceilf(rw / Pw);
                const float inv_count = 1.0f / (float)(gz_ * gy_ * gx_);
#pragma unroll
                for (int q = 0; q < RGQ; ++q) {
                    const int gz = t0[0] + sz + 4 * (q >> 1), gy = t0[1] + sy + 4 * (q & 1);
                    if (gz < bx.lo[0] || gz > bx.hi[0] || gy < bx.lo[1] || gy > bx.hi[1]) continue;
                    const float* wz = tab + (size_t)gz * 8, *wy = tab + (size_t)(S + gy) * 8;
#pragma unroll
                    for (int x = 0; x < RGT; ++x) {
                        const int gx = t0[2] + x;
                        if (gx < xa || gx > xb) continue;
                        const float* wxp = tab + (size_t)(S + H + gx) * 8;
                        float a3 = 0.f;
                        for (int ps = 0; ps < Ps; ++ps) {
                            const float a = wz[ps];
                            if (a == 0.f) continue;
                            float s2 = 0.f;
                            for (int ph = 0; ph < Ph; ++ph) {
                                const float bq = wy[ph];
                                if (bq == 0.f) continue;
                                float s1 = 0.f;
                                for (int pw = 0; pw < Pw; ++pw) s1 = fmaf(wxp[pw], to_f(gc[(ps * Ph + ph) * Pw + pw]), s1);
                                s2 = fmaf(bq, s1, s2);
                            }
                            a3 = fmaf(a, s2, a3);
                        }
                        acc[q][x] += a3 * inv_count;
                    }
                }
            }
        }
    }
    // every grad_in element of the tile is written exactly once: four runs of 8 consecutive x per thread
    if (c_ok) {
#pragma unroll
        for (int q = 0; q < RGQ; ++q) {
            const int gz = t0[0] + sz + 4 * (q >> 1), gy = t0[1] + sy + 4 * (q & 1);
            if (gz >= S || gy >= H) continue;
            T* dst = gin + ((((size_t)b * C + cg * RF_G + c) * S + gz) * H + gy) * W + t0[2];
            if (sizeof(T) == 4 && t0[2] + RGT <= W && (W & 3) == 0) {
                reinterpret_cast<float4*>(dst)[0] = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
                reinterpret_cast<float4*>(dst)[1] = make_float4(acc[q][4], acc[q][5], acc[q][6], acc[q][7]);
            } else {
#pragma unroll
                for (int x = 0; x < RGT; ++x) if (t0[2] + x < W) dst[x] = from_f<T>(acc[q][x]);
            }
        }
    }
}

}  // namespace b200seg

using namespace b200seg;

static size_t bwd_table_floats(int R, int S, int H, int W, int PT) { return (size_t)R * (size_t)(S + H + W) * PT; }

// layout of the fast backward's workspace: [BwdRoi x R] [adjoint tables R x (S+H+W) x 8] [footprint gradients R x NG x slot]
static size_t bwd_fast_bytes(int R, int C, int S, int H, int W, size_t* off_tables, size_t* off_df) {
    const size_t r = R > 0 ? R : 1;
    const size_t ng = (size_t)(C + RF_G - 1) / RF_G;
    size_t o = align_up(r * sizeof(BwdRoi), 256);
    if (off_tables) *off_tables = o;
    o += align_up(r * (size_t)(S + H + W) * 8 * sizeof(float), 256);
    if (off_df) *off_df = o;
    o += align_up(r * ng * (size_t)RG_SLOT * sizeof(float), 256);
    return o + 256;
}

extern "C" size_t b200seg_roialign3d_bwd_workspace_bytes(int R, int C, int S, int H, int W, int P_max) {
    const size_t base = b200seg_roialign3d_workspace_bytes(R, S, H, W, P_max);
    if (P_max > 8 || C <= 0 || S <= 0 || H <= 0 || W <= 0) return base;
    const size_t fastb = bwd_fast_bytes(R, C, S, H, W, nullptr, nullptr);
    return fastb > base ? fastb : base;
}

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
    if (pmax <= 8) {
        int cpb = C <= 96 ? C : 96;                                    // channels per CTA (one RoI per CTA): six tasks for three teams of two warps
        cpb = (cpb + RA_CC - 1) / RA_CC * RA_CC;
        dim3 grid(R, (C + cpb - 1) / cpb);
        const size_t smem = RF_POOL_FLOATS * sizeof(float);
        const bool cubic7 = Ps == 7 && Ph == 7 && Pw == 7;             // the box head's pooled size
        auto launch = [&](auto kern) -> int {
            B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, RF_THREADS, smem, stream>>>((const T*)features, rois, (T*)output, C, S, H, W, Ps, Ph, Pw, scale, sr, cpb);
            return 0;
        };
        int e;
        if (cubic7) e = layout == 0 ? launch(roialign3d_fwd_fast_kernel<T, 7, true>) : launch(roialign3d_fwd_fast_kernel<T, 7, false>);
        else e = layout == 0 ? launch(roialign3d_fwd_fast_kernel<T, 0, true>) : launch(roialign3d_fwd_fast_kernel<T, 0, false>);
        if (e) return e;
        B200_LAUNCH_CHECK("roialign3d_fwd_fast_kernel");
        return 0;
    }
    dim3 grid(R, (C + RA_CC - 1) / RA_CC);
    const size_t smem = FwdCfg<16>::SMEM_FLOATS * sizeof(float);
    B200_CUDA(cudaFuncSetAttribute(roialign3d_fwd_kernel<T, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roialign3d_fwd_kernel<T, 16><<<grid, RA_THREADS, smem, stream>>>((const T*)features, rois, (T*)output, C, S, H, W,
                                                                     Ps, Ph, Pw, scale, sr, layout);
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
    if (pmax <= 8 && B <= 65535 && (C + RF_G - 1) / RF_G <= 65535) {
        size_t off_tables = 0, off_df = 0;
        const size_t need = bwd_fast_bytes(R, C, S, H, W, &off_tables, &off_df);
        if (workspace_bytes >= need) {                      // sized with b200seg_roialign3d_bwd_workspace_bytes: two-launch fast path
            BwdRoi* broi = (BwdRoi*)ws;
            float* tables = (float*)(ws + off_tables);
            float* dF = (float*)(ws + off_df);
            const double zguard = layout == 0 ? -0.1 : -1.0;
            const bool cubic7 = Ps == 7 && Ph == 7 && Pw == 7;
            const int tiles_x = (W + RGT - 1) / RGT, tiles_y = (H + RGT - 1) / RGT, tiles_z = (S + RGZ - 1) / RGZ;
            auto run = [&](auto kroi, auto kgather, auto tag) -> int {
                using T = decltype(tag);
                if (R > 0) {
                    int cpb = C <= 96 ? C : 96;
                    cpb = (cpb + RF_G - 1) / RF_G * RF_G;
                    dim3 grid(R, (C + cpb - 1) / cpb);
                    const size_t smem = RF_POOL_FLOATS * sizeof(float);
                    B200_CUDA(cudaFuncSetAttribute(kroi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    kroi<<<grid, RF_THREADS, smem, stream>>>((const T*)grad_out, rois, broi, tables, dF, C, S, H, W, Ps, Ph, Pw,
                                                             spatial_scale, sampling_ratio, zguard, cpb);
                    B200_LAUNCH_CHECK("roialign3d_bwd_roi_kernel");
                }
                dim3 g2(tiles_x * tiles_y * tiles_z, (C + RF_G - 1) / RF_G, B);
                kgather<<<g2, 256, 0, stream>>>((const T*)grad_out, broi, tables, rois, dF, (T*)grad_in, C, S, H, W, R, Ps, Ph, Pw,
                                                spatial_scale, sampling_ratio, tiles_x, tiles_y);
                B200_LAUNCH_CHECK("roialign3d_bwd_gather_kernel");
                return 0;
            };
            if (dtype == B200SEG_F32)
                return cubic7 ? run(roialign3d_bwd_roi_kernel<float, 7>, roialign3d_bwd_gather_kernel<float, 7>, float())
                              : run(roialign3d_bwd_roi_kernel<float, 0>, roialign3d_bwd_gather_kernel<float, 0>, float());
            return cubic7 ? run(roialign3d_bwd_roi_kernel<__nv_bfloat16, 7>, roialign3d_bwd_gather_kernel<__nv_bfloat16, 7>, __nv_bfloat16())
                          : run(roialign3d_bwd_roi_kernel<__nv_bfloat16, 0>, roialign3d_bwd_gather_kernel<__nv_bfloat16, 0>, __nv_bfloat16());
        }
    }
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
