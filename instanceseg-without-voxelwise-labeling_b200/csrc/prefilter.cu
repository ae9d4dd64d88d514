// prefilter.cu -- whole-volume prefilters and normalisations that bracket the hot path.
//
//   gaussian3d : scipy.ndimage.gaussian_filter(img, sigma) on an INTEGER volume as the reference calls it
//                (tools/binarization_nuclei.py:43).  scipy filters axis 0, then 1, then 2; every pass accumulates
//                in fp64 in a fixed order ( w[R]*x0, then += (x[-j]+x[+j])*w[R-j] from the outermost pair inwards,
//                multiply and add rounded separately ) and stores the result TRUNCATED to the volume's integer
//                dtype before the next pass reads it.  The three passes are reproduced bit for bit in one kernel:
//                a CTA marches along z through a (32+2R) x (64+2R) haloed footprint, keeps the last 2R+1 input
//                planes in shared memory, and runs z-pass -> y-pass -> x-pass on each plane with the integer
//                intermediates in shared memory, so the volume is read once and written once.
//   median3d   : scipy.ndimage.median_filter(img, size=3) (tools/binarization_nuclei.py:44): rank 13 of the 27
//                neighbours, 'reflect' borders (= clamp for radius 1).  Threads own an (y, x-pair) column, march
//                along z with the voxel pair packed in one u16x2 register, and use the generated min/max networks of
//                median27_net.cuh (one 9-sort per plane, one pruned 9+9 merge per two outputs, one 14-op select).
//   zscore     : (im - mean(im[im>0])) / std(im[im>0])  (tools/infer_simple.py:180-183, lib/utils/blob.py:180-184).
//   prm_to_u8  : per-channel  fm -= min; fm /= max; fm *= 255; astype(uint8)  (tools/infer_simple.py:233-238), fp32
//                operations in the reference's order, bit exact.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "median27_net.cuh"

namespace b200seg {

// =====================================================================================================
// gaussian3d
// =====================================================================================================

constexpr int GA_TY = 32, GA_TX = 64, GA_THREADS = 288, GA_RY = 4, GA_RX = 8, GA_RMAX = 8;
struct GaussW {
    double w[GA_RMAX + 1];                            // w[0..R]: w[R] is the centre tap, w[0] the outermost
    unsigned wi[GA_RMAX + 1];                         // min(round(w * 2^32), 2^32 - 1): weights of the fixed-point filter
};

// scipy 'reflect' (d c b a | a b c d | d c b a), any distance
__device__ __forceinline__ int reflect_idx(int i, int n) {
    if ((unsigned)i < (unsigned)n) return i;
    if (n == 1) return 0;
    const int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - 1 - i;
}

// exact unsigned (< 2^32) -> double without the conversion pipe: (2^52 + v) - 2^52
__device__ __forceinline__ double u2d(unsigned v) {
    return __dadd_rn(__hiloint2double(0x43300000, (int)v), -4503599627370496.0);
}

// One output of scipy's symmetric correlate1d over the taps v[0..2R] (values < 2^16).
//
// Exact path: fp64 in scipy's order, floor (== C truncation, the sum is >= 0).
// Fast path: the same sum in 32.32 fixed point, X = sum(wi_j * s_j) / 2^32.  With T the real-valued sum of the double
// weights, |X - T| <= 2^-32 * sum(s_j) <= 2^-32 * 17 * 65535 < 2.6e-4 and the fp64 result F has |F - T| < 3e-10, so
// whenever the fraction of X lies in [2^-11, 1 - 2^-11) both floors agree; otherwise (about 1 output in 1000, and
// every output of a constant non-zero neighbourhood) the exact path decides.  All-zero taps give 0 on both paths.
template <int R>
__device__ __forceinline__ unsigned gfilt_exact(const unsigned* v, const GaussW& gw) {
    double acc = __dmul_rn(gw.w[R], u2d(v[R]));
#pragma unroll
    for (int j = 0; j < R; ++j) acc = __dadd_rn(acc, __dmul_rn(u2d(v[j] + v[2 * R - j]), gw.w[j]));
    return (unsigned)__double2loint(__dadd_rd(acc, 4503599627370496.0));
}
template <int R, int ST>
__device__ __forceinline__ unsigned gfilt(const unsigned* v, const GaussW& gw) {
    static_assert(ST == 1, "contiguous windows only");
    unsigned long long acc = (unsigned long long)gw.wi[R] * v[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc += (unsigned long long)gw.wi[j] * (v[j] + v[2 * R - j]);
    const unsigned lo = (unsigned)acc;
    if ((lo - (1u << 21)) < (0u - (1u << 22)) || acc == 0ull) return (unsigned)(acc >> 32);
    return gfilt_exact<R>(v, gw);
}

template <int R> struct GaussCfg {
    static constexpr int NS = 2 * R + 1, FY = GA_TY + 2 * R, FX = GA_TX + 2 * R, FXW = FX / 2, BW = FXW | 1;
    static constexpr int NE = (FY * FX + GA_THREADS - 1) / GA_THREADS;
    static constexpr size_t SMEM = (size_t)(NS * FY * FXW + FY * FXW + GA_TY * BW + FX + FY) * 4;
};

template <typename T, int R>
__global__ void __launch_bounds__(GA_THREADS, (R <= 4 ? 3 : 1)) gauss3d_kernel(const T* __restrict__ in, T* __restrict__ out, int S, int H, int W, int zc,
                                                             const GaussW gw) {
    using C = GaussCfg<R>;
    constexpr int NS = C::NS, FY = C::FY, FX = C::FX, FXW = C::FXW, BW = C::BW, NE = C::NE;
    extern __shared__ uint32_t ga_sm[];
    uint32_t* ring = ga_sm;                               // [NS][FY][FXW]  raw input planes, two voxels per word
    uint32_t* A = ring + NS * FY * FXW;                   // [FY][FXW]      after the z pass
    uint32_t* B = A + FY * FXW;                           // [TY][BW]       after the y pass (odd word stride)
    int* gx = (int*)(B + GA_TY * BW);                     // [FX]  reflected source column
    int* gy = gx + FX;                                    // [FY]  reflected source row * W
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * GA_TX, y0 = blockIdx.y * GA_TY, z0 = blockIdx.z * zc;
    for (int i = tid; i < FX; i += GA_THREADS) gx[i] = reflect_idx(x0 - R + i, W);
    for (int i = tid; i < FY; i += GA_THREADS) gy[i] = reflect_idx(y0 - R + i, H) * W;
    __syncthreads();
    int off[NE];
#pragma unroll
    for (int k = 0; k < NE; ++k) {
        const int e = min(tid + k * GA_THREADS, FY * FX - 1);
        off[k] = gy[e / FX] + gx[e % FX];
    }
    const size_t plane = (size_t)H * W;
    uint16_t* ring16 = (uint16_t*)ring;
    // prologue: planes c = 0 .. 2R-1  (z = z0 - R + c)
    for (int c = 0; c < 2 * R; ++c) {
        const T* p = in + (size_t)reflect_idx(z0 - R + c, S) * plane;
#pragma unroll
        for (int k = 0; k < NE; ++k) {
            const int e = tid + k * GA_THREADS;
            if (e < FY * FX) ring16[c * FY * FX + e] = (uint16_t)__ldg(p + off[k]);
        }
    }
    unsigned pv[NE];
    {
        const T* p = in + (size_t)reflect_idx(z0 + R, S) * plane;
#pragma unroll
        for (int k = 0; k < NE; ++k) pv[k] = __ldg(p + off[k]);
    }
    for (int i = 0; i < zc; ++i) {
        const int z = z0 + i;
        if (z >= S) break;
        {
            const int slot = (i + 2 * R) % NS;
#pragma unroll
            for (int k = 0; k < NE; ++k) {
                const int e = tid + k * GA_THREADS;
                if (e < FY * FX) ring16[slot * FY * FX + e] = (uint16_t)pv[k];
            }
        }
        __syncthreads();
        if (i + 1 < zc && z + 1 < S) {                    // next plane's loads fly during this plane's arithmetic
            const T* p = in + (size_t)reflect_idx(z + 1 + R, S) * plane;
#pragma unroll
            for (int k = 0; k < NE; ++k) pv[k] = __ldg(p + off[k]);
        }
        // ---- z pass: ring -> A -------------------------------------------------------------------
        {
            int so[NS];                                   // ring slot of tap k, as a word offset
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                int s = i % NS + k;
                if (s >= NS) s -= NS;
                so[k] = s * FY * FXW;
            }
            for (int t = tid; t < FY * FXW; t += GA_THREADS) {
                unsigned lo[NS], hi[NS];
#pragma unroll
                for (int k = 0; k < NS; ++k) {
                    const unsigned w = ring[so[k] + t];
                    lo[k] = w & 0xFFFFu;
                    hi[k] = w >> 16;
                }
                A[t] = gfilt<R, 1>(lo, gw) | (gfilt<R, 1>(hi, gw) << 16);
            }
        }
        __syncthreads();
        // ---- y pass: A -> B (sliding window down a column pair) ------------------------------------------
        for (int t = tid; t < FXW * (GA_TY / GA_RY); t += GA_THREADS) {
            const int p = t % FXW, r = t / FXW;
            unsigned lo[GA_RY + 2 * R], hi[GA_RY + 2 * R];
#pragma unroll
            for (int k = 0; k < GA_RY + 2 * R; ++k) {
                const unsigned w = A[(r * GA_RY + k) * FXW + p];
                lo[k] = w & 0xFFFFu;
                hi[k] = w >> 16;
            }
#pragma unroll
            for (int o = 0; o < GA_RY; ++o)
                B[(r * GA_RY + o) * BW + p] = gfilt<R, 1>(lo + o, gw) | (gfilt<R, 1>(hi + o, gw) << 16);
        }
        __syncthreads();
        // ---- x pass: B -> global --------------------------------------------------------------------
        for (int t = tid; t < GA_TY * (GA_TX / GA_RX); t += GA_THREADS) {
            const int ty = t / (GA_TX / GA_RX), run = t % (GA_TX / GA_RX);
            unsigned v[GA_RX + 2 * R];
#pragma unroll
            for (int k = 0; k < (GA_RX + 2 * R) / 2; ++k) {
                const unsigned w = B[ty * BW + run * (GA_RX / 2) + k];
                v[2 * k] = w & 0xFFFFu;
                v[2 * k + 1] = w >> 16;
            }
            unsigned r[GA_RX];
#pragma unroll
            for (int o = 0; o < GA_RX; ++o) r[o] = gfilt<R, 1>(v + o, gw);
            const int y = y0 + ty, x = x0 + run * GA_RX;
            if (y < H && x < W) {
                T* q = out + (size_t)z * plane + (size_t)y * W + x;
                if (x + GA_RX <= W && (reinterpret_cast<uintptr_t>(q) & (sizeof(T) * GA_RX - 1)) == 0) {
                    if (sizeof(T) == 2) {
                        uint4 u;
                        u.x = r[0] | (r[1] << 16); u.y = r[2] | (r[3] << 16); u.z = r[4] | (r[5] << 16); u.w = r[6] | (r[7] << 16);
                        *reinterpret_cast<uint4*>(q) = u;
                    } else {
                        uint2 u;
                        u.x = r[0] | (r[1] << 8) | (r[2] << 16) | (r[3] << 24);
                        u.y = r[4] | (r[5] << 8) | (r[6] << 16) | (r[7] << 24);
                        *reinterpret_cast<uint2*>(q) = u;
                    }
                } else {
#pragma unroll
                    for (int o = 0; o < GA_RX; ++o)
                        if (x + o < W) q[o] = (T)r[o];
                }
            }
        }
    }
}

template <typename T, int R>
static int launch_gauss(const void* in, void* out, int S, int H, int W, const GaussW& gw, cudaStream_t stream) {
    using C = GaussCfg<R>;
    B200_CUDA(cudaFuncSetAttribute(gauss3d_kernel<T, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    const int gx = (W + GA_TX - 1) / GA_TX, gy = (H + GA_TY - 1) / GA_TY;
    const int per_sm = (int)(200 * 1024 / (C::SMEM + 1024)) > 0 ? (int)(200 * 1024 / (C::SMEM + 1024)) : 1;
    static const int waves = getenv("B200SEG_GAUSS_WAVES") ? atoi(getenv("B200SEG_GAUSS_WAVES")) : 3;
    const long long want = (long long)waves * num_sms() * per_sm;      // >= 3 waves of CTAs so the tail is short
    long long chunks = (want + (long long)gx * gy - 1) / ((long long)gx * gy);
    if (chunks < 1) chunks = 1;
    int zc = (int)((S + chunks - 1) / chunks);
    if (zc < 4) zc = 4;                                   // each chunk pays 2R planes of warm-up loads
    if (zc > S) zc = S;
    dim3 grid(gx, gy, (S + zc - 1) / zc);
    gauss3d_kernel<T, R><<<grid, GA_THREADS, C::SMEM, stream>>>((const T*)in, (T*)out, S, H, W, zc, gw);
    B200_LAUNCH_CHECK("gauss3d_kernel");
    return 0;
}

template <typename T>
static int dispatch_gauss(int R, const void* in, void* out, int S, int H, int W, const GaussW& gw, cudaStream_t stream) {
    switch (R) {
        case 1: return launch_gauss<T, 1>(in, out, S, H, W, gw, stream);
        case 2: return launch_gauss<T, 2>(in, out, S, H, W, gw, stream);
        case 3: return launch_gauss<T, 3>(in, out, S, H, W, gw, stream);
        case 4: return launch_gauss<T, 4>(in, out, S, H, W, gw, stream);
        case 5: return launch_gauss<T, 5>(in, out, S, H, W, gw, stream);
        case 6: return launch_gauss<T, 6>(in, out, S, H, W, gw, stream);
        case 7: return launch_gauss<T, 7>(in, out, S, H, W, gw, stream);
        case 8: return launch_gauss<T, 8>(in, out, S, H, W, gw, stream);
    }
    set_error("gaussian3d: radius %d not in 1..%d", R, GA_RMAX);
    return B200SEG_EINVAL;
}

// =====================================================================================================
// median3d (3x3x3)
// =====================================================================================================

constexpr int MD_THREADS = 256, MD_ROWS = MD_THREADS / 32, MD_LANES = 30;     // lanes 0 and 31 of a warp only carry the x halo

template <typename T, bool EVENW>
__device__ __forceinline__ unsigned md_pair(const T* p, bool has_hi) {
    if (EVENW) {
        if (sizeof(T) == 2) return __ldg(reinterpret_cast<const unsigned*>(p));
        const unsigned w = __ldg(reinterpret_cast<const uint16_t*>(p));
        return (w & 0xFFu) | ((w >> 8) << 16);
    }
    const unsigned lo = __ldg(p);
    return lo | ((has_hi ? (unsigned)__ldg(p + 1) : lo) << 16);
}

struct MdRaw { unsigned r[3]; };

// raw voxel pairs (x, x+1) of rows y-1, y, y+1 of one plane
template <typename T, bool EVENW>
__device__ __forceinline__ MdRaw md_load(const T* const (&rows)[3], size_t plane_off, bool has_hi) {
    MdRaw m;
#pragma unroll
    for (int k = 0; k < 3; ++k) m.r[k] = md_pair<T, EVENW>(rows[k] + plane_off, has_hi);
    return m;
}

// the nine in-plane neighbours of every lane's voxel pair, sorted per 16-bit half
__device__ __forceinline__ void md_sort_plane(const MdRaw& m, bool first, bool last, unsigned (&o)[9]) {
    unsigned v[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const unsigned c = m.r[k];
        unsigned prev = __shfl_up_sync(0xFFFFFFFFu, c, 1), next = __shfl_down_sync(0xFFFFFFFFu, c, 1);
        if (first) prev = c << 16;                                        // clamp: left neighbour of voxel 0 is voxel 0
        if (last) next = c >> 16;                                         // clamp: right neighbour of the last voxel
        v[3 * k + 0] = __funnelshift_l(prev, c, 16);                      // (x-1, x)
        v[3 * k + 1] = c;                                                 // (x,   x+1)
        v[3 * k + 2] = __funnelshift_r(c, next, 16);                      // (x+1, x+2)
    }
    median27_sort9(v, o);
}

template <typename T, bool EVENW>
__device__ __forceinline__ void md_store(T* q, unsigned r, bool has_hi) {
    if (EVENW) {
        if (sizeof(T) == 2) *reinterpret_cast<unsigned*>(q) = r;
        else *reinterpret_cast<uint16_t*>(q) = (uint16_t)((r & 0xFFu) | ((r >> 16) << 8));
    } else {
        q[0] = (T)(r & 0xFFFFu);
        if (has_hi) q[1] = (T)(r >> 16);
    }
}

template <typename T, bool EVENW>
__global__ void __launch_bounds__(MD_THREADS, 2) median3d_kernel(const T* __restrict__ in, T* __restrict__ out, int S, int H, int W, int zc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int y = blockIdx.y * MD_ROWS + warp;
    if (y >= H) return;                                   // whole warp; no block barriers below
    const int npairs = (W + 1) >> 1;
    const int pi = blockIdx.x * MD_LANES + lane - 1;      // this lane's voxel pair; lanes 0 / 31 hold the neighbours' halo
    const bool live = lane >= 1 && lane <= MD_LANES && pi < npairs;
    const int pc = min(max(pi, 0), npairs - 1);
    const int x = 2 * pc;
    const bool first = pc == 0, last = pc == npairs - 1, has_hi = x + 1 < W;
    const size_t ps = (size_t)H * W;
    const T* rows[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) rows[k] = in + (size_t)min(max(y + k - 1, 0), H - 1) * W + x;
    T* orow = out + (size_t)y * W + x;
    const int z0 = blockIdx.z * zc;
    auto plane = [&](int k) { return (size_t)min(max(z0 - 1 + k, 0), S - 1) * ps; };       // k-th plane this CTA touches
    unsigned sa[9], sb[9], sc[9], sd[9], m[10];
    MdRaw ra, rb;
    {
        const MdRaw r0 = md_load<T, EVENW>(rows, plane(0), has_hi), r1 = md_load<T, EVENW>(rows, plane(1), has_hi);
        ra = md_load<T, EVENW>(rows, plane(2), has_hi);
        rb = md_load<T, EVENW>(rows, plane(3), has_hi);
        md_sort_plane(r0, first, last, sa);
        md_sort_plane(r1, first, last, sb);
    }
    for (int i = 0; i < zc; i += 2) {
        const int z = z0 + i;
        if (z >= S) break;
        const MdRaw rc = ra, rd = rb;
        ra = md_load<T, EVENW>(rows, plane(i + 4), has_hi);               // two planes ahead of the arithmetic
        rb = md_load<T, EVENW>(rows, plane(i + 5), has_hi);
        md_sort_plane(rc, first, last, sc);
        median27_merge_mid(sb, sc, m);
        const unsigned r0 = median27_select(sa, m);
        if (live) md_store<T, EVENW>(orow + (size_t)z * ps, r0, has_hi);
        if (i + 1 >= zc || z + 1 >= S) break;
        md_sort_plane(rd, first, last, sd);
        const unsigned r1 = median27_select(sd, m);
        if (live) md_store<T, EVENW>(orow + (size_t)(z + 1) * ps, r1, has_hi);
#pragma unroll
        for (int k = 0; k < 9; ++k) { sa[k] = sc[k]; sb[k] = sd[k]; }
    }
}

template <typename T, bool EVENW>
static int launch_median(const void* in, void* out, int S, int H, int W, cudaStream_t stream) {
    const int gx = ((W + 1) / 2 + MD_LANES - 1) / MD_LANES, gy = (H + MD_ROWS - 1) / MD_ROWS;
    const long long want = 8ll * num_sms() * 2;
    long long chunks = (want + (long long)gx * gy - 1) / ((long long)gx * gy);
    if (chunks < 1) chunks = 1;
    int zc = (int)((S + chunks - 1) / chunks);
    if (zc < 16) zc = 16;                                 // each chunk pays two planes of warm-up sorts
    zc += zc & 1;
    if (zc > S) zc = S + (S & 1);
    dim3 grid(gx, gy, (S + zc - 1) / zc);
    median3d_kernel<T, EVENW><<<grid, MD_THREADS, 0, stream>>>((const T*)in, (T*)out, S, H, W, zc);
    B200_LAUNCH_CHECK("median3d_kernel");
    return 0;
}

// =====================================================================================================
// z-score normalisation over the non-zero voxels
// =====================================================================================================

constexpr int ZS_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(ZS_THREADS) zs_sums_int_kernel(const T* __restrict__ in, long long n, unsigned long long* __restrict__ acc) {
    unsigned long long s1 = 0, s2 = 0, cnt = 0;
    const long long stride = (long long)gridDim.x * ZS_THREADS;
    // one 128-bit streaming load = 16 uint8 / 8 uint16 values; their sum and non-zero count fit 32 bits per group
    constexpr int PER = 16 / (int)sizeof(T);
    const long long nvec = (reinterpret_cast<uintptr_t>(in) & 15) == 0 ? n / PER : 0;
    for (long long g = (long long)blockIdx.x * ZS_THREADS + threadIdx.x; g < nvec; g += stride) {
        alignas(16) T v[PER];
        *reinterpret_cast<uint4*>(v) = ld_stream_u4(in + g * PER);
        unsigned a = 0, c = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const unsigned x = v[k];
            a += x;
            c += x != 0;
            s2 += (unsigned long long)(x * x);                // x <= 65535: the product fits 32 bits
        }
        s1 += a;
        cnt += c;
    }
    for (long long i = nvec * PER + (long long)blockIdx.x * ZS_THREADS + threadIdx.x; i < n; i += stride) {
        const unsigned v = in[i];
        s1 += v;
        s2 += (unsigned long long)v * v;
        cnt += v != 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_down_sync(0xFFFFFFFFu, s1, o);
        s2 += __shfl_down_sync(0xFFFFFFFFu, s2, o);
        cnt += __shfl_down_sync(0xFFFFFFFFu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(acc + 0, s1);
        atomicAdd(acc + 1, s2);
        atomicAdd(acc + 2, cnt);
    }
}

// exact integer moments -> mean, population std (numpy's default ddof = 0)
__global__ void zs_finish_int_kernel(const unsigned long long* __restrict__ acc, double* __restrict__ stats) {
    const unsigned long long s1 = acc[0], s2 = acc[1], n = acc[2];
    const double mean = (double)s1 / (double)n;
    const unsigned __int128 num = (unsigned __int128)n * s2 - (unsigned __int128)s1 * s1;     // n * sum(x^2) - sum(x)^2 >= 0, exact
    const double numd = (double)(unsigned long long)(num >> 64) * 18446744073709551616.0 + (double)(unsigned long long)num;
    stats[0] = mean;
    stats[1] = sqrt(numd / ((double)n * (double)n));
    stats[2] = (double)n;
}

// fp32 input: deterministic two-pass moments (per-block partials summed in block order)
template <int PASS>
__global__ void __launch_bounds__(ZS_THREADS) zs_sums_f32_kernel(const float* __restrict__ in, long long n, const double* __restrict__ stats,
                                                                 double* __restrict__ part_s, double* __restrict__ part_n) {
    __shared__ double sh_s[ZS_THREADS / 32], sh_n[ZS_THREADS / 32];
    const double mean = PASS == 1 ? stats[0] : 0.0;
    double s = 0.0, c = 0.0;
    const long long stride = (long long)gridDim.x * ZS_THREADS;
    for (long long i = (long long)blockIdx.x * ZS_THREADS + threadIdx.x; i < n; i += stride) {
        const float v = in[i];
        if (v > 0.0f) {
            const double d = (double)v - mean;
            s += PASS == 1 ? d * d : d;
            c += 1.0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xFFFFFFFFu, s, o);
        c += __shfl_down_sync(0xFFFFFFFFu, c, o);
    }
    if ((threadIdx.x & 31) == 0) { sh_s[threadIdx.x >> 5] = s; sh_n[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tn = 0.0;
        for (int k = 0; k < ZS_THREADS / 32; ++k) { ts += sh_s[k]; tn += sh_n[k]; }
        part_s[blockIdx.x] = ts;
        part_n[blockIdx.x] = tn;
    }
}
template <int PASS>
__global__ void zs_finish_f32_kernel(const double* __restrict__ part_s, const double* __restrict__ part_n, int nblocks, double* __restrict__ stats) {
    double ts = 0.0, tn = 0.0;
    for (int k = 0; k < nblocks; ++k) { ts += part_s[k]; tn += part_n[k]; }
    if (PASS == 0) { stats[0] = ts / tn; stats[2] = tn; }
    else stats[1] = sqrt(ts / tn);
}

// (v - mean) / sd in fp64, rounded to float once (numpy keeps the float64 quotient, the network reads it as float32).
// Eight elements per thread: one 8- or 16-byte (32-byte for float input) streaming load, two 16-byte streaming stores; the
// division stays a true fp64 division (bit-identical to numpy for a given mean / sd), its cost hides behind the stores.
template <typename T>
__global__ void __launch_bounds__(ZS_THREADS) zs_apply_kernel(const T* __restrict__ in, float* __restrict__ out, long long n, const double* __restrict__ stats) {
    const double mean = stats[0], sd = stats[1];
    const long long stride = (long long)gridDim.x * ZS_THREADS;
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const long long nvec = aligned ? n >> 3 : 0;
    for (long long g = (long long)blockIdx.x * ZS_THREADS + threadIdx.x; g < nvec; g += stride) {
        alignas(16) T v[8];
        if (sizeof(T) == 1) *reinterpret_cast<uint2*>(v) = *reinterpret_cast<const uint2*>(in + g * 8);
        else if (sizeof(T) == 2) *reinterpret_cast<uint4*>(v) = ld_stream_u4(in + g * 8);
        else { reinterpret_cast<uint4*>(v)[0] = ld_stream_u4(in + g * 8); reinterpret_cast<uint4*>(v)[1] = ld_stream_u4(in + g * 8 + 4); }
        float r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = (float)(((double)v[k] - mean) / sd);
        st_stream_u4(out + g * 8, make_uint4(__float_as_uint(r[0]), __float_as_uint(r[1]), __float_as_uint(r[2]), __float_as_uint(r[3])));
        st_stream_u4(out + g * 8 + 4, make_uint4(__float_as_uint(r[4]), __float_as_uint(r[5]), __float_as_uint(r[6]), __float_as_uint(r[7])));
    }
    for (long long i = nvec * 8 + (long long)blockIdx.x * ZS_THREADS + threadIdx.x; i < n; i += stride)
        out[i] = (float)(((double)in[i] - mean) / sd);
}

constexpr int ZS_MAX_BLOCKS = 4096;

// =====================================================================================================
// PRM min-max -> uint8
// =====================================================================================================

__global__ void __launch_bounds__(256) prm_minmax_kernel(const float* __restrict__ in, long long per_map, unsigned* __restrict__ keys) {
    const int map = blockIdx.y;
    const float* p = in + (size_t)map * per_map;
    unsigned lo = 0xFFFFFFFFu, hi = 0u;
    const long long stride = (long long)gridDim.x * 256;
    const long long nvec = (reinterpret_cast<uintptr_t>(p) & 15) == 0 ? per_map >> 2 : 0;          // four floats per 128-bit load
    for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g < nvec; g += stride) {
        const uint4 v = ld_stream_u4(p + g * 4);
        const unsigned k0 = ordered_key(__uint_as_float(v.x)), k1 = ordered_key(__uint_as_float(v.y));
        const unsigned k2 = ordered_key(__uint_as_float(v.z)), k3 = ordered_key(__uint_as_float(v.w));
        lo = min(min(lo, min(k0, k1)), min(k2, k3));
        hi = max(max(hi, max(k0, k1)), max(k2, k3));
    }
    for (long long i = nvec * 4 + (long long)blockIdx.x * 256 + threadIdx.x; i < per_map; i += stride) {
        const unsigned k = ordered_key(p[i]);
        lo = min(lo, k);
        hi = max(hi, k);
    }
    lo = __reduce_min_sync(0xFFFFFFFFu, lo);
    hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(keys + 2 * map, lo);
        atomicMax(keys + 2 * map + 1, hi);
    }
}
__global__ void prm_minmax_init_kernel(unsigned* keys, int n_maps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_maps) { keys[2 * i] = 0xFFFFFFFFu; keys[2 * i + 1] = 0u; }
}
__global__ void __launch_bounds__(256) prm_scale_kernel(const float* __restrict__ in, uint8_t* __restrict__ out, long long per_map,
                                                        const unsigned* __restrict__ keys) {
    const int map = blockIdx.y;
    const float mn = key_to_float(keys[2 * map]);
    const float mx = __fsub_rn(key_to_float(keys[2 * map + 1]), mn);     // max(fm - min) == max(fm) - min (rounding is monotone)
    const float* p = in + (size_t)map * per_map;
    uint8_t* q = out + (size_t)map * per_map;
    const long long stride = (long long)gridDim.x * 256;
    auto scale = [&](float v) -> unsigned {
        const float r = __fmul_rn(__fdiv_rn(__fsub_rn(v, mn), mx), 255.0f);
        return r == r ? (unsigned)(uint8_t)(int)r : 0u;                  // constant map: 0/0 -> 0
    };
    // 16 values per thread: four 128-bit streaming loads, one 128-bit streaming store
    const long long nvec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(q)) & 15) == 0 ? per_map >> 4 : 0;
    for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g < nvec; g += stride) {
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = ld_stream_u4(p + g * 16 + k * 4);
        unsigned w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            w[k] = scale(__uint_as_float(v[k].x)) | scale(__uint_as_float(v[k].y)) << 8 | scale(__uint_as_float(v[k].z)) << 16 |
                   scale(__uint_as_float(v[k].w)) << 24;
        st_stream_u4(q + g * 16, make_uint4(w[0], w[1], w[2], w[3]));
    }
    for (long long i = nvec * 16 + (long long)blockIdx.x * 256 + threadIdx.x; i < per_map; i += stride) q[i] = (uint8_t)scale(p[i]);
}

}  // namespace b200seg

using namespace b200seg;

extern "C" int b200seg_gaussian3d_dev(const void* in, void* out, int elem_bytes, int S, int H, int W, const double* weights, int radius,
                                      b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(in && out && weights && in != out, "gaussian3d: null or aliased pointers");
    B200_CHECK_ARG(S > 0 && H > 0 && W > 0 && (long long)H * W < (1ll << 31), "gaussian3d: bad shape %d x %d x %d", S, H, W);
    B200_CHECK_ARG(elem_bytes == 1 || elem_bytes == 2, "gaussian3d: integer volumes only (uint8 / uint16), got %d-byte elements", elem_bytes);
    B200_CHECK_ARG(radius >= 1 && radius <= GA_RMAX, "gaussian3d: radius %d not in 1..%d", radius, GA_RMAX);
    GaussW gw;
    for (int j = 0; j <= GA_RMAX; ++j) {
        gw.w[j] = j <= radius ? weights[j] : 0.0;
        B200_CHECK_ARG(gw.w[j] >= 0.0 && gw.w[j] <= 1.0, "gaussian3d: weight %d = %g outside [0, 1]", j, gw.w[j]);
        const double q = nearbyint(gw.w[j] * 4294967296.0);
        gw.wi[j] = q >= 4294967295.0 ? 0xFFFFFFFFu : (unsigned)q;
    }
    return elem_bytes == 1 ? dispatch_gauss<uint8_t>(radius, in, out, S, H, W, gw, stream)
                           : dispatch_gauss<uint16_t>(radius, in, out, S, H, W, gw, stream);
}

extern "C" int b200seg_median3d_dev(const void* in, void* out, int elem_bytes, int S, int H, int W, b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(in && out && in != out, "median3d: null or aliased pointers");
    B200_CHECK_ARG(S > 0 && H > 0 && W > 0, "median3d: bad shape %d x %d x %d", S, H, W);
    B200_CHECK_ARG(elem_bytes == 1 || elem_bytes == 2, "median3d: integer volumes only (uint8 / uint16), got %d-byte elements", elem_bytes);
    const bool even = (W % 2 == 0) && W >= 2 && ((uintptr_t)in % 4 == 0) && ((uintptr_t)out % 4 == 0);
    if (elem_bytes == 2) return even ? launch_median<uint16_t, true>(in, out, S, H, W, stream) : launch_median<uint16_t, false>(in, out, S, H, W, stream);
    return even ? launch_median<uint8_t, true>(in, out, S, H, W, stream) : launch_median<uint8_t, false>(in, out, S, H, W, stream);
}

extern "C" size_t b200seg_zscore_workspace_bytes(void) { return 256 + (size_t)ZS_MAX_BLOCKS * 16 + 256; }

extern "C" int b200seg_zscore_norm_dev(const void* in, int elem_bytes, float* out, long long n, double* stats, void* workspace, size_t workspace_bytes,
                                       b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(in && stats && workspace && n > 0, "zscore_norm: null pointer or empty input");
    B200_CHECK_ARG(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4, "zscore_norm: %d-byte elements (1 = uint8, 2 = uint16, 4 = float32)", elem_bytes);
    if (workspace_bytes < b200seg_zscore_workspace_bytes()) { set_error("zscore_norm: workspace too small"); return B200SEG_EWORKSPACE; }
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    long long blocks = (n + ZS_THREADS * 8 - 1) / (ZS_THREADS * 8);
    const long long capb = (long long)num_sms() * 8;
    if (blocks > capb) blocks = capb;
    if (blocks > ZS_MAX_BLOCKS) blocks = ZS_MAX_BLOCKS;
    if (elem_bytes == 4) {
        double* part_s = (double*)ws;
        double* part_n = part_s + ZS_MAX_BLOCKS;
        zs_sums_f32_kernel<0><<<(unsigned)blocks, ZS_THREADS, 0, stream>>>((const float*)in, n, stats, part_s, part_n);
        B200_LAUNCH_CHECK("zs_sums_f32_kernel<0>");
        zs_finish_f32_kernel<0><<<1, 1, 0, stream>>>(part_s, part_n, (int)blocks, stats);
        B200_LAUNCH_CHECK("zs_finish_f32_kernel<0>");
        zs_sums_f32_kernel<1><<<(unsigned)blocks, ZS_THREADS, 0, stream>>>((const float*)in, n, stats, part_s, part_n);
        B200_LAUNCH_CHECK("zs_sums_f32_kernel<1>");
        zs_finish_f32_kernel<1><<<1, 1, 0, stream>>>(part_s, part_n, (int)blocks, stats);
        B200_LAUNCH_CHECK("zs_finish_f32_kernel<1>");
        if (out) { zs_apply_kernel<float><<<(unsigned)blocks, ZS_THREADS, 0, stream>>>((const float*)in, out, n, stats); B200_LAUNCH_CHECK("zs_apply_kernel"); }
        return 0;
    }
    unsigned long long* acc = (unsigned long long*)ws;
    B200_CUDA(cudaMemsetAsync(acc, 0, 24, stream));
    if (elem_bytes == 1) zs_sums_int_kernel<uint8_t><<<(unsigned)blocks, ZS_THREADS, 0, stream>>>((const uint8_t*)in, n, acc);
    else zs_sums_int_kernel<uint16_t><<<(unsigned)blocks, ZS_THREADS, 0, stream>>>((const uint16_t*)in, n, acc);
    B200_LAUNCH_CHECK("zs_sums_int_kernel");
    zs_finish_int_kernel<<<1, 1, 0, stream>>>(acc, stats);
    B200_LAUNCH_CHECK("zs_finish_int_kernel");
    if (out) {
        if (elem_bytes == 1) zs_apply_kernel<uint8_t><<<(unsigned)blocks, ZS_THREADS, 0, stream>>>((const uint8_t*)in, out, n, stats);
        else zs_apply_kernel<uint16_t><<<(unsigned)blocks, ZS_THREADS, 0, stream>>>((const uint16_t*)in, out, n, stats);
        B200_LAUNCH_CHECK("zs_apply_kernel");
    }
    return 0;
}

extern "C" size_t b200seg_prm_to_u8_workspace_bytes(int n_maps) { return 256 + (size_t)(n_maps > 0 ? n_maps : 0) * 8 + 256; }

extern "C" int b200seg_prm_to_u8_dev(const float* in, uint8_t* out, int n_maps, long long per_map, void* workspace, size_t workspace_bytes,
                                     b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(n_maps >= 0 && per_map >= 0 && n_maps <= 65535, "prm_to_u8: bad sizes");
    if (n_maps == 0 || per_map == 0) return 0;
    B200_CHECK_ARG(in && out && workspace, "prm_to_u8: null pointer");
    if (workspace_bytes < b200seg_prm_to_u8_workspace_bytes(n_maps)) { set_error("prm_to_u8: workspace too small"); return B200SEG_EWORKSPACE; }
    unsigned* keys = (unsigned*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    prm_minmax_init_kernel<<<(n_maps + 255) / 256, 256, 0, stream>>>(keys, n_maps);
    B200_LAUNCH_CHECK("prm_minmax_init_kernel");
    long long bx = (per_map + 256 * 8 - 1) / (256 * 8);
    const long long capb = ((long long)num_sms() * 8 + n_maps - 1) / n_maps;
    if (bx > capb) bx = capb;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, n_maps);
    prm_minmax_kernel<<<grid, 256, 0, stream>>>(in, per_map, keys);
    B200_LAUNCH_CHECK("prm_minmax_kernel");
    prm_scale_kernel<<<grid, 256, 0, stream>>>(in, out, per_map, keys);
    B200_LAUNCH_CHECK("prm_scale_kernel");
    return 0;
}
