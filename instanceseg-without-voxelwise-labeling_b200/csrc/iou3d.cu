// iou3d.cu -- 3D box IoU matrix (replaces lib/utils/cython_bbox_3d.pyx:32-80).
//
// HBM-bound on the N*K fp32 output (4 B per pair): every thread produces four consecutive
// elements of the flattened [N,K] matrix and issues one 128-bit streaming store; boxes and query
// boxes (24 B each) are re-read through L1/L2.  The reference's mixed precision is reproduced
// operation by operation (see oracle/oracle.c:oracle_bbox_overlaps_3d for the derivation from the
// Cython-generated C): fp32 differences, "+ 1.0" in fp64, an fp64 union volume and an fp64 divide
// rounded once to fp32.  (float)((double)f + 1.0) equals the fp32 add because double rounding is
// innocuous for a sum when the wide format has >= 2p+2 bits, so iw/ih/iss stay in fp32.
#include "common.cuh"

namespace b200seg {

__device__ __forceinline__ float iou_pair(const float* __restrict__ b, const float* __restrict__ q) {
    // cython_bbox_3d.pyx:57-79
    const float b0 = b[0], b1 = b[1], b2 = b[2], b3 = b[3], b4 = b[4], b5 = b[5];
    const float q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4], q5 = q[5];
    float mn = q3 < b3 ? q3 : b3;
    float mx = q0 > b0 ? q0 : b0;
    const float iw = __fadd_rn(__fsub_rn(mn, mx), 1.0f);
    if (!(iw > 0)) return 0.f;
    mn = q4 < b4 ? q4 : b4;
    mx = q1 > b1 ? q1 : b1;
    const float ih = __fadd_rn(__fsub_rn(mn, mx), 1.0f);
    if (!(ih > 0)) return 0.f;
    mn = q5 < b5 ? q5 : b5;
    mx = q2 > b2 ? q2 : b2;
    const float iss = __fadd_rn(__fsub_rn(mn, mx), 1.0f);
    if (!(iss > 0)) return 0.f;
    const float inter = __fmul_rn(__fmul_rn(iw, ih), iss);
    // query volume is stored as fp32 (DTYPE_t box_volume, :52-56); the box volume stays fp64
    const double qv64 = __dmul_rn(__dmul_rn(__dadd_rn((double)__fsub_rn(q3, q0), 1.0),
                                            __dadd_rn((double)__fsub_rn(q4, q1), 1.0)),
                                  __dadd_rn((double)__fsub_rn(q5, q2), 1.0));
    const float box_volume = __double2float_rn(qv64);
    const double bv64 = __dmul_rn(__dmul_rn(__dadd_rn((double)__fsub_rn(b3, b0), 1.0),
                                            __dadd_rn((double)__fsub_rn(b4, b1), 1.0)),
                                  __dadd_rn((double)__fsub_rn(b5, b2), 1.0));
    const double uv = __dsub_rn(__dadd_rn(bv64, (double)box_volume), (double)inter);
    return __double2float_rn(__ddiv_rn((double)inter, uv));
}

constexpr int IOU_THREADS = 256;

__global__ void __launch_bounds__(IOU_THREADS)
iou3d_kernel(const float* __restrict__ boxes, long long N, const float* __restrict__ query, long long K,
             float* __restrict__ out, bool vec_ok) {
    const long long total = N * K;
    const long long nvec = (total + 3) >> 2;
    for (long long v = (long long)blockIdx.x * IOU_THREADS + threadIdx.x; v < nvec;
         v += (long long)gridDim.x * IOU_THREADS) {
        const long long e0 = v << 2;
        long long n = e0 / K;
        long long k = e0 - n * K;
        float r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r[i] = 0.f;
            if (e0 + i < total) r[i] = iou_pair(boxes + n * 6, query + k * 6);
            if (++k == K) { k = 0; ++n; }
        }
        if (vec_ok && e0 + 3 < total) {
            st_stream_u4(out + e0, make_uint4(__float_as_uint(r[0]), __float_as_uint(r[1]),
                                              __float_as_uint(r[2]), __float_as_uint(r[3])));
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) if (e0 + i < total) out[e0 + i] = r[i];
        }
    }
}

}  // namespace b200seg

using namespace b200seg;

extern "C" int b200seg_iou3d_dev(const float* boxes, long long N, const float* query, long long K,
                                 float* overlaps, b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(N >= 0 && K >= 0, "iou3d: negative size");
    if (N == 0 || K == 0) return 0;
    B200_CHECK_ARG(boxes && query && overlaps, "iou3d: null pointer");
    const long long nvec = (N * K + 3) / 4;
    long long blocks = (nvec + IOU_THREADS - 1) / IOU_THREADS;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    const bool vec_ok = (((uintptr_t)overlaps) & 15) == 0;
    iou3d_kernel<<<(unsigned)blocks, IOU_THREADS, 0, stream>>>(boxes, N, query, K, overlaps, vec_ok);
    B200_LAUNCH_CHECK("iou3d_kernel");
    return 0;
}
