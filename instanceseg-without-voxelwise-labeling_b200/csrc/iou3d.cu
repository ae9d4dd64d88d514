// iou3d.cu -- 3D box IoU matrix (replaces lib/utils/cython_bbox_3d.pyx:32-80).
//
// HBM-bound on the N*K fp32 output (4 B per pair): every thread produces four consecutive
// elements of the flattened [N,K] matrix and issues one 128-bit streaming store; query boxes are staged
// in shared memory (conflict-free layout), row boxes (24 B) are re-read through L1/L2.  The reference's mixed precision is reproduced
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
constexpr int IOU_QMAX = 1024;            // query boxes staged in shared memory per CTA

// Query boxes live in shared memory in a layout "transposed by 4": a thread owns 4 consecutive outputs, so the lanes
// of a warp read queries k = 4*lane + i; slot(k) = (k & 3) * KQ4 + (k >> 2) makes those reads consecutive
// (conflict-free 128-bit loads).  The fp64 part (union volume + divide, cython_bbox_3d.pyx:73-79) is needed only
// for the few pairs that overlap: each thread first settles its four pairs in fp32 and then works off its
// overlapping ones in a warp-wide loop, so the fp64 code runs about once per warp instead of once per pair slot.
// IDX = int when N*K < 2^31 (all index arithmetic in 32 bits), long long otherwise.
template <typename IDX>
__global__ void __launch_bounds__(IOU_THREADS)
iou3d_kernel(const float* __restrict__ boxes, long long N_, const float* __restrict__ query, long long K_,
             float* __restrict__ out, bool vec_ok) {
    const IDX N = (IDX)N_;
    const int K = (int)K_;
    __shared__ float4 s_qa[IOU_QMAX];     // q0 q1 q2 q3
    __shared__ float2 s_qb[IOU_QMAX];     // q4 q5
    __shared__ double s_qv[IOU_QMAX];     // volume as the reference keeps it: fp64 product rounded to fp32 (:52-56)
    const bool staged = K <= IOU_QMAX - 4;                    // the transposed layout uses up to 4 * ceil(K / 4) slots
    const int KQ4 = (int)((K + 3) >> 2);
    if (staged) {
        for (int k = threadIdx.x; k < (int)K; k += IOU_THREADS) {
            const float* q = query + (size_t)k * 6;
            const int slot = (k & 3) * KQ4 + (k >> 2);
            s_qa[slot] = make_float4(q[0], q[1], q[2], q[3]);
            s_qb[slot] = make_float2(q[4], q[5]);
            const double qv64 = __dmul_rn(__dmul_rn(__dadd_rn((double)__fsub_rn(q[3], q[0]), 1.0),
                                                    __dadd_rn((double)__fsub_rn(q[4], q[1]), 1.0)),
                                          __dadd_rn((double)__fsub_rn(q[5], q[2]), 1.0));
            s_qv[slot] = (double)__double2float_rn(qv64);
        }
        __syncthreads();
    }
    const IDX total = N * (IDX)K;
    const IDX nvec = (total + 3) >> 2;
    const IDX nvec_round = (nvec + 31) & ~(IDX)31;                // whole warps stay in the loop (warp votes below)
    // (row, column) of the thread's first element, advanced per iteration in mixed radix: the 64-bit division of
    // the flat index (about a hundred instructions) happens once per thread, not once per group of four outputs
    const IDX v0 = (IDX)blockIdx.x * IOU_THREADS + threadIdx.x;
    const IDX step_v = (IDX)gridDim.x * IOU_THREADS;
    const IDX step_n = (step_v << 2) / K;
    const int step_k = (int)((step_v << 2) - step_n * K);
    IDX n_it = (v0 << 2) / K;
    int k_it = (int)((v0 << 2) - n_it * K);
    for (IDX v = v0; v < nvec_round; v += step_v) {
        const IDX e0 = v << 2;
        const bool active = v < nvec;
        const bool full = active && e0 + 3 < total;               // only the very last group of four can be partial
        IDX n = n_it;
        int k = k_it;
        n_it += step_n; k_it += step_k;
        if (k_it >= K) { k_it -= K; ++n_it; }
        float r[4] = {0.f, 0.f, 0.f, 0.f};
        if (!staged) {                                   // very wide query sets: straight evaluation
            if (active) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (e0 + i < total) r[i] = iou_pair(boxes + (size_t)n * 6, query + (size_t)k * 6);
                    if (++k == K) { k = 0; ++n; }
                }
            }
        } else {
            // fp32 stage: intersection of the four pairs; overlapping ones are remembered (bit i of `pending`)
            float inter[4];
            int slot_of[4];
            IDX row_of[4];
            unsigned pending = 0u;
            float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f, b4 = 0.f, b5 = 0.f;
            IDX cur = -1;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                inter[i] = 0.f; slot_of[i] = 0; row_of[i] = n;
                if (full || (active && e0 + i < total)) {
                    if (n != cur) {                       // at most twice per thread (row change inside the group of 4)
                        const float* b = boxes + (size_t)n * 6;
                        b0 = b[0]; b1 = b[1]; b2 = b[2]; b3 = b[3]; b4 = b[4]; b5 = b[5];
                        cur = n;
                    }
                    const int slot = (k & 3) * KQ4 + (k >> 2);
                    const float4 qa = s_qa[slot];
                    const float2 qb = s_qb[slot];
                    // cython_bbox_3d.pyx:57-72 (see iou_pair)
                    const float iw = __fadd_rn(__fsub_rn(qa.w < b3 ? qa.w : b3, qa.x > b0 ? qa.x : b0), 1.0f);
                    const float ih = __fadd_rn(__fsub_rn(qb.x < b4 ? qb.x : b4, qa.y > b1 ? qa.y : b1), 1.0f);
                    const float is = __fadd_rn(__fsub_rn(qb.y < b5 ? qb.y : b5, qa.z > b2 ? qa.z : b2), 1.0f);
                    if (iw > 0 && ih > 0 && is > 0) {
                        inter[i] = __fmul_rn(__fmul_rn(iw, ih), is);
                        slot_of[i] = slot; pending |= 1u << i;
                    }
                }
                if (++k == K) { k = 0; ++n; }
            }
            // fp64 stage: one overlapping pair per thread per round until the warp has none left
            while (__any_sync(0xffffffffu, pending != 0u)) {
                if (pending) {
                    const int i = __ffs(pending) - 1;
                    pending &= pending - 1;
                    float it = inter[0]; int sl = slot_of[0]; IDX rn = row_of[0];
                    if (i == 1) { it = inter[1]; sl = slot_of[1]; rn = row_of[1]; }
                    if (i == 2) { it = inter[2]; sl = slot_of[2]; rn = row_of[2]; }
                    if (i == 3) { it = inter[3]; sl = slot_of[3]; rn = row_of[3]; }
                    const float* b = boxes + (size_t)rn * 6;
                    const double bv64 = __dmul_rn(__dmul_rn(__dadd_rn((double)__fsub_rn(b[3], b[0]), 1.0),
                                                            __dadd_rn((double)__fsub_rn(b[4], b[1]), 1.0)),
                                                  __dadd_rn((double)__fsub_rn(b[5], b[2]), 1.0));
                    const double uv = __dsub_rn(__dadd_rn(bv64, s_qv[sl]), (double)it);
                    const float res = __double2float_rn(__ddiv_rn((double)it, uv));
                    if (i == 0) r[0] = res; else if (i == 1) r[1] = res; else if (i == 2) r[2] = res; else r[3] = res;
                }
            }
        }
        if (!active) continue;
        if (vec_ok && full) {
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
    B200_CHECK_ARG(K < (1ll << 31) && N < (1ll << 40), "iou3d: matrix too large");
    const long long nvec = (N * K + 3) / 4;
    long long blocks = (nvec + IOU_THREADS - 1) / IOU_THREADS;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    const bool vec_ok = (((uintptr_t)overlaps) & 15) == 0;
    if (N * K + 4 * (cap + 1) * IOU_THREADS < (1ll << 31))
        iou3d_kernel<int><<<(unsigned)blocks, IOU_THREADS, 0, stream>>>(boxes, N, query, K, overlaps, vec_ok);
    else
        iou3d_kernel<long long><<<(unsigned)blocks, IOU_THREADS, 0, stream>>>(boxes, N, query, K, overlaps, vec_ok);
    B200_LAUNCH_CHECK("iou3d_kernel");
    return 0;
}
