// peaks3d.cu -- PRM peak stimulation (replaces lib/prm/peak_stimulation_3d.py:9-48 and the median
// filter of lib/prm/peak_response_mapping_3d.py:45-49).
//
// The reference runs pad(-inf) -> max_pool3d(return_indices) -> (indices == own index) ->
// (input >= median) -> nonzero, i.e. five full-volume torch kernels plus an int64 index volume.
// Here (round 2 layout: one memset + four launches, no host round trip, every map of the batch per launch):
//   peaks_scan3_kernel     window 3: one pass over the maps with NO shared-memory tile.  A warp owns a strip of
//                          32 x-positions x 4 rows and marches along z; rows arrive as coalesced 128-byte loads
//                          straight into registers, x-neighbours by warp shuffle, and the ATen arg-max rule is
//                          evaluated on ordered integer keys with three-input maxima.  The candidate bits of a row
//                          are one warp ballot = one 32-bit word of the 1 bit/voxel mask (plain store), and the
//                          level-1 (top 12 key bits) radix histogram of the exact median lives in shared memory.
//                          The last CTA of each map selects the level-1 bin (no separate select launch).
//   peaks_scan_kernel<5|7> generic halo-tile kernel for the larger windows.
//   peaks_refine_kernel    two more radix levels (12 + 8 bits) over the L2-resident maps with 128-bit streaming
//                          loads; the last CTA of a map selects the next prefix / the final threshold.
//   peaks_finalize_kernel  filter (input >= threshold) + count + single-pass chained scan (decoupled look-back,
//                          work items handed out by ticket) + ordered emission of the (b,a,z,y,x) int64 rows +
//                          aggregation partial sums; the last CTA reduces them per map in a fixed order.
// Exact lower median = element of rank (V-1)/2 of the ascending order (torch.median).
#include "common.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace b200seg {

constexpr int PK_TX = 32, PK_TY = 8, PK_TZ = 8;
constexpr int PK_THREADS = 256;
constexpr int PK_NW = PK_THREADS / 32;
constexpr int PK_BINS1 = 4096;
constexpr int PK_CHUNK_WORDS = 1024;               // bitmask words (32 voxels each) per finalize work item: four per thread
constexpr int PK_RF_THREADS = 512;                 // refine kernels

struct PeakWs {
    uint32_t* hist1;      // [BA][4096]   key bits 31..20
    uint32_t* hist2;      // [BA][4096]   key bits 19..8 of the members of the level-1 bin
    uint32_t* hist3;      // [BA][256]    key bits 7..0 of the members of the level-2 bin
    uint32_t* ticket;     // [BA][4]      per-map CTA tickets of the scan / refine<2> / refine<3> kernels
    uint32_t* gticket;    // [4]          finalize: work-item ticket, done counter
    unsigned long long* look;   // [BA * nchunks]  per work item: READY flag | number of peaks (chained scan of the finalize kernel)
    unsigned long long* map_tot;// [BA]  READY flag | peaks of the whole map, published by the map's last finishing work item
    uint32_t* map_cnt;    // [BA]  finished work items per map
    uint32_t* bits;       // [BA][words]
    unsigned long long* sel;    // [BA][2]  (prefix, remaining rank) handed from one radix level to the next
    uint32_t* list;       // [BA][list_cap]  low 20 key bits of the members of the level-1 bin (list mode)
    uint32_t* list_n;     // [BA]  (zeroed per call)
    struct PeakIv* iv;    // [BA]  sampled interval of the median (interval path)
    struct PeakIvCnt* ivcnt;   // [BA]  exact counters of the interval pass (zeroed per call)
    long long list_cap;
    uint32_t* chunk_cnt;  // [BA][nchunks]
    float* chunk_sum;     // [BA][nchunks]
    float* thr;           // [BA]
    size_t zero_bytes;    // prefix of the workspace that must be cleared per call
    size_t total_bytes;
    int words, nchunks;
};

static PeakWs peak_layout(void* base, int BA, long long V) {
    PeakWs w;
    char* p = (char*)base;
    w.words = (int)((V + 31) / 32);
    w.nchunks = (w.words + PK_CHUNK_WORDS - 1) / PK_CHUNK_WORDS;
    auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes, 256); return r; };
    w.hist1 = (uint32_t*)take((size_t)BA * PK_BINS1 * 4);
    w.hist2 = (uint32_t*)take((size_t)BA * PK_BINS1 * 4);
    w.hist3 = (uint32_t*)take((size_t)BA * 256 * 4);
    w.ticket = (uint32_t*)take((size_t)BA * 4 * 4);
    w.gticket = (uint32_t*)take(16);
    w.list_n = (uint32_t*)take((size_t)BA * 4);
    w.ivcnt = (PeakIvCnt*)take((size_t)BA * 16);
    w.look = (unsigned long long*)take((size_t)BA * w.nchunks * 8);
    w.map_tot = (unsigned long long*)take((size_t)BA * 8);
    w.map_cnt = (uint32_t*)take((size_t)BA * 4);
    w.bits = (uint32_t*)take((size_t)BA * w.words * 4);
    w.zero_bytes = (size_t)(p - (char*)base);
    w.sel = (unsigned long long*)take((size_t)BA * 2 * 8);
    w.chunk_cnt = (uint32_t*)take((size_t)BA * w.nchunks * 4);
    w.chunk_sum = (float*)take((size_t)BA * w.nchunks * 4);
    w.thr = (float*)take((size_t)BA * 4);
    w.iv = (PeakIv*)take((size_t)BA * 16);
    // member list of the level-1 bin: room for a quarter of the map (a map whose median bin holds more falls back to
    // full passes over the map, decided on the device)
    w.list_cap = V < 65536 ? V : (V / 4 < 65536 ? 65536 : (V / 4 > 131072 ? 131072 : V / 4));
    w.list = (uint32_t*)take((size_t)BA * (size_t)w.list_cap * 4);
    w.total_bytes = (size_t)(p - (char*)base);
    return w;
}

// Finds the bin holding ascending rank k in hist[nbins] (read through L2: the counts were produced by atomics of
// other CTAs); returns bin and the rank inside it.  Must be called by all NT threads of the CTA; s_tmp has NT/32+2
// entries.  Block-wide prefix sum of per-thread bin groups (warp shuffles + one shared-memory hop); the single
// thread whose group contains rank k resolves it.
template <int NT>
__device__ void select_bin(const uint32_t* hist, int nbins, unsigned long long k,
                           unsigned long long* s_tmp, int* out_bin, unsigned long long* out_k) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    constexpr int PER = (PK_BINS1 + NT - 1) / NT;          // nbins <= PK_BINS1: a thread's bins sit in registers
    const int per = (nbins + NT - 1) / NT;
    const int b0 = min(nbins, tid * per);
    uint32_t h[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) h[j] = (j < per && b0 + j < nbins) ? __ldcg(hist + b0 + j) : 0u;     // independent loads
    unsigned long long local = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) local += h[j];
    unsigned long long incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long a = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += a;
    }
    if (lane == 31) s_tmp[warp] = incl;
    if (tid == 0) { s_tmp[NW] = (unsigned long long)(nbins - 1); s_tmp[NW + 1] = 0ull; }   // fallback: rank beyond the total
    __syncthreads();
    unsigned long long before = 0;
    for (int w = 0; w < warp; ++w) before += s_tmp[w];
    const unsigned long long excl = before + incl - local;
    if (k >= excl && k < excl + local) {                   // exactly one thread (groups are disjoint, local > 0 here)
        unsigned long long acc = excl;
        int b = b0;
        bool done = false;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (!done) { if (acc + h[j] > k) done = true; else { acc += h[j]; ++b; } }
        }
        s_tmp[NW] = (unsigned long long)b;
        s_tmp[NW + 1] = k - acc;
    }
    __syncthreads();
    *out_bin = (int)s_tmp[NW];
    *out_k = s_tmp[NW + 1];
    __syncthreads();
}

template <int NT>
__device__ void select_bin_smem(const uint32_t* hist, int nbins, unsigned long long k,
                           unsigned long long* s_tmp, int* out_bin, unsigned long long* out_k) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    constexpr int PER = (PK_BINS1 + NT - 1) / NT;          // nbins <= PK_BINS1: a thread's bins sit in registers
    const int per = (nbins + NT - 1) / NT;
    const int b0 = min(nbins, tid * per);
    uint32_t h[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) h[j] = (j < per && b0 + j < nbins) ? hist[b0 + j] : 0u;     // independent loads
    unsigned long long local = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) local += h[j];
    unsigned long long incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long a = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += a;
    }
    if (lane == 31) s_tmp[warp] = incl;
    if (tid == 0) { s_tmp[NW] = (unsigned long long)(nbins - 1); s_tmp[NW + 1] = 0ull; }   // fallback: rank beyond the total
    __syncthreads();
    unsigned long long before = 0;
    for (int w = 0; w < warp; ++w) before += s_tmp[w];
    const unsigned long long excl = before + incl - local;
    if (k >= excl && k < excl + local) {                   // exactly one thread (groups are disjoint, local > 0 here)
        unsigned long long acc = excl;
        int b = b0;
        bool done = false;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (!done) { if (acc + h[j] > k) done = true; else { acc += h[j]; ++b; } }
        }
        s_tmp[NW] = (unsigned long long)b;
        s_tmp[NW + 1] = k - acc;
    }
    __syncthreads();
    *out_bin = (int)s_tmp[NW];
    *out_k = s_tmp[NW + 1];
    __syncthreads();
}

// "last CTA of the map" protocol: every CTA adds its shared-memory histogram to the map's global one, then takes
// a ticket; the CTA that draws the last ticket sees every contribution (fence + atomics are ordered through L2).
template <int NT>
__device__ bool flush_hist_and_vote(const uint32_t* s_hist, int nbins, uint32_t* gh, uint32_t* ticket, int* s_flag) {
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += NT) { const uint32_t c = s_hist[i]; if (c) atomicAdd(&gh[i], c); }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t t = atomicAdd(ticket, 1u);
        *s_flag = (t == gridDim.x - 1);
        __threadfence();
    }
    __syncthreads();
    return *s_flag != 0;
}

// epilogue of the scan kernels: publish the level-1 histogram; the last CTA selects the bin of the median
template <int NT = PK_THREADS>
__device__ void flush_level1(const uint32_t* s_hist, int ba, long long V, PeakWs ws) {
    __shared__ unsigned long long s_tmp[NT / 32 + 2];
    __shared__ int s_flag;
    uint32_t* gh = ws.hist1 + (size_t)ba * PK_BINS1;
    if (!flush_hist_and_vote<NT>(s_hist, PK_BINS1, gh, ws.ticket + ba * 4 + 0, &s_flag)) return;
    int bin1; unsigned long long k1;
    select_bin<NT>(gh, PK_BINS1, (unsigned long long)((V - 1) / 2), s_tmp, &bin1, &k1);
    if (threadIdx.x == 0) { ws.sel[ba * 2] = (unsigned long long)bin1; ws.sel[ba * 2 + 1] = k1; }
}

template <int WIN>
__global__ void __launch_bounds__(PK_THREADS)
peaks_scan_kernel(const float* __restrict__ in, int S, int H, int W, int tiles_x, int tiles_y, int tiles_z,
                  int do_hist, PeakWs ws) {
    constexpr int O = (WIN - 1) / 2;
    constexpr int HX = PK_TX + 2 * O, HY = PK_TY + 2 * O, HZ = PK_TZ + 2 * O;
    __shared__ float s_tile[HZ * HY * HX];
    __shared__ uint32_t s_hist[PK_BINS1];
    const int ba = blockIdx.y;
    const long long V = (long long)S * H * W;
    const float* vol = in + (size_t)ba * V;
    uint32_t* bits = ws.bits + (size_t)ba * ws.words;
    const int tid = threadIdx.x, lane = tid & 31;
    if (do_hist) for (int i = tid; i < PK_BINS1; i += PK_THREADS) s_hist[i] = 0u;
    const int ntiles = tiles_x * tiles_y * tiles_z;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, tz = tile / (tiles_x * tiles_y);
        const int x0 = tx * PK_TX, y0 = ty * PK_TY, z0 = tz * PK_TZ;
        __syncthreads();
        for (int i = tid; i < HZ * HY * HX; i += PK_THREADS) {
            const int hx = i % HX, hy = (i / HX) % HY, hz = i / (HX * HY);
            const int gx = x0 + hx - O, gy = y0 + hy - O, gz = z0 + hz - O;
            float v = -CUDART_INF_F;
            if (gx >= 0 && gx < W && gy >= 0 && gy < H && gz >= 0 && gz < S) v = vol[((size_t)gz * H + gy) * W + gx];
            s_tile[i] = v;
        }
        __syncthreads();
        const int lx = tid % PK_TX, ly = tid / PK_TX;       // 32 x 8 threads, each walks PK_TZ voxels in z
        const int gx = x0 + lx, gy = y0 + ly;
        for (int lz = 0; lz < PK_TZ; ++lz) {
            const int gz = z0 + lz;
            const bool inb = gx < W && gy < H && gz < S;
            const float v = s_tile[((lz + O) * HY + (ly + O)) * HX + (lx + O)];
            bool peak = false;
            if (inb) {
                // ATen max_pool3d arg-max rule on the -inf padded volume (oracle.c: oracle_peak_stimulation)
                peak = true;
                const bool vnan = v != v;
                if (!vnan && v == -CUDART_INF_F) peak = false;
                for (int dz = 0; dz < WIN && peak; ++dz)
                    for (int dy = 0; dy < WIN && peak; ++dy)
#pragma unroll
                        for (int dx = 0; dx < WIN; ++dx) {
                            const int rel = (dz - O) * 9 * WIN * WIN + (dy - O) * 3 * WIN + (dx - O);   // sign = raster order
                            if (rel == 0) continue;
                            const float q = s_tile[((lz + dz) * HY + (ly + dy)) * HX + (lx + dx)];
                            if (vnan) { if (rel > 0 && q != q) peak = false; }
                            else if (rel < 0) { if (!(q < v)) peak = false; }
                            else { if (!(q <= v)) peak = false; }
                        }
            }
            if (do_hist) {                                   // warp-uniform: every lane reaches the match
                const uint32_t bin = inb ? (ordered_key(v) >> 20) : 0xFFFFFFFFu;
                // warp aggregation for the contended case (all lanes in one bin); otherwise plain atomics
                const uint32_t b0 = __shfl_sync(0xffffffffu, bin, 0);
                if (__all_sync(0xffffffffu, bin == b0)) { if (lane == 0 && b0 != 0xFFFFFFFFu) atomicAdd(&s_hist[b0], 32u); }
                else if (inb) atomicAdd(&s_hist[bin], 1u);
            }
            if (peak) {
                const long long flat = ((long long)gz * H + gy) * W + gx;
                atomicOr(&bits[flat >> 5], 1u << (flat & 31));
            }
        }
    }
    if (do_hist) flush_level1(s_hist, ba, V, ws);
}

// ---- window 3 (the reference default, peak_response_mapping_3d.py:29): register marching, no shared tile ----
// Values are handled as ORDERED KEYS (monotone uint32 image of the float, every NaN = 0xFFFFFFFF, -0 = +0), so
// the ATen arg-max rule becomes integer arithmetic: with Kb / Ka the unsigned maxima over the 13 neighbours that
// come earlier / later in the (z,y,x) window scan,
//     v not NaN:  peak <=> v != -inf  and  Kb < key(v)  and  Ka <= key(v)      (a NaN neighbour has the largest key)
//     v NaN:      peak <=> Ka != key(NaN)                                       (the LAST NaN of a window wins)
// The 13 + 13 split is separable: with m3(z,y,x) = max over x-1..x+1 and m9(z,y,x) = max over y-1..y+1 of m3,
//     Kb = max(m9(z-1,y,x), m3(z,y-1,x), key(z,y,x-1)),   Ka = max(m9(z+1,y,x), m3(z,y+1,x), key(z,y,x+1)).
// A warp owns 32 consecutive x (one lane each) x P3_R rows and marches along z.  Per plane it loads P3_R + 2 rows
// (coalesced 128-byte segments; lanes 0 / 31 also fetch the two x-halo elements), converts them to keys once,
// gets the x-neighbours by shuffle, and keeps m9 of the previous plane, the centre plane and the next plane in
// registers; the next plane's loads are issued one iteration ahead.  Out-of-volume elements are -inf (the padding
// of peak_stimulation_3d.py:14-16).
constexpr uint32_t KEY_NAN = 0xFFFFFFFFu, KEY_NINF = 0x007FFFFFu;          // ordered_key(-inf) = ~0xFF800000

// Same map as ordered_key() in three instructions: the add canonicalises -0 to +0 AND every NaN (any sign, any
// payload) to the GPU's canonical NaN 0x7FFFFFFF, whose key is 0xFFFFFFFF.
__device__ __forceinline__ uint32_t fkey(float f) {
    const uint32_t u = __float_as_uint(__fadd_rn(f, 0.0f));
    return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}

template <int P3_R> struct P3Raw { float v[P3_R + 2], h[P3_R + 2]; };
template <int P3_R> struct P3Plane { uint32_t k[P3_R], l[P3_R], r[P3_R], m3[P3_R + 2], m9[P3_R]; };

// per-unit constants of a lane: row offsets inside a z plane and which rows / halo elements exist
template <int P3_R> struct P3Unit {
    int roff[P3_R + 2];          // (y0 - 1 + j) * W + gx
    int mywoff;                  // ALIGNED: lane i < P3_R stores the mask word of row y0 + i: its word index inside a z plane (-1: none)
    uint32_t vmask, hmask;       // bit j: row j is inside the map and this lane's element / halo element exists
    int hoff;
    bool interior, h_ok;         // interior (warp-uniform): every row and every lane of the unit is inside the map
};

template <int P3_R>
__device__ __forceinline__ void p3_load(P3Raw<P3_R>& rw, const float* __restrict__ vol, int z, int S, size_t HW,
                                        const P3Unit<P3_R>& u) {
    if (z < 0 || z >= S) {                                  // warp-uniform: the -inf padding plane
#pragma unroll
        for (int j = 0; j < P3_R + 2; ++j) { rw.v[j] = -CUDART_INF_F; rw.h[j] = -CUDART_INF_F; }
        return;
    }
    const float* p = vol + (size_t)z * HW;
    const float* ph = p + u.hoff;
    if (u.interior) {
#pragma unroll
        for (int j = 0; j < P3_R + 2; ++j) {
            rw.v[j] = __ldg(p + u.roff[j]);
            rw.h[j] = u.h_ok ? __ldg(ph + u.roff[j]) : -CUDART_INF_F;
        }
    } else {
#pragma unroll
        for (int j = 0; j < P3_R + 2; ++j) {
            rw.v[j] = (u.vmask >> j) & 1u ? __ldg(p + u.roff[j]) : -CUDART_INF_F;
            rw.h[j] = (u.hmask >> j) & 1u ? __ldg(ph + u.roff[j]) : -CUDART_INF_F;
        }
    }
}

template <int P3_R>
__device__ __forceinline__ void p3_process(P3Plane<P3_R>& pl, const P3Raw<P3_R>& rw, int lane) {
#pragma unroll
    for (int j = 0; j < P3_R + 2; ++j) {
        const uint32_t kk = fkey(rw.v[j]), hk = fkey(rw.h[j]);
        uint32_t l = __shfl_up_sync(0xffffffffu, kk, 1), r = __shfl_down_sync(0xffffffffu, kk, 1);
        l = lane == 0 ? hk : l;
        r = lane == 31 ? hk : r;
        pl.m3[j] = __vimax3_u32(l, kk, r);
        if (j >= 1 && j <= P3_R) { pl.k[j - 1] = kk; pl.l[j - 1] = l; pl.r[j - 1] = r; }
    }
#pragma unroll
    for (int i = 0; i < P3_R; ++i) pl.m9[i] = __vimax3_u32(pl.m3[i], pl.m3[i + 1], pl.m3[i + 2]);
}

// outputs of plane z (centre plane `cur`, previous plane summarised by m9p, next plane `nxt`).
// peak <=> v != -inf, no later NaN unless... in key arithmetic (see above):
//   kv != key(-inf)  and  (Kb < kv or kv is NaN)  and  Ka <= kv  and  Ka is not NaN
// (for a non-NaN centre Ka <= kv already implies that Ka is not NaN; for a NaN centre Ka <= kv always holds).
template <int P3_R, bool ALIGNED>
__device__ __forceinline__ void p3_outputs(const uint32_t (&m9p)[P3_R], const P3Plane<P3_R>& cur, const P3Plane<P3_R>& nxt,
                                           int z, int y0, int x0, int H, int W, size_t HW, int lane, uint32_t omask,
                                           uint32_t rowmask, const P3Unit<P3_R>& u, uint32_t* __restrict__ bits,
                                           uint32_t* __restrict__ bz, uint32_t* s_hist, int do_hist) {
    uint32_t myword = 0;
#pragma unroll
    for (int i = 0; i < P3_R; ++i) {
        const uint32_t kb = __vimax3_u32(m9p[i], cur.m3[i], cur.l[i]);
        const uint32_t ka = __vimax3_u32(nxt.m9[i], cur.m3[i + 2], cur.r[i]);
        const uint32_t kv = cur.k[i];
        const bool inb = (omask >> i) & 1u;
        const bool peak = inb & (kv != KEY_NINF) & ((kb < kv) | (kv == KEY_NAN)) & (ka <= kv) & (ka != KEY_NAN);
        const uint32_t word = __ballot_sync(0xffffffffu, peak);
        const bool row = (rowmask >> i) & 1u;               // warp-uniform: row y0 + i exists
        if (ALIGNED) {                                      // W % 32 == 0: the word belongs to this row segment alone
            myword = lane == i ? word : myword;
        } else if (row && word) {
            const long long flat0 = ((long long)z * H + (y0 + i)) * W + x0;
            uint32_t* wp = bits + (flat0 >> 5);
            const int sh = (int)(flat0 & 31);
            if (lane == 0) atomicOr(wp, word << sh);
            if (lane == 1 && sh && (word >> (32 - sh))) atomicOr(wp + 1, word >> (32 - sh));
        }
        if (do_hist && inb) atomicAdd(&s_hist[kv >> 20], 1u);           // NaN keys land in the last bin
    }
    if (ALIGNED && u.mywoff >= 0) bz[u.mywoff] = myword;
}

// grid (ctas per map, BA).  Work unit of a warp = (z chunk, x strip, row group); adjacent warps take adjacent row
// groups of one strip, so the two halo rows they share are L1 hits.  H * W < 2^31 (checked by the launcher).
template <int P3_R, bool ALIGNED, int MINB>
__global__ void __launch_bounds__(PK_THREADS, MINB)
peaks_scan3_kernel(const float* __restrict__ in, int S, int H, int W, int nstrips, int nrowg, int nzc, int tz,
                   int do_hist, PeakWs ws) {
    __shared__ uint32_t s_hist[PK_BINS1];
    const int ba = blockIdx.y;
    const size_t HW = (size_t)H * W;
    const long long V = (long long)S * (long long)HW;
    const float* vol = in + (size_t)ba * V;
    uint32_t* bits = ws.bits + (size_t)ba * ws.words;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (do_hist) for (int i = tid; i < PK_BINS1; i += PK_THREADS) s_hist[i] = 0u;
    __syncthreads();
    const int nunits = nstrips * nrowg * nzc;
    for (int unit = blockIdx.x * PK_NW + warp; unit < nunits; unit += gridDim.x * PK_NW) {
        const int rowg = unit % nrowg;
        const int t = unit / nrowg;
        const int strip = t % nstrips, zc = t / nstrips;
        const int x0 = strip * 32, y0 = rowg * P3_R, z0 = zc * tz, z1 = min(S, z0 + tz);
        const int gx = x0 + lane;
        const bool col_ok = gx < W;
        P3Unit<P3_R> u;
        u.h_ok = (lane == 0 && x0 > 0) || (lane == 31 && x0 + 32 < W);
        u.hoff = lane == 0 ? -1 : 1;
        uint32_t rows = 0;                                  // bit j: row y0 - 1 + j is inside the map
#pragma unroll
        for (int j = 0; j < P3_R + 2; ++j) {
            const int gy = y0 - 1 + j;
            u.roff[j] = gy * W + gx;
            if (gy >= 0 && gy < H) rows |= 1u << j;
        }
        u.mywoff = (lane < P3_R && y0 + lane < H) ? ((y0 + lane) * W + x0) >> 5 : -1;
        u.vmask = col_ok ? rows : 0u;
        u.hmask = u.h_ok ? rows : 0u;
        u.interior = rows == (1u << (P3_R + 2)) - 1u && x0 + 32 <= W;
        const uint32_t rowmask = rows >> 1;                 // bit i: centre row y0 + i
        const uint32_t omask = col_ok ? rowmask : 0u;

        P3Raw<P3_R> ra, rb;
        P3Plane<P3_R> pa, pb;
        uint32_t m9p[P3_R];
        p3_load(ra, vol, z0 - 1, S, HW, u);
        p3_process(pa, ra, lane);
#pragma unroll
        for (int i = 0; i < P3_R; ++i) m9p[i] = pa.m9[i];
        p3_load(ra, vol, z0, S, HW, u);
        p3_process(pa, ra, lane);
        p3_load(ra, vol, z0 + 1, S, HW, u);
        int z = z0;
        const size_t HW32 = HW >> 5;
        uint32_t* bz = bits + (size_t)z0 * HW32;            // ALIGNED: first word of plane z
        while (true) {                                      // two planes per trip: the plane / raw buffers swap roles
            if (z + 2 <= z1) p3_load(rb, vol, z + 2, S, HW, u);                 // one plane ahead
            p3_process(pb, ra, lane);
            p3_outputs<P3_R, ALIGNED>(m9p, pa, pb, z, y0, x0, H, W, HW, lane, omask, rowmask, u, bits, bz, s_hist, do_hist);
#pragma unroll
            for (int i = 0; i < P3_R; ++i) m9p[i] = pa.m9[i];
            bz += HW32;
            if (++z >= z1) break;
            if (z + 2 <= z1) p3_load(ra, vol, z + 2, S, HW, u);
            p3_process(pa, rb, lane);
            p3_outputs<P3_R, ALIGNED>(m9p, pb, pa, z, y0, x0, H, W, HW, lane, omask, rowmask, u, bits, bz, s_hist, do_hist);
#pragma unroll
            for (int i = 0; i < P3_R; ++i) m9p[i] = pb.m9[i];
            bz += HW32;
            if (++z >= z1) break;
        }
    }
    if (do_hist) flush_level1(s_hist, ba, V, ws);
}

// ---- window 3, rows that are multiples of 128 voxels: four consecutive x per lane ---------------------------------
// Same algorithm as peaks_scan3_kernel with a 128-voxel strip per warp: a lane owns x0 + 4*lane .. +3, a row is ONE
// 128-bit load per lane, three of the four x-neighbour pairs live in the lane's own registers and only the two
// edge elements travel by shuffle (the two strip-edge lanes fetch one halo element).  Per plane a warp loads
// PW_R + 2 rows and emits PW_R rows of 128 voxels.  Planes without a NaN in reach (one vote per plane) use the
// two-compare form of the predicate: Kb < kv (which already excludes kv = -inf: every Kb >= key(-inf)) and Ka <= kv.
constexpr int PW_R = 2;
constexpr int PW_THREADS = 128, PW_NW = PW_THREADS / 32, PW_MINB = 3;

struct PWRaw { uint4 v[PW_R + 2]; float h[PW_R + 2]; };
struct PWPlane {
    uint32_t k[PW_R][4], le[PW_R], re[PW_R];   // centre rows: keys, left neighbour of element 0, right neighbour of element 3
    uint32_t m3[PW_R + 2][4], m9[PW_R][4];
};

// pz = first voxel of the plane to load (only dereferenced when z_ok); all_rows: every row of the unit is inside the map
__device__ __forceinline__ void pw_load(PWRaw& rw, const float* __restrict__ pz, bool z_ok, const int (&roff)[PW_R + 2],
                                        uint32_t rows, bool all_rows, bool h_ok, int hoff) {
    const uint32_t NINF_BITS = 0xFF800000u;
#pragma unroll
    for (int j = 0; j < PW_R + 2; ++j) { rw.v[j] = make_uint4(NINF_BITS, NINF_BITS, NINF_BITS, NINF_BITS); rw.h[j] = -CUDART_INF_F; }
    if (!z_ok) return;                                      // warp-uniform: the -inf padding plane
    if (all_rows) {
#pragma unroll
        for (int j = 0; j < PW_R + 2; ++j) {
            const float* q;
            asm("mad.wide.s32 %0, %1, 4, %2;" : "=l"(q) : "r"(roff[j]), "l"(pz));
            rw.v[j] = __ldg(reinterpret_cast<const uint4*>(q));
            if (h_ok) rw.h[j] = __ldg(q + hoff);
        }
    } else {
#pragma unroll
        for (int j = 0; j < PW_R + 2; ++j) {
            if ((rows >> j) & 1u) {                         // warp-uniform
                const float* q;
                asm("mad.wide.s32 %0, %1, 4, %2;" : "=l"(q) : "r"(roff[j]), "l"(pz));
                rw.v[j] = __ldg(reinterpret_cast<const uint4*>(q));
                if (h_ok) rw.h[j] = __ldg(q + hoff);
            }
        }
    }
}

// returns (warp-uniform) whether some key of the plane's PW_R + 2 rows (with the x halo) is NaN
__device__ __forceinline__ bool pw_process(PWPlane& pl, const PWRaw& rw, int lane) {
    uint32_t mx = 0u;
#pragma unroll
    for (int j = 0; j < PW_R + 2; ++j) {
        const uint32_t k0 = fkey(__uint_as_float(rw.v[j].x)), k1 = fkey(__uint_as_float(rw.v[j].y));
        const uint32_t k2 = fkey(__uint_as_float(rw.v[j].z)), k3 = fkey(__uint_as_float(rw.v[j].w));
        const uint32_t hk = fkey(rw.h[j]);
        uint32_t le = __shfl_up_sync(0xffffffffu, k3, 1), re = __shfl_down_sync(0xffffffffu, k0, 1);
        le = lane == 0 ? hk : le;
        re = lane == 31 ? hk : re;
        pl.m3[j][0] = __vimax3_u32(le, k0, k1);
        pl.m3[j][1] = __vimax3_u32(k0, k1, k2);
        pl.m3[j][2] = __vimax3_u32(k1, k2, k3);
        pl.m3[j][3] = __vimax3_u32(k2, k3, re);
        mx = __vimax3_u32(mx, pl.m3[j][0], pl.m3[j][3]);    // covers le, k0..k3, re
        if (j >= 1 && j <= PW_R) {
            pl.k[j - 1][0] = k0; pl.k[j - 1][1] = k1; pl.k[j - 1][2] = k2; pl.k[j - 1][3] = k3;
            pl.le[j - 1] = le; pl.re[j - 1] = re;
        }
    }
#pragma unroll
    for (int i = 0; i < PW_R; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) pl.m9[i][e] = __vimax3_u32(pl.m3[i][e], pl.m3[i + 1][e], pl.m3[i + 2][e]);
    return __any_sync(0xffffffffu, mx == KEY_NAN);
}

template <bool HIST, bool SLOW>
__device__ __forceinline__ void pw_outputs(const uint32_t (&m9p)[PW_R][4], const PWPlane& cur, const PWPlane& nxt,
                                           uint32_t rowmask, int lane, int mywoff, int w32, uint32_t* __restrict__ bz, uint32_t hist_addr) {
#pragma unroll
    for (int i = 0; i < PW_R; ++i) {
        if (!((rowmask >> i) & 1u)) continue;               // warp-uniform: row y0 + i is outside the map
        uint32_t nib = 0u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint32_t left = e == 0 ? cur.le[i] : cur.k[i][e - 1], right = e == 3 ? cur.re[i] : cur.k[i][e + 1];
            const uint32_t kb = __vimax3_u32(m9p[i][e], cur.m3[i][e], left);
            const uint32_t ka = __vimax3_u32(nxt.m9[i][e], cur.m3[i + 2][e], right);
            const uint32_t kv = cur.k[i][e];
            bool peak;
            if (!SLOW) peak = (kb < kv) & (ka <= kv);
            else peak = ((kb < kv) | (kv == KEY_NAN)) & (ka <= kv) & (ka != KEY_NAN) & (kv != KEY_NINF);
            nib |= peak ? (1u << e) : 0u;
            if (HIST) {                                     // NaN keys land in the last bin
                const uint32_t a = hist_addr + ((kv >> 18) & 0x3FFCu);
                asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(a), "r"(1u) : "memory");
            }
        }
        uint32_t w = nib << (4 * (lane & 7));               // 8 lanes = one 32-voxel word of the mask
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        w |= __shfl_xor_sync(0xffffffffu, w, 4);
        if ((lane & 7) == 0) bz[mywoff + i * w32] = w;
    }
}

// grid (ctas per map, BA); W % 128 == 0, maps 16-byte aligned, H * W < 2^31 (checked by the launcher).
// Work unit of a warp = (z chunk, 128-voxel strip, pair of rows); adjacent warps take adjacent row pairs.
// Per trip: issue the loads of plane z + 2, emit plane z from the planes already in registers (this hides the
// load latency), then turn the loaded rows into plane z + 2 in the registers plane z just vacated.
template <bool HIST, int MINB>
__global__ void __launch_bounds__(PW_THREADS, MINB)
peaks_scan3w_kernel(const float* __restrict__ in, int S, int H, int W, int nstrips, int nrowg, int nzc, int tz, PeakWs ws) {
    __shared__ uint32_t s_hist[PK_BINS1];
    const int ba = blockIdx.y;
    const size_t HW = (size_t)H * W;
    const long long V = (long long)S * (long long)HW;
    const float* vol = in + (size_t)ba * V;
    uint32_t* bits = ws.bits + (size_t)ba * ws.words;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (HIST) for (int i = tid; i < PK_BINS1; i += PW_THREADS) s_hist[i] = 0u;
    __syncthreads();
    const uint32_t hist_addr = (uint32_t)__cvta_generic_to_shared(s_hist);
    const int w32 = W >> 5;
    const size_t HW32 = HW >> 5;
    const int nunits = nstrips * nrowg * nzc;
    for (int unit = blockIdx.x * PW_NW + warp; unit < nunits; unit += gridDim.x * PW_NW) {
        const int rowg = unit % nrowg;
        const int t = unit / nrowg;
        const int strip = t % nstrips, zc = t / nstrips;
        const int x0 = strip * 128, y0 = rowg * PW_R, z0 = zc * tz, z1 = min(S, z0 + tz);
        const bool h_ok = (lane == 0 && x0 > 0) || (lane == 31 && x0 + 128 < W);
        const int hoff = lane == 0 ? -1 : 4;
        int roff[PW_R + 2];
        uint32_t rows = 0;                                  // bit j: row y0 - 1 + j is inside the map
#pragma unroll
        for (int j = 0; j < PW_R + 2; ++j) {
            const int gy = y0 - 1 + j;
            roff[j] = gy * W + x0 + 4 * lane;
            if (gy >= 0 && gy < H) rows |= 1u << j;
        }
        const bool all_rows = rows == (1u << (PW_R + 2)) - 1u;
        const uint32_t rowmask = rows >> 1;                 // bit i: centre row y0 + i
        const int mywoff = ((y0 * W + x0) >> 5) + (lane >> 3);

        PWRaw rw;
        PWPlane pa, pb;
        uint32_t m9p[PW_R][4];
        uint32_t nanbits;                                   // bit 0: plane z - 1, bit 1: plane z, bit 2: plane z + 1 holds a NaN
        pw_load(rw, vol + (ptrdiff_t)(z0 - 1) * (ptrdiff_t)HW, z0 >= 1, roff, rows, all_rows, h_ok, hoff);
        nanbits = pw_process(pa, rw, lane) ? 1u : 0u;
#pragma unroll
        for (int i = 0; i < PW_R; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) m9p[i][e] = pa.m9[i][e];
        pw_load(rw, vol + (size_t)z0 * HW, true, roff, rows, all_rows, h_ok, hoff);
        nanbits |= pw_process(pa, rw, lane) ? 2u : 0u;
        pw_load(rw, vol + (size_t)(z0 + 1) * HW, z0 + 1 < S, roff, rows, all_rows, h_ok, hoff);
        nanbits |= pw_process(pb, rw, lane) ? 4u : 0u;
        int z = z0;
        const float* pz = vol + (size_t)(z0 + 2) * HW;     // plane z + 2
        uint32_t* bz = bits + (size_t)z0 * HW32;            // first mask word of plane z
        while (true) {                                      // two planes per trip: pa / pb swap roles
            const bool more = z + 1 < z1;                   // plane z + 2 is needed (as the next plane of plane z + 1)
            if (more) pw_load(rw, pz, z + 2 < S, roff, rows, all_rows, h_ok, hoff);
            if (nanbits) pw_outputs<HIST, true>(m9p, pa, pb, rowmask, lane, mywoff, w32, bz, hist_addr);     // warp-uniform
            else pw_outputs<HIST, false>(m9p, pa, pb, rowmask, lane, mywoff, w32, bz, hist_addr);
            if (!more) break;
#pragma unroll
            for (int i = 0; i < PW_R; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) m9p[i][e] = pa.m9[i][e];
            nanbits = (nanbits >> 1) | (pw_process(pa, rw, lane) ? 4u : 0u);
            ++z; pz += HW; bz += HW32;
            const bool more2 = z + 1 < z1;
            if (more2) pw_load(rw, pz, z + 2 < S, roff, rows, all_rows, h_ok, hoff);
            if (nanbits) pw_outputs<HIST, true>(m9p, pb, pa, rowmask, lane, mywoff, w32, bz, hist_addr);
            else pw_outputs<HIST, false>(m9p, pb, pa, rowmask, lane, mywoff, w32, bz, hist_addr);
            if (!more2) break;
#pragma unroll
            for (int i = 0; i < PW_R; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) m9p[i][e] = pb.m9[i][e];
            nanbits = (nanbits >> 1) | (pw_process(pb, rw, lane) ? 4u : 0u);
            ++z; pz += HW; bz += HW32;
        }
    }
    if (HIST) flush_level1<PW_THREADS>(s_hist, ba, V, ws);
}

// ---- window 3, wide rows, FLOAT domain (round 2) ----------------------------------------------------------------------
// Same strip / register-marching layout as peaks_scan3w_kernel, but without ordered keys and without the histogram:
// the maxima are taken with the NaN-propagating three-input float maximum of sm_100 (FMNMX3.NAN), so that with
// Kb / Ka = maxima over the 13 neighbours that come earlier / later in the window scan
//     v not NaN:  peak <=> Kb < v and Ka <= v     (a NaN neighbour makes Kb / Ka NaN and both compares false; v = -inf
//                                                  fails Kb < v because Kb >= -inf; -0 == +0 as in the key order)
//     v NaN:      peak <=> Ka is not NaN          (the LAST NaN of a window wins the ATen arg-max)
// Three instructions per loaded element (key conversion) and the per-voxel shared-memory histogram update are gone; the
// exact median now comes from peaks_sample_kernel + peaks_interval_kernel.
__device__ __forceinline__ float fmax3n(float a, float b, float c) {
    float d;
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

struct PFRaw { float4 v[PW_R + 2]; float h[PW_R + 2]; };
struct PFPlane {
    float k[PW_R][4], le[PW_R], re[PW_R];       // centre rows: values, left neighbour of element 0, right neighbour of element 3
    float m3[PW_R + 2][4], m9[PW_R][4];
};

__device__ __forceinline__ void pf_load(PFRaw& rw, const float* __restrict__ pz, bool z_ok, const int (&roff)[PW_R + 2],
                                        uint32_t rows, bool all_rows, bool h_ok, int hoff) {
    const float4 ninf4 = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    if (z_ok && all_rows) {                                 // the common case carries no initialisation moves
#pragma unroll
        for (int j = 0; j < PW_R + 2; ++j) {
            const float* q;
            asm("mad.wide.s32 %0, %1, 4, %2;" : "=l"(q) : "r"(roff[j]), "l"(pz));
            rw.v[j] = __ldg(reinterpret_cast<const float4*>(q));
            rw.h[j] = -CUDART_INF_F;
            if (h_ok) rw.h[j] = __ldg(q + hoff);
        }
    } else if (z_ok) {
#pragma unroll
        for (int j = 0; j < PW_R + 2; ++j) {
            rw.v[j] = ninf4; rw.h[j] = -CUDART_INF_F;
            if ((rows >> j) & 1u) {                         // warp-uniform
                const float* q;
                asm("mad.wide.s32 %0, %1, 4, %2;" : "=l"(q) : "r"(roff[j]), "l"(pz));
                rw.v[j] = __ldg(reinterpret_cast<const float4*>(q));
                if (h_ok) rw.h[j] = __ldg(q + hoff);
            }
        }
    } else {                                                // warp-uniform: the -inf padding plane
#pragma unroll
        for (int j = 0; j < PW_R + 2; ++j) { rw.v[j] = ninf4; rw.h[j] = -CUDART_INF_F; }
    }
}

// returns (warp-uniform) whether some value of the plane's PW_R + 2 rows (with the x halo) is NaN
__device__ __forceinline__ bool pf_process(PFPlane& pl, const PFRaw& rw, int lane) {
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < PW_R + 2; ++j) {
        const float k0 = rw.v[j].x, k1 = rw.v[j].y, k2 = rw.v[j].z, k3 = rw.v[j].w;
        float le = __shfl_up_sync(0xffffffffu, k3, 1), re = __shfl_down_sync(0xffffffffu, k0, 1);
        le = lane == 0 ? rw.h[j] : le;
        re = lane == 31 ? rw.h[j] : re;
        pl.m3[j][0] = fmax3n(le, k0, k1);
        pl.m3[j][1] = fmax3n(k0, k1, k2);
        pl.m3[j][2] = fmax3n(k1, k2, k3);
        pl.m3[j][3] = fmax3n(k2, k3, re);
        mx = fmax3n(mx, pl.m3[j][0], pl.m3[j][3]);          // covers le, k0..k3, re
        if (j >= 1 && j <= PW_R) {
            pl.k[j - 1][0] = k0; pl.k[j - 1][1] = k1; pl.k[j - 1][2] = k2; pl.k[j - 1][3] = k3;
            pl.le[j - 1] = le; pl.re[j - 1] = re;
        }
    }
#pragma unroll
    for (int i = 0; i < PW_R; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) pl.m9[i][e] = fmax3n(pl.m3[i][e], pl.m3[i + 1][e], pl.m3[i + 2][e]);
    return __any_sync(0xffffffffu, mx != mx);
}

template <bool SLOW>
__device__ __forceinline__ void pf_outputs(const float (&m9p)[PW_R][4], const PFPlane& cur, const PFPlane& nxt,
                                           uint32_t rowmask, int lane, int mywoff, int w32, uint32_t* __restrict__ bz) {
#pragma unroll
    for (int i = 0; i < PW_R; ++i) {
        if (!((rowmask >> i) & 1u)) continue;               // warp-uniform: row y0 + i is outside the map
        uint32_t nib = 0u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float left = e == 0 ? cur.le[i] : cur.k[i][e - 1], right = e == 3 ? cur.re[i] : cur.k[i][e + 1];
            const float kb = fmax3n(m9p[i][e], cur.m3[i][e], left);
            const float ka = fmax3n(nxt.m9[i][e], cur.m3[i + 2][e], right);
            const float v = cur.k[i][e];
            bool peak = (kb < v) & (ka <= v);
            if (SLOW) peak = (v != v) ? !(ka != ka) : peak;
            nib |= peak ? (1u << e) : 0u;
        }
        uint32_t w = nib << (4 * (lane & 7));               // 8 lanes = one 32-voxel word of the mask
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        w |= __shfl_xor_sync(0xffffffffu, w, 4);
        if ((lane & 7) == 0) bz[mywoff + i * w32] = w;
    }
}

// grid (ctas per map, BA); W % 128 == 0, maps 16-byte aligned, H * W < 2^31 (checked by the launcher).
template <int MINB>
__global__ void __launch_bounds__(PW_THREADS, MINB)
peaks_scan3f_kernel(const float* __restrict__ in, int S, int H, int W, int nstrips, int nrowg, int nzc, int tz, PeakWs ws) {
    const int ba = blockIdx.y;
    const size_t HW = (size_t)H * W;
    const long long V = (long long)S * (long long)HW;
    const float* vol = in + (size_t)ba * V;
    uint32_t* bits = ws.bits + (size_t)ba * ws.words;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int w32 = W >> 5;
    const size_t HW32 = HW >> 5;
    const int nunits = nstrips * nrowg * nzc;
    for (int unit = blockIdx.x * PW_NW + warp; unit < nunits; unit += gridDim.x * PW_NW) {
        const int rowg = unit % nrowg;
        const int t = unit / nrowg;
        const int strip = t % nstrips, zc = t / nstrips;
        const int x0 = strip * 128, y0 = rowg * PW_R, z0 = zc * tz, z1 = min(S, z0 + tz);
        const bool h_ok = (lane == 0 && x0 > 0) || (lane == 31 && x0 + 128 < W);
        const int hoff = lane == 0 ? -1 : 4;
        int roff[PW_R + 2];
        uint32_t rows = 0;                                  // bit j: row y0 - 1 + j is inside the map
#pragma unroll
        for (int j = 0; j < PW_R + 2; ++j) {
            const int gy = y0 - 1 + j;
            roff[j] = gy * W + x0 + 4 * lane;
            if (gy >= 0 && gy < H) rows |= 1u << j;
        }
        const bool all_rows = rows == (1u << (PW_R + 2)) - 1u;
        const uint32_t rowmask = rows >> 1;                 // bit i: centre row y0 + i
        const int mywoff = ((y0 * W + x0) >> 5) + (lane >> 3);

        PFRaw rw;
        PFPlane pa, pb;
        float m9p[PW_R][4];
        uint32_t nanbits;                                   // bit 0: plane z - 1, bit 1: plane z, bit 2: plane z + 1 holds a NaN
        pf_load(rw, vol + (ptrdiff_t)(z0 - 1) * (ptrdiff_t)HW, z0 >= 1, roff, rows, all_rows, h_ok, hoff);
        nanbits = pf_process(pa, rw, lane) ? 1u : 0u;
#pragma unroll
        for (int i = 0; i < PW_R; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) m9p[i][e] = pa.m9[i][e];
        pf_load(rw, vol + (size_t)z0 * HW, true, roff, rows, all_rows, h_ok, hoff);
        nanbits |= pf_process(pa, rw, lane) ? 2u : 0u;
        pf_load(rw, vol + (size_t)(z0 + 1) * HW, z0 + 1 < S, roff, rows, all_rows, h_ok, hoff);
        nanbits |= pf_process(pb, rw, lane) ? 4u : 0u;
        int z = z0;
        const float* pz = vol + (size_t)(z0 + 2) * HW;     // plane z + 2
        uint32_t* bz = bits + (size_t)z0 * HW32;            // first mask word of plane z
        while (true) {                                      // two planes per trip: pa / pb swap roles
            const bool more = z + 1 < z1;                   // plane z + 2 is needed (as the next plane of plane z + 1)
            if (more) pf_load(rw, pz, z + 2 < S, roff, rows, all_rows, h_ok, hoff);
            if (nanbits) pf_outputs<true>(m9p, pa, pb, rowmask, lane, mywoff, w32, bz);     // warp-uniform
            else pf_outputs<false>(m9p, pa, pb, rowmask, lane, mywoff, w32, bz);
            if (!more) break;
#pragma unroll
            for (int i = 0; i < PW_R; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) m9p[i][e] = pa.m9[i][e];
            nanbits = (nanbits >> 1) | (pf_process(pa, rw, lane) ? 4u : 0u);
            ++z; pz += HW; bz += HW32;
            const bool more2 = z + 1 < z1;
            if (more2) pf_load(rw, pz, z + 2 < S, roff, rows, all_rows, h_ok, hoff);
            if (nanbits) pf_outputs<true>(m9p, pb, pa, rowmask, lane, mywoff, w32, bz);
            else pf_outputs<false>(m9p, pb, pa, rowmask, lane, mywoff, w32, bz);
            if (!more2) break;
#pragma unroll
            for (int i = 0; i < PW_R; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) m9p[i][e] = pb.m9[i][e];
            nanbits = (nanbits >> 1) | (pf_process(pb, rw, lane) ? 4u : 0u);
            ++z; pz += HW; bz += HW32;
        }
    }
}

// LEVEL 2: bins = key bits 19..8 of elements whose top 12 bits match the level-1 bin.
// LEVEL 3: bins = key bits 7..0 of elements whose top 24 bits match.  The prefix and the remaining rank come from
// ws.sel (written by the last CTA of the previous level); the last CTA of this level writes the next one, and at
// level 3 the final threshold (NaN when the map holds a NaN: torch.median propagates it).
template <int LEVEL>
__global__ void __launch_bounds__(PK_RF_THREADS)
peaks_refine_kernel(const float* __restrict__ in, long long V, PeakWs ws, float* __restrict__ thr_out) {
    constexpr int NB = LEVEL == 2 ? PK_BINS1 : 256;
    __shared__ uint32_t s_hist[NB];
    __shared__ unsigned long long s_tmp[PK_RF_THREADS / 32 + 2];
    __shared__ int s_flag;
    const int ba = blockIdx.y;
    const int tid = threadIdx.x;
    if (ws.ticket[ba * 4 + 1]) return;                      // list mode: the collect kernel has produced this map's threshold
    const uint32_t prefix = (uint32_t)ws.sel[ba * 2];
    constexpr int shift = LEVEL == 2 ? 20 : 8;
    for (int i = tid; i < NB; i += PK_RF_THREADS) s_hist[i] = 0u;
    __syncthreads();
    const float* vol = in + (size_t)ba * V;
    const bool vec = ((((uintptr_t)vol) & 15) == 0);
    const long long nvec = vec ? (V >> 2) : 0;
    const uint32_t hist_addr = (uint32_t)__cvta_generic_to_shared(s_hist);
    auto add = [&](float f) {                              // predicated shared-memory reduction, no branch
        const uint32_t key = fkey(f);
        const uint32_t a = hist_addr + (LEVEL == 2 ? ((key >> 6) & 0x3FFCu) : ((key << 2) & 0x3FCu));
        asm volatile("{ .reg .pred p; setp.eq.u32 p, %0, %1; @p red.shared.add.u32 [%2], %3; }"
                     :: "r"(key >> shift), "r"(prefix), "r"(a), "r"(1u) : "memory");
    };
    {
        const long long stride = (long long)gridDim.x * PK_RF_THREADS;
        long long i = (long long)blockIdx.x * PK_RF_THREADS + tid;
        for (; i + 3 * stride < nvec; i += 4 * stride) {           // 4 independent 128-bit loads in flight
            uint4 u[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) u[j] = ld_stream_u4(vol + ((i + j * stride) << 2));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                add(__uint_as_float(u[j].x)); add(__uint_as_float(u[j].y)); add(__uint_as_float(u[j].z)); add(__uint_as_float(u[j].w));
            }
        }
        for (; i < nvec; i += stride) {
            const uint4 u = ld_stream_u4(vol + (i << 2));
            add(__uint_as_float(u.x)); add(__uint_as_float(u.y)); add(__uint_as_float(u.z)); add(__uint_as_float(u.w));
        }
    }
    for (long long i = (nvec << 2) + (long long)blockIdx.x * PK_RF_THREADS + tid; i < V; i += (long long)gridDim.x * PK_RF_THREADS)
        add(vol[i]);
    uint32_t* gh = (LEVEL == 2 ? ws.hist2 + (size_t)ba * PK_BINS1 : ws.hist3 + (size_t)ba * 256);
    if (!flush_hist_and_vote<PK_RF_THREADS>(s_hist, NB, gh, ws.ticket + ba * 4 + LEVEL, &s_flag)) return;
    int bin; unsigned long long kk;
    select_bin<PK_RF_THREADS>(gh, NB, ws.sel[ba * 2 + 1], s_tmp, &bin, &kk);
    if (tid == 0) {
        if (LEVEL == 2) {
            ws.sel[ba * 2] = ((unsigned long long)prefix << 12) | (unsigned long long)bin;
            ws.sel[ba * 2 + 1] = kk;
        } else {
            const uint32_t key = (prefix << 8) | (uint32_t)bin;
            const uint32_t nan_cnt = __ldcg(ws.hist1 + (size_t)ba * PK_BINS1 + (PK_BINS1 - 1));   // only NaN keys reach the last bin
            const float thr = nan_cnt ? CUDART_NAN_F : key_to_float(key);
            ws.thr[ba] = thr;
            if (thr_out) thr_out[ba] = thr;
        }
    }
}

// ---- list mode of the exact median (the common case) -------------------------------------------------------------
// After the scan the level-1 bin of the median and its population are known.  When the population fits the list
// (a quarter of the map, at most 131072 entries: one CTA finishes it), ONE more pass over the (L2-resident) map gathers the low 20 key bits of the bin's members:
// matches are staged in shared memory (slot = shared-memory atomic) and appended to the map's list with one global
// atomic per flush.  The last CTA of the map then finishes the selection on the list alone: 12-bit histogram, bin,
// 8-bit histogram, bin -> threshold.  Maps whose bin is too crowded are left to the two full refinement passes
// (peaks_refine_kernel<2>, <3>), which return at once for list-mode maps.
constexpr int PK_STAGE = 16 * PK_RF_THREADS;          // worst case of one round: 4 x 128-bit loads per thread, all matching

__device__ __forceinline__ bool list_mode(const PeakWs& ws, int ba) {
    const uint32_t bin1 = (uint32_t)ws.sel[ba * 2];
    return (long long)__ldcg(ws.hist1 + (size_t)ba * PK_BINS1 + bin1) <= ws.list_cap;
}

__global__ void __launch_bounds__(PK_RF_THREADS)
peaks_collect_kernel(const float* __restrict__ in, long long V, PeakWs ws, float* __restrict__ thr_out) {
    __shared__ uint32_t s_stage[PK_STAGE];                 // 32 KB; reused as the histograms of the finishing step
    __shared__ unsigned long long s_tmp[PK_RF_THREADS / 32 + 2];
    __shared__ uint32_t s_n, s_base;
    __shared__ int s_flag;
    uint32_t* s_hist = s_stage;
    const int ba = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t prefix = (uint32_t)ws.sel[ba * 2];
    const float* vol = in + (size_t)ba * V;
    const bool vec = ((((uintptr_t)vol) & 15) == 0);
    const long long nvec = vec ? (V >> 2) : 0;
    const long long stride = (long long)gridDim.x * PK_RF_THREADS;
    if (!list_mode(ws, ba)) {
        // ---- crowded level-1 bin (uniform per map): this pass is the level-2 histogram (key bits 19..8); the last CTA
        // selects the level-2 bin and peaks_refine_kernel<3> finishes with one more pass over the map
        for (int i = tid; i < PK_BINS1; i += PK_RF_THREADS) s_hist[i] = 0u;
        __syncthreads();
        const uint32_t hist_addr = (uint32_t)__cvta_generic_to_shared(s_hist);
        auto add2 = [&](float f) {
            const uint32_t key = fkey(f);
            if ((key >> 20) == prefix) atomicAdd(&s_hist[(key >> 8) & 0xFFFu], 1u);
        };
        (void)hist_addr;
        for (long long i = (long long)blockIdx.x * PK_RF_THREADS + tid; i < nvec; i += stride) {
            const uint4 u = ld_stream_u4(vol + (i << 2));
            add2(__uint_as_float(u.x)); add2(__uint_as_float(u.y)); add2(__uint_as_float(u.z)); add2(__uint_as_float(u.w));
        }
        for (long long i = (nvec << 2) + (long long)blockIdx.x * PK_RF_THREADS + tid; i < V; i += stride) add2(vol[i]);
        uint32_t* gh = ws.hist2 + (size_t)ba * PK_BINS1;
        if (!flush_hist_and_vote<PK_RF_THREADS>(s_hist, PK_BINS1, gh, ws.ticket + ba * 4 + 2, &s_flag)) return;
        int bin; unsigned long long kk;
        select_bin<PK_RF_THREADS>(gh, PK_BINS1, ws.sel[ba * 2 + 1], s_tmp, &bin, &kk);
        if (tid == 0) {
            ws.sel[ba * 2] = ((unsigned long long)prefix << 12) | (unsigned long long)bin;
            ws.sel[ba * 2 + 1] = kk;
        }
        return;
    }
    uint32_t* list = ws.list + (size_t)ba * (size_t)ws.list_cap;
    if (tid == 0) s_n = 0u;
    __syncthreads();
    // One round = 4 x 128-bit loads per thread.  A thread first marks its matches among the 16 values, the warp agrees on
    // slots with one prefix sum and ONE shared-memory atomic per round, then the matches are written to the stage.
    auto round16 = [&](const uint32_t (&val)[16], uint32_t okmask) {
        uint32_t mm = 0u;
#pragma unroll
        for (int e = 0; e < 16; ++e) mm |= ((fkey(__uint_as_float(val[e])) >> 20) == prefix) ? (1u << e) : 0u;
        mm &= okmask;
        const uint32_t c = __popc(mm);
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t a = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += a; }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total) {                                        // warp-uniform
            uint32_t base = 0;
            if (lane == 31) base = atomicAdd(&s_n, total);
            uint32_t pos = __shfl_sync(0xffffffffu, base, 31) + incl - c;
#pragma unroll
            for (int e = 0; e < 16; ++e)
                if ((mm >> e) & 1u) s_stage[pos++] = fkey(__uint_as_float(val[e])) & 0xFFFFFu;
        }
    };
    auto flush = [&]() {                                    // called by all threads: stage -> the map's list
        __syncthreads();
        const uint32_t n = s_n;
        if (n) {                                            // uniform
            if (tid == 0) s_base = atomicAdd(ws.list_n + ba, n);
            __syncthreads();
            const uint32_t base = s_base;
            for (uint32_t i = tid; i < n; i += PK_RF_THREADS) list[base + i] = s_stage[i];
            __syncthreads();
            if (tid == 0) s_n = 0u;
            __syncthreads();
        }
    };
    {
        const long long rounds = (nvec + 4 * stride - 1) / (4 * stride);
        long long i = (long long)blockIdx.x * PK_RF_THREADS + tid;
        for (long long r = 0; r < rounds; ++r, i += 4 * stride) {           // 4 independent 128-bit loads in flight
            uint32_t val[16];
            uint32_t okmask = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long idx = i + j * stride;
                uint4 u = make_uint4(0u, 0u, 0u, 0u);
                if (idx < nvec) { u = ld_stream_u4(vol + (idx << 2)); okmask |= 0xFu << (4 * j); }
                val[4 * j] = u.x; val[4 * j + 1] = u.y; val[4 * j + 2] = u.z; val[4 * j + 3] = u.w;
            }
            round16(val, okmask);
            flush();
        }
    }
    for (long long i0 = (nvec << 2) + (long long)blockIdx.x * PK_RF_THREADS; i0 < V; i0 += stride * 16) {     // unaligned maps / tails
        uint32_t val[16];
        uint32_t okmask = 0u;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const long long i = i0 + e * stride + tid;
            val[e] = 0u;
            if (i < V) { val[e] = __float_as_uint(vol[i]); okmask |= 1u << e; }
        }
        round16(val, okmask);
        flush();
    }
    // ---- last CTA of the map: finish the selection on the list ---------------------------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) { s_flag = (atomicAdd(ws.ticket + ba * 4 + 1, 1u) == gridDim.x - 1); __threadfence(); }
    __syncthreads();
    if (!s_flag) return;
    const uint32_t n = __ldcg(ws.list_n + ba);
    __syncthreads();
    for (int i = tid; i < PK_BINS1; i += PK_RF_THREADS) s_hist[i] = 0u;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += PK_RF_THREADS) atomicAdd(&s_hist[__ldcg(list + i) >> 8], 1u);
    __syncthreads();
    int bin2; unsigned long long k2;
    select_bin_smem<PK_RF_THREADS>(s_hist, PK_BINS1, ws.sel[ba * 2 + 1], s_tmp, &bin2, &k2);
    for (int i = tid; i < 256; i += PK_RF_THREADS) s_hist[i] = 0u;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += PK_RF_THREADS) {
        const uint32_t lo = __ldcg(list + i);
        if ((lo >> 8) == (uint32_t)bin2) atomicAdd(&s_hist[lo & 0xFFu], 1u);
    }
    __syncthreads();
    int bin3; unsigned long long k3;
    select_bin_smem<PK_RF_THREADS>(s_hist, 256, k2, s_tmp, &bin3, &k3);
    if (tid == 0) {
        const uint32_t key = (prefix << 20) | ((uint32_t)bin2 << 8) | (uint32_t)bin3;
        const uint32_t nan_cnt = __ldcg(ws.hist1 + (size_t)ba * PK_BINS1 + (PK_BINS1 - 1));   // only NaN keys reach the last bin
        const float thr = nan_cnt ? CUDART_NAN_F : key_to_float(key);
        ws.thr[ba] = thr;
        if (thr_out) thr_out[ba] = thr;
    }
}

// ---- exact median by sampled interval (round 2; maps of 2^16 .. 2^21 elements) -----------------------------------------
// peaks_sample_kernel: one CTA per map sorts a stratified sample of PK_NS keys and publishes the sample order statistics
//   PK_ND ranks (5 sigma of the sample median's rank) below / above the middle as the interval [lo, hi]; it holds the map's
//   median unless the sample is wildly unrepresentative (probability ~3e-7 per map for any distribution), and about 8 % of the map.
//   Values that repeat in the sample next to lo / hi mark heavy ties (flag bits): their copies are then counted instead of listed.
// peaks_interval_kernel: ONE streaming pass over the map counts the keys below lo exactly and lists the keys inside the
//   interval (per-warp shared-memory stages, one global atomic per flush, no block barrier in the loop).  The last CTA of the
//   map verifies count_below <= rank < count_below + count_inside -- which makes the result exact, not probabilistic -- and
//   selects the median from the list (radix select on key - lo); if the check fails, or the list overflowed, the same CTA
//   falls back to a full three-level radix select over the map (slow, never seen on continuous data).
__device__ unsigned long long pk_fallbacks;           // maps whose sampled interval missed the median (b200seg_peaks3d_fallback_count)
constexpr int PK_NS = 4096, PK_ND = 160;
constexpr int PK_IV_THREADS = 512, PK_IV_NW = PK_IV_THREADS / 32;
constexpr int PK_IV_WSTAGE = 1024;                      // list entries a warp stages before it flushes (a round adds at most 512)
constexpr uint32_t PK_TIE_LO = 1u, PK_TIE_HI = 2u, PK_FORCE_FAIL = 4u;

struct PeakIv { uint32_t lo, hi, flags, pad; };
struct PeakIvCnt { uint32_t n_lt, n_eqlo, n_eqhi, has_nan; };

__global__ void __launch_bounds__(512)
peaks_sample_kernel(const float* __restrict__ in, long long V, PeakWs ws, int force_fail) {
    __shared__ uint32_t s_h[2][PK_BINS1];
    __shared__ unsigned long long s_tmp[512 / 32 + 2];
    const int ba = blockIdx.x, tid = threadIdx.x;
    const float* vol = in + (size_t)ba * V;
    const uint32_t st = (uint32_t)(V / PK_NS);              // stratum length (V >= 16 * PK_NS)
    uint32_t key[PK_NS / 512];
#pragma unroll
    for (int u = 0; u < PK_NS / 512; ++u) {                 // independent gathers: one element per stratum
        const int i = tid + u * 512;
        const uint32_t h = ((uint32_t)i * 2654435761u) >> 7;    // fixed pseudo-random offset inside the stratum
        key[u] = fkey(__ldg(vol + (size_t)i * st + (h % st)));
    }
    // three radix levels (12 + 12 + 8 bits) for the two ranks at once: exact order statistics and their multiplicity
    uint32_t pre[2] = {0u, 0u};
    unsigned long long rk[2] = {(unsigned long long)(PK_NS / 2 - PK_ND), (unsigned long long)(PK_NS / 2 + PK_ND)};
    uint32_t mult[2] = {0u, 0u};
#pragma unroll
    for (int level = 0; level < 3; ++level) {
        const int nb = level < 2 ? PK_BINS1 : 256;
        const int sh_pre = level == 0 ? 32 : (level == 1 ? 20 : 8), sh_bin = level == 0 ? 20 : (level == 1 ? 8 : 0);
        const bool same = level > 0 && pre[0] == pre[1];    // both ranks still in one bin: one histogram serves both
        __syncthreads();
        for (int i = tid; i < nb; i += 512) { s_h[0][i] = 0u; s_h[1][i] = 0u; }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < PK_NS / 512; ++u) {
            const uint32_t k = key[u], bin = (k >> sh_bin) & (uint32_t)(nb - 1);
            if (level == 0) atomicAdd(&s_h[0][bin], 1u);
            else {
                if ((k >> sh_pre) == pre[0]) atomicAdd(&s_h[0][bin], 1u);
                if (!same && (k >> sh_pre) == pre[1]) atomicAdd(&s_h[1][bin], 1u);
            }
        }
        __syncthreads();
        const int which1 = (level == 0 || same) ? 0 : 1;
        int b0, b1; unsigned long long r0, r1;
        select_bin_smem<512>(s_h[0], nb, rk[0], s_tmp, &b0, &r0);
        select_bin_smem<512>(s_h[which1], nb, rk[1], s_tmp, &b1, &r1);
        mult[0] = s_h[0][b0]; mult[1] = s_h[which1][b1];
        pre[0] = (level == 0 ? 0u : pre[0] << (level == 1 ? 12 : 8)) | (uint32_t)b0;
        pre[1] = (level == 0 ? 0u : pre[1] << (level == 1 ? 12 : 8)) | (uint32_t)b1;
        rk[0] = r0; rk[1] = r1;
    }
    if (tid == 0) {
        PeakIv iv;
        iv.lo = pre[0]; iv.hi = pre[1];
        // a value that repeats inside a 4096-element sample is a heavy tie of the map
        iv.flags = (mult[0] >= 2u ? PK_TIE_LO : 0u) | (mult[1] >= 2u ? PK_TIE_HI : 0u) | (force_fail ? PK_FORCE_FAIL : 0u);
        iv.pad = 0u;
        ws.iv[ba] = iv;
    }
}

// Radix select of the element of ascending rank `rank` among get(0..n-1) (uint32 values below 2^top_bits), 12 bits per
// level, by ONE CTA; s_hist has PK_BINS1 entries.  Must be called by all NT threads.
template <int NT, typename F>
__device__ uint32_t cta_select_u32(F get, long long n, unsigned long long rank, int top_bits, uint32_t* s_hist, unsigned long long* s_tmp) {
    uint32_t prefix = 0u;
    int shift = top_bits;                                   // candidates: x >> shift == prefix (everything while shift == 32)
    while (shift > 0) {
        const int nb = shift < 12 ? shift : 12, s2 = shift - nb;
        __syncthreads();
        for (int i = threadIdx.x; i < (1 << nb); i += NT) s_hist[i] = 0u;
        __syncthreads();
        {
            long long i = threadIdx.x;
            for (; i + 7ll * NT < n; i += 8ll * NT) {           // eight independent loads in flight per thread
                uint32_t x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = get(i + (long long)u * NT);
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (shift >= 32 || (x[u] >> shift) == prefix) atomicAdd(&s_hist[(x[u] >> s2) & ((1u << nb) - 1u)], 1u);
            }
            for (; i < n; i += NT) {
                const uint32_t x = get(i);
                if (shift >= 32 || (x >> shift) == prefix) atomicAdd(&s_hist[(x >> s2) & ((1u << nb) - 1u)], 1u);
            }
        }
        __syncthreads();
        int bin; unsigned long long r2;
        select_bin_smem<NT>(s_hist, 1 << nb, rank, s_tmp, &bin, &r2);
        prefix = (shift >= 32 ? 0u : (prefix << nb)) | (uint32_t)bin;
        rank = r2;
        shift = s2;
    }
    return prefix;
}

// One streaming pass over (this CTA's share of) a map.  Matches are queued per LANE (slot s of lane l lives at
// s_stage_w[s * 32 + l]: conflict free, no prefix sum and no key recomputation per match); a warp flushes its 32 queues
// with one prefix sum and one global atomic when some lane could overflow in the next round.
constexpr int PK_IV_QCAP = PK_IV_WSTAGE / 32;           // 32 slots per lane

template <bool TIES>
__device__ __forceinline__ void interval_stream(const float* __restrict__ vol, long long V, uint32_t lo, uint32_t hi, uint32_t flags,
                                                uint32_t* s_stage_w, uint32_t* __restrict__ list, uint32_t* __restrict__ list_n,
                                                long long list_cap, uint32_t& n_lt, uint32_t& n_eqlo, uint32_t& n_eqhi, uint32_t& mx) {
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t w = hi - lo;
    const bool tlo = TIES && (flags & PK_TIE_LO), thi = TIES && (flags & PK_TIE_HI) && w != 0u;
    const bool vec = ((((uintptr_t)vol) & 15) == 0);
    const long long nvec = vec ? (V >> 2) : 0;
    const long long stride = (long long)gridDim.x * PK_IV_THREADS;
    uint32_t* q = s_stage_w + lane;                         // next free slot of my queue
    uint32_t qn = 0;
    auto flush = [&]() {
        uint32_t incl = qn;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t a = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += a; }
        uint32_t base = 0;
        if (lane == 31) base = atomicAdd(list_n, incl);
        base = __shfl_sync(0xffffffffu, base, 31) + incl - qn;
        for (uint32_t s2 = 0; s2 < qn; ++s2)
            if ((long long)(base + s2) < list_cap) list[base + s2] = s_stage_w[s2 * 32 + lane];
        qn = 0;
        q = s_stage_w + lane;
    };
    auto element = [&](uint32_t raw) {
        const uint32_t key = fkey(__uint_as_float(raw));
        const uint32_t t = key - lo;
        n_lt += key < lo ? 1u : 0u;
        bool in = t <= w;                                   // lo <= key <= hi (a key below lo wraps far above w)
        if (TIES) {
            if (tlo) { n_eqlo += t == 0u ? 1u : 0u; in = in && t != 0u; }
            if (thi) { n_eqhi += t == w ? 1u : 0u; in = in && t != w; }
        }
        mx = max(mx, key);
        if (in) { *q = key; q += 32; ++qn; }
    };
    {
        // loop conditions are evaluated on the warp's first index so that every lane makes the same trips (the warp votes inside)
        long long ib = (long long)blockIdx.x * PK_IV_THREADS + (tid & ~31);
        for (; ib + 31 + 3 * stride < nvec; ib += 4 * stride) {            // full rounds: 4 independent 128-bit loads in flight
            const long long i = ib + lane;
            uint4 u[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) u[j] = ld_stream_u4(vol + ((i + j * stride) << 2));
#pragma unroll
            for (int j = 0; j < 4; ++j) { element(u[j].x); element(u[j].y); element(u[j].z); element(u[j].w); }
            if (__any_sync(0xffffffffu, qn > (uint32_t)(PK_IV_QCAP - 16))) flush();
        }
        for (; ib < nvec; ib += stride) {
            const long long i = ib + lane;
            if (i < nvec) {
                const uint4 u = ld_stream_u4(vol + (i << 2));
                element(u.x); element(u.y); element(u.z); element(u.w);
            }
            if (__any_sync(0xffffffffu, qn > (uint32_t)(PK_IV_QCAP - 16))) flush();
        }
    }
    // unaligned maps / tails
    for (long long ib = (nvec << 2) + (long long)blockIdx.x * PK_IV_THREADS + (tid & ~31); ib < V; ib += stride) {
        const long long i = ib + lane;
        if (i < V) element(__float_as_uint(vol[i]));
        if (__any_sync(0xffffffffu, qn > (uint32_t)(PK_IV_QCAP - 16))) flush();
    }
    if (__any_sync(0xffffffffu, qn > 0u)) flush();
}

__global__ void __launch_bounds__(PK_IV_THREADS, 2)
peaks_interval_kernel(const float* __restrict__ in, long long V, PeakWs ws, float* __restrict__ thr_out) {
    extern __shared__ __align__(16) uint32_t s_dyn[];       // PK_IV_NW warp stages; reused as the histogram of the finishing step
    __shared__ unsigned long long s_tmp[PK_IV_THREADS / 32 + 2];
    __shared__ uint32_t s_red[4][PK_IV_NW];
    __shared__ int s_flag;
    const int ba = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* vol = in + (size_t)ba * V;
    const PeakIv iv = ws.iv[ba];
    uint32_t* list = ws.list + (size_t)ba * (size_t)ws.list_cap;
    uint32_t n_lt = 0, n_eqlo = 0, n_eqhi = 0, mx = 0;
    if (iv.flags & (PK_TIE_LO | PK_TIE_HI))
        interval_stream<true>(vol, V, iv.lo, iv.hi, iv.flags, s_dyn + warp * PK_IV_WSTAGE, list, ws.list_n + ba, ws.list_cap, n_lt, n_eqlo, n_eqhi, mx);
    else
        interval_stream<false>(vol, V, iv.lo, iv.hi, iv.flags, s_dyn + warp * PK_IV_WSTAGE, list, ws.list_n + ba, ws.list_cap, n_lt, n_eqlo, n_eqhi, mx);
    // CTA totals -> the map's counters
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        n_lt += __shfl_xor_sync(0xffffffffu, n_lt, o); n_eqlo += __shfl_xor_sync(0xffffffffu, n_eqlo, o);
        n_eqhi += __shfl_xor_sync(0xffffffffu, n_eqhi, o); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { s_red[0][warp] = n_lt; s_red[1][warp] = n_eqlo; s_red[2][warp] = n_eqhi; s_red[3][warp] = mx; }
    __syncthreads();
    PeakIvCnt* cnt = ws.ivcnt + ba;
    if (tid == 0) {
        uint32_t a = 0, b = 0, c = 0, m = 0;
        for (int i = 0; i < PK_IV_NW; ++i) { a += s_red[0][i]; b += s_red[1][i]; c += s_red[2][i]; m = max(m, s_red[3][i]); }
        if (a) atomicAdd(&cnt->n_lt, a);
        if (b) atomicAdd(&cnt->n_eqlo, b);
        if (c) atomicAdd(&cnt->n_eqhi, c);
        if (m == KEY_NAN) atomicOr(&cnt->has_nan, 1u);
        __threadfence();
        s_flag = (atomicAdd(ws.ticket + ba * 4 + 1, 1u) == gridDim.x - 1);
        __threadfence();
    }
    __syncthreads();
    if (!s_flag) return;
    // ---- last CTA of the map: verify the interval and finish the selection ---------------------------------------
    uint32_t* s_hist = s_dyn;
    const unsigned long long k = (unsigned long long)((V - 1) / 2);
    const unsigned long long lt = __ldcg(&cnt->n_lt), eqlo = __ldcg(&cnt->n_eqlo), eqhi = __ldcg(&cnt->n_eqhi);
    const unsigned long long nl = __ldcg(ws.list_n + ba);
    const bool has_nan = __ldcg(&cnt->has_nan) != 0u;
    uint32_t key = 0u;
    int how;                                                // 0: lo, 1: from the list, 2: hi, 3: fall back
    if (iv.flags & PK_FORCE_FAIL) how = 3;
    else if (k < lt) how = 3;
    else if (k < lt + eqlo) how = 0;
    else if (k < lt + eqlo + nl) how = (long long)nl <= ws.list_cap ? 1 : 3;
    else if (k < lt + eqlo + nl + eqhi) how = 2;
    else how = 3;
    if (has_nan) how = 0;                                   // the threshold is NaN whatever the interval says
    if (how == 0) key = iv.lo;
    else if (how == 2) key = iv.hi;
    else if (how == 1) {
        const uint32_t lo = iv.lo, w = iv.hi - iv.lo;
        const int bits_w = 32 - __clz(w | 1u);
        key = lo + cta_select_u32<PK_IV_THREADS>([&](long long i) { return __ldcg(list + i) - lo; }, (long long)nl, k - lt - eqlo,
                                                 bits_w, s_hist, s_tmp);
    } else {
        key = cta_select_u32<PK_IV_THREADS>([&](long long i) { return fkey(__ldg(vol + i)); }, V, k, 32, s_hist, s_tmp);
        if (tid == 0) atomicAdd(&pk_fallbacks, 1ull);       // statistics: maps that took the fallback
    }
    if (tid == 0) {
        const float thr = has_nan ? CUDART_NAN_F : key_to_float(key);
        ws.thr[ba] = thr;
        if (thr_out) thr_out[ba] = thr;
    }
}

// One launch for filter + count + scan + emit.  Work item = PK_ITEM_WORDS consecutive words of one map's candidate
// mask (4 per thread), items ordered (map, chunk) = the lexicographic order of the output rows.  A CTA draws its item
// by ticket, so every predecessor of a running item is running or finished; an item publishes its count as soon as
// it is known and obtains its output offset by summing the counts of ALL its predecessors (a few hundred words
// read by the whole CTA in parallel) -- no chain of dependent look-backs.
// filter_mode 1: threshold = ws.thr (median, written by the refine<3> kernel); 2: thr_in; 0: no filter.
constexpr int PK_PER = 4;
constexpr int PK_ITEM_WORDS = PK_PER * PK_THREADS;
constexpr unsigned long long LB_READY = 1ull << 63, LB_MASK = (1ull << 63) - 1;

__global__ void __launch_bounds__(PK_THREADS)
peaks_finalize_kernel(const float* __restrict__ in, int BA, int A, int H, int W, long long V, int filter_mode,
                      const float* __restrict__ thr_in, PeakWs ws, int64_t* __restrict__ peaks, int cap,
                      int32_t* __restrict__ n_peaks, float* __restrict__ agg, float* __restrict__ thr_out) {
    __shared__ int s_item, s_last;
    __shared__ uint32_t s_cnt[PK_NW];
    __shared__ float s_sum[PK_NW];
    __shared__ unsigned long long s_part[PK_NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nitems = BA * ws.nchunks;
    if (tid == 0) s_item = (int)atomicAdd(ws.gticket, 1u);
    __syncthreads();
    const int item = s_item;
    const int ba = item / ws.nchunks, chunk = item - ba * ws.nchunks;
    float thr = 0.f;
    if (filter_mode == 1) thr = ws.thr[ba];
    else if (filter_mode == 2) thr = thr_in[ba];
    if (filter_mode != 1 && chunk == 0 && tid == 0) { ws.thr[ba] = thr; if (thr_out) thr_out[ba] = thr; }

    const float* vol = in + (size_t)ba * V;
    const uint32_t* bits = ws.bits + (size_t)ba * ws.words;
    const int w0 = chunk * PK_ITEM_WORDS + tid * PK_PER;   // a thread owns PK_PER consecutive words
    uint32_t g[PK_PER], keep[PK_PER];
    float sm[PK_PER];
#pragma unroll
    for (int j = 0; j < PK_PER; ++j) { g[j] = (w0 + j) < ws.words ? __ldg(bits + w0 + j) : 0u; keep[j] = 0u; sm[j] = 0.f; }
    while (g[0] | g[1] | g[2] | g[3]) {                     // one candidate of every word per trip: 4 independent loads in flight
        float v[PK_PER];
        int bit[PK_PER];
#pragma unroll
        for (int j = 0; j < PK_PER; ++j) {
            bit[j] = __ffs(g[j]) - 1;
            v[j] = 0.f;
            if (g[j]) { v[j] = __ldg(vol + (long long)(w0 + j) * 32 + bit[j]); g[j] &= g[j] - 1; }
        }
#pragma unroll
        for (int j = 0; j < PK_PER; ++j)
            if (bit[j] >= 0 && (filter_mode == 0 || v[j] >= thr)) { keep[j] |= 1u << bit[j]; sm[j] += v[j]; }
    }
    uint32_t cnt = 0;
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < PK_PER; ++j) { cnt += __popc(keep[j]); sum += sm[j]; }
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t a = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += a; }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);          // deterministic tree
    if (lane == 31) s_cnt[warp] = incl;
    if (lane == 0) s_sum[warp] = sum;
    __syncthreads();
    volatile unsigned long long* look = ws.look;
    volatile unsigned long long* map_tot = ws.map_tot;
    if (tid == 0) {
        uint32_t T = 0; float t = 0.f;
#pragma unroll
        for (int i = 0; i < PK_NW; ++i) { T += s_cnt[i]; t += s_sum[i]; }
        ws.chunk_cnt[item] = T;
        ws.chunk_sum[item] = t;
        look[item] = LB_READY | (unsigned long long)T;     // the count is the whole message: one 64-bit store
        __threadfence();
        if (atomicAdd(ws.map_cnt + ba, 1u) == (uint32_t)(ws.nchunks - 1)) {     // last work item of this map: publish the map's total
            __threadfence();
            unsigned long long tot = 0;
            for (int j = 0; j < ws.nchunks; ++j) tot += look[(size_t)ba * ws.nchunks + j] & LB_MASK;
            map_tot[ba] = LB_READY | tot;
        }
    }
    // output offset = peaks of all earlier maps (one word per map, published by the map's last item) + peaks of the earlier
    // items of this map; a few hundred words read by the whole CTA in parallel, spinning until each has been published
    unsigned long long part = 0;
    for (int j = tid; j < ba + chunk; j += PK_THREADS) {
        volatile unsigned long long* src = j < ba ? map_tot + j : look + ((size_t)ba * ws.nchunks + (j - ba));
        unsigned long long st;
        do { st = *src; } while (!(st & LB_READY));
        part += st & LB_MASK;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_part[warp] = part;
    __syncthreads();
    unsigned long long prefix = 0;
#pragma unroll
    for (int i = 0; i < PK_NW; ++i) prefix += s_part[i];
    if (cnt && cap > 0) {
        unsigned long long pos = prefix + (incl - cnt);
        for (int i = 0; i < warp; ++i) pos += s_cnt[i];
        const int b = ba / A, a = ba - b * A;
        // (z, y, x) of the first voxel of the thread's first word: one division per thread, then increments (a word never
        // leaves its row when W % 32 == 0; in general the carry loops run a few times)
        const long long flat0 = (long long)w0 * 32;
        int x = (int)(flat0 % W);
        const long long t0 = flat0 / W;
        int y = (int)(t0 % H), z = (int)(t0 / H);
#pragma unroll
        for (int j = 0; j < PK_PER; ++j) {
            for (uint32_t q = keep[j]; q;) {
                const int bit = __ffs(q) - 1;
                q &= q - 1;
                if (pos < (unsigned long long)cap) {
                    int xx = x + bit, yy = y, zz = z;
                    while (xx >= W) { xx -= W; if (++yy == H) { yy = 0; ++zz; } }
                    int64_t* row = peaks + pos * 5;
                    row[0] = b; row[1] = a; row[2] = zz; row[3] = yy; row[4] = xx;
                }
                ++pos;
            }
            x += 32;
            while (x >= W) { x -= W; if (++y == H) { y = 0; ++z; } }
        }
    }
    // ---- the last CTA reduces the per-chunk sums per map (fixed order) and publishes the total ----------
    __threadfence();
    __syncthreads();
    if (tid == 0) { s_last = (atomicAdd(ws.gticket + 1, 1u) == (uint32_t)(nitems - 1)); __threadfence(); }
    __syncthreads();
    if (!s_last) return;
    unsigned long long tot = 0;
    for (int m2 = tid; m2 < BA; m2 += PK_THREADS) {
        float s2 = 0.f, c = 0.f;
        for (int jc = 0; jc < ws.nchunks; ++jc) {
            const uint32_t cc = __ldcg(ws.chunk_cnt + (size_t)m2 * ws.nchunks + jc);
            s2 += __ldcg(ws.chunk_sum + (size_t)m2 * ws.nchunks + jc);
            c += (float)cc;
            tot += cc;
        }
        if (agg) agg[m2] = s2 / c;      // 0/0 = NaN when the map has no peak, as in the reference
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    __syncthreads();
    if (lane == 0) s_part[warp] = tot;
    __syncthreads();
    if (tid == 0) {
        unsigned long long total = 0;
        for (int i = 0; i < PK_NW; ++i) total += s_part[i];
        *n_peaks = (int32_t)(total > 0x7FFFFFFFull ? 0x7FFFFFFFull : total);
    }
}

__global__ void peaks_bwd_scatter_kernel(const int64_t* __restrict__ peaks, const int32_t* __restrict__ n_peaks, int cap,
                                         const float* __restrict__ grad_agg, float* __restrict__ grad_in,
                                         int A, int S, int H, int W) {
    const int n = min(*n_peaks, cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int64_t* r = peaks + (size_t)i * 5;
        const size_t ba = (size_t)r[0] * A + r[1];
        grad_in[(ba * S + r[2]) * H * W + r[3] * W + r[4]] = grad_agg[ba];
    }
}

}  // namespace b200seg

using namespace b200seg;

extern "C" size_t b200seg_peaks3d_workspace_bytes(int B, int A, int S, int H, int W) {
    if (B <= 0 || A <= 0 || S <= 0 || H <= 0 || W <= 0) return 256;
    PeakWs w = peak_layout(nullptr, B * A, (long long)S * H * W);
    return w.total_bytes + 256;
}

extern "C" int b200seg_peaks3d_dev(const float* input, int B, int A, int S, int H, int W, int win,
                                   int filter_mode, const float* thr_in, int64_t* peaks, int cap,
                                   int32_t* n_peaks, float* agg, float* thr_out, void* workspace,
                                   size_t workspace_bytes, b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(B >= 0 && A >= 0 && S >= 0 && H >= 0 && W >= 0 && cap >= 0, "peaks3d: negative size");
    B200_CHECK_ARG(win == 3 || win == 5 || win == 7, "peaks3d: win must be 3, 5 or 7 (got %d)", win);
    B200_CHECK_ARG(filter_mode >= 0 && filter_mode <= 2, "peaks3d: bad filter_mode");
    B200_CHECK_ARG(n_peaks, "peaks3d: null n_peaks");
    const long long V = (long long)S * H * W;
    const int BA = B * A;
    if (BA == 0 || V == 0) {
        B200_CUDA(cudaMemsetAsync(n_peaks, 0, 4, stream));
        return 0;
    }
    B200_CHECK_ARG(input && workspace && (peaks || cap == 0), "peaks3d: null pointer");
    B200_CHECK_ARG(filter_mode != 2 || thr_in, "peaks3d: filter_mode 2 needs thr_in");
    B200_CHECK_ARG(BA <= 65535, "peaks3d: B*A too large");
    B200_CHECK_ARG(V < (1ll << 36), "peaks3d: volume too large");
    char* base = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    PeakWs ws = peak_layout(base, BA, V);
    if ((size_t)(base - (char*)workspace) + ws.total_bytes > workspace_bytes) {
        set_error("peaks3d: workspace too small");
        return B200SEG_EWORKSPACE;
    }
    B200_CHECK_ARG((long long)BA * ws.nchunks < (1ll << 31), "peaks3d: too many work items");
    B200_CUDA(cudaMemsetAsync(base, 0, ws.zero_bytes, stream));
    const int sms = num_sms();
    // exact median: sampled interval + one streaming pass for maps of 2^16 .. 2^21 elements, else the level-1 histogram of
    // the scan + list / refinement passes (option "peaks_median_mode": 0 auto, 1 always the histogram path, 2 auto with the
    // interval forced to miss, which exercises the fallback selection)
    const int median_mode = opt_peaks_median_mode();
    const bool interval = filter_mode == 1 && median_mode != 1 && V >= 16ll * PK_NS && V <= (1ll << 21);
    const int do_hist = filter_mode == 1 && !interval;
    const int stop_after = opt_peaks_stop_after();        // profiling aid (b200seg_set_option): run only the first k kernels
    if (stop_after < 1) return 0;
    if (win == 3) {
        static const int variant = getenv("B200SEG_PEAKS_VARIANT") ? atoi(getenv("B200SEG_PEAKS_VARIANT")) : 0;
        B200_CHECK_ARG((long long)H * W < (1ll << 31) - 256, "peaks3d: H * W too large");
        const bool wide = variant < 10 && (W % 128) == 0 && (((uintptr_t)input) & 15) == 0;
        // big batches (many maps per call) run a little faster at 4 resident CTAs per SM despite a few spilled registers
        const bool wide0 = variant < 10 && (W % 128) == 0;
        const long long units_guess = (long long)BA * ((W + 127) / 128) * ((H + PW_R - 1) / PW_R) * ((S + 15) / 16);
        const bool dense4 = variant == 4 || (variant == 0 && wide0 && units_guess >= (long long)sms * 4 * PW_NW * 2 && BA >= 32);
        const bool fdom = wide && !do_hist && variant != 7;           // float-domain kernel (no histogram)
        const int fminb = variant == 5 ? 5 : (variant == 3 ? 3 : 4);
        const int wminb = fdom ? fminb : (dense4 ? 4 : PW_MINB);
        const int xs = wide ? 128 : 32, R = wide ? PW_R : 4, minb = wide ? wminb : 2, nw = wide ? PW_NW : PK_NW;
        const int nstrips = (W + xs - 1) / xs, nrowg = (H + R - 1) / R;
        // persistent CTAs: 2 resident CTAs per SM in total, spread over the B*A maps
        int per_map = (sms * minb) / BA;                   // never more CTAs than resident slots: a second wave would double the time
        if (per_map < 1) per_map = 1;
        // z chunk: a warp's time is (units per warp) x (planes per unit + 2 halo planes); pick the cheapest split
        int best_tz = S;
        long long best_cost = -1;
        for (int tz = S;; tz = (tz + 1) / 2) {
            const long long nzc = (S + tz - 1) / tz, units = (long long)nstrips * nrowg * nzc;
            const long long warps = (long long)per_map * nw;
            const long long cost = ((units + warps - 1) / warps) * (tz + 2);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_tz = tz; }
            if (tz <= 4) break;
        }
        const int tz = best_tz, nzc = (S + tz - 1) / tz;
        const long long units = (long long)nstrips * nrowg * nzc;
        B200_CHECK_ARG(units < (1ll << 31), "peaks3d: too many work units");
        if ((long long)per_map * nw > units) per_map = (int)((units + nw - 1) / nw);
        dim3 g1((unsigned)per_map, BA);
        if (fdom) {
            if (fminb == 5) peaks_scan3f_kernel<5><<<g1, PW_THREADS, 0, stream>>>(input, S, H, W, nstrips, nrowg, nzc, tz, ws);
            else if (fminb == 3) peaks_scan3f_kernel<3><<<g1, PW_THREADS, 0, stream>>>(input, S, H, W, nstrips, nrowg, nzc, tz, ws);
            else peaks_scan3f_kernel<4><<<g1, PW_THREADS, 0, stream>>>(input, S, H, W, nstrips, nrowg, nzc, tz, ws);
        } else if (wide) {
            if (dense4) {
                if (do_hist) peaks_scan3w_kernel<true, 4><<<g1, PW_THREADS, 0, stream>>>(input, S, H, W, nstrips, nrowg, nzc, tz, ws);
                else peaks_scan3w_kernel<false, 4><<<g1, PW_THREADS, 0, stream>>>(input, S, H, W, nstrips, nrowg, nzc, tz, ws);
            } else {
                if (do_hist) peaks_scan3w_kernel<true, PW_MINB><<<g1, PW_THREADS, 0, stream>>>(input, S, H, W, nstrips, nrowg, nzc, tz, ws);
                else peaks_scan3w_kernel<false, PW_MINB><<<g1, PW_THREADS, 0, stream>>>(input, S, H, W, nstrips, nrowg, nzc, tz, ws);
            }
        } else if ((W % 32) == 0) peaks_scan3_kernel<4, true, 2><<<g1, PK_THREADS, 0, stream>>>(input, S, H, W, nstrips, nrowg, nzc, tz, do_hist, ws);
        else peaks_scan3_kernel<4, false, 2><<<g1, PK_THREADS, 0, stream>>>(input, S, H, W, nstrips, nrowg, nzc, tz, do_hist, ws);
        B200_LAUNCH_CHECK("peaks_scan3_kernel");
    } else {
        const int tiles_x = (W + PK_TX - 1) / PK_TX, tiles_y = (H + PK_TY - 1) / PK_TY, tiles_z = (S + PK_TZ - 1) / PK_TZ;
        const long long ntiles = (long long)tiles_x * tiles_y * tiles_z;
        B200_CHECK_ARG(ntiles < (1ll << 31), "peaks3d: too many tiles");
        long long per_map = ((long long)sms * 5 + BA - 1) / BA;
        if (per_map > ntiles) per_map = ntiles;
        if (per_map < 1) per_map = 1;
        dim3 g1((unsigned)per_map, BA);
        if (win == 5) peaks_scan_kernel<5><<<g1, PK_THREADS, 0, stream>>>(input, S, H, W, tiles_x, tiles_y, tiles_z, do_hist, ws);
        else peaks_scan_kernel<7><<<g1, PK_THREADS, 0, stream>>>(input, S, H, W, tiles_x, tiles_y, tiles_z, do_hist, ws);
        B200_LAUNCH_CHECK("peaks_scan_kernel");
    }
    if (stop_after < 2) return 0;
    if (interval) {
        peaks_sample_kernel<<<BA, 512, 0, stream>>>(input, V, ws, median_mode == 2);
        B200_LAUNCH_CHECK("peaks_sample_kernel");
        if (stop_after < 3) return 0;
        static OncePerDevice iv_attr;
        int iv_dev;
        constexpr int iv_smem = PK_IV_NW * PK_IV_WSTAGE * 4;
        if (iv_attr.needed(&iv_dev)) {
            B200_CUDA(cudaFuncSetAttribute(peaks_interval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, iv_smem));
            iv_attr.mark(iv_dev);
        }
        // CTAs per map: enough to fill the machine twice over when the batch is small, one fat CTA per map when it is large
        long long want = (V / 4 + (long long)PK_IV_THREADS * 16 - 1) / ((long long)PK_IV_THREADS * 16);     // >= 4 rounds per thread
        long long lim = ((long long)sms * 2 + BA - 1) / BA;
        int gi = (int)(want < 1 ? 1 : (want > lim ? lim : want));
        if (gi < 1) gi = 1;
        dim3 g2(gi, BA);
        peaks_interval_kernel<<<g2, PK_IV_THREADS, iv_smem, stream>>>(input, V, ws, thr_out);
        B200_LAUNCH_CHECK("peaks_interval_kernel");
    } else if (filter_mode == 1) {
        // a few fat CTAs per map: the fixed cost per CTA (clear + publish a 4096-bin histogram) must stay small
        // against its share of the stream
        long long want = (V / 4 + (long long)PK_RF_THREADS * 8 - 1) / ((long long)PK_RF_THREADS * 8);
        long long lim = ((long long)sms * 2) / BA;
        int gr = (int)(want < 1 ? 1 : (want > lim ? lim : want));
        if (gr < 1) gr = 1;
        dim3 g2(gr, BA);
        peaks_collect_kernel<<<g2, PK_RF_THREADS, 0, stream>>>(input, V, ws, thr_out);
        B200_LAUNCH_CHECK("peaks_collect_kernel");
        if (stop_after < 3) return 0;
        peaks_refine_kernel<3><<<g2, PK_RF_THREADS, 0, stream>>>(input, V, ws, thr_out);
        B200_LAUNCH_CHECK("peaks_refine_kernel<3>");
    }
    if (stop_after < 4) return 0;
    peaks_finalize_kernel<<<BA * ws.nchunks, PK_THREADS, 0, stream>>>(input, BA, A, H, W, V, filter_mode, thr_in, ws, peaks, cap,
                                                                     n_peaks, agg, thr_out);
    B200_LAUNCH_CHECK("peaks_finalize_kernel");
    return 0;
}

extern "C" int b200seg_peaks3d_fallback_count(unsigned long long* count) {
    B200_CHECK_ARG(count, "peaks3d_fallback_count: null pointer");
    B200_CUDA(cudaMemcpyFromSymbol(count, pk_fallbacks, sizeof(unsigned long long)));
    return 0;
}

extern "C" int b200seg_peaks3d_bwd_dev(const int64_t* peaks, const int32_t* n_peaks, int cap,
                                       const float* grad_agg, float* grad_in, int B, int A, int S, int H, int W,
                                       b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(B >= 0 && A >= 0 && S >= 0 && H >= 0 && W >= 0 && cap >= 0, "peaks3d_bwd: negative size");
    const size_t total = (size_t)B * A * S * H * W;
    if (total == 0) return 0;
    B200_CHECK_ARG(grad_in && grad_agg && n_peaks && (peaks || cap == 0), "peaks3d_bwd: null pointer");
    B200_CUDA(cudaMemsetAsync(grad_in, 0, total * sizeof(float), stream));
    if (cap > 0) {
        int blocks = (cap + 255) / 256;
        if (blocks > num_sms() * 8) blocks = num_sms() * 8;
        peaks_bwd_scatter_kernel<<<blocks, 256, 0, stream>>>(peaks, n_peaks, cap, grad_agg, grad_in, A, S, H, W);
        B200_LAUNCH_CHECK("peaks_bwd_scatter_kernel");
    }
    return 0;
}
