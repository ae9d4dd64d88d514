// peaks3d.cu -- PRM peak stimulation (replaces lib/prm/peak_stimulation_3d.py:9-48 and the median
// filter of lib/prm/peak_response_mapping_3d.py:45-49).
//
// The reference runs pad(-inf) -> max_pool3d(return_indices) -> (indices == own index) ->
// (input >= median) -> nonzero, i.e. five full-volume torch kernels plus an int64 index volume.
// Here:
//   peaks_scan_kernel    one pass over the map: halo tile in shared memory, the ATen arg-max rule
//                        evaluated directly as a raster-order-aware local-max predicate ("earlier
//                        neighbours strictly smaller, later neighbours not larger"), candidate bits
//                        OR-ed into a 1 bit/voxel mask, and the level-1 (top 12 key bits) radix
//                        histogram for the exact median accumulated in shared memory.
//   peaks_refine_kernel  two more radix levels (12 + 8 bits) with 128-bit streaming loads; every CTA
//                        re-derives the previous level's selected bin from the global histogram
//                        (no separate "select" launches).
//   peaks_filter_kernel  applies input >= threshold to the candidate bits, counts per chunk and
//                        accumulates the aggregation partial sums (deterministic order).
//   peaks_offsets_kernel exclusive scan of the chunk counts, final aggregation, total count.
//   peaks_emit_kernel    expands bits to (b,a,z,y,x) int64 rows in lexicographic order --
//                        ordered stream compaction, no sort, no int64 index volume.
// Exact lower median = element of rank (V-1)/2 of the ascending order (torch.median).
#include "common.cuh"
#include <math_constants.h>

namespace b200seg {

constexpr int PK_TX = 32, PK_TY = 8, PK_TZ = 8;
constexpr int PK_THREADS = 256;
constexpr int PK_BINS1 = 4096;
constexpr int PK_CHUNK_WORDS = 1024;               // bitmask words (32 voxels each) per filter/emit CTA

struct PeakWs {
    uint32_t* hist1;      // [BA][4096]
    uint32_t* hist2;      // [BA][4096]
    uint32_t* hist3;      // [BA][256]
    uint32_t* nan_cnt;    // [BA]
    uint32_t* bits;       // [BA][words]
    uint32_t* chunk_cnt;  // [BA][nchunks]
    float* chunk_sum;     // [BA][nchunks]
    uint32_t* chunk_off;  // [BA][nchunks]
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
    w.nan_cnt = (uint32_t*)take((size_t)BA * 4);
    w.bits = (uint32_t*)take((size_t)BA * w.words * 4);
    w.zero_bytes = (size_t)(p - (char*)base);
    w.chunk_cnt = (uint32_t*)take((size_t)BA * w.nchunks * 4);
    w.chunk_sum = (float*)take((size_t)BA * w.nchunks * 4);
    w.chunk_off = (uint32_t*)take((size_t)BA * w.nchunks * 4);
    w.thr = (float*)take((size_t)BA * 4);
    w.total_bytes = (size_t)(p - (char*)base);
    return w;
}

// Finds the bin holding ascending rank k in hist[nbins]; returns bin and the rank inside it.
// Must be called by all PK_THREADS threads; s_tmp has PK_THREADS+2 entries.  Block-wide prefix sum of per-thread
// bin groups (warp shuffles + one shared-memory hop); the single thread whose group contains rank k resolves it.
__device__ void select_bin(const uint32_t* __restrict__ hist, int nbins, unsigned long long k,
                           unsigned long long* s_tmp, int* out_bin, unsigned long long* out_k) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = PK_THREADS / 32;
    const int per = (nbins + PK_THREADS - 1) / PK_THREADS;
    const int b0 = min(nbins, tid * per), b1 = min(nbins, b0 + per);
    unsigned long long local = 0;
    for (int b = b0; b < b1; ++b) local += hist[b];
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
        for (; b < b1; ++b) { const unsigned long long h = hist[b]; if (acc + h > k) break; acc += h; }
        s_tmp[NW] = (unsigned long long)b;
        s_tmp[NW + 1] = k - acc;
    }
    __syncthreads();
    *out_bin = (int)s_tmp[NW];
    *out_k = s_tmp[NW + 1];
    __syncthreads();
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
    unsigned nan_local = 0;
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
                if (inb && v != v) ++nan_local;
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
    if (do_hist) {
        __syncthreads();
        uint32_t* gh = ws.hist1 + (size_t)ba * PK_BINS1;
        for (int i = tid; i < PK_BINS1; i += PK_THREADS) { const uint32_t c = s_hist[i]; if (c) atomicAdd(&gh[i], c); }
        if (nan_local) atomicAdd(&ws.nan_cnt[ba], nan_local);
    }
}

// ---- window 3 (the reference default, peak_response_mapping_3d.py:29): register sliding window ----------
// The tile is stored as ORDERED KEYS (monotone uint32 image of the float, every NaN = 0xFFFFFFFF), so the ATen
// arg-max rule becomes integer arithmetic: with Kb / Ka the unsigned maxima over the 13 neighbours that come
// earlier / later in the (z,y,x) window scan,
//     v not NaN:  peak <=> v != -inf  and  Kb < key(v)  and  Ka <= key(v)      (a NaN neighbour has the largest key)
//     v NaN:      peak <=> Ka != key(NaN)                                       (the LAST NaN of a window wins)
// A thread owns one (x,y) column of the tile and marches along z keeping three plane summaries in registers
// (max of the 9, max of the upper row, max of the lower row, left, centre, right): 9 shared loads, 6 three-input
// integer maxima and two compares per voxel, no branches.  The level-1 radix histogram of the exact median
// uses the key's top 12 bits directly.
constexpr int P3_TX = 32, P3_TY = 8, P3_TZ = 16;
constexpr int P3_HX = P3_TX + 2, P3_HY = P3_TY + 2, P3_HZ = P3_TZ + 2;

struct PlaneStat { uint32_t m9, top3, bot3, l, c, r; };

__device__ __forceinline__ PlaneStat plane_stat(const uint32_t* __restrict__ p) {   // p -> key (dy=0, dx=0) of the 3x3 patch
    PlaneStat s;
    s.top3 = __vimax3_u32(p[0], p[1], p[2]);
    s.l = p[P3_HX]; s.c = p[P3_HX + 1]; s.r = p[P3_HX + 2];
    s.bot3 = __vimax3_u32(p[2 * P3_HX], p[2 * P3_HX + 1], p[2 * P3_HX + 2]);
    s.m9 = __vimax3_u32(s.top3, s.bot3, __vimax3_u32(s.l, s.c, s.r));
    return s;
}

__global__ void __launch_bounds__(PK_THREADS)
peaks_scan3_kernel(const float* __restrict__ in, int S, int H, int W, int tiles_x, int tiles_y, int tiles_z,
                   int do_hist, PeakWs ws) {
    __shared__ uint32_t s_key[P3_HZ * P3_HY * P3_HX];
    __shared__ uint32_t s_hist[PK_BINS1];
    const int ba = blockIdx.y;
    const long long V = (long long)S * H * W;
    const float* vol = in + (size_t)ba * V;
    uint32_t* bits = ws.bits + (size_t)ba * ws.words;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t KEY_NAN = 0xFFFFFFFFu, KEY_NINF = 0x007FFFFFu;          // ordered_key(-inf) = ~0xFF800000
    if (do_hist) for (int i = tid; i < PK_BINS1; i += PK_THREADS) s_hist[i] = 0u;
    unsigned nan_local = 0;
    const int ntiles = tiles_x * tiles_y * tiles_z;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, tz = tile / (tiles_x * tiles_y);
        const int x0 = tx * P3_TX, y0 = ty * P3_TY, z0 = tz * P3_TZ;
        __syncthreads();
        // ---- load: one warp per halo row; lanes 0,1 also fetch the two halo columns --------------------
        constexpr int NWARP = PK_THREADS / 32, RU = 4;    // RU rows per warp in flight: the loads of a batch are independent
        for (int rb = warp; rb < P3_HY * P3_HZ; rb += NWARP * RU) {
            float v[RU], vh[RU];
            bool ok[RU], okh[RU];
#pragma unroll
            for (int j = 0; j < RU; ++j) {
                const int rr = rb + j * NWARP;
                const int hz = rr / P3_HY, hy = rr - hz * P3_HY;
                const int gy = y0 + hy - 1, gz = z0 + hz - 1;
                const bool row_ok = rr < P3_HY * P3_HZ && gy >= 0 && gy < H && gz >= 0 && gz < S;   // warp-uniform
                const float* row = vol + ((size_t)(row_ok ? gz : 0) * H + (row_ok ? gy : 0)) * W;
                const int gx = x0 + lane;
                const int hxg = lane == 0 ? x0 - 1 : x0 + P3_TX;
                ok[j] = row_ok && gx < W;
                okh[j] = row_ok && lane < 2 && hxg >= 0 && hxg < W;
                v[j] = ok[j] ? row[gx] : 0.f;
                vh[j] = okh[j] ? row[hxg] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < RU; ++j) {
                const int rr = rb + j * NWARP;
                if (rr < P3_HY * P3_HZ) {
                    uint32_t* dst = s_key + rr * P3_HX;
                    dst[1 + lane] = ok[j] ? ordered_key(v[j]) : KEY_NINF;
                    if (lane < 2) dst[lane == 0 ? 0 : P3_HX - 1] = okh[j] ? ordered_key(vh[j]) : KEY_NINF;
                }
            }
        }
        __syncthreads();
        // ---- compute: thread = (x, y) column, march along z -------------------------------------------
        const int lx = lane, ly = warp;                   // 32 x 8 threads
        const int gx = x0 + lx, gy = y0 + ly;
        const bool col_ok = gx < W && gy < H;
        const uint32_t* colp = s_key + ly * P3_HX + lx;   // (dy = 0, dx = 0) of the patch in halo plane 0
        PlaneStat prev = plane_stat(colp), cur = plane_stat(colp + P3_HY * P3_HX);
#pragma unroll 4
        for (int lz = 0; lz < P3_TZ; ++lz) {
            const PlaneStat next = plane_stat(colp + (lz + 2) * P3_HY * P3_HX);
            const uint32_t kb = __vimax3_u32(prev.m9, cur.top3, cur.l);
            const uint32_t ka = __vimax3_u32(next.m9, cur.bot3, cur.r);
            const uint32_t kv = cur.c;
            const int gz = z0 + lz;
            const bool inb = col_ok && gz < S;
            const bool peak = inb && (kv == KEY_NAN ? ka != KEY_NAN : (kv != KEY_NINF && kb < kv && ka <= kv));
            if (do_hist && inb) {
                if (kv == KEY_NAN) ++nan_local;
                atomicAdd(&s_hist[kv >> 20], 1u);
            }
            if (peak) {
                const long long flat = ((long long)gz * H + gy) * W + gx;
                atomicOr(&bits[flat >> 5], 1u << (flat & 31));
            }
            prev = cur; cur = next;
        }
    }
    if (do_hist) {
        __syncthreads();
        uint32_t* gh = ws.hist1 + (size_t)ba * PK_BINS1;
        for (int i = tid; i < PK_BINS1; i += PK_THREADS) { const uint32_t c = s_hist[i]; if (c) atomicAdd(&gh[i], c); }
        if (nan_local) atomicAdd(&ws.nan_cnt[ba], nan_local);
    }
}

// LEVEL 2: bins = key bits 19..8 of elements whose top 12 bits match the level-1 bin.
// LEVEL 3: bins = key bits 7..0 of elements whose top 24 bits match.
template <int LEVEL>
__global__ void __launch_bounds__(PK_THREADS)
peaks_refine_kernel(const float* __restrict__ in, long long V, PeakWs ws) {
    __shared__ uint32_t s_hist[PK_BINS1];
    __shared__ unsigned long long s_tmp[PK_THREADS + 2];
    const int ba = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned long long k = (unsigned long long)((V - 1) / 2);
    int bin1; unsigned long long k1;
    select_bin(ws.hist1 + (size_t)ba * PK_BINS1, PK_BINS1, k, s_tmp, &bin1, &k1);
    uint32_t prefix = (uint32_t)bin1;
    int shift = 20;
    if (LEVEL == 3) {
        int bin2; unsigned long long k2;
        select_bin(ws.hist2 + (size_t)ba * PK_BINS1, PK_BINS1, k1, s_tmp, &bin2, &k2);
        prefix = ((uint32_t)bin1 << 12) | (uint32_t)bin2;
        shift = 8;
    }
    constexpr int NB = LEVEL == 2 ? PK_BINS1 : 256;
    for (int i = tid; i < NB; i += PK_THREADS) s_hist[i] = 0u;
    __syncthreads();
    const float* vol = in + (size_t)ba * V;
    const bool vec = ((((uintptr_t)vol) & 15) == 0);
    const long long nvec = vec ? (V >> 2) : 0;
    auto add = [&](float f) {
        const uint32_t key = ordered_key(f);
        if ((key >> shift) == prefix) {
            const uint32_t bin = LEVEL == 2 ? ((key >> 8) & 0xFFFu) : (key & 0xFFu);
            atomicAdd(&s_hist[bin], 1u);
        }
    };
    {
        const long long stride = (long long)gridDim.x * PK_THREADS;
        long long i = (long long)blockIdx.x * PK_THREADS + tid;
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
    for (long long i = (nvec << 2) + (long long)blockIdx.x * PK_THREADS + tid; i < V; i += (long long)gridDim.x * PK_THREADS)
        add(vol[i]);
    (void)lane;
    __syncthreads();
    uint32_t* gh = (LEVEL == 2 ? ws.hist2 + (size_t)ba * PK_BINS1 : ws.hist3 + (size_t)ba * 256);
    for (int i = tid; i < NB; i += PK_THREADS) { const uint32_t c = s_hist[i]; if (c) atomicAdd(&gh[i], c); }
}

// grid (nchunks, BA).  filter_mode 1: derive the median threshold from the three histograms;
// 2: thresholds in thr_in; 0: no filter.
__global__ void __launch_bounds__(PK_THREADS)
peaks_filter_kernel(const float* __restrict__ in, long long V, int filter_mode, const float* __restrict__ thr_in,
                    PeakWs ws, float* __restrict__ thr_out) {
    __shared__ unsigned long long s_tmp[PK_THREADS + 2];
    __shared__ float s_thr;
    __shared__ uint32_t s_cnt[PK_THREADS / 32];
    __shared__ float s_sum[PK_THREADS / 32];
    const int ba = blockIdx.y, chunk = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float thr = 0.f;
    if (filter_mode == 1) {
        const unsigned long long k = (unsigned long long)((V - 1) / 2);
        int b1, b2, b3; unsigned long long k1, k2, k3;
        select_bin(ws.hist1 + (size_t)ba * PK_BINS1, PK_BINS1, k, s_tmp, &b1, &k1);
        select_bin(ws.hist2 + (size_t)ba * PK_BINS1, PK_BINS1, k1, s_tmp, &b2, &k2);
        select_bin(ws.hist3 + (size_t)ba * 256, 256, k2, s_tmp, &b3, &k3);
        if (tid == 0) {
            const uint32_t key = ((uint32_t)b1 << 20) | ((uint32_t)b2 << 8) | (uint32_t)b3;
            s_thr = ws.nan_cnt[ba] ? CUDART_NAN_F : key_to_float(key);
        }
        __syncthreads();
        thr = s_thr;
    } else if (filter_mode == 2) {
        thr = thr_in[ba];
    }
    if (chunk == 0 && tid == 0) { ws.thr[ba] = thr; if (thr_out) thr_out[ba] = thr; }

    const float* vol = in + (size_t)ba * V;
    uint32_t* bits = ws.bits + (size_t)ba * ws.words;
    uint32_t cnt = 0;
    float sum = 0.f;
    const int w0 = chunk * PK_CHUNK_WORDS;
    for (int wi = tid; wi < PK_CHUNK_WORDS; wi += PK_THREADS) {
        const int w = w0 + wi;
        if (w >= ws.words) break;
        uint32_t m = bits[w];
        if (m) {
            uint32_t keepm = 0;
            uint32_t g = m;
            while (g) {
                const int bit = __ffs(g) - 1;
                g &= g - 1;
                const float v = vol[(long long)w * 32 + bit];
                if (filter_mode == 0 || v >= thr) { keepm |= 1u << bit; sum += v; ++cnt; }
            }
            if (keepm != m) bits[w] = keepm;
        }
    }
    // deterministic tree reduction
#pragma unroll
    for (int o = 16; o; o >>= 1) { cnt += __shfl_down_sync(0xffffffffu, cnt, o); sum += __shfl_down_sync(0xffffffffu, sum, o); }
    if (lane == 0) { s_cnt[warp] = cnt; s_sum[warp] = sum; }
    __syncthreads();
    if (tid == 0) {
        uint32_t c = 0; float s = 0.f;
        for (int i = 0; i < PK_THREADS / 32; ++i) { c += s_cnt[i]; s += s_sum[i]; }
        ws.chunk_cnt[(size_t)ba * ws.nchunks + chunk] = c;
        ws.chunk_sum[(size_t)ba * ws.nchunks + chunk] = s;
    }
}

// single CTA: exclusive scan over [BA][nchunks] counts in lexicographic order, aggregation per map
__global__ void __launch_bounds__(PK_THREADS)
peaks_offsets_kernel(int BA, PeakWs ws, int32_t* __restrict__ n_peaks, float* __restrict__ agg) {
    __shared__ unsigned long long s_scan[PK_THREADS];
    const int tid = threadIdx.x;
    const long long total = (long long)BA * ws.nchunks;
    unsigned long long base = 0;
    for (long long i0 = 0; i0 < total; i0 += PK_THREADS) {
        const long long i = i0 + tid;
        const unsigned long long c = i < total ? ws.chunk_cnt[i] : 0u;
        s_scan[tid] = c;
        __syncthreads();
        for (int d = 1; d < PK_THREADS; d <<= 1) {
            const unsigned long long v = tid >= d ? s_scan[tid - d] : 0ull;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        if (i < total) {
            const unsigned long long off = base + s_scan[tid] - c;
            ws.chunk_off[i] = (uint32_t)(off > 0xFFFFFFFFull ? 0xFFFFFFFFull : off);
        }
        base += s_scan[PK_THREADS - 1];
        __syncthreads();
    }
    if (tid == 0) *n_peaks = (int32_t)(base > 0x7FFFFFFFull ? 0x7FFFFFFFull : base);
    if (agg) {
        for (int ba = tid; ba < BA; ba += PK_THREADS) {
            float s = 0.f, c = 0.f;
            for (int j = 0; j < ws.nchunks; ++j) {
                s += ws.chunk_sum[(size_t)ba * ws.nchunks + j];
                c += (float)ws.chunk_cnt[(size_t)ba * ws.nchunks + j];
            }
            agg[ba] = s / c;            // 0/0 = NaN when the map has no peak, as in the reference
        }
    }
}

__global__ void __launch_bounds__(PK_THREADS)
peaks_emit_kernel(int A, int H, int W, PeakWs ws, int64_t* __restrict__ peaks, int cap) {
    __shared__ uint32_t s_scan[PK_THREADS];
    const int ba = blockIdx.y, chunk = blockIdx.x;
    const int tid = threadIdx.x;
    const size_t ci = (size_t)ba * ws.nchunks + chunk;
    if (ws.chunk_cnt[ci] == 0) return;
    const uint32_t* bits = ws.bits + (size_t)ba * ws.words;
    constexpr int PER = PK_CHUNK_WORDS / PK_THREADS;      // consecutive words per thread
    const int w0 = chunk * PK_CHUNK_WORDS + tid * PER;
    uint32_t m[PER];
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) { m[j] = (w0 + j) < ws.words ? bits[w0 + j] : 0u; c += __popc(m[j]); }
    s_scan[tid] = c;
    __syncthreads();
    for (int d = 1; d < PK_THREADS; d <<= 1) {
        const uint32_t v = tid >= d ? s_scan[tid - d] : 0u;
        __syncthreads();
        s_scan[tid] += v;
        __syncthreads();
    }
    long long pos = (long long)ws.chunk_off[ci] + s_scan[tid] - c;
    const int b = ba / A, a = ba % A;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        uint32_t g = m[j];
        while (g) {
            const int bit = __ffs(g) - 1;
            g &= g - 1;
            if (pos < cap) {
                const long long flat = (long long)(w0 + j) * 32 + bit;
                const int x = (int)(flat % W);
                const long long t = flat / W;
                const int y = (int)(t % H), z = (int)(t / H);
                int64_t* row = peaks + pos * 5;
                row[0] = b; row[1] = a; row[2] = z; row[3] = y; row[4] = x;
            }
            ++pos;
        }
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
    B200_CUDA(cudaMemsetAsync(base, 0, ws.zero_bytes, stream));
    const int sms = num_sms();
    const int do_hist = filter_mode == 1;
    {
        const int TX = win == 3 ? P3_TX : PK_TX, TY = win == 3 ? P3_TY : PK_TY, TZ = win == 3 ? P3_TZ : PK_TZ;
        const int tiles_x = (W + TX - 1) / TX, tiles_y = (H + TY - 1) / TY, tiles_z = (S + TZ - 1) / TZ;
        const long long ntiles = (long long)tiles_x * tiles_y * tiles_z;
        B200_CHECK_ARG(ntiles < (1ll << 31), "peaks3d: too many tiles");
        // persistent CTAs: about 5 resident CTAs per SM in total, spread over the B*A maps
        long long per_map = ((long long)sms * 5 + BA - 1) / BA;
        if (per_map > ntiles) per_map = ntiles;
        if (per_map < 1) per_map = 1;
        dim3 g1((unsigned)per_map, BA);
        if (win == 3) peaks_scan3_kernel<<<g1, PK_THREADS, 0, stream>>>(input, S, H, W, tiles_x, tiles_y, tiles_z, do_hist, ws);
        else if (win == 5) peaks_scan_kernel<5><<<g1, PK_THREADS, 0, stream>>>(input, S, H, W, tiles_x, tiles_y, tiles_z, do_hist, ws);
        else peaks_scan_kernel<7><<<g1, PK_THREADS, 0, stream>>>(input, S, H, W, tiles_x, tiles_y, tiles_z, do_hist, ws);
        B200_LAUNCH_CHECK("peaks_scan_kernel");
    }
    if (filter_mode == 1) {
        long long want = (V / 4 + PK_THREADS - 1) / PK_THREADS;
        int gr = (int)(want < 1 ? 1 : (want > (long long)sms * 8 / BA + 1 ? (long long)sms * 8 / BA + 1 : want));
        dim3 g2(gr, BA);
        peaks_refine_kernel<2><<<g2, PK_THREADS, 0, stream>>>(input, V, ws);
        B200_LAUNCH_CHECK("peaks_refine_kernel<2>");
        peaks_refine_kernel<3><<<g2, PK_THREADS, 0, stream>>>(input, V, ws);
        B200_LAUNCH_CHECK("peaks_refine_kernel<3>");
    }
    dim3 g3(ws.nchunks, BA);
    peaks_filter_kernel<<<g3, PK_THREADS, 0, stream>>>(input, V, filter_mode, thr_in, ws, thr_out);
    B200_LAUNCH_CHECK("peaks_filter_kernel");
    peaks_offsets_kernel<<<1, PK_THREADS, 0, stream>>>(BA, ws, n_peaks, agg);
    B200_LAUNCH_CHECK("peaks_offsets_kernel");
    if (cap > 0) {
        peaks_emit_kernel<<<g3, PK_THREADS, 0, stream>>>(A, H, W, ws, peaks, cap);
        B200_LAUNCH_CHECK("peaks_emit_kernel");
    }
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
