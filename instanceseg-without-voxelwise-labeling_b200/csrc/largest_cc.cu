// largest_cc.cu -- keep the largest 26-connected component of every instance mask, in place
// (tools/binarization_soma.py:97-99: `labels = label(box_bi)`, skimage.measure.label, full connectivity;
//  `largestCC = labels == argsort(bincount(labels.flat)[1:])[-1] + 1`).  The step sits between the 2D-Otsu
// binarization and the label paste-back of the per-volume chain.
//
// One CTA per instance crop, run-based labelling: a mask row (x fastest) is a bit vector, its maximal runs of
// set bits are the union-find nodes (a few per row instead of one per voxel).
//   1. row bit vectors from the packed mask bytes (shared memory, global scratch when the crop is large)
//   2. runs per row + block prefix sum -> run ids in raster order of their first voxel
//   3. union of every run with the runs it touches in the four earlier neighbour rows (z-1,y-1) (z-1,y) (z-1,y+1)
//      (z,y-1), x ranges dilated by one: exactly 26-connectivity.  Lock-free union-find, smaller id = root, so the
//      root of a component is the run that holds its first voxel in raster order -- roots are ordered like
//      skimage's labels.
//   4. component sizes (sum of run lengths at the root), arg-max, and the runs of every other component are
//      cleared in the mask.
// Tie rule (documented; numpy's default argsort is not stable): equal sizes -> the component with the LATER
// first voxel wins (what a stable argsort()[-1] returns).  A mask without foreground makes the reference raise
// (IndexError on the empty bincount); here it is reported as status 5 and nothing is pasted.
#include "common.cuh"

namespace b200seg {

constexpr int CC_THREADS = 256;
// dynamic shared memory (72 KB, three CTAs per SM); larger crops spill to their slice of the global scratch
constexpr int CC_SMEM_WORDS = 3072;        // row bit vectors kept in shared memory (64-bit words)
constexpr int CC_SMEM_ROWS = 3072;         // per-row run offsets kept in shared memory
constexpr int CC_SMEM_RUNS = 2816;         // union-find nodes kept in shared memory (parent, size, interval, row)
constexpr size_t CC_DYN_BYTES = (size_t)CC_SMEM_WORDS * 8 + (size_t)(CC_SMEM_ROWS + 2) * 4 + (size_t)CC_SMEM_RUNS * 12;

struct CcShared {
    int scan[CC_THREADS];
    unsigned long long best;               // (size << 32) | root
    int total_runs;
};

// find with path halving.  Parent links only ever move towards smaller ids (roots are component minima), so a
// racy shortcut store always installs a valid ancestor; roots are never written here (uf_union's CAS owns them).
__device__ __forceinline__ int uf_find(volatile int* parent, int x) {
    int p = parent[x];
    while (p != x) {
        const int gp = parent[p];
        if (gp == p) return p;
        parent[x] = gp;
        x = gp;
        p = parent[x];
    }
    return x;
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a); b = uf_find(parent, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }
        const int old = atomicCAS(&parent[b], b, a);        // b was a root: hang it under the smaller id
        if (old == b) return;
        b = old;
    }
}

// next maximal run of set bits at or after position p in a W64-word row of sx valid bits; false when none
__device__ __forceinline__ bool next_run(const unsigned long long* __restrict__ row, int W64, int sx, int p, int& xs, int& xe) {
    int w = p >> 6;
    if (w >= W64) return false;
    unsigned long long cur = row[w] & (~0ull << (p & 63));
    while (cur == 0ull) { if (++w >= W64) return false; cur = row[w]; }
    xs = (w << 6) + __ffsll((long long)cur) - 1;
    // end of the run: first clear bit after xs
    unsigned long long inv = ~row[w] & (~0ull << (xs & 63));
    while (inv == 0ull) { if (++w >= W64) { xe = sx - 1; return true; } inv = ~row[w]; }
    xe = min(sx, (w << 6) + __ffsll((long long)inv) - 1) - 1;
    return true;
}
// Union phase on INTERVAL LISTS (rows of <= 256 voxels, runs and rows fit the shared-memory tables).  The bit
// vectors are only used to enumerate the runs once; afterwards a run is the 16-bit interval (first x << 8) | last x
// and "run A touches run B of a neighbour row" is two integer compares (x ranges dilated by one = 26-connectivity).
// Components are formed by hooking and pointer jumping (Shiloach-Vishkin style) instead of find() walks -- with
// every thread of the CTA joining runs at once, find() chains grow as long as the crop is deep and dominate:
//   (0) nodes: parent = self, size = run length, interval, row
//   repeat
//     (H) every (run, earlier neighbour row) pair: for each touching run, hook the larger of the two ROOTS under the
//         smaller one with atomicMin (parents are roots here: the forest is flat after (J));
//     (J) pointer jumping until the forest is flat again (~log2 of the longest chain, uniform rounds);
//   until no hook happened.  The root of a component ends up as its smallest run id = its first run in raster order.
__device__ __forceinline__ void cc_unions_intervals(const unsigned long long* __restrict__ bits, const int* __restrict__ run_off,
                                                    int* parent, int* size, unsigned short* iv, unsigned short* rrow,
                                                    int rows, int sx, int sy, int W64, int T) {
    const int tid = threadIdx.x;
    for (int r = tid; r < rows; r += CC_THREADS) {            // (0)
        const unsigned long long* row = bits + (size_t)r * W64;
        int id = run_off[r], p = 0, xs, xe;
        while (next_run(row, W64, sx, p, xs, xe)) {
            parent[id] = id; size[id] = xe - xs + 1; iv[id] = (unsigned short)((xs << 8) | xe); rrow[id] = (unsigned short)r;
            ++id; p = xe + 2;
        }
    }
    __syncthreads();
    const unsigned int m_sy = 0xFFFFFFFFu / (unsigned)sy + 1u;                // r / sy for r * sy < 2^32
    for (int iter = 0; iter < 64; ++iter) {                   // 64 >> the handful of iterations real masks need; see below
        int hooked = 0;
        for (int task = tid; task < 4 * T; task += CC_THREADS) {  // (H)
            const int id = task >> 2, nb = task & 3;
            const int r = rrow[id];
            const int z = sy == 1 ? r : (int)__umulhi((unsigned)r, m_sy), y = r - z * sy;
            const int zz = nb < 3 ? z - 1 : z, yy = nb < 3 ? y - 1 + nb : y - 1;
            if (zz < 0 || yy < 0 || yy >= sy) continue;
            const int r2 = zz * sy + yy;
            const unsigned int a = iv[id];
            const int o2 = run_off[r2], e2 = run_off[r2 + 1];
            for (int j = o2; j < e2; ++j) {
                const unsigned int b = iv[j];
                if ((b >> 8) > (a & 255u) + 1u || (b & 255u) + 1u < (a >> 8)) continue;      // x ranges do not touch
                const int ra = parent[id], rb = parent[j];
                if (ra != rb) { atomicMin(&parent[max(ra, rb)], min(ra, rb)); hooked = 1; }
            }
        }
        if (!__syncthreads_or(hooked)) break;
        for (int round = 0; round < 32; ++round) {            // (J)
            int changed = 0;
            for (int id = tid; id < T; id += CC_THREADS) {
                const int p = parent[id];
                const int gp = parent[p];
                if (gp != p) { parent[id] = gp; changed = 1; }
            }
            if (!__syncthreads_or(changed)) break;
        }
    }
}

// joins every run of row r with the runs of row r2 it touches (x ranges dilated by one voxel)
__device__ __forceinline__ void union_rows(const unsigned long long* __restrict__ bits, const int* __restrict__ run_off,
                                           int* parent, int W64, int sx, int r, int r2) {
    if (run_off[r + 1] == run_off[r] || run_off[r2 + 1] == run_off[r2]) return;
    if (W64 == 1) {
        // single-word rows: carry / mask arithmetic, a couple of iterations in practice
        const unsigned long long b2 = bits[r2];
        const unsigned long long starts2 = b2 & ~(b2 << 1);
        unsigned long long b = bits[r];
        int ida = run_off[r];
        while (b) {
            const unsigned long long run = b & ~(b + (b & (0ull - b)));                  // lowest maximal run of ones
            unsigned long long touch = b2 & (run | (run << 1) | (run >> 1));
            while (touch) {
                const int pos = __ffsll((long long)touch) - 1;
                const unsigned long long below = ~b2 & ((1ull << pos) - 1ull);          // zeros of b2 below pos
                const int bs = below ? 64 - __clzll((long long)below) : 0;               // first bit of that run
                const unsigned long long runb = b2 & ~(b2 + (1ull << bs));
                uf_union(parent, ida, run_off[r2] + __popcll(starts2 & ((1ull << bs) - 1ull)));
                touch &= ~runb;
            }
            b &= ~run; ++ida;
        }
        return;
    }
    const unsigned long long* row = bits + (size_t)r * W64;
    const unsigned long long* row2 = bits + (size_t)r2 * W64;
    // two-pointer walk over the runs of both rows (increasing x)
    int ida = run_off[r], as, ae, bs, be;
    bool ha = next_run(row, W64, sx, 0, as, ae), hb = next_run(row2, W64, sx, 0, bs, be);
    int idb = run_off[r2];
    while (ha && hb) {
        if (be < as - 1) { hb = next_run(row2, W64, sx, be + 2, bs, be); ++idb; }
        else if (bs > ae + 1) { ha = next_run(row, W64, sx, ae + 2, as, ae); ++ida; }
        else {
            uf_union(parent, ida, idb);
            if (be < ae) { hb = next_run(row2, W64, sx, be + 2, bs, be); ++idb; }
            else { ha = next_run(row, W64, sx, ae + 2, as, ae); ++ida; }
        }
    }
}

// grid = (slots, volumes), like the binarization kernel.  scratch: 8 bytes per mask byte (see the launcher).
__global__ void __launch_bounds__(CC_THREADS, 3)
largest_cc_kernel(uint8_t* __restrict__ mask, const int64_t* __restrict__ crop_off, int n_crops,
                  const int32_t* __restrict__ det_off, const int32_t* __restrict__ boxes,
                  const int32_t* __restrict__ order, const int32_t* __restrict__ n_valid,
                  int32_t* __restrict__ status, unsigned long long* __restrict__ scratch, long long total_mask_bytes) {
    __shared__ CcShared sh;
    extern __shared__ __align__(16) unsigned char cc_dyn[];
    unsigned long long* sm_bits = reinterpret_cast<unsigned long long*>(cc_dyn);
    int* sm_run_off = reinterpret_cast<int*>(sm_bits + CC_SMEM_WORDS);
    int* sm_parent = sm_run_off + CC_SMEM_ROWS + 2;
    int* sm_size = sm_parent + CC_SMEM_RUNS;
    unsigned short* sm_iv = reinterpret_cast<unsigned short*>(sm_size + CC_SMEM_RUNS);    // run interval (first x << 8) | last x
    unsigned short* sm_row = sm_iv + CC_SMEM_RUNS;                                           // row of the run
    const int slot = blockIdx.x, vol = blockIdx.y;
    const int base = det_off ? det_off[vol] : 0;
    const int n_here = det_off ? det_off[vol + 1] - base : n_crops;
    if (slot >= n_here || (n_valid && slot >= n_valid[vol])) return;
    const int inst = base + (order ? order[base + slot] : slot);
    if (status && status[inst] != 0) return;                 // skipped / failed instances keep their (empty) mask
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t off = crop_off[inst];
    const int n = (int)(crop_off[inst + 1] - off);
    const int32_t* bb = boxes + 6 * (size_t)inst;
    const int sx = bb[3] - bb[0] + 1, sy = bb[4] - bb[1] + 1;
    if (n <= 0 || sx <= 0 || sy <= 0) return;
    uint8_t* m = mask + off;
    const uint8_t* mask_safe_end = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(mask + total_mask_bytes) & ~(uintptr_t)7);
    const int rows = n / sx;
    const int W64 = (sx + 63) >> 6;

    // ---- storage: shared memory when the crop is small enough, else this crop's slice of the global scratch ----
    // scratch slice (8 bytes per voxel, 8-byte aligned because off*8 is): [bits rows*W64 u64][run_off rows+1 i32]
    // [parent n/2+1 i32][size n/2+1 i32] <= 8n bytes for sx >= 2; rows of one voxel are handled in shared memory
    // sizes only (their count is bounded by the same 8n).
    unsigned long long* gscr = scratch + off;
    const int n_words = rows * W64;
    // degenerate crops (rows of 1-3 voxels, thousands of rows) could outgrow the 8 bytes per voxel of scratch:
    // they are reported as status 6 and pasted as empty rather than silently left unfiltered
    if ((n_words > CC_SMEM_WORDS || rows > CC_SMEM_ROWS) && (size_t)n_words * 8 + (size_t)(rows + 2) * 4 > (size_t)n * 8) {
        for (int j = tid; j < n; j += CC_THREADS) m[j] = 0;
        if (tid == 0 && status) status[inst] = 6;
        return;
    }
    unsigned long long* bits = n_words <= CC_SMEM_WORDS ? sm_bits : gscr;
    int* run_off = rows <= CC_SMEM_ROWS ? sm_run_off : reinterpret_cast<int*>(gscr + n_words);
    for (int i = tid; i < n_words; i += CC_THREADS) bits[i] = 0ull;
    __syncthreads();

    // ---- 1. row bit vectors: 8 mask bytes -> 8 bits per step (SWAR non-zero test + multiply gather) -----------
    for (int j0 = tid * 8; j0 < n; j0 += CC_THREADS * 8) {
        const int nbytes = min(8, n - j0);
        const unsigned long long v = load8_unaligned(m + j0, nbytes, mask_safe_end);
        const unsigned long long t = ((((v & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full) | v) & 0x8080808080808080ull) >> 7;
        unsigned long long b8 = (t * 0x0102040810204080ull) >> 56;              // bit k = byte k non-zero
        if (nbytes < 8) b8 &= (1ull << nbytes) - 1ull;
        if (b8 == 0ull) continue;
        int row = j0 / sx, x = j0 - row * sx;
        if (x + nbytes <= sx && (x >> 6) == ((x + nbytes - 1) >> 6)) {
            atomicOr(&bits[row * W64 + (x >> 6)], b8 << (x & 63));
        } else {                                              // the 8 voxels straddle a row end or a 64-bit word
            for (int k = 0; k < nbytes; ++k) {
                if ((b8 >> k) & 1ull) atomicOr(&bits[row * W64 + (x >> 6)], 1ull << (x & 63));
                if (++x == sx) { x = 0; ++row; }
            }
        }
    }
    __syncthreads();

    // ---- 2. runs per row, exclusive prefix sum over rows -> run ids ----------------------------------------
    int carry_total = 0;
    for (int r0 = 0; r0 < rows; r0 += CC_THREADS) {
        const int r = r0 + tid;
        int cnt = 0;
        if (r < rows) {
            unsigned long long carry = 0ull;
            for (int w = 0; w < W64; ++w) {
                const unsigned long long b = bits[r * W64 + w];
                cnt += __popcll(b & ~((b << 1) | carry));
                carry = b >> 63;
            }
        }
        int incl = cnt;                                       // warp inclusive scan, then one hop through shared memory
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) sh.scan[warp] = incl;
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < CC_THREADS / 32; ++w) { const int v = sh.scan[w]; if (w < warp) before += v; total += v; }
        if (r < rows) run_off[r] = carry_total + before + incl - cnt;
        carry_total += total;
        __syncthreads();
    }
    const int T = carry_total;
    if (tid == 0) { run_off[rows] = T; sh.best = 0ull; }
    if (T == 0) {                                            // no foreground at all (the reference raises here)
        if (tid == 0 && status) status[inst] = 5;
        return;
    }
    int* parent; int* size;
    if (T <= CC_SMEM_RUNS) { parent = sm_parent; size = sm_size; }
    else if ((size_t)n_words * 8 + (size_t)(rows + 2) * 4 + (size_t)T * 8 > (size_t)n * 8) {
        __syncthreads();
        for (int j = tid; j < n; j += CC_THREADS) m[j] = 0;
        if (tid == 0 && status) status[inst] = 6;
        return;
    } else {
        int* gi = reinterpret_cast<int*>(gscr + n_words) + (rows + 2);
        parent = gi; size = gi + T;
    }
    __syncthreads();

    // rows of <= 256 voxels whose runs fit the shared-memory tables take the interval-list path; everything else
    // walks rows with the generic multi-word helpers (union_rows).
    const bool by_run = sx <= 256 && T <= CC_SMEM_RUNS && rows < 65536 && (unsigned long long)rows * (unsigned long long)sy < 0xFFFFFFFFull;
    const int sz = rows / sy;
    if (by_run) {
        cc_unions_intervals(bits, run_off, parent, size, sm_iv, sm_row, rows, sx, sy, W64, T);
    } else {
    // generic rows (wider than 128 voxels, or more runs than the shared-memory tables hold): nodes, then
    // (i) inside every z slice one thread walks the rows top to bottom and joins each row with the previous one
    // (a sequential union-find with path halving keeps the trees of a slice shallow), (ii) every row is joined, in
    // parallel, with the three rows it touches in the slice above.
    for (int r = tid; r < rows; r += CC_THREADS) {
        const unsigned long long* row = bits + (size_t)r * W64;
        int id = run_off[r], p = 0, xs, xe;
        while (next_run(row, W64, sx, p, xs, xe)) { parent[id] = id; size[id] = xe - xs + 1; ++id; p = xe + 2; }
    }
    __syncthreads();
    for (int z = tid; z < sz; z += CC_THREADS)
        for (int y = 1; y < sy; ++y) union_rows(bits, run_off, parent, W64, sx, z * sy + y, z * sy + y - 1);
    __syncthreads();
    for (int r = tid; r < rows; r += CC_THREADS) {
        const int z = r / sy, y = r - z * sy;
        if (z == 0 || run_off[r + 1] == run_off[r]) continue;
#pragma unroll 1
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = y + dy;
            if (yy >= 0 && yy < sy) union_rows(bits, run_off, parent, W64, sx, r, (z - 1) * sy + yy);
        }
    }
    }
    __syncthreads();
    // ---- 4a. sizes at the roots ---------------------------------------------------------------------------
    for (int t0 = 0; t0 < T; t0 += CC_THREADS) {
        const int t = t0 + tid;
        int root = -1, add = 0;
        if (t < T) { root = uf_find(parent, t); if (root != t) add = size[t]; else root = -1; }   // size[t] of a non-root is final
        // typical mask: one big component -> the whole warp adds to the same root; aggregate that case
        const int r0 = __shfl_sync(0xffffffffu, root, 0);
        if (__all_sync(0xffffffffu, root == r0 || root < 0) && r0 >= 0) {
            const int sum = __reduce_add_sync(0xffffffffu, add);
            if (lane == 0) atomicAdd(&size[r0], sum);
        } else if (root >= 0) atomicAdd(&size[root], add);
    }
    __syncthreads();
    // ---- 4b. largest component; ties -> later first voxel (larger root id) --------------------------------
    {
        unsigned long long bestk = 0ull;
        for (int t = tid; t < T; t += CC_THREADS)
            if (parent[t] == t) {
                const unsigned long long k = ((unsigned long long)(unsigned)size[t] << 32) | (unsigned)t;
                bestk = k > bestk ? k : bestk;
            }
#pragma unroll
        for (int o = 16; o; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xffffffffu, bestk, o); bestk = v > bestk ? v : bestk; }
        if (lane == 0) atomicMax(&sh.best, bestk);
        (void)warp;
    }
    __syncthreads();
    const int best_root = (int)(sh.best & 0xFFFFFFFFull);
    // ---- 4c. clear the runs of every other component ------------------------------------------------------
    if (by_run) {
        for (int id = tid; id < T; id += CC_THREADS) {
            if (uf_find(parent, id) == best_root) continue;
            const unsigned int a = sm_iv[id];
            uint8_t* dst = m + (size_t)sm_row[id] * sx;
            for (int x = (int)(a >> 8); x <= (int)(a & 255u); ++x) dst[x] = 0;
        }
    } else {
        for (int r = tid; r < rows; r += CC_THREADS) {
            const unsigned long long* row = bits + (size_t)r * W64;
            int id = run_off[r], p = 0, xs, xe;
            while (next_run(row, W64, sx, p, xs, xe)) {
                if (uf_find(parent, id) != best_root) for (int x = xs; x <= xe; ++x) m[(size_t)r * sx + x] = 0;
                ++id; p = xe + 2;
            }
        }
    }
}

}  // namespace b200seg

using namespace b200seg;

extern "C" size_t b200seg_largest_cc_workspace_bytes(long long total_mask_bytes) {
    if (total_mask_bytes < 0) return 256;
    return (size_t)total_mask_bytes * 8 + 64 * 8 + 256;
}

extern "C" int b200seg_largest_cc_dev(uint8_t* masks, const int64_t* crop_off, long long total_mask_bytes,
                                      int n_volumes, const int32_t* det_off, int n_max, const int32_t* boxes,
                                      const int32_t* order, const int32_t* n_valid, int32_t* status,
                                      void* workspace, size_t workspace_bytes, b200seg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200_CHECK_ARG(n_max >= 0 && n_volumes >= 0 && total_mask_bytes >= 0, "largest_cc: bad sizes");
    if (n_max == 0 || n_volumes == 0) return 0;
    B200_CHECK_ARG(n_volumes == 1 || det_off, "largest_cc: det_off is required for more than one volume");
    B200_CHECK_ARG(n_volumes <= 65535, "largest_cc: too many volumes in one call");
    B200_CHECK_ARG(masks && crop_off && boxes && workspace, "largest_cc: null pointer");
    if (workspace_bytes < b200seg_largest_cc_workspace_bytes(total_mask_bytes)) {
        set_error("largest_cc: workspace too small");
        return B200SEG_EWORKSPACE;
    }
    unsigned long long* scratch = (unsigned long long*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    static bool attr_set = false;
    if (!attr_set) {
        B200_CUDA(cudaFuncSetAttribute(largest_cc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CC_DYN_BYTES));
        attr_set = true;
    }
    dim3 grid(n_max, n_volumes);
    largest_cc_kernel<<<grid, CC_THREADS, CC_DYN_BYTES, stream>>>(masks, crop_off, n_max, det_off, boxes, order, n_valid, status, scratch, total_mask_bytes);
    B200_LAUNCH_CHECK("largest_cc_kernel");
    return 0;
}
