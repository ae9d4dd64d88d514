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
#include <stdlib.h>

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
// ---------------------------------------------------------------------------------------------------------
// Fast path: one small CTA per instance, bit-parallel flood fill from the crop centre.
//
// The Otsu mask of a detected instance is almost always one blob around the crop centre plus a few specks.  For
// crops with rows of <= 64 voxels and <= 32 z planes, lane z of a warp owns plane z: a row is one 64-bit word, filling
// along x is a carry trick ( up = m & ~(m + s) | s, the same on the bit-reversed words for the other direction ), and
// warp 0 sweeps the planes forwards and backwards along y in lockstep, the lanes exchanging the fill of the
// neighbouring planes by shuffle (work-efficient, but a dependent chain).  After every sweep pair all four warps test,
// on independent rows, whether anything is still reachable; a convex blob is complete after one pair.  The other
// phases (mask bytes -> row words, counting, clearing) are spread over the four warps.  The converged fill is
// exactly the 26-connected component of the seed.  If it holds MORE THAN HALF of the foreground it is the unique
// largest component: everything else is cleared and the instance is marked done (CC_DONE_BIT in status) so that the
// general kernel below only clears the flag.  Anything else -- wide or deep crops, an empty centre row, a seed component
// without the majority, a maze that needs more than CCF_MAX_PAIRS sweep pairs -- is left untouched for the general
// kernel, which also owns the tie rule and the status codes.
constexpr int CC_DONE_BIT = 0x40000000;
// path statistics since the last reset (b200seg_largest_cc_path_counts): 0 filled by the fast path, 1 geometry not
// supported, 2 no foreground, 3 empty centre row, 4 not converged, 5 no majority
__device__ unsigned long long ccf_counts[8];
__device__ __forceinline__ void ccf_count(int tid, int k) { if (tid == 0) atomicAdd(&ccf_counts[k], 1ull); }
constexpr int CCF_ROWS = 1152;                 // rows per instance kept in shared memory (mask + fill words)
constexpr int CCF_MAX_PAIRS = 6;                              // forward + backward sweeps before a (maze-like) mask is handed on

__device__ __forceinline__ unsigned long long ccf_fill_row(unsigned long long s, unsigned long long m) {
    // every run of m that contains a bit of s (s subset of m)
    const unsigned long long up = (m & ~(m + s)) | s;
    const unsigned long long rm = __brevll(m), rs = __brevll(s);
    return up | __brevll((rm & ~(rm + rs)) | rs);
}

// (volume, slot) work item -> instance index, or -1 when the slot is empty / beyond the valid detections
__device__ __forceinline__ int cc_inst_of(int slot, int vol, int n_crops, const int32_t* __restrict__ det_off,
                                          const int32_t* __restrict__ order, const int32_t* __restrict__ n_valid) {
    const int base = det_off ? det_off[vol] : 0;
    const int n_here = det_off ? det_off[vol + 1] - base : n_crops;
    if (slot >= n_here || (n_valid && slot >= n_valid[vol])) return -1;
    return base + (order ? order[base + slot] : slot);
}

// one sweep step of the fill for row y of every plane (lane = plane); returns the row's new fill word
__device__ __forceinline__ unsigned long long ccf_seeds(unsigned long long own3, unsigned long long mm) {
    const unsigned long long nb = own3 | __shfl_up_sync(0xFFFFFFFFu, own3, 1) | __shfl_down_sync(0xFFFFFFFFu, own3, 1);
    return (nb | (nb << 1) | (nb >> 1)) & mm;
}

template <int CCF_WARPS>
__global__ void __launch_bounds__(CCF_WARPS * 32)
largest_cc_fill_kernel(uint8_t* __restrict__ mask, const int64_t* __restrict__ crop_off, int n_crops,
                       const int32_t* __restrict__ det_off, const int32_t* __restrict__ boxes,
                       const int32_t* __restrict__ order, const int32_t* __restrict__ n_valid,
                       int32_t* __restrict__ status, unsigned long long* __restrict__ scratch, long long total_mask_bytes) {
    extern __shared__ __align__(16) unsigned char ccf_dyn[];
    __shared__ int s_sum[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int inst = cc_inst_of(blockIdx.x, blockIdx.y, n_crops, det_off, order, n_valid);
    if (inst < 0 || status[inst] != 0) return;                // CTA-uniform exits
    const int64_t off = crop_off[inst];
    const int n = (int)(crop_off[inst + 1] - off);
    const int32_t* bb = boxes + 6 * (size_t)inst;
    const int sx = bb[3] - bb[0] + 1, sy = bb[4] - bb[1] + 1;
    if (n <= 0 || sx <= 0 || sy <= 0 || sx > 64) { ccf_count(tid, 1); return; }
    const int rows = n / sx, sz = rows / sy;
    if (sz > 32 || sz < 1 || sz * sy != rows || rows * sx != n) { ccf_count(tid, 1); return; }
    // plane z keeps its rows at [z * ps, z * ps + sy): an odd stride spreads the lanes (= planes) over the banks
    const int ps = sy | 1, words = sz * ps;
    unsigned long long* sm_m;
    if (words <= CCF_ROWS) sm_m = reinterpret_cast<unsigned long long*>(ccf_dyn);
    else if ((size_t)words * 2 <= (size_t)n) sm_m = scratch + off;            // large crop: its slice of the global scratch (8 B per voxel)
    else { ccf_count(tid, 1); return; }
    unsigned long long* sm_f = sm_m + words;
    uint8_t* m = mask + off;
    const uint8_t* mask_safe_end = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(mask + total_mask_bytes) & ~(uintptr_t)7);
    if (tid < 2) s_sum[tid] = 0;
    __syncthreads();

    // ---- row words from the mask bytes (all warps); foreground count -------------------------------------------
    int fg = 0;
    for (int r = tid; r < rows; r += CCF_WARPS * 32) {
        const uint8_t* src = m + (size_t)r * sx;
        unsigned long long word = 0ull;
#pragma unroll
        for (int x0 = 0; x0 < 64; x0 += 8) {
            if (x0 >= sx) break;
            const int nbytes = min(8, sx - x0);
            const unsigned long long v = load8_unaligned(src + x0, nbytes, mask_safe_end);
            const unsigned long long t = ((((v & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full) | v) & 0x8080808080808080ull) >> 7;
            unsigned long long b8 = (t * 0x0102040810204080ull) >> 56;
            if (nbytes < 8) b8 &= (1ull << nbytes) - 1ull;
            word |= b8 << x0;
        }
        const int z = r / sy, w = z * ps + (r - z * sy);
        sm_m[w] = word;
        sm_f[w] = 0ull;
        fg += __popcll(word);
    }
    fg = __reduce_add_sync(0xFFFFFFFFu, fg);
    if (lane == 0 && fg) atomicAdd(&s_sum[0], fg);
    __syncthreads();
    fg = s_sum[0];
    if (fg == 0) { ccf_count(tid, 2); return; }               // status 5 is the general kernel's to report
    // ---- seed: the foreground voxel of the centre row nearest to the centre column ---------------------------
    const int yc = sy >> 1, rc = (sz >> 1) * ps + yc, xc = sx >> 1;
    const unsigned long long mc = sm_m[rc];
    if (mc == 0ull) { ccf_count(tid, 3); return; }
    if (tid == 0) {
        const unsigned long long hi = mc >> xc, lo = xc ? (mc & ((1ull << xc) - 1ull)) : 0ull;
        const int dh = hi ? __ffsll((long long)hi) - 1 : 1000, dl = lo ? xc - (63 - __clzll((long long)lo)) : 1000;
        sm_f[rc] = ccf_fill_row(1ull << (dh <= dl ? xc + dh : xc - dl), mc);
    }
    __syncthreads();
    // ---- fill: lane z owns plane z (idle lanes read plane 0 through an empty mask).  Warp 0 sweeps forwards and
    //      backwards along y (a dependent chain); whether anything is still reachable is then checked by all warps on
    //      independent rows, so a converged fill costs one sweep pair, not two. -----------------------------------------
    const bool act = lane < sz;
    const unsigned long long* pm = sm_m + (size_t)(act ? lane : 0) * ps;
    unsigned long long* pf = sm_f + (size_t)(act ? lane : 0) * ps;
    const unsigned long long live = act ? ~0ull : 0ull;
    for (int pairs = 0;; ++pairs) {
        if (warp == 0) {
#pragma unroll 1
            for (int dir = 0; dir < 2; ++dir) {
                const int step = dir ? -1 : 1;
                // nothing above the seed row can change before the first forward sweep reaches it
                int y = dir ? sy - 1 : (pairs == 0 ? max(yc - 1, 0) : 0);
                const int nsteps = dir ? sy : sy - y;
                unsigned long long prev = (y - step >= 0 && y - step < sy) ? pf[y - step] & live : 0ull;
                unsigned long long cur = pf[y] & live;
                unsigned long long next = (y + step >= 0 && y + step < sy) ? pf[y + step] & live : 0ull;
                unsigned long long mm = pm[y] & live;
                for (int k = 0; k < nsteps; ++k, y += step) {
                    // rows ahead are only ever written by this lane, later: safe to fetch them early
                    const int y2 = y + 2 * step, y1 = y + step;
                    const unsigned long long next2 = (y2 >= 0 && y2 < sy) ? pf[y2] & live : 0ull;
                    const unsigned long long mm1 = (y1 >= 0 && y1 < sy) ? pm[y1] & live : 0ull;
                    const unsigned long long seeds = ccf_seeds(prev | cur | next, mm);
                    unsigned long long fnew = cur;
                    if (seeds & ~cur) {                           // something new reaches this row
                        fnew = ccf_fill_row(seeds, mm);
                        pf[y] = fnew;
                    }
                    prev = fnew; cur = next; next = next2; mm = mm1;
                }
            }
        }
        __syncthreads();
        int open_rows = 0;
        for (int y = warp; y < sy; y += CCF_WARPS) {
            const unsigned long long cur = pf[y] & live;
            const unsigned long long own3 = cur | (y > 0 ? pf[y - 1] & live : 0ull) | (y + 1 < sy ? pf[y + 1] & live : 0ull);
            open_rows |= (ccf_seeds(own3, pm[y] & live) & ~cur) != 0ull;
        }
        if (!__syncthreads_or(open_rows)) break;
        if (pairs + 1 >= CCF_MAX_PAIRS) { ccf_count(tid, 4); return; }      // maze-like mask: the general kernel takes it
    }
    // ---- majority test; clear everything outside the fill ---------------------------------------------------
    int cnt = 0;
    for (int r = tid; r < rows; r += CCF_WARPS * 32) {
        const int z = r / sy;
        cnt += __popcll(sm_f[z * ps + (r - z * sy)]);
    }
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if (lane == 0 && cnt) atomicAdd(&s_sum[1], cnt);
    __syncthreads();
    if (2 * s_sum[1] <= fg) { ccf_count(tid, 5); return; }    // not a strict majority: the general kernel decides
    for (int r = tid; r < rows; r += CCF_WARPS * 32) {
        const int z = r / sy, w = z * ps + (r - z * sy);
        unsigned long long diff = sm_m[w] & ~sm_f[w];
        uint8_t* dst = m + (size_t)r * sx;
        while (diff) {
            const int x = __ffsll((long long)diff) - 1;
            dst[x] = 0;
            diff &= diff - 1ull;
        }
    }
    if (tid == 0) status[inst] = CC_DONE_BIT;
    ccf_count(tid, 0);
}

// one instance, whole CTA; every return below is taken by all threads of the CTA
__device__ __forceinline__ void
cc_instance(CcShared& sh, unsigned char* cc_dyn, const int inst,
            uint8_t* __restrict__ mask, const int64_t* __restrict__ crop_off, const int32_t* __restrict__ boxes,
            int32_t* __restrict__ status, unsigned long long* __restrict__ scratch, long long total_mask_bytes, const int tie_first) {
    unsigned long long* sm_bits = reinterpret_cast<unsigned long long*>(cc_dyn);
    int* sm_run_off = reinterpret_cast<int*>(sm_bits + CC_SMEM_WORDS);
    int* sm_parent = sm_run_off + CC_SMEM_ROWS + 2;
    int* sm_size = sm_parent + CC_SMEM_RUNS;
    unsigned short* sm_iv = reinterpret_cast<unsigned short*>(sm_size + CC_SMEM_RUNS);    // run interval (first x << 8) | last x
    unsigned short* sm_row = sm_iv + CC_SMEM_RUNS;                                           // row of the run
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
    // ---- 4b. largest component; ties -> later first voxel (larger root id), or the earlier one with tie_first ----
    {
        unsigned long long bestk = 0ull;
        for (int t = tid; t < T; t += CC_THREADS)
            if (parent[t] == t) {
                const unsigned long long k = ((unsigned long long)(unsigned)size[t] << 32) | (tie_first ? 0xFFFFFFFFu - (unsigned)t : (unsigned)t);
                bestk = k > bestk ? k : bestk;
            }
#pragma unroll
        for (int o = 16; o; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xffffffffu, bestk, o); bestk = v > bestk ? v : bestk; }
        if (lane == 0) atomicMax(&sh.best, bestk);
        (void)warp;
    }
    __syncthreads();
    const int best_root = tie_first ? (int)(0xFFFFFFFFu - (unsigned)(sh.best & 0xFFFFFFFFull)) : (int)(sh.best & 0xFFFFFFFFull);
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

// Persistent CTAs claim batches of (volume, slot) work items from a counter.  One thread per item resolves it: an
// instance the fill kernel finished only has its flag cleared; instances that still need the union-find path are
// collected and processed by the whole CTA one after the other.
constexpr int CC_BATCH = 16;
__global__ void __launch_bounds__(CC_THREADS, 3)
largest_cc_kernel(uint8_t* __restrict__ mask, const int64_t* __restrict__ crop_off, int n_crops, int n_work,
                  const int32_t* __restrict__ det_off, const int32_t* __restrict__ boxes,
                  const int32_t* __restrict__ order, const int32_t* __restrict__ n_valid,
                  int32_t* __restrict__ status, unsigned long long* __restrict__ scratch, long long total_mask_bytes,
                  unsigned int* __restrict__ work_counter, int tie_first) {
    __shared__ CcShared sh;
    __shared__ int s_first, s_n, s_list[CC_BATCH];
    extern __shared__ __align__(16) unsigned char cc_dyn[];
    while (true) {
        __syncthreads();                                     // the previous batch is done with shared memory
        if (threadIdx.x == 0) { s_first = (int)atomicAdd(work_counter, (unsigned)CC_BATCH); s_n = 0; }
        __syncthreads();
        const int first = s_first;
        if (first >= n_work) break;
        if (threadIdx.x < CC_BATCH && first + threadIdx.x < n_work) {
            const int w = first + threadIdx.x;
            const int inst = cc_inst_of(w % n_crops, w / n_crops, n_crops, det_off, order, n_valid);
            if (inst >= 0) {
                const int st = status ? status[inst] : 0;
                if (st & CC_DONE_BIT) status[inst] = st & ~CC_DONE_BIT;         // already filtered by largest_cc_fill_kernel
                else if (st == 0) s_list[atomicAdd(&s_n, 1)] = inst;           // skipped / failed instances keep their (empty) mask
            }
        }
        __syncthreads();
        const int cnt = s_n;
        for (int k = 0; k < cnt; ++k) {
            cc_instance(sh, cc_dyn, s_list[k], mask, crop_off, boxes, status, scratch, total_mask_bytes, tie_first);
            __syncthreads();
        }
    }
}

}  // namespace b200seg

using namespace b200seg;

extern "C" int b200seg_largest_cc_path_counts(long long* counts, int reset) {
    B200_CHECK_ARG(counts, "largest_cc_path_counts: null pointer");
    unsigned long long h[8];
    B200_CUDA(cudaMemcpyFromSymbol(h, ccf_counts, sizeof(h)));
    for (int i = 0; i < 8; ++i) counts[i] = (long long)h[i];
    if (reset) {
        memset(h, 0, sizeof(h));
        B200_CUDA(cudaMemcpyToSymbol(ccf_counts, h, sizeof(h)));
    }
    return 0;
}

extern "C" size_t b200seg_largest_cc_workspace_bytes(long long total_mask_bytes) {
    if (total_mask_bytes < 0) return 256;
    return (size_t)total_mask_bytes * 8 + 64 * 8 + 256;
}

extern "C" int b200seg_largest_cc_dev(uint8_t* masks, const int64_t* crop_off, long long total_mask_bytes,
                                      int n_volumes, const int32_t* det_off, int n_max, const int32_t* boxes,
                                      const int32_t* order, const int32_t* n_valid, int32_t* status,
                                      void* workspace, size_t workspace_bytes, b200seg_stream_t stream_) {
    return b200seg_largest_cc_ex_dev(masks, crop_off, total_mask_bytes, n_volumes, det_off, n_max, boxes, order, n_valid, status, 0,
                                     workspace, workspace_bytes, stream_);
}

extern "C" int b200seg_largest_cc_ex_dev(uint8_t* masks, const int64_t* crop_off, long long total_mask_bytes,
                                         int n_volumes, const int32_t* det_off, int n_max, const int32_t* boxes,
                                         const int32_t* order, const int32_t* n_valid, int32_t* status, int tie_first,
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
    static OncePerDevice attr;
    int attr_dev;
    if (attr.needed(&attr_dev)) {
        B200_CUDA(cudaFuncSetAttribute(largest_cc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CC_DYN_BYTES));
        attr.mark(attr_dev);
    }
    if (status) {                                            // warp-per-instance flood fill first; it marks what it finished
        const size_t fill_smem = (size_t)2 * CCF_ROWS * 8;
        static OncePerDevice fill_attr;
        int fill_dev;
        if (fill_attr.needed(&fill_dev)) {
            B200_CUDA(cudaFuncSetAttribute(largest_cc_fill_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fill_smem));
            B200_CUDA(cudaFuncSetAttribute(largest_cc_fill_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fill_smem));
            B200_CUDA(cudaFuncSetAttribute(largest_cc_fill_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fill_smem));
            B200_CUDA(cudaFuncSetAttribute(largest_cc_fill_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fill_smem));
            fill_attr.mark(fill_dev);
        }
        dim3 fgrid(n_max, n_volumes);
        static const int ccf_warps = getenv("B200SEG_CCF_WARPS") ? atoi(getenv("B200SEG_CCF_WARPS")) : 4;
        if (ccf_warps == 1) largest_cc_fill_kernel<1><<<fgrid, 32, fill_smem, stream>>>(masks, crop_off, n_max, det_off, boxes, order, n_valid, status, scratch, total_mask_bytes);
        else if (ccf_warps == 2) largest_cc_fill_kernel<2><<<fgrid, 64, fill_smem, stream>>>(masks, crop_off, n_max, det_off, boxes, order, n_valid, status, scratch, total_mask_bytes);
        else if (ccf_warps == 8) largest_cc_fill_kernel<8><<<fgrid, 256, fill_smem, stream>>>(masks, crop_off, n_max, det_off, boxes, order, n_valid, status, scratch, total_mask_bytes);
        else largest_cc_fill_kernel<4><<<fgrid, 128, fill_smem, stream>>>(masks, crop_off, n_max, det_off, boxes, order, n_valid, status, scratch, total_mask_bytes);
        B200_LAUNCH_CHECK("largest_cc_fill_kernel");
    }
    unsigned int* work_counter = reinterpret_cast<unsigned int*>(scratch + total_mask_bytes);       // inside the 64 spare words
    B200_CUDA(cudaMemsetAsync(work_counter, 0, sizeof(unsigned int), stream));
    const long long n_work = (long long)n_max * n_volumes;
    B200_CHECK_ARG(n_work < (1ll << 31), "largest_cc: too many instances");
    const long long batches = (n_work + CC_BATCH - 1) / CC_BATCH;
    const long long ctas = batches < 3ll * num_sms() ? batches : 3ll * num_sms();
    largest_cc_kernel<<<(unsigned)ctas, CC_THREADS, CC_DYN_BYTES, stream>>>(masks, crop_off, n_max, (int)n_work, det_off, boxes, order, n_valid, status, scratch,
                                                                            total_mask_bytes, work_counter, tie_first ? 1 : 0);
    B200_LAUNCH_CHECK("largest_cc_kernel");
    return 0;
}
