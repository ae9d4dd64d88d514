// paste.cu -- label paste-back (replaces the inline numpy of tools/binarization_soma.py:100-104
// and tools/binarization_nuclei.py:141-149).
//
// Reference semantics: instances are visited in order; instance i writes its label where its mask
// is set AND the label volume is still 0 ("first come wins"); afterwards `mask_id in np.unique(seg)`
// decides whether the instance survives.  The serial loop is replaced by an owner-computes gather:
// every voxel of the label volume is produced exactly once as "label of the first instance, in
// visit order, whose box covers the voxel and whose mask is set there" -- no atomics on the volume,
// no pre-clear, the zero fill is folded into the single 128-bit streaming store per 8 voxels.
//
// Two launches for a whole batch of volumes:
//   paste_bin_kernel     one CTA per (visited instance, volume): writes its box record in visit order
//                        and ORs its visit rank into the bitmap of every 4x16x64-voxel tile its box
//                        touches (bitmap = ceil(n/32) words per tile; OR is order independent).
//   paste_labels_kernel  persistent CTAs walk the tiles of all volumes; each warp fetches the tile's
//                        bitmap with ONE coalesced load (lane = word, prefetched one tile ahead), ballots
//                        the non-empty words and visits the set bits in increasing rank = visit order.
//                        No __syncthreads, no shared memory.  A thread owns 8 x-voxels on the 4 planes of
//                        the tile: an empty tile costs four 16-byte stores per thread; a covering
//                        instance costs four 8-byte mask-row reads (two aligned 64-bit loads + funnel
//                        shift) that are independent of each other.
// HBM traffic = 2 B/voxel written + the mask bytes of covered voxels read (L2-resident crops).
#include "common.cuh"

namespace b200seg {

constexpr int PT_X = 64, PT_Y = 16, PT_Z = 4;
constexpr int PASTE_THREADS = (PT_X / 8) * PT_Y;          // 128: one thread = 8 consecutive x voxels on PT_Z planes

struct __align__(8) VBox {          // box record in visit order
    int x1, y1, z1, x2, y2, z2;
    int sx, sy;
    long long moff;                 // offset of the mask crop
    long long mend;                 // one past its last byte
};

struct PasteGeom {
    int S, H, W;
    int tiles_x, tiles_y, tiles_z;
    int words;                      // bitmap words per tile
    int n_max;                      // slots per volume (vbox stride)
};

// grid (n_max, n_volumes)
__global__ void __launch_bounds__(128)
paste_bin_kernel(PasteGeom g, const int32_t* __restrict__ det_off, const int32_t* __restrict__ boxes,
                 const int64_t* __restrict__ mask_off, const int32_t* __restrict__ order,
                 const int32_t* __restrict__ n_valid, VBox* __restrict__ vbox_all, uint32_t* __restrict__ tile_bits_all) {
    const int slot = blockIdx.x, vol = blockIdx.y;
    const int base = det_off ? det_off[vol] : 0;
    const int n_here = det_off ? det_off[vol + 1] - base : g.n_max;
    const int nv = n_valid ? min(n_here, n_valid[vol]) : n_here;
    if (slot >= nv) return;
    const int inst = base + (order ? order[base + slot] : slot);
    const int32_t* b = boxes + 6 * (size_t)inst;
    const int x1 = b[0], y1 = b[1], z1 = b[2], x2 = b[3], y2 = b[4], z2 = b[5];
    if (threadIdx.x == 0) {
        VBox v;
        v.x1 = x1; v.y1 = y1; v.z1 = z1; v.x2 = x2; v.y2 = y2; v.z2 = z2;
        v.sx = x2 - x1 + 1; v.sy = y2 - y1 + 1; v.moff = mask_off[inst];
        v.mend = v.moff + (long long)v.sx * v.sy * (z2 - z1 + 1);
        vbox_all[(size_t)vol * g.n_max + slot] = v;
    }
    // tiles touched by the (volume-clipped) box
    const int cx1 = max(x1, 0), cy1 = max(y1, 0), cz1 = max(z1, 0);
    const int cx2 = min(x2, g.W - 1), cy2 = min(y2, g.H - 1), cz2 = min(z2, g.S - 1);
    if (cx2 < cx1 || cy2 < cy1 || cz2 < cz1) return;
    const int tx1 = cx1 / PT_X, tx2 = cx2 / PT_X, ty1 = cy1 / PT_Y, ty2 = cy2 / PT_Y, tz1 = cz1 / PT_Z, tz2 = cz2 / PT_Z;
    const int nx = tx2 - tx1 + 1, ny = ty2 - ty1 + 1, nz = tz2 - tz1 + 1;
    const uint32_t bit = 1u << (slot & 31);
    const int w = slot >> 5;
    const size_t ntiles = (size_t)g.tiles_x * g.tiles_y * g.tiles_z;
    uint32_t* tile_bits = tile_bits_all + (size_t)vol * ntiles * g.words;
    for (int t = threadIdx.x; t < nx * ny * nz; t += blockDim.x) {
        const int tx = tx1 + t % nx, ty = ty1 + (t / nx) % ny, tz = tz1 + t / (nx * ny);
        const size_t tile = ((size_t)tz * g.tiles_y + ty) * g.tiles_x + tx;
        atomicOr(&tile_bits[tile * g.words + w], bit);
    }
}

// 8 mask bytes starting at p (any alignment) as two little-endian 32-bit words: three aligned 32-bit loads and
// two funnel shifts (the caller guarantees that the three words lie inside the mask buffer)
__device__ __forceinline__ uint2 load8_fs(const uint8_t* p) {
    const unsigned int mis = (unsigned int)(reinterpret_cast<uintptr_t>(p) & 3);
    const unsigned int* q = reinterpret_cast<const unsigned int*>(p - mis);
    const unsigned int w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
    return make_uint2(__funnelshift_r(w0, w1, mis * 8), __funnelshift_r(w1, w2, mis * 8));
}
// 0xFF in every byte of x that is non-zero, 0x00 elsewhere
__device__ __forceinline__ unsigned int nonzero_bytes(unsigned int x) {
    unsigned int t = ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x;       // bit 7 of each byte = byte != 0
    unsigned int d;
    asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(d) : "r"(t));       // replicate each byte's sign bit over the byte
    return d;
}
// bytes (lo, lo+1) of m, each 0x00 / 0xFF, widened to two 16-bit lanes 0x0000 / 0xFFFF
__device__ __forceinline__ unsigned int widen01(unsigned int m) { return __byte_perm(m, 0u, 0x1100); }
__device__ __forceinline__ unsigned int widen23(unsigned int m) { return __byte_perm(m, 0u, 0x3322); }

// LINES (single volume, W % 32 == 0, 16-byte aligned seg): besides the dense label volume the kernel emits the volume's
// non-zero 64-byte lines (32 labels: the four 8-voxel groups of a quad of lanes) as (line index, payload) pairs in no
// particular order -- the form the host batch entry point downloads (host_batch.cu).  One global atomic per warp and tile
// that holds a label; `lines.count` keeps counting past `lines.cap`.
template <bool LINES>
__global__ void __launch_bounds__(PASTE_THREADS, 8)
paste_labels_kernel(uint16_t* __restrict__ seg_all, PasteGeom g, int n_volumes, const uint16_t* __restrict__ ids,
                    const uint8_t* __restrict__ masks, const VBox* __restrict__ vbox_all,
                    const uint32_t* __restrict__ tile_bits_all, uint8_t* __restrict__ survive_all,
                    const int32_t* __restrict__ det_off, int vec_ok, int vec_mask_ok, PasteLines lines) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int lx = tid % (PT_X / 8), ly = tid / (PT_X / 8);
    const int ntiles = g.tiles_x * g.tiles_y * g.tiles_z;
    const int total_tiles = ntiles * n_volumes;
    const int txy = g.tiles_x * g.tiles_y;
    const size_t V = (size_t)g.S * g.H * g.W;
    int gt = blockIdx.x;
    // tile coordinates (tx, ty, tz, vol) of gt, advanced by gridDim.x tiles per step in mixed radix: the
    // divisions happen once per kernel, not once per tile
    int tx, ty, tz, vol, dtx, dty, dtz, dvol;
    {
        int r = gt;
        vol = r / ntiles; r -= vol * ntiles; tz = r / txy; r -= tz * txy; ty = r / g.tiles_x; tx = r - ty * g.tiles_x;
        r = gridDim.x;
        dvol = r / ntiles; r -= dvol * ntiles; dtz = r / txy; r -= dtz * txy; dty = r / g.tiles_x; dtx = r - dty * g.tiles_x;
    }
    int surv_off = det_off ? det_off[vol] : 0, surv_vol = vol;
    // software prefetch of the next tile's first 32 bitmap words (takes the load off the per-tile chain)
    uint32_t next_word = (gt < total_tiles && lane < g.words) ? __ldg(tile_bits_all + (size_t)gt * g.words + lane) : 0u;
    for (; gt < total_tiles; gt += gridDim.x) {
        const uint32_t first_word = next_word;
        const int ngt = gt + gridDim.x;
        next_word = (ngt < total_tiles && lane < g.words) ? __ldg(tile_bits_all + (size_t)ngt * g.words + lane) : 0u;
        const int x = tx * PT_X + lx * 8, y = ty * PT_Y + ly, z0 = tz * PT_Z;
        const int cur_vol = vol;
        {   // advance to the tile of the next iteration
            tx += dtx; int c = tx >= g.tiles_x; tx -= c ? g.tiles_x : 0;
            ty += dty + c; c = ty >= g.tiles_y; ty -= c ? g.tiles_y : 0;
            tz += dtz + c; c = tz >= g.tiles_z; tz -= c ? g.tiles_z : 0;
            vol += dvol + c;
        }
        const bool inside = x < g.W && y < g.H;
        // per plane: 8 labels as four packed words (two uint16 each) + a byte mask of the voxels already labelled
        unsigned int lab[PT_Z][4], filled[PT_Z][2];
#pragma unroll
        for (int p = 0; p < PT_Z; ++p) {
#pragma unroll
            for (int k = 0; k < 4; ++k) lab[p][k] = 0u;
            filled[p][0] = filled[p][1] = 0u;
        }
        const uint32_t* bits = tile_bits_all + (size_t)gt * g.words;
        const VBox* vbox = vbox_all + (size_t)cur_vol * g.n_max;
        if (cur_vol != surv_vol) { surv_vol = cur_vol; surv_off = det_off ? det_off[cur_vol] : 0; }
        uint8_t* survive = survive_all + surv_off;
        for (int w0 = 0; w0 < g.words; w0 += 32) {
            const uint32_t myword = w0 == 0 ? first_word : ((w0 + lane < g.words) ? __ldg(bits + w0 + lane) : 0u);
            unsigned nzw = __ballot_sync(0xffffffffu, myword != 0u);
            while (nzw) {                                           // warp-uniform
                const int wl = __ffs(nzw) - 1;
                nzw &= nzw - 1;
                uint32_t m = __shfl_sync(0xffffffffu, myword, wl);
                while (m) {
                    const int rank = (w0 + wl) * 32 + (__ffs(m) - 1);
                    m &= m - 1;
                    const VBox vb = vbox[rank];                    // uniform address: one broadcast transaction
                    if (!inside || y < vb.y1 || y > vb.y2 || x + 7 < vb.x1 || x > vb.x2) continue;
                    const unsigned int id = ids[rank];
                    const unsigned int id2 = id | (id << 16);
                    // bytes k in [klo,khi] of my 8-voxel group lie inside the box
                    const int klo = max(0, vb.x1 - x), khi = min(7, vb.x2 - x);
                    const unsigned long long range = (~0ull >> (8 * (7 - khi))) & (~0ull << (8 * klo));
                    const unsigned int range_lo = (unsigned int)range, range_hi = (unsigned int)(range >> 32);
                    const long long row0 = vb.moff + (long long)(y - vb.y1) * vb.sx + (x - vb.x1);   // may start before the row
                    const long long zstride = (long long)vb.sy * vb.sx;
                    unsigned int wrote = 0u;
#pragma unroll
                    for (int p = 0; p < PT_Z; ++p) {
                        const int z = z0 + p;
                        if (z < vb.z1 || z > vb.z2) continue;
                        const long long o = row0 + (long long)(z - vb.z1) * zstride;
                        uint2 mb;
                        if (vec_mask_ok && o >= vb.moff + 3 && o + 12 <= vb.mend) {
                            mb = load8_fs(masks + o);                  // the three aligned words stay inside this crop
                        } else {                                       // crop edge: byte reads inside [moff, mend)
                            unsigned int b[2] = {0u, 0u};
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                if (k >= klo && k <= khi) b[k >> 2] |= (unsigned int)masks[o + k] << (8 * (k & 3));
                            mb = make_uint2(b[0], b[1]);
                        }
                        // voxels this instance takes: mask set, inside the box, not labelled yet ("first come wins")
                        const unsigned int set_lo = nonzero_bytes(mb.x) & range_lo, set_hi = nonzero_bytes(mb.y) & range_hi;
                        const unsigned int take_lo = set_lo & ~filled[p][0], take_hi = set_hi & ~filled[p][1];
                        if ((take_lo | take_hi) == 0u) continue;
                        filled[p][0] |= take_lo; filled[p][1] |= take_hi;
                        lab[p][0] |= id2 & widen01(take_lo); lab[p][1] |= id2 & widen23(take_lo);
                        lab[p][2] |= id2 & widen01(take_hi); lab[p][3] |= id2 & widen23(take_hi);
                        wrote = 1u;
                    }
                    if (wrote) survive[rank] = 1;                  // benign race: every writer stores the same value
                }
            }
        }
        if (LINES) {
            unsigned int lm[PT_Z], total = 0u;
#pragma unroll
            for (int p = 0; p < PT_Z; ++p) {
                const bool nz = inside && z0 + p < g.S && (lab[p][0] | lab[p][1] | lab[p][2] | lab[p][3]) != 0u;
                const unsigned int m = __ballot_sync(0xffffffffu, nz);
                lm[p] = (m | (m >> 1) | (m >> 2) | (m >> 3)) & 0x11111111u;       // bit 4q: quad q holds a non-zero line
                total += __popc(lm[p]);
            }
            if (total) {                                                          // warp-uniform
                unsigned int b = 0u;
                if (lane == 0) b = atomicAdd(lines.count, total);
                b = __shfl_sync(0xffffffffu, b, 0);
                const unsigned int q0 = lane & ~3u;
                const size_t HW = (size_t)g.H * g.W;
#pragma unroll
                for (int p = 0; p < PT_Z; ++p) {
                    if (lm[p] == 0u) continue;
                    const unsigned int pos = b + __popc(lm[p] & ((1u << q0) - 1u));
                    b += __popc(lm[p]);
                    if (((lm[p] >> q0) & 1u) && pos < lines.cap) {
                        if ((lane & 3u) == 0u) lines.idx[pos] = (uint32_t)(((size_t)(z0 + p) * HW + (size_t)y * g.W + x) >> 5);
                        lines.val[(size_t)pos * 4 + (lane & 3u)] = make_uint4(lab[p][0], lab[p][1], lab[p][2], lab[p][3]);
                    }
                }
            }
        }
        if (inside) {
            const size_t HW = (size_t)g.H * g.W;
            uint16_t* dst = seg_all + (size_t)cur_vol * V + (size_t)z0 * HW + (size_t)y * g.W + x;
#pragma unroll
            for (int p = 0; p < PT_Z; ++p, dst += HW) {
                if (z0 + p >= g.S) break;
                if (vec_ok && x + 7 < g.W) {
                    st_stream_u4(dst, make_uint4(lab[p][0], lab[p][1], lab[p][2], lab[p][3]));
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) if (x + k < g.W) dst[k] = (uint16_t)(lab[p][k >> 1] >> (16 * (k & 1)));
                }
            }
        }
    }
}

static PasteGeom paste_geom(int S, int H, int W, int n_max) {
    PasteGeom g;
    g.S = S; g.H = H; g.W = W;
    g.tiles_x = (W + PT_X - 1) / PT_X; g.tiles_y = (H + PT_Y - 1) / PT_Y; g.tiles_z = (S + PT_Z - 1) / PT_Z;
    g.words = (n_max + 31) / 32;
    if (g.words < 1) g.words = 1;
    g.n_max = n_max > 0 ? n_max : 1;
    return g;
}

}  // namespace b200seg

using namespace b200seg;

extern "C" size_t b200seg_paste_labels_workspace_bytes(int n_volumes, int S, int H, int W, int n_max) {
    if (S <= 0 || H <= 0 || W <= 0 || n_max < 0 || n_volumes <= 0) return 256;
    const PasteGeom g = paste_geom(S, H, W, n_max);
    const size_t ntiles = (size_t)g.tiles_x * g.tiles_y * g.tiles_z;
    return align_up((size_t)n_volumes * ntiles * g.words * 4, 256) + align_up((size_t)n_volumes * g.n_max * sizeof(VBox), 256) + 256;
}

extern "C" int b200seg_paste_labels_dev(uint16_t* seg, int n_volumes, int S, int H, int W,
                                        const int32_t* det_off, int n_max, const int32_t* boxes,
                                        const uint16_t* ids, const uint8_t* masks, const int64_t* mask_off,
                                        const int32_t* order, const int32_t* n_valid, uint8_t* survive,
                                        void* workspace, size_t workspace_bytes, b200seg_stream_t stream_) {
    return b200seg::paste_labels_launch(seg, n_volumes, S, H, W, det_off, n_max, boxes, ids, masks, mask_off, order, n_valid, survive,
                                        workspace, workspace_bytes, (cudaStream_t)stream_, nullptr);
}

bool b200seg::paste_lines_supported(const uint16_t* seg, int n_volumes, int S, int H, int W) {
    return n_volumes == 1 && (W % 32) == 0 && ((((uintptr_t)seg) & 15) == 0) && (((size_t)S * H * W) % 8 == 0);
}

int b200seg::paste_labels_launch(uint16_t* seg, int n_volumes, int S, int H, int W,
                                 const int32_t* det_off, int n_max, const int32_t* boxes,
                                 const uint16_t* ids, const uint8_t* masks, const int64_t* mask_off,
                                 const int32_t* order, const int32_t* n_valid, uint8_t* survive,
                                 void* workspace, size_t workspace_bytes, cudaStream_t stream, const PasteLines* lines) {
    B200_CHECK_ARG(!lines || paste_lines_supported(seg, n_volumes, S, H, W), "paste_labels: line output needs one volume with W %% 32 == 0");
    B200_CHECK_ARG(S > 0 && H > 0 && W > 0 && n_max >= 0 && n_volumes >= 0, "paste_labels: bad sizes");
    if (n_volumes == 0) return 0;
    B200_CHECK_ARG(seg && workspace, "paste_labels: null seg/workspace");
    B200_CHECK_ARG(n_volumes == 1 || det_off, "paste_labels: det_off is required for more than one volume");
    B200_CHECK_ARG(n_volumes <= 65535, "paste_labels: too many volumes in one call");
    B200_CHECK_ARG(n_max == 0 || (boxes && ids && masks && mask_off && survive), "paste_labels: null pointer");
    if (workspace_bytes < b200seg_paste_labels_workspace_bytes(n_volumes, S, H, W, n_max)) {
        set_error("paste_labels: workspace too small");
        return B200SEG_EWORKSPACE;
    }
    const PasteGeom g = paste_geom(S, H, W, n_max);
    const size_t ntiles = (size_t)g.tiles_x * g.tiles_y * g.tiles_z;
    B200_CHECK_ARG(ntiles * n_volumes < (1ull << 31), "paste_labels: too many tiles");
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    uint32_t* tile_bits = (uint32_t*)ws;
    const size_t bits_bytes = (size_t)n_volumes * ntiles * g.words * 4;
    VBox* vbox = (VBox*)(ws + align_up(bits_bytes, 256));
    B200_CUDA(cudaMemsetAsync(tile_bits, 0, bits_bytes, stream));
    if (n_max > 0) {
        // single volume: survive[n_max] is cleared here; batched (det_off given): the caller clears survive[total]
        if (!det_off) B200_CUDA(cudaMemsetAsync(survive, 0, (size_t)n_max, stream));
        dim3 gb(n_max, n_volumes);
        paste_bin_kernel<<<gb, 128, 0, stream>>>(g, det_off, boxes, mask_off, order, n_valid, vbox, tile_bits);
        B200_LAUNCH_CHECK("paste_bin_kernel");
    }
    const int vec_ok = (W % 8 == 0) && ((((uintptr_t)seg) & 15) == 0) && (((size_t)S * H * W) % 8 == 0);
    static int occ = 0, occ_lines = 0;
    if (occ == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, paste_labels_kernel<false>, PASTE_THREADS, 0) != cudaSuccess || occ < 1) occ = 4;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_lines, paste_labels_kernel<true>, PASTE_THREADS, 0) != cudaSuccess || occ_lines < 1) occ_lines = 4;
    }
    long long grid = (long long)num_sms() * (lines ? occ_lines : occ);
    if (grid > (long long)(ntiles * n_volumes)) grid = (long long)(ntiles * n_volumes);
    const int vec_mask_ok = (int)((((uintptr_t)masks) & 7) == 0);
    if (lines)
        paste_labels_kernel<true><<<(unsigned)grid, PASTE_THREADS, 0, stream>>>(seg, g, n_volumes, ids, masks, vbox, tile_bits, survive,
                                                                               det_off, vec_ok, vec_mask_ok, *lines);
    else
        paste_labels_kernel<false><<<(unsigned)grid, PASTE_THREADS, 0, stream>>>(seg, g, n_volumes, ids, masks, vbox, tile_bits, survive,
                                                                                det_off, vec_ok, vec_mask_ok, PasteLines{nullptr, nullptr, nullptr, 0u});
    B200_LAUNCH_CHECK("paste_labels_kernel");
    return 0;
}
