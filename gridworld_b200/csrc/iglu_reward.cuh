// iglu_reward.cuh -- warp-cooperative Task.maximal_intersection (reference gridworld/tasks/task.py:147-161,
// argmax 121-136) for one env / one (grid, target) pair per warp.
//
// Restatement used (verified against the reference by tests/golden):
//   max_int = max over rotation k, admissible (dx, dz) of  #{(y,x,z): T_k[y,x,z] != 0 and
//             T_k[y,x,z] == G[y, x-dx, z-dz]}
// where T_k[:, z, 10-x] = T_{k-1}[:, x, z] (task.py:47-56) and the admissible set of rotation k is a
// rectangle [dx_lo, dx_hi] x [dz_lo, dz_hi] (task.py:57-72).  G and T are "difference grids"
// (values -6..6, env.py:230, 290); a value v is handled as the 4-bit class code v & 15.
//
// Formulation (bit-packed occupancy rows + vertical counters):
//   * the grid is packed per class c and (y, x) row into an 11-bit z-occupancy mask, stored
//     z-reversed (bit 10-z) as uint16 rows in shared memory;
//   * one lane owns one (rotation k, dx) pair -- "work item" -- and loops over the target blocks:
//     for a block (c, y, x_k, z_k) a single 16-bit row load + shift yields the match bits for ALL dz of
//     the rectangle at once (bit j <-> dz = dz_lo + j);
//   * the <= 11 one-bit results are added into eleven 5-bit counters packed in one 64-bit register
//     through a 2048-entry spread table in shared memory (one LDS.64 + one 64-bit add per block),
//     flushed to 32-bit counters every 31 blocks;
//   * when fewer than 17 work items exist the blocks are split over lane groups and combined with
//     shuffles; the final max over lanes is __reduce_max_sync.
#pragma once
#include <cstdint>

namespace iglu {

constexpr int ROWS_PER_CLASS = 99;               // 9 y-levels x 11 x-rows
constexpr int NROWS = 16 * ROWS_PER_CLASS;       // 4-bit class codes
constexpr int ROWS_WORDS = NROWS / 2;            // uint16 rows packed in uint32 words (792)
constexpr int LIST_CHUNK = 64;                   // target blocks expanded per pass
constexpr int SPREAD_N = 2048;

// Per-warp shared-memory scratch (16-byte aligned pieces).
struct __align__(16) WarpScratch {
    uint32_t grid[GSTRIDE / 4];          // int8 grid of the current env (1089 cells + zero pad)
    uint32_t init[GSTRIDE / 4];          // starting grid (only when the env has one)
    uint32_t rows[ROWS_WORDS];           // z-reversed occupancy rows, two uint16 per word
    uint32_t list[4][LIST_CHUNK];        // expanded target blocks per rotation
};

// spread table: bit j of v -> bit 5*j of the result (eleven 5-bit fields in a uint64).
__device__ __forceinline__ void build_spread_table(unsigned long long *tbl, int tid, int nthreads) {
    for (int v = tid; v < SPREAD_N; v += nthreads) {
        unsigned long long r = 0;
#pragma unroll
        for (int j = 0; j < 11; j++) r |= (unsigned long long)((v >> j) & 1) << (5 * j);
        tbl[v] = r;
    }
}

// Pack the (difference) grid held in ws.grid (minus ws.init when has_init) into ws.rows.
__device__ __forceinline__ void build_rows(WarpScratch &ws, bool has_init, int lane) {
    for (int w = lane; w < ROWS_WORDS; w += 32) ws.rows[w] = 0u;
    __syncwarp();
    for (int w = lane; w < (CELLS + 3) / 4; w += 32) {
        uint32_t gv = ws.grid[w];
        uint32_t iv = has_init ? ws.init[w] : 0u;
        if ((gv | iv) == 0u) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int d = (int)(int8_t)(gv >> (8 * j)) - (int)(int8_t)(iv >> (8 * j));
            int cell = 4 * w + j;
            if (d != 0 && cell < CELLS) {
                int y = cell / 121, r = cell - 121 * y, x = r / 11, z = r - 11 * x;
                int row = (d & 15) * ROWS_PER_CLASS + y * 11 + x;
                atomicOr(&ws.rows[row >> 1], (1u << (10 - z)) << (16 * (row & 1)));
            }
        }
    }
    __syncwarp();
}

struct Rects {
    int dx_lo[4], dx_hi[4], dz_lo[4], dz_hi[4];
    int n_rot;
};

// Rectangles of the four rotations from the rotation-0 bounding box of the non-zero cells
// (x_min..x_max, z_min..z_max; empty = (10, 0, 10, 0) which yields all 441 shifts, task.py:67-72).
__device__ __forceinline__ void rects_from_bbox(int xmin, int xmax, int zmin, int zmax, int8_t out[16]) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        out[4 * k + 0] = (int8_t)(xmax - 10);
        out[4 * k + 1] = (int8_t)xmin;
        out[4 * k + 2] = (int8_t)(zmax - 10);
        out[4 * k + 3] = (int8_t)zmin;
        // rotate the box: x' = z, z' = 10 - x  (task.py:52-53)
        int nxmin = zmin, nxmax = zmax, nzmin = 10 - xmax, nzmax = 10 - xmin;
        xmin = nxmin; xmax = nxmax; zmin = nzmin; zmax = nzmax;
    }
}

// Evaluate max intersection for the env whose rows are in ws.rows.
//   tlist : nblk packed target blocks (y | x<<4 | z<<8 | code<<12), global or shared memory
//   rc    : 16 int8 rectangle bounds; n_rot rotations are evaluated
//   CLIP  : true when shifts may move target cells outside the grid (full_grid rectangles)
// Returns (warp-uniform) best count; if key_out != nullptr also the first (rot, dx, dz) reaching
// it in the reference's iteration order, encoded k*441 + (dx+10)*21 + (dz+10), or -1 if best == 0.
template <bool CLIP>
__device__ __forceinline__ int eval_max_intersection(WarpScratch &ws, const unsigned long long *__restrict__ tbl,
                                                     const uint16_t *__restrict__ tlist, int nblk,
                                                     const int8_t *__restrict__ rc, int n_rot, int lane,
                                                     int *key_out) {
    // ---- enumerate work items (k, dx, dz-chunk of <= 11) ----
    int nx[4], nch[4], base[5];
    base[0] = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int wx = (int)rc[4 * k + 1] - (int)rc[4 * k + 0] + 1;
        int wz = (int)rc[4 * k + 3] - (int)rc[4 * k + 2] + 1;
        nx[k] = (k < n_rot && wx > 0 && wz > 0) ? wx : 0;
        nch[k] = CLIP ? (wz + 10) / 11 : 1;
        base[k + 1] = base[k] + nx[k] * nch[k];
    }
    const int T = base[4];
    int best = 0;  // packed (count << 12 | 4095 - key) when key_out, else count
    if (T == 0 || nblk == 0) {
        if (key_out) *key_out = -1;
        return 0;
    }
    int ppad = 32;
    if (T <= 16) { ppad = 1; while (ppad < T) ppad <<= 1; }
    const int G = 32 / ppad;
    const int p = lane & (ppad - 1), g = lane / ppad;

    for (int pass0 = 0; pass0 < T; pass0 += 32) {
        const int item = pass0 + p;
        const bool valid = item < T;
        int k = 0;
#pragma unroll
        for (int q = 1; q < 4; q++) if (item >= base[q]) k = q;
        int local = item - base[k];
        int nchk = 1, nxk = 1, dxlo = 0, dzlo0 = 0, dzhi = 0;
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (q == k) { nchk = nch[q]; nxk = nx[q]; dxlo = rc[4 * q]; dzlo0 = rc[4 * q + 2]; dzhi = rc[4 * q + 3]; }
        (void)nxk;
        const int ch = CLIP ? local % nchk : 0;
        const int dx = dxlo + (CLIP ? local / nchk : local);
        const int dzlo = dzlo0 + 11 * ch;
        int W = dzhi - dzlo + 1;
        if (W > 11) W = 11;
        const int cz = 10 + dzlo;  // shift = cz - z_k

        int cnt[11];
#pragma unroll
        for (int j = 0; j < 11; j++) cnt[j] = 0;
        unsigned long long acc = 0ull;
        int nacc = 0;

        for (int c0 = 0; c0 < nblk; c0 += LIST_CHUNK) {
            const int cn = min(LIST_CHUNK, nblk - c0);
            // expand the chunk for the four rotations: rowoff2 | z_k << 12 | x_k << 16
            for (int i = lane; i < 4 * cn; i += 32) {
                int kk = i / cn, b = i - kk * cn;
                uint32_t e = tlist[c0 + b];
                int y = e & 15, x = (e >> 4) & 15, z = (e >> 8) & 15, code = e >> 12;
                int xk = (kk == 0) ? x : (kk == 1) ? z : (kk == 2) ? 10 - x : 10 - z;
                int zk = (kk == 0) ? z : (kk == 1) ? 10 - x : (kk == 2) ? 10 - z : x;
                ws.list[kk][b] = (uint32_t)((code * ROWS_PER_CLASS + y * 11 + xk) * 2) | (zk << 12) | (xk << 16);
            }
            __syncwarp();
            if (valid) {
                const unsigned char *rows8 = reinterpret_cast<const unsigned char *>(ws.rows);
                for (int b = g; b < cn; b += G) {
                    uint32_t e = ws.list[k][b];
                    int roff = (int)(e & 0xFFFu) - 2 * dx;
                    int sh = cz - (int)((e >> 12) & 15u);
                    uint32_t v;
                    if (CLIP) {
                        int xs = (int)(e >> 16) - dx;
                        uint32_t R = ((unsigned)xs <= 10u) ? *reinterpret_cast<const unsigned short *>(rows8 + roff) : 0u;
                        v = sh >= 0 ? (sh > 31 ? 0u : R >> sh) : (sh < -31 ? 0u : R << (-sh));
                        v &= 0x7FFu;
                    } else {
                        uint32_t R = *reinterpret_cast<const unsigned short *>(rows8 + roff);
                        v = R >> sh;
                    }
                    acc += tbl[v];
                    if (++nacc == 31) {
#pragma unroll
                        for (int j = 0; j < 11; j++) cnt[j] += (int)((acc >> (5 * j)) & 31ull);
                        acc = 0ull;
                        nacc = 0;
                    }
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 11; j++) cnt[j] += (int)((acc >> (5 * j)) & 31ull);
        // combine the block slices of the G lane groups
        for (int s = ppad; s < 32; s <<= 1) {
#pragma unroll
            for (int j = 0; j < 11; j++) cnt[j] += __shfl_xor_sync(0xffffffffu, cnt[j], s);
        }
        if (valid) {
            const int kbase = k * 441 + (dx + 10) * 21 + (dzlo + 10);
#pragma unroll
            for (int j = 0; j < 11; j++) {
                if (j < W) {
                    int sc = key_out ? ((cnt[j] << 12) | (4095 - (kbase + j))) : cnt[j];
                    best = max(best, sc);
                }
            }
        }
    }
    best = __reduce_max_sync(0xffffffffu, best);
    if (key_out) {
        int count = best >> 12;
        *key_out = count > 0 ? 4095 - (best & 4095) : -1;
        return count;
    }
    return best;
}

}  // namespace iglu
