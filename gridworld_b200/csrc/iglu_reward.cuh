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
//     flushed to 32-bit counters after every chunk of <= 28 blocks;
//   * when at most 16 work items exist the blocks are split over lane groups and combined with
//     shuffles; the final max over lanes is __reduce_max_sync.
#pragma once
#include <cstdint>

namespace iglu {

constexpr int ROWS_PER_CLASS = 99;               // 9 y-levels x 11 x-rows
constexpr int NCLASS = 13;                       // class 0 = never-set dummy rows, 1..6 colours, 7..12 = values -6..-1
constexpr int NROWS = NCLASS * ROWS_PER_CLASS;   // 1287 uint16 rows
constexpr int ROWS_WORDS = (NROWS + 1 + 7) / 8 * 4;  // uint16 rows packed in uint32 words, padded to 16 bytes (648)
constexpr int LIST_CHUNK = 28;                   // target blocks per pass (<= 31 adds per 5-bit field)
constexpr int LIST_CAP = 32;                     // chunk padded up to a multiple of 4*G <= 32
constexpr int SPREAD_N = 2048;
constexpr uint32_t DUMMY_ENTRY = 2 * 50;         // row 50 of class 0 (never set), shift 0 -> adds nothing

// Grid / target values are differences of colour ids, i.e. -6..6 (env.py:230, 290). code_of() maps a value
// to a 4-bit code (value & 15; 8 = "not a valid value", never matches), class_of() a code to its row class:
// 1..6 -> 1..6, 10..15 (values -6..-1) -> 7..12, anything else -> 0 (the never-set dummy class).
__device__ __forceinline__ int code_of(int v) { return (v >= -6 && v <= 6) ? (v & 15) : 8; }
__device__ __forceinline__ int class_of(int code) { return code < 7 ? code : (code >= 10 ? code - 3 : 0); }

// Per-warp scratch for one evaluation.
struct __align__(16) EvalScratch {
    uint32_t rows[ROWS_WORDS];      // z-reversed occupancy rows, two uint16 per word
    uint32_t list[4][LIST_CAP];     // expanded target blocks per rotation
    uint32_t planes[16];            // per class code: bit y set when the grid has a block on that level
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

// Sparse traversal of a staged grid. Lane L owns the 36 cells [36 L, 36 L + 36) = nine 32-bit words (lane
// stride 9 words is bank-conflict free). Grids hold a few dozen blocks, so instead of visiting every
// byte each lane first builds a 9-bit mask of its non-zero words (3 instructions per word) and then
// loops over the set words / non-zero bytes only, in ascending cell order:  f(cell, value, sub_value).
// `g` is shared memory; `sub` (may be nullptr) is subtracted cell-wise and may be shared or global.
template <class F>
__device__ __forceinline__ void for_each_nonzero_cell(const uint32_t *g, const uint32_t *sub, int lane, F f) {
    unsigned wm = 0u;
#pragma unroll
    for (int q = 0; q < 9; q++) {
        const int i = 9 * lane + q;
        uint32_t x = 0u;
        if (i < GSTRIDE / 4) {
            x = g[i];
            if (sub) x ^= sub[i];
        }
        if (x) wm |= 1u << q;
    }
    while (wm) {
        const int q = __ffs(wm) - 1;
        wm &= wm - 1;
        const int i = 9 * lane + q;
        const uint32_t gw = g[i], sw = sub ? sub[i] : 0u;
        uint32_t x = gw ^ sw;
        while (x) {
            const int j = (__ffs(x) - 1) >> 3;
            x &= ~(0xFFu << (8 * j));
            f(4 * i + j, (int)(int8_t)(gw >> (8 * j)), (int)(int8_t)(sw >> (8 * j)));
        }
    }
}

// number of cells of lane L's segment that differ from 0 (ascending-order emit needs it for the scan)
__device__ __forceinline__ int count_nonzero_cells(const uint32_t *g, int lane) {
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < 9; q++) {
        const int i = 9 * lane + q;
        const uint32_t x = i < GSTRIDE / 4 ? g[i] : 0u;
        cnt += __popc((x | ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u);
    }
    return cnt;
}

// Pack the (difference) grid `grid` (minus `init` when has_init) into es.rows / es.planes. `grid` is a
// 16-byte aligned shared-memory copy of GSTRIDE bytes, `init` the env's starting grid in GLOBAL memory
// (same layout); pad bytes (cells >= 1089) are zero in both.
__device__ __forceinline__ void build_rows(EvalScratch &es, const uint32_t *grid, const int8_t *init,
                                           bool has_init, int lane) {
    int4 *rz = reinterpret_cast<int4 *>(es.rows);
#pragma unroll
    for (int it = 0; it < (ROWS_WORDS / 4 + 31) / 32; it++) {
        const int q = lane + 32 * it;
        if (q < ROWS_WORDS / 4) rz[q] = make_int4(0, 0, 0, 0);
    }
    if (lane < 16) es.planes[lane] = 0u;
    __syncwarp();
    const uint32_t *iw = has_init ? reinterpret_cast<const uint32_t *>(init) : nullptr;
    for_each_nonzero_cell(grid, iw, lane, [&](int cell, int gv, int sv) {
        const int d = gv - sv;
        if (d >= -6 && d <= 6) {  // other values are not colour differences and match nothing
            const int y = cell / 121, r = cell - 121 * y, x = r / 11, z = r - 11 * x;
            const int row = class_of(d & 15) * ROWS_PER_CLASS + y * 11 + x;
            atomicOr(&es.rows[row >> 1], (1u << (10 - z)) << (16 * (row & 1)));
            atomicOr(&es.planes[d & 15], 1u << y);
        }
    });
    __syncwarp();
}

// Admissible-shift rectangles of the four rotations from the rotation-0 bounding box of the non-zero
// cells (x_min..x_max, z_min..z_max; empty = (10, 0, 10, 0) which yields all 441 shifts, task.py:67-72).
// Packed as one 32-bit word per rotation: byte 0 = dx_lo (int8), byte 1 = number of dx values (bits 0-4),
// byte 2 = dz_lo (int8), byte 3 = number of dz values; bits 13-14 of word 0 hold the lane-mapping shift
// s = ceil(log2(min(8, max dx count))) so that the evaluation does not have to derive it per env.
__device__ __forceinline__ int4 rects_from_bbox(int xmin, int xmax, int zmin, int zmax) {
    int r[4], nxmax = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int nx = xmin - (xmax - 10) + 1, nz = zmin - (zmax - 10) + 1;
        r[k] = ((xmax - 10) & 0xFF) | ((nx & 0x1F) << 8) | (((zmax - 10) & 0xFF) << 16) | ((nz & 0xFF) << 24);
        nxmax = max(nxmax, nx);
        // rotate the box: x' = z, z' = 10 - x  (task.py:52-53)
        const int nxmin = zmin, nxmax2 = zmax, nzmin = 10 - xmax, nzmax = 10 - xmin;
        xmin = nxmin; xmax = nxmax2; zmin = nzmin; zmax = nzmax;
    }
    const int s = nxmax <= 1 ? 0 : (nxmax <= 2 ? 1 : (nxmax <= 4 ? 2 : 3));
    r[0] |= s << 13;
    return make_int4(r[0], r[1], r[2], r[3]);
}
// rectangle (0, 0) x (0, 0) of rotation 0 only (invariant=False, task.py:59-60)
__device__ __forceinline__ int4 rects_identity() { return make_int4(0x01000100, 0, 0, 0); }

__device__ __forceinline__ int rect_byte(int w, int i) { return (int)(int8_t)((unsigned)w >> (8 * i)); }

// Evaluate max intersection for the env whose rows are in es.rows.
//   tlist : nblk packed target blocks (y | x<<4 | z<<8 | code<<12), shared or global memory
//   rc    : the four rotation rectangles (see rects_from_bbox); n_rot rotations are evaluated
//   CLIP  : true when shifts may move target cells outside the grid (full_grid rectangles)
// Target blocks whose (class, y) plane holds no grid block cannot match under any shift; they are
// dropped while the list is expanded (es.planes, filled by build_rows), so the cost follows the
// number of target blocks that CAN match rather than the target size.
// Returns (warp-uniform) best count; if WANT_KEY also the first (rot, dx, dz) reaching it in the
// reference's iteration order, encoded k*441 + (dx+10)*21 + (dz+10), or -1 if best == 0.
template <bool CLIP, bool WANT_KEY>
__device__ __forceinline__ int eval_max_intersection(EvalScratch &es, const unsigned long long *__restrict__ tbl,
                                                     const uint16_t *__restrict__ tlist, int nblk, int4 rc,
                                                     int n_rot, int lane, int *key_out) {
    if (WANT_KEY) *key_out = -1;
    if (nblk == 0) return 0;
    const int rw[4] = {rc.x, rc.y, rc.z, rc.w};
    // ---- work items: lane -> (rotation k, dx index, block-slice g) ----
    const int s = (rc.x >> 13) & 3;
    const int G = 8 >> s;                      // lane groups sharing the blocks of one work item
    const int dxi0 = lane & ((1 << s) - 1);
    const int k = (lane >> s) & 3;
    const int g = lane >> (s + 2);
    const int rk = k == 0 ? rw[0] : (k == 1 ? rw[1] : (k == 2 ? rw[2] : rw[3]));
    const int dx_lo = rect_byte(rk, 0), nxk = (rk >> 8) & 0x1F;
    const int dz_lo = rect_byte(rk, 2), wz = (rk >> 24) & 0xFF;
    int nxmax = 8;
    if (CLIP) nxmax = max(max((rw[0] >> 8) & 0x1F, (rw[1] >> 8) & 0x1F), max((rw[2] >> 8) & 0x1F, (rw[3] >> 8) & 0x1F));
    else if (s == 3) nxmax = max((rw[0] >> 8) & 0x1F, (rw[1] >> 8) & 0x1F);  // rotations 2, 3 repeat 0, 1
    const uint32_t *mylist = es.list[k];
    const unsigned char *rows8 = reinterpret_cast<const unsigned char *>(es.rows);
    const int n_dxpass = (nxmax + 7) >> 3;               // 2 only when some rotation has 9..11 dx values
    const int n_zpass = CLIP ? (21 + 10) / 11 : 1;       // dz chunks of 11 (up to 21 values when CLIP)
    int best = 0;  // WANT_KEY: count << 12 | 4095 - key

    for (int pass = 0; pass < n_dxpass * n_zpass; pass++) {
        const int dxi = dxi0 + 8 * (CLIP ? pass / n_zpass : pass);
        const int ch = CLIP ? pass % n_zpass : 0;
        const int dx = dx_lo + dxi;
        const int dzlo = dz_lo + 11 * ch;
        int W = wz - 11 * ch;
        if (W > 11) W = 11;
        const bool valid = k < n_rot && dxi < nxk && W > 0;
        const int cz = 10 + dzlo;  // CLIP: shift = cz - z_k
        // running totals: fifteen-bit fields, field j of wide[j % 3] at bit 15 * (j / 3)
        unsigned long long wide0 = 0ull, wide1 = 0ull, wide2 = 0ull;
        const unsigned long long M = 0x1Full | (0x1Full << 15) | (0x1Full << 30) | (0x1Full << 45);

        for (int c0 = 0; c0 < nblk; c0 += LIST_CHUNK) {
            const int cn = min(LIST_CHUNK, nblk - c0);
            // expand: lane b computes the four rotated entries of block b; blocks of empty planes are dropped
            uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0;
            bool keep = false;
            if (lane < cn) {
                const uint32_t e = tlist[c0 + lane];
                const int y = e & 15, x = (e >> 4) & 15, z = (e >> 8) & 15, code = e >> 12;
                keep = (es.planes[code] >> y) & 1u;
                const int rb = (class_of(code) * ROWS_PER_CLASS + y * 11) * 2;
                if (CLIP) {  // rowoff2 | z_k << 12 | x_k << 16
                    e0 = (uint32_t)(rb + 2 * x) | (z << 12) | (x << 16);
                    e1 = (uint32_t)(rb + 2 * z) | ((10 - x) << 12) | (z << 16);
                    e2 = (uint32_t)(rb + 2 * (10 - x)) | ((10 - z) << 12) | ((10 - x) << 16);
                    e3 = (uint32_t)(rb + 2 * (10 - z)) | (x << 12) | ((10 - z) << 16);
                } else {  // rowoff2 | shift << 16 with shift = 10 + dz_lo(k) - z_k
                    e0 = (uint32_t)(rb + 2 * x) | ((10 + rect_byte(rw[0], 2) - z) << 16);
                    e1 = (uint32_t)(rb + 2 * z) | ((10 + rect_byte(rw[1], 2) - (10 - x)) << 16);
                    e2 = (uint32_t)(rb + 2 * (10 - x)) | ((10 + rect_byte(rw[2], 2) - (10 - z)) << 16);
                    e3 = (uint32_t)(rb + 2 * (10 - z)) | ((10 + rect_byte(rw[3], 2) - x) << 16);
                }
            }
            const unsigned km = __ballot_sync(0xffffffffu, keep);
            const int ck = __popc(km);
            if (ck == 0) continue;  // warp-uniform
            const int cnp = (ck + 4 * G - 1) & ~(4 * G - 1);  // padded with entries that add nothing
            // kept blocks are compacted to slots [0, ck); slots [ck, cnp) get entries that add nothing
            if (lane >= ck && lane < cnp) {
                es.list[0][lane] = DUMMY_ENTRY; es.list[1][lane] = DUMMY_ENTRY;
                es.list[2][lane] = DUMMY_ENTRY; es.list[3][lane] = DUMMY_ENTRY;
            }
            if (keep) {
                const int pos = __popc(km & ((1u << lane) - 1u));
                es.list[0][pos] = e0; es.list[1][pos] = e1; es.list[2][pos] = e2; es.list[3][pos] = e3;
            }
            __syncwarp();
            if (valid) {
                unsigned long long acc = 0ull;
                for (int b = g; b < cnp; b += 4 * G) {
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t e = mylist[b + u * G];
                        uint32_t v;
                        if (CLIP) {
                            const int roff = (int)(e & 0xFFFu) - 2 * dx;
                            const int sh = cz - (int)((e >> 12) & 15u);
                            const int xs = (int)(e >> 16) - dx;
                            const uint32_t R = ((unsigned)xs <= 10u)
                                                   ? *reinterpret_cast<const unsigned short *>(rows8 + roff) : 0u;
                            v = sh >= 0 ? (sh > 31 ? 0u : R >> sh) : (sh < -31 ? 0u : R << (-sh));
                            v &= 0x7FFu;
                        } else {
                            const uint32_t ed = e - 2u * (uint32_t)dx;  // row offset never borrows
                            const uint32_t R = *reinterpret_cast<const unsigned short *>(rows8 + (ed & 0xFFFFu));
                            v = R >> (ed >> 16);
                        }
                        acc += tbl[v];
                    }
                }
                wide0 += acc & M;
                wide1 += (acc >> 5) & M;
                wide2 += (acc >> 10) & M;
            }
            __syncwarp();
        }
        // combine the block slices of the G lane groups
        for (int st = 4 << s; st < 32; st <<= 1) {
            wide0 += __shfl_xor_sync(0xffffffffu, wide0, st);
            wide1 += __shfl_xor_sync(0xffffffffu, wide1, st);
            wide2 += __shfl_xor_sync(0xffffffffu, wide2, st);
        }
        if (valid) {
            const int kbase = k * 441 + (dx + 10) * 21 + (dzlo + 10);
#pragma unroll
            for (int j = 0; j < 11; j++) {
                if (j < W) {
                    const unsigned long long wj = (j % 3 == 0) ? wide0 : (j % 3 == 1 ? wide1 : wide2);
                    const int c = (int)((wj >> (15 * (j / 3))) & 0x7FFFull);
                    const int sc = WANT_KEY ? ((c << 12) | (4095 - (kbase + j))) : c;
                    best = max(best, sc);
                }
            }
        }
    }
    best = __reduce_max_sync(0xffffffffu, best);
    if (WANT_KEY) {
        const int count = best >> 12;
        *key_out = count > 0 ? 4095 - (best & 4095) : -1;
        return count;
    }
    return best;
}

// Direct evaluation for tasks whose every rotation admits exactly ONE shift (targets spanning the whole zone in
// x and z, or invariant=False): per target block and rotation one byte of the staged grid is compared -- no
// occupancy rows, no shift vectors. Lane L takes blocks L, L + 32, ...
//   g8 : staged grid (shared memory, dense 1089 bytes); sub : starting grid to subtract cell-wise (global memory)
//        or nullptr; tlist / rc / n_rot as for eval_max_intersection
__device__ __forceinline__ bool single_shift_rects(int4 rc, int n_rot) {
    const int rw[4] = {rc.x, rc.y, rc.z, rc.w};
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k < n_rot) ok = ok && ((rw[k] >> 8) & 0x1F) == 1 && ((rw[k] >> 24) & 0xFF) == 1;
    return ok;
}

template <bool WANT_KEY>
__device__ __forceinline__ int eval_single_shift(const int8_t *g8, const int8_t *__restrict__ sub,
                                                 const uint16_t *__restrict__ tlist, int nblk, int4 rc, int n_rot,
                                                 int lane, int *key_out) {
    auto at = [&](int idx) -> int { return sub ? (int)g8[idx] - (int)sub[idx] : (int)g8[idx]; };
    const int dx0 = rect_byte(rc.x, 0), dz0 = rect_byte(rc.x, 2), dx1 = rect_byte(rc.y, 0), dz1 = rect_byte(rc.y, 2);
    const int dx2 = rect_byte(rc.z, 0), dz2 = rect_byte(rc.z, 2), dx3 = rect_byte(rc.w, 0), dz3 = rect_byte(rc.w, 2);
    // cell index of (y, x_k - dx_k, z_k - dz_k) = 121 y + a_k x + b_k z + o_k  with the rotations of task.py:52-53
    const int o0 = -11 * dx0 - dz0;                       // x_0 = x,      z_0 = z
    const int o1 = -11 * dx1 + 10 - dz1;                  // x_1 = z,      z_1 = 10 - x
    const int o2 = 11 * (10 - dx2) + 10 - dz2;            // x_2 = 10 - x, z_2 = 10 - z
    const int o3 = 11 * (10 - dx3) - dz3;                 // x_3 = 10 - z, z_3 = x
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    for (int b = lane; b < nblk; b += 32) {
        const uint32_t e = tlist[b];
        const int y = e & 15, x = (e >> 4) & 15, z = (e >> 8) & 15, code = e >> 12;
        if (code == 8) continue;                          // not a colour difference: matches nothing
        const int v = code < 8 ? code : code - 16;
        const int base = 121 * y;
        c0 += at(base + 11 * x + z + o0) == v;
        if (n_rot > 1) {
            c1 += at(base + 11 * z - x + o1) == v;
            c2 += at(base - 11 * x - z + o2) == v;
            c3 += at(base - 11 * z + x + o3) == v;
        }
    }
    c0 = __reduce_add_sync(0xffffffffu, c0);
    int best = c0, kbest = 0;
    if (n_rot > 1) {
        c1 = __reduce_add_sync(0xffffffffu, c1);
        c2 = __reduce_add_sync(0xffffffffu, c2);
        c3 = __reduce_add_sync(0xffffffffu, c3);
        if (c1 > best) { best = c1; kbest = 1; }          // strict: the first rotation reaching the max wins
        if (c2 > best) { best = c2; kbest = 2; }
        if (c3 > best) { best = c3; kbest = 3; }
    }
    if (WANT_KEY) {
        const int dx = kbest == 0 ? dx0 : (kbest == 1 ? dx1 : (kbest == 2 ? dx2 : dx3));
        const int dz = kbest == 0 ? dz0 : (kbest == 1 ? dz1 : (kbest == 2 ? dz2 : dz3));
        *key_out = best > 0 ? kbest * 441 + (dx + 10) * 21 + (dz + 10) : -1;
    }
    return best;
}

}  // namespace iglu
