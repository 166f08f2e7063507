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
// Formulation (round 2): SHIFT VOTING. A grid block g = (y, xg, zg, c) and a block t = (y, x_k, z_k, c) of the
// rotated target with the same level and the same class contribute to exactly one shift of rotation k, namely
// (dx, dz) = (x_k - xg, z_k - zg). So instead of testing every shift against every block (the reference: 4 x S
// dense compares; round 1: one bit-row operation per (block, rotation, dx)), every MATCHING (grid block, target
// block) pair casts one vote per rotation into a 4 x 21 x 21 table of 16-bit counters in shared memory, and the
// answer is the maximum of the table over the admissible rectangles. Work is proportional to the number of
// same-(level, class) block pairs -- a few dozen for real structures -- and it needs no clipping case: a target
// cell that a shift moves outside the zone simply has no grid partner.
//   * matching pairs are found through a per-(level, class) bit mask over the (<= 32 per pass) target blocks:
//     lane t registers its block with one shared-memory atomicOr; a lane that holds a grid block reads the mask
//     of its key and walks the set bits;
//   * the grid is traversed sparsely: lane L owns the words L, L + 32, ... of the staged int8 grid and only
//     visits non-zero words / bytes; (y, 22 xg + zg) of a cell comes from a 1089-entry table in shared memory;
//   * votes are shared-memory atomic adds of 1 << 16 (idx & 1) on word idx >> 1; the table is cleared with
//     16-byte stores before every evaluation; the scan runs over (rotation, dx) rows with __reduce_max_sync at
//     the end (and the reference's first-max tie-break encoded in the key when argmax is wanted).
// Tasks whose every rotation admits exactly ONE shift (zone-spanning targets, invariant=False) keep the direct
// byte-compare path (eval_single_shift).
#pragma once
#include <cstdint>

namespace iglu {

constexpr int HROW = 22;                         // 16-bit counters per (rotation, dx) row: 21 dz values + 1 pad
constexpr int HROT = 21 * HROW;                  // counters per rotation (462)
constexpr int HIST_WORDS = 4 * HROT / 2;         // 924 words = 3696 bytes
constexpr int NKEY = 144;                        // (level y) * 16 + class code
constexpr int CELL_LUT_N = 1092;                 // 1089 cells (+ pad to a multiple of 4)

// Grid / target values are differences of colour ids, i.e. -6..6 (env.py:230, 290). code_of() maps a value
// to a 4-bit code (value & 15; 8 = "not a valid value", never matches).
__device__ __forceinline__ int code_of(int v) { return (v >= -6 && v <= 6) ? (v & 15) : 8; }

// Per-warp scratch for one evaluation.
struct __align__(16) EvalScratch {
    uint32_t hist[HIST_WORDS];   // vote counters, idx16 = k * HROT + (dx + 10) * HROW + (dz + 10)
    uint32_t kmask[NKEY];        // per (y, class): bit t set when target block t of the current pass has that key
                                 // (all zero between evaluations: every pass clears what it set)
    uint2 tb[32];                // per target block of the pass: the four vote bases (16 bits each)
};

// cell index -> y << 12 | (HROW * x + z); built once per CTA
__device__ __forceinline__ void build_cell_lut(uint16_t *lut, int tid, int nthreads) {
    for (int c = tid; c < CELL_LUT_N; c += nthreads) {
        const int y = c / 121, r = c - 121 * y, x = r / 11, z = r - 11 * x;
        lut[c] = (uint16_t)(c < 1089 ? ((y << 12) | (HROW * x + z)) : 0);
    }
}

// A staged int8 grid in shared memory. ALIGNED: 16-byte aligned copy of a GSTRIDE-byte row (the env state layout).
// Otherwise the 1089 cells start at an arbitrary byte offset of a 16-byte aligned window (the stateless sweep brings
// the unaligned rows of a dense [count][1089] array in with aligned TMA bulk copies): a logical word is then two
// aligned shared-memory words and a funnel shift, and bytes past cell 1088 are masked to zero.
template <bool ALIGNED>
struct Staged {
    const uint32_t *w;   // word holding the first cell
    int sh;              // 8 * (byte offset of the first cell inside that word)
    __device__ __forceinline__ uint32_t word(int i, bool last_q) const {
        uint32_t v = ALIGNED ? w[i] : __funnelshift_r(w[i], w[i + 1], sh);
        if (last_q && i >= 272) v = i == 272 ? (v & 0xFFu) : 0u;   // cell 1088 | bytes that are not cells
        return v;
    }
    __device__ __forceinline__ const int8_t *bytes() const {
        return reinterpret_cast<const int8_t *>(w) + (sh >> 3);
    }
};

// Sparse traversal of a staged grid. Lane L owns the 32-bit words L, L + 32, ..., L + 256 (consecutive lanes read
// consecutive words: conflict free; a clump of blocks spreads over many lanes). Grids hold a few dozen blocks, so
// each lane first builds a 9-bit mask of its non-zero words and then loops over the set words / non-zero bytes
// only:  f(cell, value, sub_value).  `sub` (may be nullptr; 16-byte aligned row, shared or global) is subtracted
// cell-wise. Pad bytes (cells >= 1089) are zero in both. No ordering guarantee.
template <bool ALIGNED, class F>
__device__ __forceinline__ void for_each_nonzero_cell(const Staged<ALIGNED> g, const uint32_t *sub, int lane, F f) {
    unsigned wm = 0u;
#pragma unroll
    for (int q = 0; q < 9; q++) {
        const int i = lane + 32 * q;
        uint32_t x = 0u;
        if (q < 8 || i < GSTRIDE / 4) {
            x = g.word(i, q == 8);
            if (sub) x ^= sub[i];
        }
        if (x) wm |= 1u << q;
    }
    while (wm) {
        const int q = __ffs(wm) - 1;
        wm &= wm - 1;
        const int i = lane + 32 * q;
        const uint32_t gw = g.word(i, true), sw = sub ? sub[i] : 0u;
        uint32_t x = gw ^ sw;
        while (x) {
            const int j = (__ffs(x) - 1) >> 3;
            x &= ~(0xFFu << (8 * j));
            f(4 * i + j, (int)(int8_t)(gw >> (8 * j)), (int)(int8_t)(sw >> (8 * j)));
        }
    }
}

// number of non-zero cells among the words lane L owns
template <bool ALIGNED>
__device__ __forceinline__ int count_nonzero_cells(const Staged<ALIGNED> g, int lane) {
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < 9; q++) {
        const int i = lane + 32 * q;
        const uint32_t x = (q < 8 || i < GSTRIDE / 4) ? g.word(i, q == 8) : 0u;
        cnt += __popc((x | ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u);
    }
    return cnt;
}

// Admissible-shift rectangles of the four rotations from the rotation-0 bounding box of the non-zero
// cells (x_min..x_max, z_min..z_max; empty = (10, 0, 10, 0) which yields all 441 shifts, task.py:67-72).
// Packed as one 32-bit word per rotation: byte 0 = dx_lo (int8), byte 1 = number of dx values (bits 0-4),
// byte 2 = dz_lo (int8), byte 3 = number of dz values.
__device__ __forceinline__ int4 rects_from_bbox(int xmin, int xmax, int zmin, int zmax) {
    int r[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int nx = xmin - (xmax - 10) + 1, nz = zmin - (zmax - 10) + 1;
        r[k] = ((xmax - 10) & 0xFF) | ((nx & 0x1F) << 8) | (((zmax - 10) & 0xFF) << 16) | ((nz & 0xFF) << 24);
        // rotate the box: x' = z, z' = 10 - x  (task.py:52-53)
        const int nxmin = zmin, nxmax2 = zmax, nzmin = 10 - xmax, nzmax = 10 - xmin;
        xmin = nxmin; xmax = nxmax2; zmin = nzmin; zmax = nzmax;
    }
    return make_int4(r[0], r[1], r[2], r[3]);
}
// rectangle (0, 0) x (0, 0) of rotation 0 only (invariant=False, task.py:59-60)
__device__ __forceinline__ int4 rects_identity() { return make_int4(0x01000100, 0, 0, 0); }

__device__ __forceinline__ int rect_byte(int w, int i) { return (int)(int8_t)((unsigned)w >> (8 * i)); }

// One target block of the current pass: slot t (0..31) gets its four vote bases and its (level, class) key mask bit.
// Returns the key (for un-registering), or -1 for a value that is no colour difference (matches nothing).
__device__ __forceinline__ int register_target(EvalScratch &es, int t, int y, int x, int z, int code) {
    if (code == 8) return -1;
    // vote base of rotation k: k * HROT + (x_k + 10) * HROW + (z_k + 10) with the rotations of task.py:52-53
    const uint32_t b0 = 0 * HROT + (x + 10) * HROW + (z + 10);                 // (x, z)
    const uint32_t b1 = 1 * HROT + (z + 10) * HROW + (20 - x);                 // (z, 10 - x)
    const uint32_t b2 = 2 * HROT + (20 - x) * HROW + (20 - z);                 // (10 - x, 10 - z)
    const uint32_t b3 = 3 * HROT + (20 - z) * HROW + (x + 10);                 // (10 - z, x)
    es.tb[t] = make_uint2(b0 | (b1 << 16), b2 | (b3 << 16));
    const int key = (y << 4) | code;
    atomicOr(&es.kmask[key], 1u << t);
    return key;
}

// Evaluate max intersection of the staged (difference) grid with a set of target blocks.
//   grid  : staged int8 grid (shared memory)
//   init  : the env's starting grid (16-byte aligned row, global or shared memory) subtracted cell-wise, or nullptr
//   lut   : build_cell_lut table (shared memory)
//   nblk  : number of target blocks; they are handed over in passes of 32: reg(c0) makes every lane register the
//           blocks it holds among slots [c0, c0 + 32) (register_target with t = slot - c0) and unreg(c0) makes it
//           clear es.kmask[key] for each of them again
//   rc    : the four rotation rectangles (see rects_from_bbox); n_rot rotations are evaluated
// Returns (warp-uniform) best count; if WANT_KEY also the first (rot, dx, dz) reaching it in the
// reference's iteration order, encoded k*441 + (dx+10)*21 + (dz+10), or -1 if best == 0.
template <bool WANT_KEY, bool ALIGNED, class RegF, class UnregF>
__device__ __forceinline__ int eval_votes(EvalScratch &es, const Staged<ALIGNED> grid, const int8_t *init,
                                          const uint16_t *__restrict__ lut, int nblk, RegF reg, UnregF unreg,
                                          int4 rc, int n_rot, int lane, int *key_out) {
    if (WANT_KEY) *key_out = -1;
    if (nblk == 0) return 0;
    // ---- clear the vote table ----
    {
        int4 *hz = reinterpret_cast<int4 *>(es.hist);
#pragma unroll
        for (int it = 0; it < (HIST_WORDS / 4 + 31) / 32; it++) {
            const int q = lane + 32 * it;
            if (q < HIST_WORDS / 4) hz[q] = make_int4(0, 0, 0, 0);
        }
    }
    const uint32_t *iw = reinterpret_cast<const uint32_t *>(init);
    const uint32_t hist_base = (uint32_t)__cvta_generic_to_shared(es.hist);
    for (int c0 = 0; c0 < nblk; c0 += 32) {
        reg(c0);
        __syncwarp();
        // ---- every grid block votes once per rotation with each target block of its (level, class) ----
        for_each_nonzero_cell(grid, iw, lane, [&](int cell, int gv, int sv) {
            const int d = gv - sv;
            if (d >= -6 && d <= 6) {  // other values are not colour differences and match nothing
                const uint32_t l = lut[cell];
                unsigned m = es.kmask[((l >> 12) << 4) | (d & 15)];
                const uint32_t gb = l & 0xFFFu;
                while (m) {
                    const int t = __ffs(m) - 1;
                    m &= m - 1;
                    const uint2 b = es.tb[t];
                    auto vote = [&](uint32_t v) {
                        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hist_base + ((v >> 1) << 2)),
                                     "r"(1u << ((v & 1u) << 4)) : "memory");
                    };
                    vote((b.x & 0xFFFFu) - gb);
                    if (n_rot > 1) {
                        vote((b.x >> 16) - gb);
                        vote((b.y & 0xFFFFu) - gb);
                        vote((b.y >> 16) - gb);
                    }
                }
            }
        });
        __syncwarp();
        unreg(c0);   // leave the key masks all-zero for the next pass / evaluation
        __syncwarp();
    }
    // ---- maximum over the admissible rectangles: one (rotation, dx) row per lane and pass ----
    const int rw[4] = {rc.x, rc.y, rc.z, rc.w};
    const int n0 = (rw[0] >> 8) & 0x1F, n1 = n_rot > 1 ? (rw[1] >> 8) & 0x1F : 0;
    const int n2 = n_rot > 1 ? (rw[2] >> 8) & 0x1F : 0, n3 = n_rot > 1 ? (rw[3] >> 8) & 0x1F : 0;
    const int c1 = n0, c2 = n0 + n1, c3 = c2 + n2, total = c3 + n3;
    const unsigned short *h16 = reinterpret_cast<const unsigned short *>(es.hist);
    int best = 0;  // WANT_KEY: count << 12 | 4095 - key
    for (int r = lane; r < total; r += 32) {
        const int k = (r >= c1) + (r >= c2) + (r >= c3);
        const int rk = k == 0 ? rw[0] : (k == 1 ? rw[1] : (k == 2 ? rw[2] : rw[3]));
        const int dxi = r - (k == 0 ? 0 : (k == 1 ? c1 : (k == 2 ? c2 : c3)));
        const int dx = rect_byte(rk, 0) + dxi, dz_lo = rect_byte(rk, 2), nz = (rk >> 24) & 0xFF;
        const unsigned short *row = h16 + k * HROT + (dx + 10) * HROW + (dz_lo + 10);
        const int kbase = k * 441 + (dx + 10) * 21 + (dz_lo + 10);
        for (int j = 0; j < nz; j++) {
            const int c = row[j];
            best = max(best, WANT_KEY ? ((c << 12) | (4095 - (kbase + j))) : c);
        }
    }
    best = __reduce_max_sync(0xffffffffu, best);
    __syncwarp();  // every lane is done with the table before the next evaluation clears it
    if (WANT_KEY) {
        const int count = best >> 12;
        *key_out = count > 0 ? 4095 - (best & 4095) : -1;
        return count;
    }
    return best;
}

// The env path: the target blocks are a packed list (y | x<<4 | z<<8 | code<<12; shared or global memory) and lane t
// registers entry c0 + t.
template <bool WANT_KEY>
__device__ __forceinline__ int eval_max_intersection(EvalScratch &es, const uint32_t *grid, const int8_t *init,
                                                     const uint16_t *__restrict__ lut,
                                                     const uint16_t *__restrict__ tlist, int nblk, int4 rc,
                                                     int n_rot, int lane, int *key_out) {
    int mykey = -1;
    auto reg = [&](int c0) {
        mykey = -1;
        if (c0 + lane < nblk) {
            const uint32_t e = tlist[c0 + lane];
            mykey = register_target(es, lane, e & 15, (e >> 4) & 15, (e >> 8) & 15, e >> 12);
        }
    };
    auto unreg = [&](int) { if (mykey >= 0) es.kmask[mykey] = 0u; };
    return eval_votes<WANT_KEY, true>(es, Staged<true>{grid, 0}, init, lut, nblk, reg, unreg, rc, n_rot, lane, key_out);
}

// Direct evaluation for tasks whose every rotation admits exactly ONE shift (targets spanning the whole zone in
// x and z, or invariant=False): per target block and rotation one byte of the staged grid is compared -- no
// vote table. Lane L takes blocks L, L + 32, ...
//   g8 : staged grid (shared memory, dense 1089 bytes); sub : starting grid to subtract cell-wise (global memory)
//        or nullptr; rc / n_rot as for eval_votes
__device__ __forceinline__ bool single_shift_rects(int4 rc, int n_rot) {
    const int rw[4] = {rc.x, rc.y, rc.z, rc.w};
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k < n_rot) ok = ok && ((rw[k] >> 8) & 0x1F) == 1 && ((rw[k] >> 24) & 0xFF) == 1;
    return ok;
}

// for_each_target(f) calls f(y, x, z, code) for the target blocks this lane is responsible for.
template <bool WANT_KEY, class ForEachT>
__device__ __forceinline__ int eval_single_shift(const int8_t *g8, const int8_t *__restrict__ sub,
                                                 ForEachT for_each_target, int4 rc, int n_rot, int *key_out) {
    auto at = [&](int idx) -> int { return sub ? (int)g8[idx] - (int)sub[idx] : (int)g8[idx]; };
    const int dx0 = rect_byte(rc.x, 0), dz0 = rect_byte(rc.x, 2), dx1 = rect_byte(rc.y, 0), dz1 = rect_byte(rc.y, 2);
    const int dx2 = rect_byte(rc.z, 0), dz2 = rect_byte(rc.z, 2), dx3 = rect_byte(rc.w, 0), dz3 = rect_byte(rc.w, 2);
    // cell index of (y, x_k - dx_k, z_k - dz_k) = 121 y + a_k x + b_k z + o_k  with the rotations of task.py:52-53
    const int o0 = -11 * dx0 - dz0;                       // x_0 = x,      z_0 = z
    const int o1 = -11 * dx1 + 10 - dz1;                  // x_1 = z,      z_1 = 10 - x
    const int o2 = 11 * (10 - dx2) + 10 - dz2;            // x_2 = 10 - x, z_2 = 10 - z
    const int o3 = 11 * (10 - dx3) - dz3;                 // x_3 = 10 - z, z_3 = x
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    for_each_target([&](int y, int x, int z, int code) {
        if (code == 8) return;                            // not a colour difference: matches nothing
        const int v = code < 8 ? code : code - 16;
        const int base = 121 * y;
        c0 += at(base + 11 * x + z + o0) == v;
        if (n_rot > 1) {
            c1 += at(base + 11 * z - x + o1) == v;
            c2 += at(base - 11 * x - z + o2) == v;
            c3 += at(base - 11 * z + x + o3) == v;
        }
    });
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
