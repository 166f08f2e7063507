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
//   * the staged int8 row is COMPACTED first: every lane reads its share with 16-byte shared-memory loads, finds
//     its non-zero bytes with three logic operations per word, a shuffle scan gives it its slots, and the blocks
//     land in a short list (cell | class << 11) -- everything after that runs in lock step over list entries
//     instead of diverging over bytes;
//   * matching pairs are found through a per-(level, class) bit mask over the (<= 32 per pass) target blocks:
//     lane t registers target block t with one shared-memory atomicOr; the lane that holds a grid block reads the
//     mask of its key and walks the set bits;
//   * votes are shared-memory reductions of 1 << 16 (idx & 1) on word idx >> 1; the table is cleared with
//     16-byte stores before every evaluation; the scan runs over (rotation, dx) rows with __reduce_max_sync at
//     the end (and the reference's first-max tie-break encoded in the key when argmax is wanted).
// Tasks whose every rotation admits exactly ONE shift (zone-spanning targets, invariant=False) keep the direct
// byte-compare path (eval_single_shift); in the stateless sweep, crowded ones compare whole (x, z) columns packed into
// nibbles under the four rotations (pack_columns / eval_zero_shift_packed).
#pragma once
#include <cstdint>

namespace iglu {

constexpr int HROW = 22;                         // 16-bit counters per (rotation, dx) row: 21 dz values + 1 pad
constexpr int HROT = 21 * HROW;                  // counters per rotation (462)
constexpr int HIST_WORDS = 4 * HROT / 2;         // 924 words = 3696 bytes
constexpr int NCLS = 13;                         // class of a code: 1..6 colours, 7..12 the values -6..-1 (0 unused)
constexpr int NKEY = 120;                        // 9 levels x 13 classes = 117 key masks, padded to 16 bytes
constexpr int CELL_LUT_N = 1092;                 // 1089 cells (+ pad to a multiple of 4)
constexpr int ROW_WORDS = GSTRIDE / 4;           // a staged row / window is 276 words
constexpr int WL_CAP = 32;                       // non-zero words per compaction round (one per lane)
constexpr int GL_CAP = 4 * WL_CAP;               // grid blocks per voting round

// Grid / target values are differences of colour ids, i.e. -6..6 (env.py:230, 290). code_of() maps a value
// to a 4-bit code (value & 15; 8 = "not a valid value", never matches), class_of() a valid code to 1..12.
__device__ __forceinline__ int code_of(int v) { return (v >= -6 && v <= 6) ? (v & 15) : 8; }
__device__ __forceinline__ int class_of(int code) { return code < 7 ? code : code - 3; }

// Per-warp scratch for one evaluation.
struct __align__(16) EvalScratch {
    uint32_t hist[HIST_WORDS];   // vote counters, idx16 = k * HROT + (dx + 10) * HROW + (dz + 10)
    uint32_t kmask[NKEY];        // per (level, class): bit t set when target block t of the current pass has that key
                                 // (all zero between evaluations: every pass clears what it set)
    uint2 tb[32];                // per target block of the pass: the four vote bases (16 bits each)
    uint16_t glist[GL_CAP];      // grid blocks of the current round: cell | code << 11
    uint16_t wlist[WL_CAP];      // indices of the non-zero words of the current round
};

// cell index -> (13 * y) << 8 | (HROW * x + z); built once per CTA
__device__ __forceinline__ void build_cell_lut(uint16_t *lut, int tid, int nthreads) {
    for (int c = tid; c < CELL_LUT_N; c += nthreads) {
        const int y = c / 121, r = c - 121 * y, x = r / 11, z = r - 11 * x;
        lut[c] = (uint16_t)(c < 1089 ? ((NCLS * y) << 8 | (HROW * x + z)) : 0);
    }
}
__device__ __forceinline__ int lut_x(int l) { return ((l & 0xFF) * 2979) >> 16; }   // (22 x + z) / 22 for values <= 230

// ---- compaction of a staged row -----------------------------------------------------------------------------------
// `win` is a 16-byte aligned shared-memory window of GSTRIDE bytes whose byte `off` (0..15) is cell 0: off = 0 for
// rows in the env state layout, the source address & 15 for the unaligned rows of a dense [count][1089] array that
// the stateless sweep brings in with aligned TMA bulk copies. `init` (nullptr, or a 16-byte aligned row in the env
// layout in global / shared memory; needs off == 0) is subtracted cell-wise: a "block" is a cell whose value differs
// from init's. Window bytes that are not cells (before `off`, from off + 1089 on) are masked out.
// Two lock-step phases, no per-lane loops:
//   word_scan   lane L looks at the words L, L + 32, ... (conflict-free 4-byte loads); one ballot per round tells every
//               lane where its non-zero word goes (running count + popc of the lower lanes: no shuffle scan), and the
//               indices of the non-zero words with slots [w0, w0 + 32) land in a 32-entry word list;
//   block_emit  lane i takes word i of that list, finds its non-zero bytes with three logic operations, gets its output
//               position from three ballots over the bits of its byte count and writes up to four list entries
//               (cell | code << 11; code 8 = value that is no colour difference).
// Rows with more than 32 non-zero words are handled in rounds (word_scan again with the next w0).
__device__ __forceinline__ uint32_t keep_from(int n) { return n <= 0 ? 0xFFFFFFFFu : (n >= 4 ? 0u : 0xFFFFFFFFu << (8 * n)); }
__device__ __forceinline__ uint32_t keep_below(int n) { return n >= 4 ? 0xFFFFFFFFu : (n <= 0 ? 0u : ~(0xFFFFFFFFu << (8 * n))); }
// bit 8 j + 7 <-> byte j of x is non-zero
__device__ __forceinline__ uint32_t nz_bytes(uint32_t x) { return (x | ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u; }

// grid ^ init of window word w, non-cell bytes cleared. EDGE says which non-cell bytes w may hold: 0 none, 1 bytes before
// cell 0 (words 0..3), 2 bytes after the last cell (words 272..275), 3 either.
template <int EDGE>
__device__ __forceinline__ uint32_t diff_word(const uint32_t *win, int off, const uint32_t *init32, int w) {
    uint32_t x = win[w];
    if (init32) x ^= init32[w];
    const int lo = off - 4 * w;   // bytes [lo, lo + CELLS) of this word are cells
    if (EDGE & 1) x &= keep_from(lo);
    if (EDGE & 2) x &= keep_below(lo + CELLS);
    return x;
}

// Returns the number of non-zero words of the row (warp-uniform); the indices of those with slots [w0, w0 + WL_CAP)
// are written to wl. COUNT: *nblocks = number of blocks (non-zero cells) of the row.
// EARLY (stateless sweep, w0 = 0): give up with -1 when the first 64 words already hold EARLY_WORDS non-zero ones -- a
// crowded row, for which the caller first tries the packed-column comparison that needs no word list.
constexpr int EARLY_WORDS = 24;
template <bool COUNT, bool EARLY = false>
__device__ __forceinline__ int word_scan(const uint32_t *win, int off, const int8_t *init, int lane, int w0,
                                         uint16_t *wl, int *nblocks) {
    const uint32_t *i32 = reinterpret_cast<const uint32_t *>(init);
    const unsigned lt = (1u << lane) - 1u;
    int base = -w0, nb = 0;
#pragma unroll
    for (int q = 0; q < (ROW_WORDS + 31) / 32; q++) {
        const int w = lane + 32 * q;
        uint32_t x = 0;
        if (32 * q + 32 <= ROW_WORDS || w < ROW_WORDS)   // only words 0..3 and the last ones can hold non-cell bytes
            x = q == 0 ? diff_word<1>(win, off, i32, w)
                       : (32 * q + 32 > ROW_WORDS - 4 ? diff_word<2>(win, off, i32, w) : diff_word<0>(win, off, i32, w));
        const unsigned b = __ballot_sync(0xffffffffu, x != 0);
        if (b) {   // (warp-uniform: sparse rows skip most rounds)
            if (COUNT) nb += __popc(nz_bytes(x));
            const int slot = base + __popc(b & lt);
            if (x != 0 && (unsigned)slot < (unsigned)WL_CAP) wl[slot] = (uint16_t)w;
            base += __popc(b);
        }
        if (EARLY && q == 1 && base >= EARLY_WORDS) return -1;
    }
    if (COUNT) *nblocks = __reduce_add_sync(0xffffffffu, nb);
    return base + w0;
}

// The blocks of the words wl[0 .. nw) (nw <= 32) -> gl[0 .. returned count) (<= 4 nw entries).
// The caller synchronises the warp between word_scan, block_emit and the readers of gl.
__device__ __forceinline__ int block_emit(const uint32_t *win, int off, const int8_t *init, int lane, int nw,
                                          const uint16_t *wl, uint16_t *gl) {
    uint32_t g = 0, s = 0, m = 0;
    int w = 0;
    if (lane < nw) {
        w = wl[lane];
        g = win[w];
        if (init) s = reinterpret_cast<const uint32_t *>(init)[w];
        uint32_t x = g ^ s;
        const int lo = off - 4 * w;
        if (w < 4) x &= keep_from(lo);                                // only the first / last words hold non-cell bytes
        else if (w >= (CELLS >> 2)) x &= keep_below(lo + CELLS);
        m = nz_bytes(x);
    }
    const int cnt = __popc(m);   // 0..4
    const unsigned lt = (1u << lane) - 1u;
    const unsigned b0 = __ballot_sync(0xffffffffu, cnt & 1), b1 = __ballot_sync(0xffffffffu, cnt & 2),
                   b2 = __ballot_sync(0xffffffffu, cnt & 4);
    int pos = __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
    const int cell0 = 4 * w - off;
    // the four codes at once (code_of per byte): d = g - s bytewise; valid iff d + 6 <= 12 (unsigned byte compare)
    const uint32_t d4 = __vsub4(g, s);
    const uint32_t ok = __vcmpleu4(__vadd4(d4, 0x06060606u), 0x0C0C0C0Cu);   // 0xFF per valid byte
    const uint32_t codes = (d4 & 0x0F0F0F0Fu & ok) | (0x08080808u & ~ok);
#pragma unroll
    for (int b = 0; b < 4; b++) {
        if (m & (0x80u << (8 * b))) gl[pos++] = (uint16_t)((cell0 + b) | (((codes >> (8 * b)) & 15u) << 11));
    }
    return __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
}

// Number of blocks of a staged row (warp-uniform).
__device__ __forceinline__ int count_blocks(const uint32_t *win, int off, const int8_t *init, int lane) {
    const uint32_t *i32 = reinterpret_cast<const uint32_t *>(init);
    int nb = 0;
#pragma unroll
    for (int q = 0; q < (ROW_WORDS + 31) / 32; q++) {
        const int w = lane + 32 * q;
        if (32 * q + 32 <= ROW_WORDS || w < ROW_WORDS) nb += __popc(nz_bytes(diff_word<3>(win, off, i32, w)));
    }
    return __reduce_add_sync(0xffffffffu, nb);
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

// ---- the vote ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void hist_clear(EvalScratch &es, int lane) {
    int4 *hz = reinterpret_cast<int4 *>(es.hist);
#pragma unroll
    for (int it = 0; it < (HIST_WORDS / 4 + 31) / 32; it++) {
        const int q = lane + 32 * it;
        if (q < HIST_WORDS / 4) hz[q] = make_int4(0, 0, 0, 0);
    }
}

// Target block (y13 = 13 * level, x, z, code) becomes slot t (0..31) of the current pass: its four vote bases and
// its (level, class) key mask bit. Returns the key (for un-registering), or -1 for a code that matches nothing.
// SAFE: entries that are not cells of the zone match nothing (k_widen of the overlapped step may look at a block list
// while a reset rewrites it; everything it derives from such a list is discarded, but it must stay inside its tables).
template <bool SAFE = false>
__device__ __forceinline__ int register_target(EvalScratch &es, int t, int y13, int x, int z, int code) {
    if (code == 8) return -1;
    if (SAFE && ((unsigned)x > 10u || (unsigned)z > 10u || (unsigned)y13 > 8u * NCLS)) return -1;
    // vote base of rotation k: k * HROT + (x_k + 10) * HROW + (z_k + 10) with the rotations of task.py:52-53
    const uint32_t b0 = 0 * HROT + (x + 10) * HROW + (z + 10);                 // (x, z)
    const uint32_t b1 = 1 * HROT + (z + 10) * HROW + (20 - x);                 // (z, 10 - x)
    const uint32_t b2 = 2 * HROT + (20 - x) * HROW + (20 - z);                 // (10 - x, 10 - z)
    const uint32_t b3 = 3 * HROT + (20 - z) * HROW + (x + 10);                 // (10 - z, x)
    es.tb[t] = make_uint2(b0 | (b1 << 16), b2 | (b3 << 16));
    const int key = y13 + class_of(code);
    atomicOr(&es.kmask[key], 1u << t);
    return key;
}

// The blocks glist[0 .. n) (lock step: lane i takes entries i, i + 32, ...) vote once per rotation with every
// registered target block of their (level, class).
__device__ __forceinline__ void vote_round(EvalScratch &es, const uint16_t *__restrict__ lut, int n, int n_rot, int lane) {
    const uint32_t hist_base = (uint32_t)__cvta_generic_to_shared(es.hist);
    for (int i = lane; i < n; i += 32) {
        const uint32_t it = es.glist[i];
        const int code = it >> 11;
        if (code == 8) continue;
        const uint32_t l = lut[it & 0x7FFu];
        unsigned m = es.kmask[(l >> 8) + class_of(code)];
        const uint32_t gb = l & 0xFFu;
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
}

// Maximum of the vote table over the admissible rectangles: one (rotation, dx) row per lane and pass. Returns
// (warp-uniform) the best count; if WANT_KEY also the first (rot, dx, dz) reaching it in the reference's iteration
// order, encoded k*441 + (dx+10)*21 + (dz+10), or -1 if the best is 0.
template <bool WANT_KEY, bool SAFE = false>
__device__ __forceinline__ int hist_max(EvalScratch &es, int4 rc, int n_rot, int lane, int *key_out) {
    const int rw[4] = {rc.x, rc.y, rc.z, rc.w};
    // number of dx rows of a rotation; SAFE: 0 for a rectangle that does not lie inside the 21 x 21 table
    auto rows = [](int r) -> int {
        const int dx = rect_byte(r, 0), nx = (r >> 8) & 0x1F, dz = rect_byte(r, 2), nz = (r >> 24) & 0xFF;
        return (!SAFE || (dx >= -10 && dx + nx <= 11 && dz >= -10 && dz + nz <= 11)) ? nx : 0;
    };
    const int n0 = rows(rw[0]), n1 = n_rot > 1 ? rows(rw[1]) : 0;
    const int n2 = n_rot > 1 ? rows(rw[2]) : 0, n3 = n_rot > 1 ? rows(rw[3]) : 0;
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

// Evaluate max intersection of a staged (difference) grid with a set of target blocks.
//   win / off / init : the staged grid (see the compaction above)
//   tg    : the target blocks, handed over in rounds (tg.rounds(); tg.load(r) makes round r current and returns its
//           number of blocks) and, inside a round, in passes of 32: tg.reg(c0) makes lane t register block c0 + t of the
//           round (register_target) and returns its key, or -1
//   rc    : the four rotation rectangles (see rects_from_bbox); n_rot rotations are evaluated
template <bool WANT_KEY, bool SAFE = false, class Targets>
__device__ __forceinline__ int eval_votes(EvalScratch &es, const uint32_t *win, int off, const int8_t *init,
                                          const uint16_t *__restrict__ lut, Targets &tg, int4 rc, int n_rot, int lane,
                                          int *key_out) {
    if (WANT_KEY) *key_out = -1;
    if (tg.empty()) return 0;
    hist_clear(es, lane);
    const int nwords = word_scan<false>(win, off, init, lane, 0, es.wlist, nullptr);
    if (nwords == 0) { __syncwarp(); return 0; }
    for (int w0 = 0; w0 < nwords; w0 += WL_CAP) {   // crowded grids (> 32 non-zero words): several rounds
        if (w0) {
            __syncwarp();
            word_scan<false>(win, off, init, lane, w0, es.wlist, nullptr);
        }
        __syncwarp();
        const int ng = block_emit(win, off, init, lane, min(WL_CAP, nwords - w0), es.wlist, es.glist);
        __syncwarp();
        for (int r = 0; r < tg.rounds(); r++) {
            const int nt = tg.load(r);
            for (int c0 = 0; c0 < nt; c0 += 32) {
                const int mykey = tg.reg(es, c0, lane);
                __syncwarp();
                vote_round(es, lut, ng, n_rot, lane);
                __syncwarp();
                if (mykey >= 0) es.kmask[mykey] = 0u;   // leave the key masks all-zero for the next pass / evaluation
                __syncwarp();
            }
        }
    }
    return hist_max<WANT_KEY, SAFE>(es, rc, n_rot, lane, key_out);
}

// The env path: the target blocks are a packed list (y | x<<4 | z<<8 | code<<12; shared or global memory) in one
// round; the grid is a row in the env layout (off = 0).
template <bool SAFE = false>
struct ListTargets {
    const uint16_t *__restrict__ tlist;
    int nblk;
    uint32_t e0;   // entry of the first pass, fetched when the object is made: the latency hides behind the compaction
    __device__ __forceinline__ ListTargets(const uint16_t *tl, int n, int lane) : tlist(tl), nblk(n) {
        e0 = lane < n ? tl[lane] : 0u;
    }
    __device__ __forceinline__ bool empty() const { return nblk == 0; }
    __device__ __forceinline__ int rounds() const { return 1; }
    __device__ __forceinline__ int load(int) const { return nblk; }
    __device__ __forceinline__ int reg(EvalScratch &es, int c0, int lane) const {
        if (c0 + lane >= nblk) return -1;
        const uint32_t e = c0 == 0 ? e0 : (uint32_t)tlist[c0 + lane];
        return register_target<SAFE>(es, lane, NCLS * (e & 15), (e >> 4) & 15, (e >> 8) & 15, e >> 12);
    }
};

template <bool WANT_KEY, bool SAFE = false>
__device__ __forceinline__ int eval_max_intersection(EvalScratch &es, const uint32_t *grid, const int8_t *init,
                                                     const uint16_t *__restrict__ lut,
                                                     const uint16_t *__restrict__ tlist, int nblk, int4 rc,
                                                     int n_rot, int lane, int *key_out) {
    ListTargets<SAFE> tg(tlist, nblk, lane);
    return eval_votes<WANT_KEY, SAFE>(es, grid, 0, init, lut, tg, rc, n_rot, lane, key_out);
}

// Direct evaluation for tasks whose every rotation admits exactly ONE shift (targets spanning the whole zone in
// x and z, or invariant=False): per target block and rotation one byte of the staged grid is compared -- no
// vote table.
//   g8 : staged grid (shared memory, cell c at g8[c]); sub : starting grid to subtract cell-wise (global memory)
//        or nullptr; for_each_target(f) calls f(y, x, z, code) for the target blocks this lane is responsible for
__device__ __forceinline__ bool single_shift_rects(int4 rc, int n_rot) {
    const int rw[4] = {rc.x, rc.y, rc.z, rc.w};
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k < n_rot) ok = ok && ((rw[k] >> 8) & 0x1F) == 1 && ((rw[k] >> 24) & 0xFF) == 1;
    return ok;
}

template <bool WANT_KEY, bool SAFE = false, class ForEachT>
__device__ __forceinline__ int eval_single_shift(const int8_t *g8, const int8_t *__restrict__ sub,
                                                 ForEachT for_each_target, int4 rc, int n_rot, int *key_out) {
    // (SAFE: the index into the global-memory row is clamped; valid tasks never leave the row)
    auto at = [&](int idx) -> int {
        return sub ? (int)g8[idx] - (int)sub[SAFE ? min(max(idx, 0), CELLS - 1) : idx] : (int)g8[idx];
    };
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

// Zone-spanning targets in the stateless sweep, dense form: when every rotation admits exactly one shift, that shift is
// (0, 0) (a bounding box that leaves one admissible dx spans x = 0..10, task.py:67-72), so rotation k just compares
// every target cell with the grid cell it rotates onto -- no compacted block list (such targets hold hundreds of
// blocks). The rotations act on the 121 (x, z) columns only (task.py:52-53): column 11 x + z goes to 11 z + 10 - x
// (rotation 1), 120 - column (rotation 2) and 11 (10 - z) + x (rotation 3); a lane owns columns lane, lane + 32, ... and
// walks the nine levels.
// t8 / g8: cell 0 of the staged target / grid rows. Returns what eval_single_shift returns for shifts (0, 0).
template <bool WANT_KEY>
__device__ __forceinline__ int eval_zero_shift_dense(const int8_t *t8, const int8_t *g8, int n_rot, int lane, int *key_out) {
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int col = lane + 32 * q;
        if (col < 121) {
            const int x = (col * 373) >> 12, z = col - 11 * x;   // col / 11 for col < 121
            const int col1 = 11 * z + 10 - x, col2 = 120 - col, col3 = 11 * (10 - z) + x;
#pragma unroll
            for (int y = 0; y < 9; y++) {
                const int t = t8[121 * y + col];
                const bool on = t != 0 && (unsigned)(t + 6) <= 12u;   // values that are no colour difference match nothing
                c0 += on && g8[121 * y + col] == t;
                if (n_rot > 1) {
                    c1 += on && g8[121 * y + col1] == t;
                    c2 += on && g8[121 * y + col2] == t;
                    c3 += on && g8[121 * y + col3] == t;
                }
            }
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
    if (WANT_KEY) *key_out = best > 0 ? kbest * 441 + 10 * 21 + 10 : -1;
    return best;
}

// ---- packed columns (round 2, second form of the dense zero-shift path) ----------------------------------------------
// The rotations permute the 121 (x, z) columns of the zone as a whole, so a column's nine levels can be compared in one
// go once they sit in one register pair: nibble y of `lo` (levels 0..7) / `hi` (level 8) holds value + 8, i.e. 2..14 for
// the colour differences -6..6 and 8 for an empty cell. A lane owns the columns lane, lane + 32, lane + 64, lane + 96
// (conflict-free byte loads: a warp reads 32 consecutive bytes per level), the sum is built with one multiply-add per
// byte (borrows of negative values cancel against the +8 bias), and the running min / max of the bytes tells whether
// the row holds anything outside -6..6 (then the nibbles are meaningless and the caller takes the byte-wise path).
struct PackedCols {
    uint32_t lo[4], hi[4];
    int mn, mx;
    __device__ __forceinline__ bool in_range() const { return mn >= -6 && mx <= 6; }   // per lane
};

__device__ __forceinline__ void pack_columns(const int8_t *r8, int lane, PackedCols &p) {
    p.mn = 0; p.mx = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int col = lane + 32 * q;
        uint32_t lo = 0x88888888u, hi = 8u;
        if (q < 3 || col < 121) {
#pragma unroll
            for (int y = 0; y < 8; y++) {
                const int v = r8[121 * y + col];
                lo += (uint32_t)v << (4 * y);
                p.mn = min(p.mn, v); p.mx = max(p.mx, v);
            }
            const int v = r8[121 * 8 + col];
            hi += (uint32_t)v;
            p.mn = min(p.mn, v); p.mx = max(p.mx, v);
        }
        p.lo[q] = lo; p.hi[q] = hi;
        asm volatile("" ::: "memory");   // one column's nine loads in flight at a time: hoisting all 36 costs spills
    }
}

// x / z extent of the occupied columns of a packed row. Warp-uniform result.
__device__ __forceinline__ void packed_bbox(const PackedCols &p, int lane, int &x0, int &x1, int &z0, int &z1) {
    x0 = 10; x1 = 0; z0 = 10; z1 = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int col = lane + 32 * q;
        if (p.lo[q] != 0x88888888u || p.hi[q] != 8u) {
            const int x = (col * 373) >> 12, z = col - 11 * x;
            x0 = min(x0, x); x1 = max(x1, x); z0 = min(z0, z); z1 = max(z1, z);
        }
    }
    x0 = __reduce_min_sync(0xffffffffu, x0); x1 = __reduce_max_sync(0xffffffffu, x1);
    z0 = __reduce_min_sync(0xffffffffu, z0); z1 = __reduce_max_sync(0xffffffffu, z1);
}

// bit 4 y set <-> nibble y of x is non-zero
__device__ __forceinline__ uint32_t nz_nibbles(uint32_t x) {
    const uint32_t a = x | (x >> 1);
    return (a | (a >> 2)) & 0x11111111u;
}

// eval_zero_shift_dense on packed rows (both in range): `xch` is per-warp shared scratch of >= 121 uint2 through which
// the lanes hand each other the grid columns their target columns rotate onto. Same result, same tie-break.
template <bool WANT_KEY>
__device__ __forceinline__ int eval_zero_shift_packed(const PackedCols &t, const PackedCols &g, int n_rot, int lane,
                                                      uint2 *xch, int *key_out) {
    if (n_rot > 1) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int col = lane + 32 * q;
            if (q < 3 || col < 121) xch[col] = make_uint2(g.lo[q], g.hi[q]);
        }
        __syncwarp();
    }
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int col = lane + 32 * q;
        if (q < 3 || col < 121) {
            const uint32_t tl = t.lo[q], th = t.hi[q];
            const uint32_t on = nz_nibbles(tl ^ 0x88888888u);   // levels where the target column holds a block
            const bool on8 = th != 8u;
            auto count = [&](uint32_t gl, uint32_t gh) {
                const uint32_t a = tl ^ gl, b = a | (a >> 1);
                return __popc(on & ~(b | (b >> 2))) + (int)(on8 && gh == th);
            };
            c0 += count(g.lo[q], g.hi[q]);
            if (n_rot > 1) {
                const int x = (col * 373) >> 12, z = col - 11 * x;
                const uint2 g1 = xch[11 * z + 10 - x], g2 = xch[120 - col], g3 = xch[11 * (10 - z) + x];
                c1 += count(g1.x, g1.y);
                c2 += count(g2.x, g2.y);
                c3 += count(g3.x, g3.y);
            }
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
    if (WANT_KEY) *key_out = best > 0 ? kbest * 441 + 10 * 21 + 10 : -1;
    return best;
}

}  // namespace iglu
