// iglu_b200.cu -- sm_100a kernels + C ABI (include/iglu_b200.h) of the IGLU Gridworld step path.
//
// k_step<KIND, PHASES>: one source, three instantiations (launch_step picks):
//   PHASES = 1  physics launch: phase A only (one thread per env, no shared memory);
//   PHASES = 2  observe launch: phases B + C, persistent 16-warp CTAs, env groups handed out dynamically;
//   PHASES = 3  fused launch:   A + B + C, the TMA ring filled BEFORE phase A so that the HBM latency hides
//               behind the physics.
//   phase A (one thread per env): auto-reset (drawing from the task pool when one is attached), action parse,
//            World.step in fp64 (iglu_physics.cuh), small observations (agentPos / inventory / compass);
//   phase B (warp-cooperative, one env at a time): the env's int8 grid and its target block list are brought
//            into shared memory by TMA bulk copies (cp.async.bulk + mbarrier) through a per-warp ring; the
//            grid is widened into the int32 obs['grid'] with 16-byte streaming stores and
//            Task.maximal_intersection is evaluated (iglu_reward.cuh: shift voting) when required;
//   phase C (one thread per env): Task.step_intersection bookkeeping, reward, done (task.py:103-119,
//            env.py:290-296).
// k_task_prep (Task.__init__), k_reset (GridWorld.reset), k_max_intersection (stateless reward sweep).
// No collective, no host round trip; every entry point only enqueues work on the caller's stream
// (iglu_step_walking_host additionally synchronises it).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/iglu_b200.h"
#include "iglu_physics.cuh"
#include "iglu_reward.cuh"
#include "iglu_render.cuh"

namespace iglu {

#ifndef IGLU_WARPS_PER_CTA
#define IGLU_WARPS_PER_CTA 8
#endif
#ifndef IGLU_RING
#define IGLU_RING 4
#endif
constexpr int WARPS_PER_CTA = IGLU_WARPS_PER_CTA;
#ifndef IGLU_MI_WARPS
#define IGLU_MI_WARPS 24
#endif
#ifndef IGLU_MI_CTAS
#define IGLU_MI_CTAS 1
#endif
// stateless reward sweep: one 24-warp CTA per SM (216 KB of shared memory)
constexpr int MI_WARPS = IGLU_MI_WARPS;
constexpr int MI_CTAS = IGLU_MI_CTAS;
constexpr int CTAS_PER_SM = 2;   // 16 warps / SM: measured faster than 20 / 24 / 32 warps with fewer registers (profiles/)
constexpr int NB = IGLU_RING;    // TMA ring depth per warp
constexpr int TL_HDR = 8;    // uint16 slots holding the 16-byte rectangle header of a tlist row

struct __align__(16) Slot {
    uint32_t grid[GSTRIDE / 4];        // int8 grid of one env (1089 cells + zero pad)
};

constexpr size_t TABLE_BYTES = (CELL_LUT_N * sizeof(uint16_t) + 15) & ~(size_t)15;  // cell -> (y, 22 x + z) table per CTA

// cell -> (13 * y) << 8 | (22 x + z), filled by iglu_ctx_create (same contents as build_cell_lut)
__device__ uint16_t g_cell_lut[CELL_LUT_N];

struct StepParams {
    iglu_state st;
    iglu_obs obs;
    iglu_config cfg;
    const int32_t *actions;    // walking discrete
    const float *movement;     // flying [n][3]
    const float *camera;       // flying / walking dict [n][2]
    const int32_t *inventory;  // flying [n]
    const int32_t *placement;  // flying [n]
    const int32_t *buttons;    // walking dict [n][8]
    iglu_task_pool pool;       // device-resident task generator (m == 0: none)
    double *reward_host;       // mapped pinned host copies of obs.reward / obs.done (iglu_step_walking_host), or null
    uint8_t *done_host;
    int dynamic;               // observe launch: groups handed out through st.sched[0] instead of a static deal
    int group;                 // envs per group (<= 32): a warp handles one group at a time, lane j <-> env j of it
    int tiles;                 // tiled observe launch: tiles per CTA (group = envs per tile)
    int64_t tail_begin;        // ... and the first env of the dynamically scheduled tail
    int overlap;               // physics launch of the overlapped step (k_widen alongside, k_commit afterwards)
};

// istate w7 as the physics launch of the overlapped step leaves it for k_commit:
// wrong_placement (2 bits, -1..1) | (edited cell + 1) << 2 | (new value of that cell) << 13 | reset << 17
__device__ __forceinline__ int pack_w7(int wrong, int edit_cell, int edit_val, bool reset) {
    return (wrong & 3) | ((edit_cell + 1) << 2) | ((edit_val & 15) << 13) | ((int)reset << 17);
}

// ---- TMA (bulk async copy) + mbarrier helpers -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    }
}

__device__ __forceinline__ int pack16(int a, int b) { return (a & 0xFFFF) | (b << 16); }
__device__ __forceinline__ int lo16(int w) { return (int)(int16_t)(w & 0xFFFF); }
__device__ __forceinline__ int hi16(int w) { return (int)(int16_t)((unsigned)w >> 16); }

// Widen one env's int8 grid (shared memory) into its int32 observation row. The row starts at element
// e0 of the obs array (4-byte aligned only); the body is written as 16-byte streaming stores, fully
// unrolled with all shared-memory loads first. (A variant that staged the widened row in shared memory and
// wrote it with one TMA bulk store -- cp.async.bulk.global.shared::cta -- was measured and was not faster:
// profiles/r01_launch_policy.txt.)
__device__ __forceinline__ void write_grid_obs(int32_t *__restrict__ obs_grid, int64_t e0, const uint32_t *sg,
                                               int lane) {
    const int8_t *sg8 = reinterpret_cast<const int8_t *>(sg);
    int32_t *og = obs_grid + e0;
    const int h = (int)((4 - (e0 & 3)) & 3);  // elements before the first 16-byte boundary
    if (lane < h) og[lane] = (int32_t)sg8[lane];
    const int nbody = (CELLS - h) >> 2;        // 271 or 272 sixteen-byte chunks
    int4 *o4 = reinterpret_cast<int4 *>(og + h);
    uint32_t lo[9], hi[9];
#pragma unroll
    for (int it = 0; it < 9; it++) {
        const int m = lane + 32 * it;
        const bool on = it < 8 || m < nbody;
        lo[it] = on ? sg[m] : 0u;
        hi[it] = on ? sg[m + 1] : 0u;
    }
    // byte h + k of the pair {lo, hi}, sign-extended to 32 bits, is ONE prmt: selector nibble 8 | i replicates the
    // sign of byte i (PTX prmt default mode; the __byte_perm intrinsic masks that bit, hence the inline PTX)
    const uint32_t sel0 = 0x8880u + 0x1111u * (uint32_t)h;
#pragma unroll
    for (int it = 0; it < 9; it++) {
        const int m = lane + 32 * it;
        int4 v;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(v.x) : "r"(lo[it]), "r"(hi[it]), "r"(sel0));
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(v.y) : "r"(lo[it]), "r"(hi[it]), "r"(sel0 + 0x1111u));
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(v.z) : "r"(lo[it]), "r"(hi[it]), "r"(sel0 + 0x2222u));
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(v.w) : "r"(lo[it]), "r"(hi[it]), "r"(sel0 + 0x3333u));
        if (it < 8 || m < nbody) __stcs(o4 + m, v);
    }
    const int t0 = h + 4 * nbody;
    if (lane < CELLS - t0) og[t0 + lane] = (int32_t)sg8[t0 + lane];
}

// ---- device-resident task generator (include/iglu_b200.h iglu_task_pool) ---------------------------
// Uniform draw of a pool index for the k-th episode of (global) env `env`: splitmix64 finaliser of a
// per-(env, episode) counter, mapped to [0, m) by multiply-shift. Stateless, so shards reproduce it.
__device__ __forceinline__ int pool_draw(uint64_t seed, int64_t env, uint32_t episode, int64_t m) {
    uint64_t x = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(env + 1) + 0xD1B54A32D192ED03ull * (uint64_t)episode;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (int)__umulhi((uint32_t)(x >> 32), (uint32_t)m);   // ((x >> 32) * m) >> 32, m < 2^31
}

// Per-env part of a draw (one thread): choose the index, advance the cursor, copy the tmeta column.
// Returns the pool index.
__device__ __noinline__ int pool_next(const iglu_task_pool &pool, int32_t *__restrict__ tmeta, int64_t n, int64_t i) {
    const int64_t m = pool.m;
    const int ep = pool.cursor[n + i];
    int j = -1;
    if (pool.next_task) {
        const int nt = pool.next_task[i];
        if (nt >= 0 && nt < m) { j = nt; pool.next_task[i] = -1; }
    }
    if (j < 0) j = pool_draw(pool.seed, pool.env_offset + i, (uint32_t)ep, m);
    pool.cursor[i] = j;
    pool.cursor[n + i] = ep + 1;
#pragma unroll
    for (int c = 0; c < 6; c++) tmeta[c * n + i] = pool.tmeta[c * m + j];
    tmeta[6 * n + i] = 1;
    return j;
}

// Warp-cooperative part: pool row j -> the env's own target / init / tlist rows (+ obs['target_grid']).
__device__ __noinline__ void pool_fetch(const iglu_task_pool &pool, int cap, int j, int nblk, int8_t *__restrict__ target,
                                        int8_t *__restrict__ init, uint16_t *__restrict__ tlist,
                                        int32_t *__restrict__ target_obs, int lane) {
    const int4 *st = reinterpret_cast<const int4 *>(pool.target + (int64_t)j * GSTRIDE);
    const int4 *si = reinterpret_cast<const int4 *>(pool.init + (int64_t)j * GSTRIDE);
    const int4 *sl = reinterpret_cast<const int4 *>(pool.tlist + (int64_t)j * cap);
    for (int q = lane; q < GSTRIDE / 16; q += 32) {
        reinterpret_cast<int4 *>(target)[q] = st[q];
        reinterpret_cast<int4 *>(init)[q] = si[q];
    }
    const int nl = (2 * (TL_HDR + nblk) + 15) >> 4;
    for (int q = lane; q < nl; q += 32) reinterpret_cast<int4 *>(tlist)[q] = sl[q];
    if (target_obs) {
        const int8_t *t8 = pool.target + (int64_t)j * GSTRIDE;
        for (int c = lane; c < CELLS; c += 32) target_obs[c] = (int32_t)t8[c];
    }
}

// PHASES selects what one launch does (same source, three instantiations):
//   3  fused:        phase A (physics) + B (observation / reward) + C (bookkeeping) -- one launch per step
//   1  physics only: phase A; leaves wrong_placement in istate w7 for the second launch
//   2  observe only: phases B + C on the state left by a PHASES = 1 launch. It needs no fp64 physics
//      registers, so it runs at 64 registers / 32 warps per SM with small (<= 32 env) chunks per warp,
//      which is what the latency-bound per-env chain of phase B wants.
#ifndef IGLU_OBS_WARPS
#define IGLU_OBS_WARPS 16
#endif
#ifndef IGLU_OBS_RING
#define IGLU_OBS_RING 2
#endif
#ifndef IGLU_OBS_CTAS
#define IGLU_OBS_CTAS 2        // resident CTAs per SM of the observe launch
#endif
#ifndef IGLU_OBS_DYN_GROUP
#define IGLU_OBS_DYN_GROUP 8   // envs per dynamically scheduled group of the observe launch
#endif
#ifndef IGLU_PHYS_WARPS
#define IGLU_PHYS_WARPS 2      // warps per CTA of the physics launch (small CTAs fill the registers k_widen leaves best)
#endif
#ifndef IGLU_PHYS_CTAS
#define IGLU_PHYS_CTAS 10      // physics CTAs per SM the register budget is set for (10 x 64 threads: 96 registers, no
                               // spills; 128 registers run at the same speed alone but leave fewer CTAs next to k_widen)
#endif
template <int PHASES> struct StepCfg {
    static constexpr int WPC = PHASES == 2 ? IGLU_OBS_WARPS : (PHASES == 1 ? IGLU_PHYS_WARPS : WARPS_PER_CTA);
    static constexpr int NBUF = PHASES == 2 ? IGLU_OBS_RING : NB;
};
template <int NBUF> struct __align__(16) WarpScratchT {
    Slot slot[NBUF];
    EvalScratch es;
    unsigned long long bar[NBUF];      // mbarriers of the ring
};
template <int PHASES> constexpr size_t step_smem_bytes() {
    return PHASES == 1 ? 0 : TABLE_BYTES + StepCfg<PHASES>::WPC * sizeof(WarpScratchT<StepCfg<PHASES>::NBUF>);
}

template <int KIND, int PHASES>
__global__ void __launch_bounds__(StepCfg<PHASES>::WPC * 32, PHASES == 1 ? IGLU_PHYS_CTAS : (PHASES == 2 ? IGLU_OBS_CTAS : CTAS_PER_SM)) k_step(const __grid_constant__ StepParams p) {
    constexpr int WPC = StepCfg<PHASES>::WPC;
    constexpr int NBUF = StepCfg<PHASES>::NBUF;
    constexpr bool HAS_A = PHASES != 2, HAS_B = PHASES != 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint16_t *lut = reinterpret_cast<uint16_t *>(smem_raw);
    WarpScratchT<NBUF> *scratch = reinterpret_cast<WarpScratchT<NBUF> *>(smem_raw + TABLE_BYTES);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpScratchT<NBUF> &ws = scratch[HAS_B ? wid : 0];
    if (PHASES == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // let the observe launch get resident early
    if (HAS_B) {
        build_cell_lut(lut, threadIdx.x, WPC * 32);
        for (int q = lane; q < NKEY; q += 32) ws.es.kmask[q] = 0u;  // invariant of eval_max_intersection
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < NBUF; s++) mbar_init(&ws.bar[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    // Programmatic dependent launch: the prologue above overlaps the tail of the previous launch on the stream;
    // nothing that launch wrote is read before this point. (A no-op when launched without the PDL attribute.)
    // (The physics launch of the overlapped step is itself a programmatic dependent of k_widen and must NOT wait for
    // it here -- it runs next to it; its last CTA waits at the very end instead, see below.)
    if (!(PHASES == 1 && p.overlap)) asm volatile("griddepcontrol.wait;" ::: "memory");
    // the fused launch is the programmatic dependent of whatever precedes it on the stream (the previous step's fused
    // launch, say): it may become resident while that grid drains -- and waits, like this one, before it reads.
    // (Measured: the same chaining costs 3-5 % for the physics launch of the two-launch form -- early-resident physics
    // CTAs get in the way of the observe launch's tail -- so that launch stays a normal one.)
    if (PHASES == 3) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int64_t n = p.st.n;
    const int GS = p.group;
    const int64_t ngroups = (n + GS - 1) / GS;
    const bool flying = KIND == IGLU_FLYING;
    const bool every_step = p.cfg.reward_mode == IGLU_REWARD_EVERY_STEP;
    const bool grid_obs = p.cfg.write_grid_obs == 1 && p.obs.grid != nullptr;  // rewrite obs['grid'] every step
    const bool inc_obs = p.cfg.write_grid_obs == 2 && p.obs.grid != nullptr;   // keep it up to date cell by cell
    const bool early = grid_obs || every_step;  // the set of envs phase B touches is known before phase A
    const bool alternate = (p.cfg.reserved & 1) == 0;  // reserved bit 0 = 1 disables the phase alternation
    const int cap = p.cfg.tlist_cap;
    uint32_t seq_issue = 0, seq_cons = 0;  // ring sequence numbers (warp-uniform, live across groups)
    int evals_done = 0;                    // committed-candidate evaluations of this warp (measurement: sched[2..3])
    uint32_t slot_addr[NBUF], bar_addr[NBUF];  // shared-window addresses of the ring (computed once)
#pragma unroll
    for (int s = 0; s < NBUF; s++) {
        slot_addr[s] = HAS_B ? smem_u32(&ws.slot[s]) : 0u;
        bar_addr[s] = HAS_B ? smem_u32(&ws.bar[s]) : 0u;
    }

    // groups are dealt warp-major (warp w of every CTA before warp w + 1 of any) so that a partial last
    // wave spreads evenly over the SMs
    // The observe launch hands groups out dynamically (one atomic per group on a counter the physics launch
    // zeroed; the next index is fetched while the current group is processed): per-env cost varies with the
    // evaluation, and a static deal leaves the SMs unevenly loaded in the single wave.
    // Every warp's FIRST group is its static index (thousands of warps asking one counter at kernel start would
    // serialise in the L2: 22 % of round 1's stall samples); the counter only deals the groups after those.
    const bool dynamic = PHASES == 2 && p.dynamic;
    if (PHASES == 1 && !p.overlap && p.st.sched != nullptr && blockIdx.x == 0 && threadIdx.x == 0) p.st.sched[0] = 0;
    const int total_warps = (int)gridDim.x * WPC;
    int64_t grp = (int64_t)wid * gridDim.x + blockIdx.x;
    while (grp < ngroups) {
        int grp_next = 0;
        if (dynamic && lane == 0) grp_next = total_warps + atomicAdd(p.st.sched, 1);
        auto env_of = [&](int j) -> int64_t { return grp * GS + j; };
        const int64_t i = env_of(lane);
        const bool valid = lane < GS && i < n;
        const int64_t ii = valid ? i : n - 1;  // clamp: invalid lanes read a real env, write nothing

        // ------------------------------------------------------------- phase A
        int w[8], tm[6];
#pragma unroll
        for (int c = 0; c < 8; c++) w[c] = (HAS_A || c >= 3) ? p.st.istate[c * n + ii] : 0;  // w0..w2: inventory, physics only
#pragma unroll
        for (int c = 0; c < 6; c++) tm[c] = (HAS_A || c < 2 || c == 5) ? p.st.tmeta[c * n + ii] : 0;
        Agent a;
        if (HAS_A) {
            a.x = p.st.pose[0 * n + ii]; a.y = p.st.pose[1 * n + ii]; a.z = p.st.pose[2 * n + ii];
            a.yaw = p.st.pose[3 * n + ii]; a.pitch = p.st.pose[4 * n + ii]; a.dy = p.st.pose[5 * n + ii];
        }
        a.inv01 = w[0]; a.inv23 = w[1]; a.inv45 = w[2];
        a.active = w[3] & 0xFF; a.tis = (w[3] >> 8) & 0xFF; a.err = (w[3] >> 24) & 0xFF;
        const bool done_prev = ((w[3] >> 16) & 0xFF) != 0;
        int step_no = w[4], dsize = w[5], max_int = w[6];

        // auto-reset (env.py:206-261): grid <- starting grid, cooperatively, 16 bytes per lane
        const bool do_reset = HAS_A && valid && p.cfg.auto_reset && done_prev;
        int pool_j = -1;  // task generator: the reset first draws the env's next task (env.py:213-216)
        if (HAS_A && p.pool.m > 0 && do_reset) {
            pool_j = pool_next(p.pool, p.st.tmeta, n, i);
#pragma unroll
            for (int c = 0; c < 6; c++) tm[c] = p.pool.tmeta[c * p.pool.m + pool_j];
        }
        const int nblk = tm[0], target_size = tm[1];
        const bool has_start = tm[5] != 0;
        if (HAS_A) {
            unsigned rm = __ballot_sync(0xffffffffu, do_reset);
            if (rm) {
                while (rm) {
                    const int j = __ffs(rm) - 1;
                    rm &= rm - 1;
                    const int64_t ej = env_of(j);
                    const bool hs = __shfl_sync(0xffffffffu, (int)has_start, j) != 0;
                    const int pj = __shfl_sync(0xffffffffu, pool_j, j);
                    const int8_t *si = p.st.init + ej * GSTRIDE;
                    if (pj >= 0) {
                        pool_fetch(p.pool, cap, pj, __shfl_sync(0xffffffffu, nblk, j), p.st.target + ej * GSTRIDE,
                                   p.st.init + ej * GSTRIDE, p.st.tlist + ej * cap,
                                   p.obs.target_grid ? p.obs.target_grid + ej * CELLS : nullptr, lane);
                        si = p.pool.init + (int64_t)pj * GSTRIDE;
                    }
                    int4 *dst = reinterpret_cast<int4 *>(p.st.grid + ej * GSTRIDE);
                    const int4 *src = reinterpret_cast<const int4 *>(si);
                    for (int q = lane; q < GSTRIDE / 16; q += 32) dst[q] = hs ? src[q] : make_int4(0, 0, 0, 0);
                    if (inc_obs)  // incremental observation: a reset rewrites the whole row
                        for (int c = lane; c < CELLS; c += 32) p.obs.grid[ej * CELLS + c] = hs ? (int32_t)si[c] : 0;
                }
                if (HAS_B) {  // the TMA reads below go through the async proxy: order them after these writes
                    __threadfence();
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                __syncwarp();
            }
            if (do_reset) {
                a.x = p.st.init_pose[0 * n + i]; a.y = p.st.init_pose[1 * n + i]; a.z = p.st.init_pose[2 * n + i];
                a.yaw = p.st.init_pose[3 * n + i]; a.pitch = p.st.init_pose[4 * n + i];
                a.inv01 = tm[2]; a.inv23 = tm[3]; a.inv45 = tm[4];
                step_no = 0; dsize = 0; max_int = 0;  // dy / tis / active survive (reference quirk)
            }
        }

        // ring producer: lane 0 issues the bulk copy of env j's grid row into the next slot
        auto issue = [&](int j) {
            if (lane == 0) {
                const int64_t ej = env_of(j);
                uint32_t sa = slot_addr[0], ba = bar_addr[0];
#pragma unroll
                for (int s = 1; s < NBUF; s++) if ((int)(seq_issue % NBUF) == s) { sa = slot_addr[s]; ba = bar_addr[s]; }
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"((uint32_t)GSTRIDE) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(sa), "l"(p.st.grid + ej * GSTRIDE), "r"((uint32_t)GSTRIDE), "r"(ba) : "memory");
            }
            seq_issue++;
        };

        int dsize_old = dsize;
        int edit_cell = -1, edit_val = 0;
        auto run_physics = [&]() {
            Action act;
            act.s0 = act.s1 = act.jump = act.cam0 = act.cam1 = 0.0;
            act.hotbar = 0; act.remove = act.add = false;
            if (KIND == IGLU_WALKING_DISCRETE) {  // world.py:360-394
                const int ac = p.actions[ii];
                if (ac == 1) act.s0 = -1.0;
                else if (ac == 2) act.s0 = 1.0;
                else if (ac == 3) act.s1 = -1.0;
                else if (ac == 4) act.s1 = 1.0;
                else if (ac == 5) act.jump = 1.0;
                else if (ac >= 6 && ac <= 11) act.hotbar = ac - 5;
                else if (ac == 12) act.cam0 = -5.0;
                else if (ac == 13) act.cam0 = 5.0;
                else if (ac == 14) act.cam1 = -5.0;
                else if (ac == 15) act.cam1 = 5.0;
                else if (ac == 16) act.remove = true;
                else if (ac == 17) act.add = true;
            } else if (KIND == IGLU_FLYING) {  // world.py:416-432 (float32 widened to fp64)
                act.s0 = (double)p.movement[3 * ii + 0];
                act.s1 = (double)p.movement[3 * ii + 1];
                act.jump = (double)p.movement[3 * ii + 2];
                act.cam0 = (double)p.camera[2 * ii + 0];
                act.cam1 = (double)p.camera[2 * ii + 1];
                act.hotbar = p.inventory[ii];
                const int pl = p.placement[ii];
                act.add = pl == 1;
                act.remove = pl == 2;
            } else {  // world.py:396-414
                const int32_t *b = p.buttons + 8 * ii;
                if (b[0]) act.s0 += -1.0;
                if (b[1]) act.s0 += 1.0;
                if (b[2]) act.s1 += -1.0;
                if (b[3]) act.s1 += 1.0;
                act.jump = (double)b[4];
                act.remove = b[5] != 0;
                act.add = b[6] != 0;
                act.hotbar = b[7];
                act.cam0 = (double)p.camera[2 * ii + 0];
                act.cam1 = (double)p.camera[2 * ii + 1];
            }
            if (valid) {
                step_no += 1;  // env.py:276
                world_step(p.st.grid + i * GSTRIDE, p.st.init + i * GSTRIDE, has_start, a, act, flying,
                           p.cfg.select_and_place != 0, dsize, edit_cell, edit_val);
                p.st.pose[0 * n + i] = a.x; p.st.pose[1 * n + i] = a.y; p.st.pose[2 * n + i] = a.z;
                p.st.pose[3 * n + i] = a.yaw; p.st.pose[4 * n + i] = a.pitch; p.st.pose[5 * n + i] = a.dy;
                // env.py:283-289 (float32 casts round to nearest even)
                if (p.obs.agent_pos) {
                    float *ap = p.obs.agent_pos + 5 * i;
                    ap[0] = (float)a.x; ap[1] = (float)a.y; ap[2] = (float)a.z;
                    ap[3] = (float)a.pitch; ap[4] = (float)a.yaw;
                }
                if (p.obs.inventory) {
                    float *iv = p.obs.inventory + 6 * i;
                    iv[0] = (float)lo16(a.inv01); iv[1] = (float)hi16(a.inv01); iv[2] = (float)lo16(a.inv23);
                    iv[3] = (float)hi16(a.inv23); iv[4] = (float)lo16(a.inv45); iv[5] = (float)hi16(a.inv45);
                }
                if (p.obs.compass) p.obs.compass[i] = (float)(a.yaw - 180.0);
            }
            __syncwarp();
        };

        // ring consumer: one env at a time -- wait for its slot, optionally apply this step's edit to the
        // staged grid, widen it into obs['grid'], evaluate maximal_intersection, refill the slot.
        unsigned pend = 0, todo = 0;
        const unsigned hm = __ballot_sync(0xffffffffu, has_start);
        int mi_eval = 0;
        auto ring_start = [&](unsigned mask) {
            todo = pend = mask;
            for (int f = 0; f < NBUF && pend; f++) {
                issue(__ffs(pend) - 1);
                pend &= pend - 1;
            }
        };
        auto ring_drain = [&](bool obs_on, unsigned em, bool patch, bool keep) {
            while (todo) {
                const int j = __ffs(todo) - 1;
                todo &= todo - 1;
                const int64_t ej = env_of(j);
                const bool ev = (em >> j) & 1u, hs = (hm >> j) & 1u;
                Slot &sl = ws.slot[seq_cons % NBUF];
                // the task's rectangles (16-byte header of its block list, one broadcast load): asked for before the
                // wait so that the latency hides behind the grid copy and the observation write
                const uint16_t *trow = p.st.tlist + ej * cap;
                int4 rc = make_int4(0, 0, 0, 0);
                if (ev) rc = __ldg(reinterpret_cast<const int4 *>(trow));
                mbar_wait(&ws.bar[seq_cons % NBUF], (seq_cons / NBUF) & 1u);
                if (patch) {  // the copy may have raced with this step's edit of the same env
                    const int ec = __shfl_sync(0xffffffffu, edit_cell, j);
                    const int evl = __shfl_sync(0xffffffffu, edit_val, j);
                    if (ec >= 0 && lane == 0) reinterpret_cast<int8_t *>(sl.grid)[ec] = (int8_t)evl;
                    __syncwarp();
                }
                if (obs_on) write_grid_obs(p.obs.grid, ej * CELLS, sl.grid, lane);
                if (ev) {
                    const int nb = __shfl_sync(0xffffffffu, nblk, j);
                    const uint16_t *tl = trow + TL_HDR;
                    int r;
                    if (single_shift_rects(rc, 4)) {  // zone-spanning target: one shift per rotation, compare bytes
                        auto each = [&](auto f) {
                            for (int b = lane; b < nb; b += 32) {
                                const uint32_t e = tl[b];
                                f(e & 15, (e >> 4) & 15, (e >> 8) & 15, e >> 12);
                            }
                        };
                        r = eval_single_shift<false>(reinterpret_cast<const int8_t *>(sl.grid),
                                                     hs ? p.st.init + ej * GSTRIDE : nullptr, each, rc, 4, nullptr);
                    } else {
                        r = eval_max_intersection<false>(ws.es, sl.grid, hs ? p.st.init + ej * GSTRIDE : nullptr, lut, tl,
                                                         nb, rc, 4, lane, nullptr);
                    }
                    if (keep && lane == j) mi_eval = r;
                    if (keep) evals_done++;
                }
                __syncwarp();  // every lane is done with the slot before it is refilled
                seq_cons++;
                if (pend) {
                    issue(__ffs(pend) - 1);
                    pend &= pend - 1;
                }
            }
        };

        if (PHASES == 1) {
            run_physics();
            if (inc_obs && edit_cell >= 0) p.obs.grid[i * CELLS + edit_cell] = edit_val;
        } else if (PHASES == 2) {
            // state was advanced by the physics launch: wrong_placement (task.py:111) is in w7
            // an all-zero difference grid (dsize == 0) scores 0 without looking: no target cell can equal 0
            const unsigned em = __ballot_sync(0xffffffffu, valid && nblk > 0 && dsize != 0 && (every_step || w[7] != 0));
            ring_start(em | (grid_obs ? __ballot_sync(0xffffffffu, valid) : 0u));
            ring_drain(grid_obs, em, false, true);
        } else {
            // Half of the warps stream the observations first and run the physics afterwards, the other
            // half the other way round (scheduling option). Observation-first warps stage the PRE-step
            // grid: the (at most one) cell this step edits is patched into obs['grid'] by its own lane
            // afterwards and only envs whose block count changed are evaluated again -- for all others the
            // pre-step grid IS the post-step grid.
            const bool b_first = early && alternate && (wid & 1);
#pragma unroll 1
            for (int stage = 0; stage < 3; stage++) {  // a loop so that physics / ring code exist once
                if (stage == 1) {
                    run_physics();
                    if (((b_first && grid_obs) || inc_obs) && edit_cell >= 0) p.obs.grid[i * CELLS + edit_cell] = edit_val;
                    continue;
                }
                const bool changed = dsize != dsize_old;  // task.py:111-112 gate (wrong_placement != 0)
                unsigned mask, em;
                bool start, drain = true, obs_on, patch, keep;
                if (stage == 0) {
                    if (b_first) {  // pre-step grids: observations + (every_step) evaluations, result unused
                        em = every_step ? __ballot_sync(0xffffffffu, valid && nblk > 0 && dsize != 0) : 0u;
                        mask = em | (grid_obs ? __ballot_sync(0xffffffffu, valid) : 0u);
                        start = true; obs_on = grid_obs; patch = false; keep = false;
                    } else {        // only start the prefetch; consumed after the physics
                        em = 0u;
                        mask = __ballot_sync(0xffffffffu, valid && (grid_obs || nblk > 0));
                        start = early; drain = false; obs_on = false; patch = false; keep = false;
                    }
                } else {
                    if (b_first) {  // re-evaluate the envs whose block count changed
                        em = mask = __ballot_sync(0xffffffffu, valid && nblk > 0 && changed && dsize != 0);
                        start = true; obs_on = false; patch = true; keep = true;
                    } else {
                        em = mask = __ballot_sync(0xffffffffu, valid && nblk > 0 && dsize != 0 && (every_step || changed));
                        start = !early; obs_on = grid_obs; patch = true; keep = true;
                    }
                }
                if (start) ring_start(mask);
                if (drain) ring_drain(obs_on, em, patch, keep);
            }
        }

        // ------------------------------------------------------------- phase C
        if (valid) {
            if (HAS_A) {
                p.st.istate[0 * n + i] = a.inv01;
                p.st.istate[1 * n + i] = a.inv23;
                p.st.istate[2 * n + i] = a.inv45;
                p.st.istate[4 * n + i] = step_no;
                p.st.istate[5 * n + i] = dsize;
            }
            const int wrong = HAS_A ? dsize_old - dsize : w[7];                // task.py:111
            if (PHASES == 1 && p.overlap && !(do_reset || edit_cell >= 0)) {
                // overlapped step, nothing edited and no reset: the cached max_int stands (task.py:112, wrong == 0)
                // and the observation row k_widen writes from the unchanged grid is final -- finish the env here
                const bool done = max_int == target_size || step_no == p.cfg.max_steps;   // task.py:113, env.py:292
                const double reward = (double)wrong * p.cfg.wrong_placement_scale;      // right == 0 (env.py:293-296)
                p.obs.reward[i] = reward;
                p.obs.done[i] = (uint8_t)done;
                if (p.reward_host) {
                    p.reward_host[i] = reward;
                    p.done_host[i] = (uint8_t)done;
                }
                p.st.istate[3 * n + i] = (a.active & 0xFF) | ((a.tis & 0xFF) << 8) | ((int)done << 16) | ((a.err & 0xFF) << 24);
                p.st.istate[6 * n + i] = max_int;
                p.st.istate[7 * n + i] = 0;
            } else if (PHASES == 1) {  // bookkeeping is finished by the observe launch / by k_commit
                p.st.istate[3 * n + i] = (a.active & 0xFF) | ((a.tis & 0xFF) << 8) | ((a.err & 0xFF) << 24);
                p.st.istate[6 * n + i] = max_int;
                p.st.istate[7 * n + i] = p.overlap ? pack_w7(wrong, edit_cell, edit_val, do_reset) : wrong;
            } else {
                const int mi = wrong != 0 ? mi_eval : max_int;                 // task.py:112
                bool done = mi == target_size;                                 // task.py:113
                const int right = mi - max_int;                                // task.py:115
                done = done || (step_no == p.cfg.max_steps);                   // env.py:292
                const double reward = right == 0 ? (double)wrong * p.cfg.wrong_placement_scale
                                                 : (double)right * p.cfg.right_placement_scale;
                p.obs.reward[i] = reward;
                p.obs.done[i] = (uint8_t)done;
                if (p.reward_host) {  // host-buffer step: results also go straight to mapped host memory over PCIe
                    p.reward_host[i] = reward;
                    p.done_host[i] = (uint8_t)done;
                }
                p.st.istate[3 * n + i] = (a.active & 0xFF) | ((a.tis & 0xFF) << 8) | ((int)done << 16) | ((a.err & 0xFF) << 24);
                p.st.istate[6 * n + i] = mi;
            }
        }
        if (PHASES == 1 && p.overlap) {   // envs k_commit has to visit: edited or reset this step
            const bool special = valid && (do_reset || edit_cell >= 0);
            const unsigned sm = __ballot_sync(0xffffffffu, special);
            if (sm) {
                int b = 0;
                if (lane == 0) b = atomicAdd(p.st.sched + 4, __popc(sm));
                b = __shfl_sync(0xffffffffu, b, 0);
                if (special) {   // one 32-byte record per listed env: everything k_commit needs except the grid
                    int4 *rec = reinterpret_cast<int4 *>(p.st.worklist) + 2 * (int64_t)(b + __popc(sm & ((1u << lane) - 1u)));
                    rec[0] = make_int4((int)i, (a.active & 0xFF) | ((a.tis & 0xFF) << 8) | ((a.err & 0xFF) << 24), step_no, dsize);
                    rec[1] = make_int4(max_int, pack_w7(dsize_old - dsize, edit_cell, edit_val, do_reset),
                                       nblk | ((int)has_start << 16), target_size);
                }
            }
        }
        grp = dynamic ? (int64_t)__shfl_sync(0xffffffffu, grp_next, 0) : grp + (int64_t)gridDim.x * WPC;
    }
    // how many maximal_intersection evaluations really ran (bench.py's evaluated_frac): one fire-and-forget
    // 64-bit add per warp and launch
    if (HAS_B && p.st.sched != nullptr && lane == 0 && evals_done)
        atomicAdd(reinterpret_cast<unsigned long long *>(p.st.sched + 2), (unsigned long long)evals_done);
    if (PHASES == 1 && p.overlap) {
        // Overlapped step: this launch follows k_widen on the same stream as its programmatic dependent and was allowed to
        // start as soon as k_widen's CTAs were resident. k_commit, launched normally behind this grid, must not start
        // before k_widen has finished as well: the LAST CTA of this grid to get here (ticket in st.sched[6]) waits for the
        // prerequisite grid before it exits, which holds this grid's completion back until then. (Only that one CTA:
        // CTAs of the first wave that waited here would keep the second wave from ever starting.)
        __syncthreads();
        if (threadIdx.x == 0) {
            const int t = atomicAdd(p.st.sched + 6, 1);
            if (t == (int)gridDim.x - 1) {
                p.st.sched[6] = 0;
                asm volatile("griddepcontrol.wait;" ::: "memory");
            }
        }
    }
}

// ------------------------------------------------------------------ observe launch (tiled)
// Phases B + C of a step for the configurations in which (almost) every env has ring work -- every-step reward or a
// fully rewritten obs['grid'] -- after a physics launch (k_step<KIND, 1>). Persistent CTAs take TILES of up to
// OBS_TILE consecutive envs:
//   1. one thread per env of the tile loads the env's bookkeeping words (full lanes, once per env), decides what the
//      env needs and the envs with ring work are compacted into a work list in shared memory;
//   2. the warps claim work-list entries one at a time through a shared-memory counter (no global atomic, balanced to
//      within one env) and run the TMA ring: stage the grid, widen it into obs['grid'], evaluate maximal_intersection;
//   3. one thread per env finishes Task.step_intersection / reward / done (task.py:103-119, env.py:290-296).
// Env -> CTA mapping: the batch is cut into chunks of OBS_CHUNK consecutive envs and chunk c belongs to CTA c % grid;
// a tile is the next K chunks of the CTA. All CTAs therefore work on the same narrow band of the observation buffer
// at the same time (296 independent write streams spread over the whole buffer measured 5 % slower: DRAM page locality).
constexpr int OBS_TILE = 256;
constexpr int OBS_CHUNK = 8;
constexpr int OBS_TAIL_GROUP = 2;
struct __align__(16) ObsSlot {
    uint32_t grid[GSTRIDE / 4];        // int8 grid of one env (1089 cells + zero pad)
    int4 rects;                        // rectangle header of the env's block list (evaluated envs only)
};
struct __align__(16) ObsWarp {
    ObsSlot slot[IGLU_OBS_RING];
    EvalScratch es;
    unsigned long long bar[IGLU_OBS_RING];  // mbarriers of the ring
};
struct __align__(16) ObsCta {
    uint16_t info[OBS_TILE];     // per env of the tile: nblk | evaluate << 11 | has_start << 12
    uint16_t result[OBS_TILE];   // maximal_intersection of the envs that were evaluated
    uint8_t work[OBS_TILE];      // local indices of the envs with ring work
    int work_n[2], work_next[2]; // by tile parity: entries in the work list, next entry to claim
};
constexpr size_t observe_smem_bytes() {
    return TABLE_BYTES + sizeof(ObsCta) + IGLU_OBS_WARPS * sizeof(ObsWarp);
}
static_assert(IGLU_OBS_CTAS * (observe_smem_bytes() + 1024) <= 233472, "observe launch: shared memory of the resident CTAs");
static_assert(IGLU_OBS_WARPS * 32 >= OBS_TILE, "observe launch: one thread per env of a tile");

__global__ void __launch_bounds__(IGLU_OBS_WARPS * 32, IGLU_OBS_CTAS) k_observe(const __grid_constant__ StepParams p) {
    constexpr int WPC = IGLU_OBS_WARPS, NBUF = IGLU_OBS_RING;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint16_t *lut = reinterpret_cast<uint16_t *>(smem_raw);
    ObsCta &cta = *reinterpret_cast<ObsCta *>(smem_raw + TABLE_BYTES);
    ObsWarp *scratch = reinterpret_cast<ObsWarp *>(smem_raw + TABLE_BYTES + sizeof(ObsCta));
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    ObsWarp &ws = scratch[wid];
    for (int q = tid; q < CELL_LUT_N / 2; q += WPC * 32)   // the table was built once by iglu_ctx_create
        reinterpret_cast<uint32_t *>(lut)[q] = reinterpret_cast<const uint32_t *>(g_cell_lut)[q];
    for (int q = lane; q < NKEY; q += 32) ws.es.kmask[q] = 0u;  // invariant of eval_votes
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NBUF; s++) mbar_init(&ws.bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid == 0) { cta.work_n[0] = cta.work_n[1] = 0; cta.work_next[0] = cta.work_next[1] = 0; }
    __syncthreads();
    // programmatic dependent launch: everything above overlaps the tail of the physics launch
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const int64_t n = p.st.n;
    const int T = p.group;                       // envs per tile (<= OBS_TILE, a multiple of OBS_CHUNK)
    const bool every_step = p.cfg.reward_mode == IGLU_REWARD_EVERY_STEP;
    const bool grid_obs = p.cfg.write_grid_obs == 1 && p.obs.grid != nullptr;
    const int cap = p.cfg.tlist_cap;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t seq_issue = 0, seq_cons = 0;        // ring sequence numbers (warp-uniform, live across tiles)
    int evals_done = 0;
    uint32_t slot_addr[NBUF], bar_addr[NBUF];
#pragma unroll
    for (int s = 0; s < NBUF; s++) { slot_addr[s] = smem_u32(&ws.slot[s]); bar_addr[s] = smem_u32(&ws.bar[s]); }

    // ring producer (lane 0): bulk copies of env ej's grid row and, for an env that is evaluated, of its task's
    // rectangle header into the next slot
    auto issue_env = [&](int64_t ej, bool evj) {
        if (lane == 0) {
            uint32_t sa = slot_addr[0], ba = bar_addr[0];
#pragma unroll
            for (int q = 1; q < NBUF; q++) if ((int)(seq_issue % NBUF) == q) { sa = slot_addr[q]; ba = bar_addr[q]; }
            const uint32_t bytes = GSTRIDE + (evj ? 16u : 0u);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sa), "l"(p.st.grid + ej * GSTRIDE), "r"((uint32_t)GSTRIDE), "r"(ba) : "memory");
            if (evj)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(sa + (uint32_t)GSTRIDE), "l"(p.st.tlist + ej * cap), "r"(16u), "r"(ba) : "memory");
        }
        seq_issue++;
    };
    // ring consumer: wait for the oldest slot, widen it into obs['grid'], evaluate; returns maximal_intersection (or 0)
    auto consume_env = [&](int64_t ej, bool evj, bool hs, int nb) -> int {
        ObsSlot &sl = ws.slot[seq_cons % NBUF];
        mbar_wait(&ws.bar[seq_cons % NBUF], (seq_cons / NBUF) & 1u);
        if (grid_obs) write_grid_obs(p.obs.grid, ej * CELLS, sl.grid, lane);
        int r = 0;
        if (evj) {
            const int4 rc = sl.rects;   // the task's rectangles (16-byte header of its block list), came with the grid
            const uint16_t *tl = p.st.tlist + ej * cap + TL_HDR;
            if (single_shift_rects(rc, 4)) {  // zone-spanning target: one shift per rotation, compare bytes
                auto each = [&](auto f) {
                    for (int b = lane; b < nb; b += 32) {
                        const uint32_t e = tl[b];
                        f(e & 15, (e >> 4) & 15, (e >> 8) & 15, e >> 12);
                    }
                };
                r = eval_single_shift<false>(reinterpret_cast<const int8_t *>(sl.grid),
                                             hs ? p.st.init + ej * GSTRIDE : nullptr, each, rc, 4, nullptr);
            } else {
                r = eval_max_intersection<false>(ws.es, sl.grid, hs ? p.st.init + ej * GSTRIDE : nullptr, lut, tl, nb, rc, 4,
                                                 lane, nullptr);
            }
            evals_done++;
        }
        __syncwarp();  // every lane is done with the slot before it is refilled
        seq_cons++;
        return r;
    };
    // Task.step_intersection / reward / done of env i (task.py:103-119, env.py:290-296), one thread
    auto finish_env = [&](int64_t i, int w3, int step_no, int max_int, int wrong, int target_size, int mi_eval) {
        const int mi = wrong != 0 ? mi_eval : max_int;                           // task.py:112
        bool done = mi == target_size;                                          // task.py:113
        const int right = mi - max_int;                                         // task.py:115
        done = done || (step_no == p.cfg.max_steps);                            // env.py:292
        const double reward = right == 0 ? (double)wrong * p.cfg.wrong_placement_scale
                                         : (double)right * p.cfg.right_placement_scale;
        p.obs.reward[i] = reward;
        p.obs.done[i] = (uint8_t)done;
        if (p.reward_host) {  // host-buffer step: results also go straight to mapped host memory over PCIe
            p.reward_host[i] = reward;
            p.done_host[i] = (uint8_t)done;
        }
        p.st.istate[3 * n + i] = (w3 & 0xFF00FFFF) | ((int)done << 16);
        p.st.istate[6 * n + i] = mi;
    };

    const int K = T / OBS_CHUNK;                 // chunks per tile
    const uint32_t G32 = gridDim.x, B32 = blockIdx.x;
    auto env_of = [&](uint32_t tile, int j) -> int64_t {   // global env of local index j of this CTA's tile (32-bit maths)
        return (int64_t)((((tile * (uint32_t)K + ((uint32_t)j / OBS_CHUNK)) * G32 + B32) * OBS_CHUNK) + ((uint32_t)j % OBS_CHUNK));
    };
    int par = 0;
    for (uint32_t tile = 0; tile < (uint32_t)p.tiles; tile++, par ^= 1) {
        // ------------------------------------------------------------- 1: per-env words, work list
        const int64_t i = env_of(tile, tid);
        const bool valid = tid < T && i < n;
        if (tid == 0) {
            cta.work_n[par ^ 1] = 0;             // nobody uses the other parity's counters during this tile
            cta.work_next[par ^ 1] = 0;
        }
        int w3 = 0, step_no = 0, dsize = 0, max_int = 0, wrong = 0, nblk = 0, target_size = 0, has_start = 0;
        if (valid) {
            w3 = p.st.istate[3 * n + i]; step_no = p.st.istate[4 * n + i]; dsize = p.st.istate[5 * n + i];
            max_int = p.st.istate[6 * n + i]; wrong = p.st.istate[7 * n + i];
            nblk = p.st.tmeta[0 * n + i]; target_size = p.st.tmeta[1 * n + i]; has_start = p.st.tmeta[5 * n + i];
        }
        // an all-zero difference grid (dsize == 0) scores 0 without looking: no target cell can equal 0
        const bool ev = valid && nblk > 0 && dsize != 0 && (every_step || wrong != 0);
        const bool need = ev || (grid_obs && valid);
        if (tid < OBS_TILE) cta.info[tid] = (uint16_t)((nblk & 0x7FF) | ((int)ev << 11) | ((has_start != 0) << 12));
        {
            const unsigned m = __ballot_sync(0xffffffffu, need);
            if (m) {
                int b = 0;
                if (lane == 0) b = atomicAdd(&cta.work_n[par], __popc(m));
                b = __shfl_sync(0xffffffffu, b, 0);
                if (need) cta.work[b + __popc(m & lt)] = (uint8_t)tid;
            }
        }
        __syncthreads();
        // ------------------------------------------------------------- 2: the ring over claimed envs
        const int work_n = cta.work_n[par];
        auto claim = [&]() -> int {              // next work-list entry for this warp (local env index), or -1
            int k = 0;
            if (lane == 0) k = atomicAdd(&cta.work_next[par], 1);
            k = __shfl_sync(0xffffffffu, k, 0);
            return k < work_n ? (int)cta.work[k] : -1;
        };
        uint32_t ring_ids = 0;                   // local env indices in flight, 8 bits per ring slot
        auto issue = [&](int j) {
            const uint32_t s = seq_issue % NBUF;
            ring_ids = (ring_ids & ~(0xFFu << (8 * s))) | ((uint32_t)j << (8 * s));
            issue_env(env_of(tile, j), (cta.info[j] >> 11) & 1);
        };
        int inflight = 0;
        bool more = true;
#pragma unroll
        for (int s = 0; s < NBUF; s++) {
            if (more) {
                const int c = claim();
                if (c >= 0) { issue(c); inflight++; } else more = false;
            }
        }
        while (inflight) {
            const int c_next = more ? claim() : -1;   // claimed before the current env is processed: latency hidden
            if (c_next < 0) more = false;
            const int j = (int)((ring_ids >> (8 * (seq_cons % NBUF))) & 0xFFu);
            const int inf = cta.info[j];
            const bool evj = (inf >> 11) & 1;
            const int r = consume_env(env_of(tile, j), evj, (inf >> 12) & 1, inf & 0x7FF);
            if (evj && lane == 0) cta.result[j] = (uint16_t)r;
            inflight--;
            if (c_next >= 0) { issue(c_next); inflight++; }
        }
        __syncthreads();
        // ------------------------------------------------------------- 3: bookkeeping (phase C)
        if (valid) finish_env(i, w3, step_no, max_int, wrong, target_size, ev ? (int)cta.result[tid] : 0);
    }

    // ------------------------------------------------------------- dynamic tail
    // The envs from p.tail_begin on are not tiled: warps that are done with their CTA's tiles take them in groups of
    // OBS_TAIL_GROUP from the counter st.sched[0] (zeroed by the physics launch), which evens out the CTAs' different
    // amounts of evaluation work (static tiles alone: +5 % on the headline configuration).
    if (p.tail_begin < n) {
        int g = 0;
        if (lane == 0) g = atomicAdd(p.st.sched, 1);
        g = __shfl_sync(0xffffffffu, g, 0);
        while (true) {
            const int64_t e0 = p.tail_begin + (int64_t)g * OBS_TAIL_GROUP;
            if (e0 >= n) break;
            int g_next = 0;
            if (lane == 0) g_next = atomicAdd(p.st.sched, 1);   // fetched while this group is processed
            const int64_t i = e0 + lane;
            const bool valid = lane < OBS_TAIL_GROUP && i < n;
            int w3 = 0, step_no = 0, dsize = 0, max_int = 0, wrong = 0, nblk = 0, target_size = 0, has_start = 0;
            if (valid) {
                w3 = p.st.istate[3 * n + i]; step_no = p.st.istate[4 * n + i]; dsize = p.st.istate[5 * n + i];
                max_int = p.st.istate[6 * n + i]; wrong = p.st.istate[7 * n + i];
                nblk = p.st.tmeta[0 * n + i]; target_size = p.st.tmeta[1 * n + i]; has_start = p.st.tmeta[5 * n + i];
            }
            const bool ev = valid && nblk > 0 && dsize != 0 && (every_step || wrong != 0);
            const unsigned em = __ballot_sync(0xffffffffu, ev);
            const unsigned nm = em | (grid_obs ? __ballot_sync(0xffffffffu, valid) : 0u);
            const unsigned hm = __ballot_sync(0xffffffffu, has_start != 0);
            unsigned pend = nm, todo = nm;
            for (int f = 0; f < NBUF && pend; f++) {
                const int j = __ffs(pend) - 1;
                pend &= pend - 1;
                issue_env(e0 + j, (em >> j) & 1u);
            }
            int mi_eval = 0;
            while (todo) {
                const int j = __ffs(todo) - 1;
                todo &= todo - 1;
                const int r = consume_env(e0 + j, (em >> j) & 1u, (hm >> j) & 1u, __shfl_sync(0xffffffffu, nblk, j));
                if (lane == j) mi_eval = r;
                if (pend) {
                    const int jn = __ffs(pend) - 1;
                    pend &= pend - 1;
                    issue_env(e0 + jn, (em >> jn) & 1u);
                }
            }
            if (valid) finish_env(i, w3, step_no, max_int, wrong, target_size, mi_eval);
            g = __shfl_sync(0xffffffffu, g_next, 0);
        }
    }
    if (p.st.sched != nullptr && lane == 0 && evals_done)
        atomicAdd(reinterpret_cast<unsigned long long *>(p.st.sched + 2), (unsigned long long)evals_done);
}

// ------------------------------------------------------------------ overlapped step: k_widen || physics, then k_commit
// A step's physics launch is a latency chain that moves almost no bytes, and the observation rewrite is a stream of
// bytes that needs few issue slots; they only depend on each other through the (at most one) cell an env edits and
// through resets. The overlapped step therefore runs
//   k_widen   obs['grid'] rows and (every-step reward) maximal_intersection from the grids AS THEY
//             ARE WHEN THE STEP STARTS, one 16-warp CTA per SM at 64 registers, next to
//   physics   (k_step<KIND, 1>, same stream, launched as k_widen's PROGRAMMATIC DEPENDENT: it starts as soon as every
//             k_widen CTA is resident and never waits for that grid's results): two 4-warp CTAs per SM in the registers
//             k_widen leaves; it finishes reward / done itself for every env it neither edited nor reset (the cached
//             max_int stands, task.py:112) and lists the others in st.worklist; its last CTA waits for k_widen's
//             completion before exiting, so that
//   k_commit  (same stream, normal launch) runs after both: for the listed envs only -- patch the edited cell into
//             obs['grid'] or rewrite the row after a reset, evaluate maximal_intersection on the new grid, reward / done.
// k_widen reads state the physics launch may be rewriting (an edited byte, a reset env's rows): what it writes for
// such an env is replaced by k_commit, and the evaluation tolerates arbitrary bytes (register_target / hist_max
// check their inputs).
#ifndef IGLU_WIDEN_RING
#define IGLU_WIDEN_RING 4
#endif
struct __align__(16) WidenWarp {
    ObsSlot slot[IGLU_WIDEN_RING];
    EvalScratch es;
    unsigned long long bar[IGLU_WIDEN_RING];
};
#ifndef IGLU_WIDEN_WARPS
#define IGLU_WIDEN_WARPS 20          // warps of the widen launch's one CTA per SM when it evaluates (every-step reward)
#endif
#ifndef IGLU_WIDEN_WARPS_NOEVAL
#define IGLU_WIDEN_WARPS_NOEVAL 16   // ... and when it does not (reward on change)
#endif
constexpr int widen_warps(bool eval) { return eval ? IGLU_WIDEN_WARPS : IGLU_WIDEN_WARPS_NOEVAL; }
constexpr size_t widen_smem_bytes(bool eval) { return TABLE_BYTES + widen_warps(eval) * sizeof(WidenWarp); }

// Shared by k_widen / k_commit: ring producer and consumer over a warp's slots (see k_observe).
template <int NBUF, class Warp, bool SAFE>
struct Ring {
    const StepParams &p;
    Warp &ws;
    const uint16_t *lut;
    int lane;
    uint32_t seq_issue = 0, seq_cons = 0;
    int evals = 0;
    __device__ __forceinline__ Ring(const StepParams &p_, Warp &ws_, const uint16_t *lut_, int lane_)
        : p(p_), ws(ws_), lut(lut_), lane(lane_) {}
    __device__ __forceinline__ void issue(int64_t ej, bool evj) {
        if (lane == 0) {
            const uint32_t sa = smem_u32(&ws.slot[seq_issue % NBUF]), ba = smem_u32(&ws.bar[seq_issue % NBUF]);
            const uint32_t bytes = GSTRIDE + (evj ? 16u : 0u);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sa), "l"(p.st.grid + ej * GSTRIDE), "r"((uint32_t)GSTRIDE), "r"(ba) : "memory");
            if (evj)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(sa + (uint32_t)GSTRIDE), "l"(p.st.tlist + ej * p.cfg.tlist_cap), "r"(16u), "r"(ba) : "memory");
        }
        seq_issue++;
    }
    __device__ __forceinline__ ObsSlot &wait() {
        mbar_wait(&ws.bar[seq_cons % NBUF], (seq_cons / NBUF) & 1u);
        return ws.slot[seq_cons % NBUF];
    }
    __device__ __forceinline__ int evaluate(ObsSlot &sl, int64_t ej, bool hs, int nb) {
        const int4 rc = sl.rects;
        const uint16_t *tl = p.st.tlist + ej * p.cfg.tlist_cap + TL_HDR;
        evals++;
        if (single_shift_rects(rc, 4)) {
            auto each = [&](auto f) {
                for (int b = lane; b < nb; b += 32) {
                    const uint32_t e = tl[b];
                    f(e & 15, (e >> 4) & 15, (e >> 8) & 15, e >> 12);
                }
            };
            return eval_single_shift<false, SAFE>(reinterpret_cast<const int8_t *>(sl.grid),
                                                  hs ? p.st.init + ej * GSTRIDE : nullptr, each, rc, 4, nullptr);
        }
        return eval_max_intersection<false, SAFE>(ws.es, sl.grid, hs ? p.st.init + ej * GSTRIDE : nullptr, lut, tl, nb, rc, 4,
                                                  lane, nullptr);
    }
    __device__ __forceinline__ void release() {
        __syncwarp();  // every lane is done with the slot before it is refilled
        seq_cons++;
    }
};

#ifndef IGLU_WIDEN_REGS
#define IGLU_WIDEN_REGS 48     // with evaluations: 20 warps x 48 registers (no spills) + five 2-warp physics CTAs
#endif
#ifndef IGLU_WIDEN_REGS_NOEVAL
#define IGLU_WIDEN_REGS_NOEVAL 32   // reward on change: no evaluation code in the widen launch (no spills at 32), so the
                                    // whole physics grid fits next to it in one wave: 0.078 -> 0.074 ms
#endif
// EVAL = false is the instantiation for reward_mode on_change (the widen launch never evaluates there)
template <bool EVAL>
__global__ void __maxnreg__(EVAL ? IGLU_WIDEN_REGS : IGLU_WIDEN_REGS_NOEVAL) k_widen(const __grid_constant__ StepParams p) {
    constexpr int NBUF = IGLU_WIDEN_RING;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint16_t *lut = reinterpret_cast<uint16_t *>(smem_raw);
    WidenWarp *scratch = reinterpret_cast<WidenWarp *>(smem_raw + TABLE_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    WidenWarp &ws = scratch[wid];
    constexpr int WIDEN_WARPS = EVAL ? IGLU_WIDEN_WARPS : IGLU_WIDEN_WARPS_NOEVAL;
    for (int q = tid; q < CELL_LUT_N / 2; q += WIDEN_WARPS * 32)
        reinterpret_cast<uint32_t *>(lut)[q] = reinterpret_cast<const uint32_t *>(g_cell_lut)[q];
    for (int q = lane; q < NKEY; q += 32) ws.es.kmask[q] = 0u;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NBUF; s++) mbar_init(&ws.bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // This launch is itself a programmatic dependent of whatever precedes it on the stream (the previous step's k_commit
    // triggers early; any other kernel triggers when it completes): the set-up above overlaps that kernel's tail.
    // Only after the wait may the physics launch -- this grid's own dependent -- be allowed to start.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int64_t n = p.st.n;
    const int GS = p.group;
    const int64_t ngroups = (n + GS - 1) / GS;
    const bool every_step = EVAL && p.cfg.reward_mode == IGLU_REWARD_EVERY_STEP;
    const bool wr = p.cfg.write_grid_obs == 1 && p.obs.grid != nullptr;   // false: evaluations only (in-place observations)
    Ring<NBUF, WidenWarp, true> ring(p, ws, lut, lane);
    const int total_warps = (int)gridDim.x * WIDEN_WARPS;
    // every warp's first group is its static index; the counter st.sched[0] (left at zero by k_commit) deals the rest
    int64_t grp = (int64_t)wid * gridDim.x + blockIdx.x;
    while (grp < ngroups) {
        int grp_next = 0;
        if (lane == 0) grp_next = total_warps + atomicAdd(p.st.sched, 1);
        const int64_t e0 = grp * GS;
        const int64_t i = e0 + lane;
        const bool valid = lane < GS && i < n;
        int w3 = 0, dsize = 0, nblk = 0, has_start = 0;
        if (valid && every_step) {   // may be mid-rewrite by the physics launch: see the note above
            w3 = p.st.istate[3 * n + i]; dsize = p.st.istate[5 * n + i];
            nblk = p.st.tmeta[0 * n + i]; has_start = p.st.tmeta[5 * n + i];
        }
        // evaluated here: envs that hold blocks and are not reset by this step (done_prev with auto-reset)
        const bool ev = valid && every_step && nblk > 0 && nblk <= p.cfg.tlist_cap - TL_HDR && dsize != 0 &&
                        !(p.cfg.auto_reset && ((w3 >> 16) & 0xFF));
        const unsigned em = __ballot_sync(0xffffffffu, ev);
        const unsigned nm = wr ? __ballot_sync(0xffffffffu, valid) : em;   // envs staged: all when obs['grid'] is rewritten
        const unsigned hm = __ballot_sync(0xffffffffu, has_start != 0);
        unsigned pend = nm, todo = nm;
        for (int f = 0; f < NBUF && pend; f++) {
            const int j = __ffs(pend) - 1;
            pend &= pend - 1;
            ring.issue(e0 + j, (em >> j) & 1u);
        }
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t ej = e0 + j;
            ObsSlot &sl = ring.wait();
            if (wr) write_grid_obs(p.obs.grid, ej * CELLS, sl.grid, lane);
            if (EVAL && ((em >> j) & 1u)) ring.evaluate(sl, ej, (hm >> j) & 1u, __shfl_sync(0xffffffffu, nblk, j));
            ring.release();
            if (pend) {
                const int jn = __ffs(pend) - 1;
                pend &= pend - 1;
                ring.issue(e0 + jn, (em >> jn) & 1u);
            }
        }
        grp = (int64_t)__shfl_sync(0xffffffffu, grp_next, 0);
    }
    if (lane == 0 && ring.evals)
        atomicAdd(reinterpret_cast<unsigned long long *>(p.st.sched + 2), (unsigned long long)ring.evals);
}

struct __align__(16) CommitWarp {
    ObsSlot slot[2];
    EvalScratch es;
    unsigned long long bar[2];
};
constexpr int COMMIT_WARPS = 4;    // small CTAs: they start and retire independently (the launch is a latency chain)
constexpr size_t commit_smem_bytes() { return TABLE_BYTES + COMMIT_WARPS * sizeof(CommitWarp); }

__global__ void __launch_bounds__(COMMIT_WARPS * 32, 7) k_commit(const __grid_constant__ StepParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint16_t *lut = reinterpret_cast<uint16_t *>(smem_raw);
    CommitWarp *scratch = reinterpret_cast<CommitWarp *>(smem_raw + TABLE_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    CommitWarp &ws = scratch[wid];
    for (int q = tid; q < CELL_LUT_N / 2; q += COMMIT_WARPS * 32)
        reinterpret_cast<uint32_t *>(lut)[q] = reinterpret_cast<const uint32_t *>(g_cell_lut)[q];
    for (int q = lane; q < NKEY; q += 32) ws.es.kmask[q] = 0u;
    if (lane == 0) {
        mbar_init(&ws.bar[0], 1);
        mbar_init(&ws.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // This launch is the programmatic dependent of the physics launch: its CTAs become resident (and run the set-up
    // above) while that grid drains; nothing it wrote -- and, through its last CTA, nothing k_widen wrote -- is touched
    // before this point.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next launch on the stream may set itself up
    const int count = p.st.sched[4];
    const int4 *recs = reinterpret_cast<const int4 *>(p.st.worklist);
    const int k0 = wid * (int)gridDim.x + (int)blockIdx.x, kstep = (int)gridDim.x * COMMIT_WARPS;
    int4 r0 = make_int4(0, 0, 0, 0), r1 = r0;
    if ((int64_t)k0 < p.st.n) { r0 = recs[2 * (int64_t)k0]; r1 = recs[2 * (int64_t)k0 + 1]; }   // valid iff k0 < count
    const int64_t n = p.st.n;
    const bool every_step = p.cfg.reward_mode == IGLU_REWARD_EVERY_STEP;
    const bool grid_full = p.cfg.write_grid_obs == 1 && p.obs.grid != nullptr;   // else obs['grid'] is not this launch's business
    Ring<2, CommitWarp, false> ring(p, ws, lut, lane);
    int counted = 0;
    if (k0 < count) ring.issue(r0.x, true);   // grid row + rectangle header of this warp's first env
    for (int k = k0; k < count; k += kstep) {
        // the next record of this warp and its bulk copies are asked for before the current env is processed
        int4 n0 = make_int4(0, 0, 0, 0), n1 = n0;
        if (k + kstep < count) {
            n0 = recs[2 * (int64_t)(k + kstep)]; n1 = recs[2 * (int64_t)(k + kstep) + 1];
            ring.issue(n0.x, true);
        }
        const int64_t e = r0.x;
        const int w3 = r0.y, step_no = r0.z, dsize = r0.w, max_int = r1.x, w7 = r1.y;
        const int nblk = r1.z & 0xFFFF, has_start = r1.z >> 16, target_size = r1.w;
        const int wrong = (w7 << 30) >> 30;              // sign-extended 2 bits
        const int edit_cell = ((w7 >> 2) & 0x7FF) - 1, edit_val = (w7 >> 13) & 15;
        const bool reset = (w7 >> 17) & 1;
        const bool ev = wrong != 0 && nblk > 0 && dsize != 0;   // task.py:111-112 (an all-zero difference grid scores 0)
        int mi_eval = 0;
        {
            ObsSlot &sl = ring.wait();
            if (reset && grid_full) write_grid_obs(p.obs.grid, e * CELLS, sl.grid, lane);   // the row k_widen wrote predates the reset
            if (ev) {
                mi_eval = ring.evaluate(sl, e, has_start != 0, nblk);
                // evaluations k_widen did not already count for this env (it skips empty grids and reset envs)
                if (!(every_step && dsize + wrong != 0 && !reset)) counted++;
            }
            ring.release();
        }
        if (lane == 0) {
            if (grid_full && !reset && edit_cell >= 0) p.obs.grid[e * CELLS + edit_cell] = edit_val;
            const int mi = wrong != 0 ? mi_eval : max_int;                           // task.py:112
            bool done = mi == target_size;                                          // task.py:113
            const int right = mi - max_int;                                         // task.py:115
            done = done || (step_no == p.cfg.max_steps);                            // env.py:292
            const double reward = right == 0 ? (double)wrong * p.cfg.wrong_placement_scale
                                             : (double)right * p.cfg.right_placement_scale;
            p.obs.reward[e] = reward;
            p.obs.done[e] = (uint8_t)done;
            if (p.reward_host) {
                p.reward_host[e] = reward;
                p.done_host[e] = (uint8_t)done;
            }
            p.st.istate[3 * n + e] = (w3 & 0xFF00FFFF) | ((int)done << 16);
            p.st.istate[6 * n + e] = mi;
            p.st.istate[7 * n + e] = wrong;
        }
        r0 = n0; r1 = n1;
    }
    // evaluations of this launch that k_widen did not count: only those (bench.py's evaluated_frac counts envs)
    if (lane == 0 && counted)
        atomicAdd(reinterpret_cast<unsigned long long *>(p.st.sched + 2), (unsigned long long)(unsigned)counted);
    ring.evals = 0;
    // the last CTA out leaves the launch-to-launch counters at zero for the next step
    __syncthreads();
    if (tid == 0) {   // (every CTA's ticket follows its own use of `count`: the last one may clear it)
        const int t = atomicAdd(p.st.sched + 5, 1);
        if (t == (int)gridDim.x - 1) { p.st.sched[0] = 0; p.st.sched[4] = 0; p.st.sched[5] = 0; }
    }
}

// ------------------------------------------------------------------ obs['pov'] (iglu_render.cuh)
// One CTA per env: the grid row is staged in shared memory, every thread traces height * width / 256 pixels.
__global__ void __launch_bounds__(256) k_render_pov(iglu_state st, uint8_t *__restrict__ pov, int height, int width,
                                                    const uint8_t *__restrict__ atlas) {
    __shared__ __align__(16) int8_t g8[GSTRIDE];
    __shared__ int box[6];   // x0, x1, y0, y1, z0, z1 of the blocks
    const int64_t e = blockIdx.x, n = st.n;
    const int4 *src = reinterpret_cast<const int4 *>(st.grid + e * GSTRIDE);
    for (int q = threadIdx.x; q < GSTRIDE / 16; q += blockDim.x) reinterpret_cast<int4 *>(g8)[q] = src[q];
    if (threadIdx.x < 6) box[threadIdx.x] = (threadIdx.x & 1) ? -1 : 16;
    const PovCamera cam = pov_camera(st.pose[0 * n + e], st.pose[1 * n + e], st.pose[2 * n + e], st.pose[3 * n + e],
                                     st.pose[4 * n + e]);
    __syncthreads();
    for (int c = threadIdx.x; c < CELLS; c += blockDim.x)
        if (g8[c] != 0) {
            const int y = c / 121, r = c - 121 * y, x = r / 11, z = r - 11 * x;
            atomicMin(&box[0], x); atomicMax(&box[1], x); atomicMin(&box[2], y); atomicMax(&box[3], y);
            atomicMin(&box[4], z); atomicMax(&box[5], z);
        }
    __syncthreads();
    const PovBox bb = {box[0], box[1], box[2], box[3], box[4], box[5]};
    const float kx = 2.0f / (float)width, ky = 2.0f / (float)height, aspect = (float)width / (float)height;
    uint8_t *out = pov + e * (int64_t)height * width * 3;
    // 256 pixels per pass: colours go through shared memory so that the frame is written with 4-byte stores of
    // consecutive lanes (the frame of env e starts at a multiple of 4 bytes when 3 * h * w is one; else byte stores)
    __shared__ __align__(16) uint8_t stage[256 * 3];
    const int npix = height * width;
    const bool wide = ((3 * npix) & 3) == 0 && ((reinterpret_cast<uintptr_t>(pov)) & 3) == 0;
    for (int p0 = 0; p0 < npix; p0 += 256) {
        const int pix = p0 + threadIdx.x;
        if (pix < npix) {
            const int j = pix / width, i = pix - j * width;
            const float sx = ((float)i + 0.5f) * kx - 1.0f, sy = 1.0f - ((float)j + 0.5f) * ky;
            const uchar3 c = pov_trace(g8, bb, atlas, cam, sx * aspect, sy);
            stage[3 * threadIdx.x + 0] = c.x; stage[3 * threadIdx.x + 1] = c.y; stage[3 * threadIdx.x + 2] = c.z;
        }
        __syncthreads();
        const int nbytes = 3 * min(256, npix - p0);
        if (wide && (nbytes & 3) == 0) {   // 3 * p0 is a multiple of 4 (p0 is a multiple of 256)
            if (threadIdx.x < nbytes / 4)
                reinterpret_cast<uint32_t *>(out + 3 * p0)[threadIdx.x] = reinterpret_cast<const uint32_t *>(stage)[threadIdx.x];
        } else {
            for (int b = threadIdx.x; b < nbytes; b += 256) out[3 * p0 + b] = stage[b];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ task preparation
// Task.__init__ of the synthetic task (env.py:227-232, task.py:9-72) + Tasks.to_dense'd starting grid.
// One warp per row.
__global__ void __launch_bounds__(256) k_task_prep(iglu_state st, iglu_obs obs, iglu_config cfg,
                                                   const int8_t *__restrict__ target,
                                                   const int8_t *__restrict__ start,
                                                   const int8_t *__restrict__ full,
                                                   const double *__restrict__ init_pose,
                                                   const int32_t *__restrict__ env_idx, int64_t count) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= count) return;
    const int64_t e = env_idx ? env_idx[row] : row;
    if (e < 0 || e >= st.n) return;
    const int64_t n = st.n;
    const int8_t *t = target + row * CELLS;
    const int8_t *s = start ? start + row * CELLS : nullptr;
    const int8_t *fg = full ? full + row * CELLS : nullptr;   // task.py:63-64: admissible shifts from full_grid
    int8_t *gt = st.target ? st.target + e * GSTRIDE : nullptr, *gi = st.init ? st.init + e * GSTRIDE : nullptr;
    uint16_t *tl = st.tlist + e * cfg.tlist_cap;
    const int room = cfg.tlist_cap - TL_HDR;
    int nout = 0, xmin = 10, xmax = 0, zmin = 10, zmax = 0, any_start = 0;
    int cnt[6] = {0, 0, 0, 0, 0, 0};
    for (int c0 = 0; c0 < GSTRIDE; c0 += 32) {
        const int c = c0 + lane;
        int tv = 0, sv = 0;
        if (c < CELLS) {
            tv = t[c];
            sv = s ? (int)s[c] : 0;
        }
        if (c < GSTRIDE && gt) { gt[c] = (int8_t)tv; gi[c] = (int8_t)sv; }
        if (c < CELLS && obs.target_grid) obs.target_grid[e * CELLS + c] = tv;
        const int d = tv - sv;  // env.py:230
        const unsigned m = __ballot_sync(0xffffffffu, d != 0);
        any_start |= __any_sync(0xffffffffu, sv != 0);
#pragma unroll
        for (int k = 0; k < 6; k++) cnt[k] += __popc(__ballot_sync(0xffffffffu, sv == k + 1));
        if (d != 0) {
            const int y = c / 121, r = c - 121 * y, x = r / 11, z = r - 11 * x;
            const int pos = nout + __popc(m & ((1u << lane) - 1u));
            if (pos < room) tl[TL_HDR + pos] = (uint16_t)(y | (x << 4) | (z << 8) | (code_of(d) << 12));
            if (!fg) { xmin = min(xmin, x); xmax = max(xmax, x); zmin = min(zmin, z); zmax = max(zmax, z); }
        }
        if (fg && c < CELLS && fg[c] != 0) {
            const int y = c / 121, r = c - 121 * y, x = r / 11, z = r - 11 * x;
            xmin = min(xmin, x); xmax = max(xmax, x); zmin = min(zmin, z); zmax = max(zmax, z);
        }
        nout += __popc(m);
    }
    xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
    zmin = __reduce_min_sync(0xffffffffu, zmin); zmax = __reduce_max_sync(0xffffffffu, zmax);
    if (lane == 0) {
        *reinterpret_cast<int4 *>(tl) = rects_from_bbox(xmin, xmax, zmin, zmax);
        st.tmeta[0 * n + e] = min(nout, room);
        st.tmeta[1 * n + e] = nout;  // target_size = nnz(target - init)
        st.tmeta[2 * n + e] = pack16(20 - cnt[0], 20 - cnt[1]);
        st.tmeta[3 * n + e] = pack16(20 - cnt[2], 20 - cnt[3]);
        st.tmeta[4 * n + e] = pack16(20 - cnt[4], 20 - cnt[5]);
        st.tmeta[5 * n + e] = any_start;
        st.tmeta[6 * n + e] = 1;
        st.tmeta[7 * n + e] = 0;
        if (st.init_pose)
            for (int k = 0; k < 5; k++) st.init_pose[k * n + e] = init_pose ? init_pose[row * 5 + k] : 0.0;
    }
}

// ------------------------------------------------------------------ random task structures on the device
// RandomTasks.sample_task (task_set.py:135-157), one thread per task, with the counter-based stream documented in
// include/iglu_b200.h (iglu_sample_tasks). Writes dense int8 [m][1089] targets: the input format of iglu_set_tasks /
// iglu_task_pool_prepare, so a task pool can be (re)filled without the host.
__device__ __forceinline__ uint32_t sample_draw(uint64_t seed, int64_t task, uint32_t &counter, uint32_t range) {
    uint64_t x = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(task + 1) + 0xD1B54A32D192ED03ull * (uint64_t)(counter++);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    // ((x >> 32) * range) >> 32 as a 32-bit multiply-high. (Written as a 64-bit product, ptxas 12.9 drops the whole
    // hash chain of block_x in the max_blocks == 1 loop -- every level then lands at x = 0; caught by the oracle.)
    return __umulhi((uint32_t)(x >> 32), range);
}

__global__ void __launch_bounds__(128) k_sample_tasks(int8_t *out, int64_t m, iglu_random_task_spec spec,
                                                      uint64_t seed, int64_t first_task) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    int8_t *g = out + t * CELLS;   // zeroed by the cudaMemsetAsync that iglu_sample_tasks enqueues before this kernel
    uint32_t ctr = 0;
    const uint32_t span = (uint32_t)(2 * spec.max_dist + 1);
    for (int h = 0; h < spec.height_levels && h < 9; h++) {
        const int bx = (int)sample_draw(seed, first_task + t, ctr, 11u);
        const int bz = (int)sample_draw(seed, first_task + t, ctr, 11u);
        int color = (int)sample_draw(seed, first_task + t, ctr, (uint32_t)spec.num_colors) + 1;
        g[h * 121 + bx * 11 + bz] = (int8_t)color;
        for (int b = 1; b < spec.max_blocks; b++) {
            int x = 0, z = 0;
            bool ok = false;
            for (int tries = 0; tries < 4096 && !ok; tries++) {   // the reference redraws without a bound
                const int ddx = (int)sample_draw(seed, first_task + t, ctr, span) - spec.max_dist;
                const int ddz = (int)sample_draw(seed, first_task + t, ctr, span) - spec.max_dist;
                color = (int)sample_draw(seed, first_task + t, ctr, (uint32_t)spec.num_colors) + 1;
                x = bx + ddx; z = bz + ddz;
                ok = !(ddx == 0 && ddz == 0) && x < 11 && z < 11 && x >= 0 && z >= 0 && g[h * 121 + x * 11 + z] == 0;
            }
            if (ok) g[h * 121 + x * 11 + z] = (int8_t)color;
        }
    }
}

// ------------------------------------------------------------------ reset
// GridWorld.reset (env.py:206-261). One warp per env.
__global__ void __launch_bounds__(256) k_reset(iglu_state st, iglu_obs obs, const uint8_t *__restrict__ mask,
                                               int hard, const __grid_constant__ iglu_task_pool pool, int cap) {
    const int lane = threadIdx.x & 31;
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n = st.n;
    if (e >= n) return;
    if (mask && !mask[e]) return;
    if (pool.m > 0) {  // task generator: draw the next task first (env.py:213-216)
        int j = 0;
        if (lane == 0) j = pool_next(pool, st.tmeta, n, e);
        j = __shfl_sync(0xffffffffu, j, 0);
        pool_fetch(pool, cap, j, pool.tmeta[j], st.target + e * GSTRIDE, st.init + e * GSTRIDE,
                   st.tlist + e * cap, obs.target_grid ? obs.target_grid + e * CELLS : nullptr, lane);
        __syncwarp();
    }
    const int4 *src = reinterpret_cast<const int4 *>(st.init + e * GSTRIDE);
    int4 *dst = reinterpret_cast<int4 *>(st.grid + e * GSTRIDE);
    for (int q = lane; q < GSTRIDE / 16; q += 32) dst[q] = src[q];
    if (obs.grid) {
        const int8_t *si = st.init + e * GSTRIDE;
        for (int c = lane; c < CELLS; c += 32) obs.grid[e * CELLS + c] = (int32_t)si[c];
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 5; k++) st.pose[k * n + e] = st.init_pose[k * n + e];
        int w3 = st.istate[3 * n + e];
        int active = w3 & 0xFF, tis = (w3 >> 8) & 0xFF;
        if (hard || active == 0) {  // fresh env: Agent.__init__ (world.py:24-29)
            active = 1; tis = 2;
            st.pose[5 * n + e] = 0.0;
        }
        st.istate[0 * n + e] = st.tmeta[2 * n + e];
        st.istate[1 * n + e] = st.tmeta[3 * n + e];
        st.istate[2 * n + e] = st.tmeta[4 * n + e];
        st.istate[3 * n + e] = active | (tis << 8);  // done_prev = 0, err = 0
        st.istate[4 * n + e] = 0;
        st.istate[5 * n + e] = 0;
        st.istate[6 * n + e] = 0;
        st.istate[7 * n + e] = 0;
        // reset observation: agentPos and compass are literal zeros (env.py:249, 254)
        if (obs.agent_pos) for (int k = 0; k < 5; k++) obs.agent_pos[5 * e + k] = 0.f;
        if (obs.compass) obs.compass[e] = 0.f;
        if (obs.inventory) {
            int t2 = st.tmeta[2 * n + e], t3 = st.tmeta[3 * n + e], t4 = st.tmeta[4 * n + e];
            float *iv = obs.inventory + 6 * e;
            iv[0] = (float)lo16(t2); iv[1] = (float)hi16(t2); iv[2] = (float)lo16(t3);
            iv[3] = (float)hi16(t3); iv[4] = (float)lo16(t4); iv[5] = (float)hi16(t4);
        }
        obs.reward[e] = 0.0;
        obs.done[e] = 0;
    }
}

// ------------------------------------------------------------------ multi-turn structures with progressive goals
// Subtasks.step_intersection (tasks/task.py:288-298) for every env, on the env's RAW grid: the score of the current
// turn's Task (target = the structure after that turn, admissible shifts from the FULL structure, task.py:63-64), and
// when that task is complete the goal moves on to the next turn (create_task + reset + a second step_intersection
// with prev_grid_size = 0). One warp per env; runs after the step kernels when staged tasks are attached.
struct __align__(16) StagedScratch {
    uint32_t grid[GSTRIDE / 4];
    uint32_t init[GSTRIDE / 4];
    EvalScratch es;
};
constexpr int STAGED_WARPS = 4;

__global__ void __launch_bounds__(STAGED_WARPS * 32) k_staged_step(iglu_state st, iglu_staged_tasks sg,
                                                                  const int32_t *__restrict__ task_row, int cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint16_t *lut = reinterpret_cast<uint16_t *>(smem_raw);
    StagedScratch *scratch = reinterpret_cast<StagedScratch *>(smem_raw + TABLE_BYTES);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    StagedScratch &ws = scratch[wid];
    build_cell_lut(lut, threadIdx.x, STAGED_WARPS * 32);
    for (int q = lane; q < NKEY; q += 32) ws.es.kmask[q] = 0u;
    __syncthreads();
    const int64_t n = st.n;
    const int64_t e = (int64_t)blockIdx.x * STAGED_WARPS + wid;
    if (e >= n) return;
    {   // stage the env's grid and starting grid
        const int4 *g4 = reinterpret_cast<const int4 *>(st.grid + e * GSTRIDE);
        const int4 *i4 = reinterpret_cast<const int4 *>(st.init + e * GSTRIDE);
        for (int q = lane; q < GSTRIDE / 16; q += 32) {
            reinterpret_cast<int4 *>(ws.grid)[q] = g4[q];
            reinterpret_cast<int4 *>(ws.init)[q] = i4[q];
        }
    }
    __syncwarp();
    const int size = count_blocks(ws.grid, 0, nullptr, lane);
    int goal = sg.state[0 * n + e], prev = sg.state[1 * n + e], max_int = sg.state[2 * n + e];
    auto eval = [&](const uint32_t *x, int row) {   // Task(stage `row`).maximal_intersection(x)
        const uint16_t *tl = sg.tlist + (int64_t)row * cap;
        const int4 rc = *reinterpret_cast<const int4 *>(tl);
        return eval_max_intersection<false>(ws.es, x, nullptr, lut, tl + TL_HDR, sg.tmeta[row], rc, 4, lane, nullptr);
    };
    auto create_task = [&](int row, int nstart) {   // create_task(...) + task.reset(): task.py:256-285, 74-86
        max_int = nstart ? eval(ws.init, row) : 0;
        prev = nstart;
    };
    const bool fresh = st.istate[4 * n + e] == 1;   // first step of an episode: Subtasks.reset() happened in env.reset()
    const int nstart = (fresh || sg.progressive) ? count_blocks(ws.init, 0, nullptr, lane) : 0;
    if (fresh) {
        const int64_t r = task_row ? task_row[e] : e;
        goal = sg.stage_of ? sg.stage_of[r] : (int)r;
        create_task(goal, nstart);
    }
    // task.py:103-119 on the raw grid
    const int wrong = prev - size;
    int mi = wrong != 0 ? eval(ws.grid, goal) : max_int;
    bool done = mi == sg.tmeta[sg.m + goal];
    const int right = mi - max_int;
    max_int = mi;
    prev = size;
    const int nxt = sg.next_stage[goal];
    if (done && sg.progressive && nxt >= 0) {   // task.py:290-297
        goal = nxt;
        create_task(goal, nstart);
        mi = size != 0 ? eval(ws.grid, goal) : max_int;   // prev_grid_size = 0: wrong_placement = -size
        done = mi == sg.tmeta[sg.m + goal];
        max_int = mi;
        prev = size;
    }
    if (lane == 0) {
        sg.state[0 * n + e] = goal; sg.state[1 * n + e] = prev; sg.state[2 * n + e] = max_int;
        sg.out[0 * n + e] = right; sg.out[1 * n + e] = wrong; sg.out[2 * n + e] = (int)done; sg.out[3 * n + e] = goal;
    }
}

// ------------------------------------------------------------------ stateless reward sweep
// Task(target, full_grid, invariant).maximal_intersection(grid) per pair (BASELINE config 4). One warp per pair.
// The rows of the [count][stride] arrays are brought into shared memory by TMA bulk copies of the 16-byte aligned
// window around each row (a granule that holds a valid byte lies in the same page: nothing outside the caller's
// pages is touched), double buffered per warp so that pair i + 1 arrives while pair i is evaluated. The consumers
// take the byte offset of cell 0 inside the window (iglu_reward.cuh, compaction), so neither a realignment pass nor
// funnel shifts are needed; rows in the env state layout (stride 1104, 16-byte aligned) have offset 0.
struct __align__(16) PairScratch {
    uint32_t grid[2][GSTRIDE / 4];
    uint32_t tgt[2][GSTRIDE / 4];
    EvalScratch es;
    uint16_t tl[GL_CAP];             // target blocks of the current round: cell | code << 11
    uint16_t twl[WL_CAP];            // non-zero words of the target's current round
    unsigned long long bar[2];
};

// The target side of one pair for eval_votes: rounds of <= 32 non-zero words (<= 128 blocks) compacted on demand.
struct WindowTargets {
    PairScratch &ps;
    const uint32_t *T;
    const uint16_t *lut;
    int toff, nwords, lane;
    int twl_round, tl_round, ntl;
    __device__ __forceinline__ WindowTargets(PairScratch &ps_, const uint32_t *T_, int toff_, const uint16_t *lut_, int nwords_,
                                             int lane_)
        : ps(ps_), T(T_), lut(lut_), toff(toff_), nwords(nwords_), lane(lane_), twl_round(0), tl_round(-1), ntl(0) {}
    __device__ __forceinline__ bool empty() const { return nwords == 0; }
    __device__ __forceinline__ int rounds() const { return (nwords + WL_CAP - 1) / WL_CAP; }
    __device__ __forceinline__ int load(int r) {
        if (tl_round != r) {
            __syncwarp();   // the readers of ps.tl / ps.twl are done
            if (twl_round != r) {
                word_scan<false>(T, toff, nullptr, lane, WL_CAP * r, ps.twl, nullptr);
                twl_round = r;
            }
            __syncwarp();
            ntl = block_emit(T, toff, nullptr, lane, min(WL_CAP, nwords - WL_CAP * r), ps.twl, ps.tl);
            tl_round = r;
            __syncwarp();
        }
        return ntl;
    }
    __device__ __forceinline__ int reg(EvalScratch &es, int c0, int) const {
        if (c0 + lane >= ntl) return -1;
        const int e = ps.tl[c0 + lane], l = lut[e & 0x7FF], x = lut_x(l);
        return register_target(es, lane, l >> 8, x, (l & 0xFF) - HROW * x, e >> 11);
    }
};

// The dense comparison of a crowded pair (iglu_reward.cuh, packed columns). from_bbox: the admissible rectangles come
// from the target's occupied columns; returns -1 when that target does not span the zone (more than one admissible
// shift: the rectangles are written to *rc and the caller carries on with them). Otherwise the caller has checked that every rotation's
// single shift is (0, 0). Whole columns at a time when both rows hold colour differences only, else cell by cell.
template <bool WANT_KEY>
__device__ __forceinline__ int dense_try(const int8_t *t8, const int8_t *g8, bool from_bbox, int n_rot, int lane, uint2 *xch,
                                      int4 *rc, int *key_out) {
    PackedCols tp, gp;
    pack_columns(t8, lane, tp);
    if (from_bbox) {
        int x0, x1, z0, z1;
        packed_bbox(tp, lane, x0, x1, z0, z1);
        if (!(x0 == 0 && x1 == 10 && z0 == 0 && z1 == 10)) {   // room to shift under some rotation (task.py:67-72)
            *rc = rects_from_bbox(x0, x1, z0, z1);
            return -1;
        }
    }
    pack_columns(g8, lane, gp);
    if (__all_sync(0xffffffffu, tp.in_range() && gp.in_range()))
        return eval_zero_shift_packed<WANT_KEY>(tp, gp, n_rot, lane, xch, key_out);
    return eval_zero_shift_dense<WANT_KEY>(t8, g8, n_rot, lane, key_out);
}

template <bool WANT_KEY>
__global__ void __launch_bounds__(MI_WARPS * 32, MI_CTAS) k_max_intersection(const int8_t *__restrict__ grids,
                                                                     const int8_t *__restrict__ targets,
                                                                     const int8_t *__restrict__ fulls,
                                                                     int64_t gstride, int64_t tstride, int64_t fstride,
                                                                     int invariant, int32_t *__restrict__ out,
                                                                     int32_t *__restrict__ argmax, int64_t count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint16_t *lut = reinterpret_cast<uint16_t *>(smem_raw);
    PairScratch *scratch = reinterpret_cast<PairScratch *>(smem_raw + TABLE_BYTES);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    PairScratch &ps = scratch[wid];
    build_cell_lut(lut, threadIdx.x, MI_WARPS * 32);
    for (int q = lane; q < NKEY; q += 32) ps.es.kmask[q] = 0u;
    if (lane == 0) {
        mbar_init(&ps.bar[0], 1);
        mbar_init(&ps.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t step = (int64_t)gridDim.x * MI_WARPS;
    auto issue = [&](int64_t pr, int b) {   // lane 0: the aligned windows around both rows of pair `pr` -> buffer b
        const uintptr_t ga = reinterpret_cast<uintptr_t>(grids + pr * gstride) & ~(uintptr_t)15;
        const uintptr_t ta = reinterpret_cast<uintptr_t>(targets + pr * tstride) & ~(uintptr_t)15;
        mbar_expect_tx(&ps.bar[b], 2u * GSTRIDE);
        tma_load(ps.grid[b], reinterpret_cast<const void *>(ga), GSTRIDE, &ps.bar[b]);
        tma_load(ps.tgt[b], reinterpret_cast<const void *>(ta), GSTRIDE, &ps.bar[b]);
    };
    int64_t pr = (int64_t)blockIdx.x * MI_WARPS + wid;
    if (pr < count && lane == 0) issue(pr, 0);
    for (uint32_t it = 0; pr < count; pr += step, it++) {
        const int b = it & 1;
        if (pr + step < count && lane == 0) issue(pr + step, b ^ 1);   // its last readers finished an iteration ago
        mbar_wait(&ps.bar[b], (it >> 1) & 1u);
        const int goff = (int)(reinterpret_cast<uintptr_t>(grids + pr * gstride) & 15);
        const int toff = (int)(reinterpret_cast<uintptr_t>(targets + pr * tstride) & 15);
        const uint32_t *G = ps.grid[b], *T = ps.tgt[b];
        const int8_t *t8 = reinterpret_cast<const int8_t *>(T) + toff, *g8 = reinterpret_cast<const int8_t *>(G) + goff;
        int4 rc = rects_identity();  // task.py:59-60 admissible = [[(0, 0)]]
        int n_rot = invariant ? 4 : 1;
        bool have_rc = !invariant;
        bool spans = !invariant;     // every evaluated rotation admits the shift (0, 0) only (bounding box = the zone)
        bool done = false;
        int key = -1, best = 0;
        // zone-spanning target with many blocks: dense comparison under the four rotations (one scope: its
        // registers are live on this path only). from_bbox: the rectangles come from the target's occupied columns.
        auto dense = [&](bool from_bbox) {
            int4 r;
            int k = -1;
            best = dense_try<WANT_KEY>(t8, g8, from_bbox, n_rot, lane, reinterpret_cast<uint2 *>(ps.es.hist), &r, &k);
            if (from_bbox) { have_rc = true; spans = best >= 0; }
            if (from_bbox && best < 0) rc = r;
            key = k;
            return best >= 0;
        };
        auto full_rects = [&]() {   // rectangles from full_grid (task.py:63-64); rare: read straight from global memory
            int x0 = 10, x1 = 0, z0 = 10, z1 = 0;
            const int8_t *f = fulls + pr * fstride;
            for (int c = lane; c < CELLS; c += 32)
                if (f[c] != 0) {
                    const int l = lut[c], x = lut_x(l), z = (l & 0xFF) - HROW * x;
                    x0 = min(x0, x); x1 = max(x1, x); z0 = min(z0, z); z1 = max(z1, z);
                }
            x0 = __reduce_min_sync(0xffffffffu, x0); x1 = __reduce_max_sync(0xffffffffu, x1);
            z0 = __reduce_min_sync(0xffffffffu, z0); z1 = __reduce_max_sync(0xffffffffu, z1);
            rc = rects_from_bbox(x0, x1, z0, z1);
            spans = x1 - x0 == 10 && z1 - z0 == 10;
            have_rc = true;
        };
        // ---- target: its non-zero words (first round in ps.twl); -1 = crowded from the start, no list made ----
        int nw = word_scan<false, true>(T, toff, nullptr, lane, 0, ps.twl, nullptr);
        if (nw < 0 && fulls) nw = word_scan<false>(T, toff, nullptr, lane, 0, ps.twl, nullptr);
        if (nw < 0 || nw > WL_CAP) {   // crowded (more than 32 non-zero words): if it spans the zone, no block list is needed
            bool attempt = true, from_bbox = false;
            if (!have_rc && fulls) full_rects();
            if (!have_rc) from_bbox = true;   // the target's own bounding box: from its packed columns
            else attempt = spans;
            if (attempt) done = dense(from_bbox);
        }
        if (!done) {
            if (nw < 0) nw = word_scan<false>(T, toff, nullptr, lane, 0, ps.twl, nullptr);
            WindowTargets tg(ps, T, toff, lut, nw, lane);   // blocks round by round
            if (!have_rc) {  // bounding box of the target (or of full_grid) -> admissible rectangles
                if (fulls) {
                    full_rects();
                } else {
                    int x0 = 10, x1 = 0, z0 = 10, z1 = 0;
                    for (int r = tg.rounds() - 1; r >= 0; r--) {   // backwards: round 0 stays loaded for the evaluation
                        const int nt = tg.load(r);
                        for (int i = lane; i < nt; i += 32) {
                            const int l = lut[ps.tl[i] & 0x7FF], x = lut_x(l), z = (l & 0xFF) - HROW * x;
                            x0 = min(x0, x); x1 = max(x1, x); z0 = min(z0, z); z1 = max(z1, z);
                        }
                    }
                    x0 = __reduce_min_sync(0xffffffffu, x0); x1 = __reduce_max_sync(0xffffffffu, x1);
                    z0 = __reduce_min_sync(0xffffffffu, z0); z1 = __reduce_max_sync(0xffffffffu, z1);
                    rc = rects_from_bbox(x0, x1, z0, z1);
                    spans = x1 - x0 == 10 && z1 - z0 == 10;
                }
            }
            if (spans) {  // one shift per rotation: compare grid bytes directly
                auto each = [&](auto f) {
                    for (int r = 0; r < tg.rounds(); r++) {
                        const int nt = tg.load(r);
                        for (int i = lane; i < nt; i += 32) {
                            const int e = ps.tl[i], l = lut[e & 0x7FF], x = lut_x(l);
                            f((l >> 8) / NCLS, x, (l & 0xFF) - HROW * x, e >> 11);
                        }
                    }
                };
                best = eval_single_shift<WANT_KEY>(g8, nullptr, each, rc, n_rot, &key);
            } else {
                best = eval_votes<WANT_KEY>(ps.es, G, goff, nullptr, lut, tg, rc, n_rot, lane, &key);
            }
        }
        if (lane == 0) {
            out[pr] = best;
            if (WANT_KEY) {
                int k = 0, dx = 0, dz = 0;
                if (key >= 0) { k = key / 441; int r = key - k * 441; dx = r / 21 - 10; dz = r % 21 - 10; }
                argmax[3 * pr + 0] = dx; argmax[3 * pr + 1] = dz; argmax[3 * pr + 2] = k;
            }
        }
        __syncwarp();  // every lane is done with buffer b before the next iteration refills it
    }
}

}  // namespace iglu

// ====================================================================== C ABI (host side)
using namespace iglu;

struct iglu_ctx {
    iglu_config cfg;
    int device;
    int num_sms;
    int last_launches;
    iglu_task_pool pool;  // copy of the attached generator (m == 0: none)
    cudaEvent_t trace[6];  // IGLU_B200_TRACE=1: timing events around the launches of the last overlapped step
    int tracing;
};

static int check(cudaError_t e) { return e == cudaSuccess ? IGLU_OK : IGLU_ERR_CUDA; }

extern "C" int iglu_abi_version(void) { return IGLU_ABI_VERSION; }

extern "C" const char *iglu_strerror(int status) {
    switch (status) {
        case IGLU_OK: return "ok";
        case IGLU_ERR_INVALID_ARG: return "invalid argument";
        case IGLU_ERR_CUDA: return cudaGetErrorString(cudaPeekAtLastError());
        case IGLU_ERR_NO_TASK:
            return "Task is not initialized! Initialize task before working with the environment using "
                   ".set_task method OR set tasks distribution using .set_task_generator method";
        case IGLU_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown status";
    }
}

extern "C" int iglu_ctx_create(const iglu_config *cfg, iglu_ctx **out) {
    if (!cfg || !out) return IGLU_ERR_INVALID_ARG;
    if (cfg->action_space < 0 || cfg->action_space > 2) return IGLU_ERR_INVALID_ARG;
    if (cfg->tlist_cap < TL_HDR + 64 || cfg->tlist_cap > IGLU_TLIST_ROW_MAX || (cfg->tlist_cap & 7))
        return IGLU_ERR_INVALID_ARG;
    iglu_ctx *c = new iglu_ctx();
    c->cfg = *cfg;
    c->last_launches = 0;
    memset(&c->pool, 0, sizeof(c->pool));
    if (cudaGetDevice(&c->device) != cudaSuccess) { delete c; return IGLU_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, c->device) != cudaSuccess) { delete c; return IGLU_ERR_CUDA; }
    c->num_sms = prop.multiProcessorCount;
    // Trig tables from the HOST libm (glibc == CPython's math module), un-reduced angles.
    static double2 trig[TRIG_N];
    static double sdeg[9];
    volatile double five = 5.0;  // keep the compiler from constant-folding libm calls
    for (int i = 0; i < TRIG_N; i++) {
        double deg = five * (double)(i + TRIG_LO);
        double rad = deg * DEG2RAD;
        trig[i].x = sin(rad);
        trig[i].y = cos(rad);
    }
    for (int a = -1; a <= 1; a++)
        for (int b = -1; b <= 1; b++) {
            volatile double s0 = a, s1 = b;
            sdeg[(a + 1) * 3 + (b + 1)] = atan2(s0, s1) * RAD2DEG;
        }
    volatile double g2 = 2.0 * 20.0 * 1.2;
    double jump = sqrt(g2);
    int rc = check(cudaMemcpyToSymbol(g_trig, trig, sizeof(trig)));
    if (!rc) rc = check(cudaMemcpyToSymbol(g_strafe_deg, sdeg, sizeof(sdeg)));
    if (!rc) rc = check(cudaMemcpyToSymbol(g_jump_speed, &jump, sizeof(jump)));
    static uint16_t cell_lut[CELL_LUT_N];
    for (int c = 0; c < CELL_LUT_N; c++) {
        const int y = c / 121, r = c - 121 * y, x = r / 11, z = r - 11 * x;
        cell_lut[c] = (uint16_t)(c < CELLS ? ((NCLS * y) << 8 | (HROW * x + z)) : 0);
    }
    if (!rc) rc = check(cudaMemcpyToSymbol(g_cell_lut, cell_lut, sizeof(cell_lut)));
    const int s3 = (int)step_smem_bytes<3>(), s2 = (int)step_smem_bytes<2>();
    if (!rc) rc = check(cudaFuncSetAttribute(k_step<0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, s3));
    if (!rc) rc = check(cudaFuncSetAttribute(k_step<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, s3));
    if (!rc) rc = check(cudaFuncSetAttribute(k_step<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, s3));
    if (!rc) rc = check(cudaFuncSetAttribute(k_step<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2));
    if (!rc) rc = check(cudaFuncSetAttribute(k_step<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2));
    if (!rc) rc = check(cudaFuncSetAttribute(k_step<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2));
    if (!rc) rc = check(cudaFuncSetAttribute(k_observe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)observe_smem_bytes()));
    if (!rc) rc = check(cudaFuncSetAttribute(k_widen<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)widen_smem_bytes(true)));
    if (!rc) rc = check(cudaFuncSetAttribute(k_widen<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)widen_smem_bytes(false)));
    if (!rc) rc = check(cudaFuncSetAttribute(k_commit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)commit_smem_bytes()));
    c->tracing = 0;
    for (int i = 0; i < 6; i++) c->trace[i] = nullptr;
    if (const char *tr = getenv("IGLU_B200_TRACE")) {
        if (tr[0] == '1') {
            c->tracing = 1;
            for (int i = 0; i < 6 && !rc; i++) rc = check(cudaEventCreate(&c->trace[i]));
        }
    }
    if (rc) { iglu_ctx_destroy(c); return rc; }
    *out = c;
    return IGLU_OK;
}

extern "C" void iglu_ctx_destroy(iglu_ctx *ctx) {
    if (!ctx) return;
    for (int i = 0; i < 6; i++) if (ctx->trace[i]) cudaEventDestroy(ctx->trace[i]);
    delete ctx;
}

extern "C" int iglu_last_launch_count(const iglu_ctx *ctx) { return ctx ? ctx->last_launches : 0; }

// Tuning aid (IGLU_B200_TRACE=1 when the context was created; synchronises): milliseconds, relative to the start of
// k_widen, of [k_widen end, physics start, physics end, k_commit start, k_commit end] of the last overlapped step.
extern "C" int iglu_debug_overlap_trace(iglu_ctx *ctx, float *out5) {
    if (!ctx || !out5 || !ctx->tracing) return IGLU_ERR_INVALID_ARG;
    if (cudaEventSynchronize(ctx->trace[5]) != cudaSuccess) return IGLU_ERR_CUDA;
    for (int i = 1; i < 6; i++)
        if (cudaEventElapsedTime(&out5[i - 1], ctx->trace[0], ctx->trace[i]) != cudaSuccess) return IGLU_ERR_CUDA;
    return IGLU_OK;
}

static bool state_ok(const iglu_state *st, const iglu_obs *obs) {
    return st && obs && st->n > 0 && st->grid && st->init && st->target && st->pose && st->init_pose &&
           st->istate && st->tlist && st->tmeta && obs->reward && obs->done;
}

extern "C" int iglu_set_tasks(iglu_ctx *ctx, const iglu_state *st, const iglu_obs *obs, const int8_t *target,
                              const int8_t *start, const double *init_pose, const int32_t *env_idx,
                              int64_t count, void *stream) {
    if (!ctx || !state_ok(st, obs) || !target || count <= 0) return IGLU_ERR_INVALID_ARG;
    const int threads = 256;
    const int64_t blocks = (count * 32 + threads - 1) / threads;
    k_task_prep<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(*st, *obs, ctx->cfg, target, start, nullptr,
                                                                         init_pose, env_idx, count);
    ctx->last_launches = 1;
    return check(cudaPeekAtLastError());
}

static bool pool_ok(const iglu_task_pool *pool) {
    return pool && pool->m > 0 && pool->m < (int64_t)1 << 31 && pool->target && pool->init && pool->tlist &&
           pool->tmeta && pool->cursor;
}

extern "C" int iglu_task_pool_prepare(iglu_ctx *ctx, const iglu_task_pool *pool, const int8_t *target,
                                      const int8_t *start, void *stream) {
    if (!ctx || !pool_ok(pool) || !target) return IGLU_ERR_INVALID_ARG;
    // the pool arrays have the layout of a state with n = m: Task.__init__ per row with the same kernel
    iglu_state st;
    iglu_obs obs;
    memset(&st, 0, sizeof(st));
    memset(&obs, 0, sizeof(obs));
    st.n = pool->m;
    st.target = pool->target; st.init = pool->init; st.tlist = pool->tlist; st.tmeta = pool->tmeta;
    const int threads = 256;
    const int64_t blocks = (pool->m * 32 + threads - 1) / threads;
    k_task_prep<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(st, obs, ctx->cfg, target, start, nullptr,
                                                                         nullptr, nullptr, pool->m);
    ctx->last_launches = 1;
    return check(cudaPeekAtLastError());
}

static bool staged_ok(const iglu_staged_tasks *sg) {
    return sg && sg->m > 0 && sg->m < (int64_t)1 << 31 && sg->tlist && sg->tmeta && sg->next_stage && sg->state && sg->out;
}

extern "C" int iglu_staged_prepare(iglu_ctx *ctx, const iglu_staged_tasks *sg, const int8_t *target,
                                   const int8_t *full, void *stream) {
    if (!ctx || !staged_ok(sg) || !target || !full) return IGLU_ERR_INVALID_ARG;
    iglu_state st;
    iglu_obs obs;
    memset(&st, 0, sizeof(st));
    memset(&obs, 0, sizeof(obs));
    st.n = sg->m;
    st.tlist = sg->tlist; st.tmeta = sg->tmeta;
    const int threads = 256;
    const int64_t blocks = (sg->m * 32 + threads - 1) / threads;
    k_task_prep<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(st, obs, ctx->cfg, target, nullptr, full, nullptr,
                                                                         nullptr, sg->m);
    ctx->last_launches = 1;
    return check(cudaPeekAtLastError());
}

extern "C" int iglu_staged_step(iglu_ctx *ctx, const iglu_state *st, const iglu_staged_tasks *sg,
                                const int32_t *task_row, void *stream) {
    if (!ctx || !st || st->n <= 0 || !st->grid || !st->init || !st->istate || !staged_ok(sg)) return IGLU_ERR_INVALID_ARG;
    const size_t smem = TABLE_BYTES + STAGED_WARPS * sizeof(StagedScratch);
    k_staged_step<<<(unsigned)((st->n + STAGED_WARPS - 1) / STAGED_WARPS), STAGED_WARPS * 32, smem,
                    (cudaStream_t)stream>>>(*st, *sg, task_row, ctx->cfg.tlist_cap);
    ctx->last_launches = 1;
    return check(cudaPeekAtLastError());
}

extern "C" int iglu_sample_tasks(int8_t *targets, int64_t m, const iglu_random_task_spec *spec, uint64_t seed,
                                 int64_t first_task, void *stream) {
    if (!targets || m <= 0 || !spec) return IGLU_ERR_INVALID_ARG;
    if (spec->max_blocks < 1 || spec->height_levels < 1 || spec->height_levels > 9 || spec->max_dist < 0 ||
        spec->max_dist > 10 || spec->num_colors < 1 || spec->num_colors > 6)
        return IGLU_ERR_INVALID_ARG;
    int rc = check(cudaMemsetAsync(targets, 0, (size_t)m * CELLS, (cudaStream_t)stream));
    if (rc) return rc;
    k_sample_tasks<<<(unsigned)((m + 127) / 128), 128, 0, (cudaStream_t)stream>>>(targets, m, *spec, seed, first_task);
    return check(cudaPeekAtLastError());
}

extern "C" int iglu_task_pool_attach(iglu_ctx *ctx, const iglu_task_pool *pool) {
    if (!ctx) return IGLU_ERR_INVALID_ARG;
    if (!pool) { memset(&ctx->pool, 0, sizeof(ctx->pool)); return IGLU_OK; }
    if (!pool_ok(pool)) return IGLU_ERR_INVALID_ARG;
    ctx->pool = *pool;
    return IGLU_OK;
}

extern "C" int iglu_reset(iglu_ctx *ctx, const iglu_state *st, const iglu_obs *obs, const uint8_t *mask,
                          int hard, void *stream) {
    if (!ctx || !state_ok(st, obs)) return IGLU_ERR_INVALID_ARG;
    const int threads = 256;
    const int64_t blocks = (st->n * 32 + threads - 1) / threads;
    k_reset<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(*st, *obs, mask, hard, ctx->pool,
                                                                     ctx->cfg.tlist_cap);
    ctx->last_launches = 1;
    return check(cudaPeekAtLastError());
}

// One step = either one fused launch (physics + observation/reward + bookkeeping) or two launches
// (physics, then observe). In the split form the physics launch needs no shared memory and the observe
// launch has no fp64 physics in it, so it fits 64 registers and runs 32 warps per SM on small chunks,
// which shortens the latency-bound per-env chain. reserved bit 1 forces the fused launch, bit 2 the split.
template <int KIND>
static int launch_step(iglu_ctx *ctx, StepParams &p, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n = p.st.n;
    // measured (profiles/r01_launch_policy.txt): the fused launch only wins for on_change reward + fully
    // rewritten obs['grid'] (its ring prefetch hides behind the physics); everything else is faster split
    // Overlapped step (k_widen || physics, then k_commit): whenever obs['grid'] is rewritten in full and the batch is
    // big enough for the overlap to pay for the third launch (measured, profiles/). reserved bit 6 = 1 turns it off.
    // The same physics + k_commit pair WITHOUT the widen launch serves reward on change when obs['grid'] is not rewritten
    // in full (in-place / incremental observations, BASELINE config 5): the physics launch finishes every env it neither
    // edited nor reset, k_commit evaluates the few others -- no pass over all envs after the physics.
    // (With every-step reward and in-place observations the widen launch still runs next to the physics -- it then only
    // stages and evaluates the envs that hold blocks.)
    const bool widen = (p.cfg.write_grid_obs == 1 && p.obs.grid) || p.cfg.reward_mode == IGLU_REWARD_EVERY_STEP;
    const bool commit_only = !widen;
    if ((widen || commit_only) && p.st.worklist && p.st.sched && (n >= (widen ? 8192 : 2048) || (p.cfg.reserved & 128)) &&
        !(p.cfg.reserved & (64 | 2 | 4))) {
        const bool w_eval = p.cfg.reward_mode == IGLU_REWARD_EVERY_STEP;
        const int WIDEN_WARPS = widen_warps(w_eval);
        int64_t chunk = (n + (int64_t)ctx->num_sms * WIDEN_WARPS - 1) / ((int64_t)ctx->num_sms * WIDEN_WARPS);
        chunk /= 6;
        chunk = chunk < 2 ? 2 : (chunk > IGLU_OBS_DYN_GROUP ? IGLU_OBS_DYN_GROUP : chunk);
        if ((p.cfg.reserved >> 8) & 63) chunk = (p.cfg.reserved >> 8) & 63;     // tuning override
        p.group = (int)(chunk > 32 ? 32 : chunk);
        p.overlap = 1;
        // one stream, three launches: k_widen; the physics launch as its PROGRAMMATIC dependent (it starts once every
        // k_widen CTA is resident -- so k_widen has its CTA per SM -- and runs next to it); k_commit behind both
        const bool tr = ctx->tracing != 0;
        const int64_t groups_w = (n + p.group - 1) / p.group;
        const int64_t ctas_w = (groups_w + WIDEN_WARPS - 1) / WIDEN_WARPS;
        if (tr) cudaEventRecord(ctx->trace[0], s);
        cudaLaunchAttribute ap[1];
        ap[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        ap[0].val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t lw = {};
        lw.gridDim = dim3((unsigned)(ctas_w < ctx->num_sms ? ctas_w : ctx->num_sms));
        lw.blockDim = dim3(WIDEN_WARPS * 32);
        lw.dynamicSmemBytes = widen_smem_bytes(w_eval);
        lw.stream = s;
        lw.attrs = ap;
        lw.numAttrs = tr ? 0 : 1;   // (a timing event in front of the launch: nothing to overlap with)
        int rc = IGLU_OK;
        if (widen) rc = w_eval ? check(cudaLaunchKernelEx(&lw, k_widen<true>, p)) : check(cudaLaunchKernelEx(&lw, k_widen<false>, p));
        if (rc) return rc;
        StepParams pa = p;
        pa.group = 32;
        cudaLaunchConfig_t lp = {};
        lp.gridDim = dim3((unsigned)(((n + 31) / 32 + StepCfg<1>::WPC - 1) / StepCfg<1>::WPC));
        lp.blockDim = dim3(StepCfg<1>::WPC * 32);
        lp.dynamicSmemBytes = 0;
        lp.stream = s;
        lp.attrs = ap;
        lp.numAttrs = widen ? 1 : 0;   // (without a widen launch in front the physics launch is a normal one)
        rc = check(cudaLaunchKernelEx(&lp, k_step<KIND, 1>, pa));
        if (rc) return rc;
        if (tr) { cudaEventRecord(ctx->trace[1], s); cudaEventRecord(ctx->trace[2], s); cudaEventRecord(ctx->trace[3], s); }
        const int64_t ctas_c = (n / 8 + COMMIT_WARPS - 1) / COMMIT_WARPS;   // ~1 listed env in 14: about one per warp
        const int64_t max_c = (p.cfg.reserved & 32768) ? 3 : (int64_t)ctx->num_sms * 7;   // bit 15: tiny grid (tests)
        if (tr) cudaEventRecord(ctx->trace[4], s);
        cudaLaunchConfig_t lcm = {};
        lcm.gridDim = dim3((unsigned)(ctas_c < max_c ? (ctas_c > 0 ? ctas_c : 1) : max_c));
        lcm.blockDim = dim3(COMMIT_WARPS * 32);
        lcm.dynamicSmemBytes = commit_smem_bytes();
        lcm.stream = s;
        lcm.attrs = ap;
        lcm.numAttrs = tr ? 0 : 1;
        rc = check(cudaLaunchKernelEx(&lcm, k_commit, p));
        if (rc) return rc;
        if (tr) cudaEventRecord(ctx->trace[5], s);
        ctx->last_launches = widen ? 3 : 2;
        return check(cudaPeekAtLastError());
    }
    cudaLaunchAttribute pdl[1];   // every step launch may overlap its set-up with the tail of the kernel before it
    pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl[0].val.programmaticStreamSerializationAllowed = 1;
    const bool fused_better = p.cfg.reward_mode != IGLU_REWARD_EVERY_STEP && p.cfg.write_grid_obs == 1 && p.obs.grid;
    const bool split = ((!fused_better) && !(p.cfg.reserved & 2)) || (p.cfg.reserved & 4);
    if (!split) {
        const int64_t resident_ctas = (int64_t)ctx->num_sms * CTAS_PER_SM;
        const int64_t warp_slots = resident_ctas * StepCfg<3>::WPC;
        p.group = (2 * ((n + 31) / 32) <= warp_slots) ? 16 : 32;
        const int64_t groups = (n + p.group - 1) / p.group;
        const unsigned grid = (unsigned)(groups < resident_ctas ? groups : resident_ctas);
        cudaLaunchConfig_t lf = {};
        lf.gridDim = dim3(grid);
        lf.blockDim = dim3(StepCfg<3>::WPC * 32);
        lf.dynamicSmemBytes = step_smem_bytes<3>();
        lf.stream = s;
        lf.attrs = pdl;
        lf.numAttrs = 1;
        const int rcf = check(cudaLaunchKernelEx(&lf, k_step<KIND, 3>, p));
        ctx->last_launches = 1;
        return rcf ? rcf : check(cudaPeekAtLastError());
    }
    // launch 1: physics, one thread per env, 32-env groups, 4-warp CTAs
    p.group = 32;
    const int64_t groups_a = (n + 31) / 32;
    k_step<KIND, 1><<<(unsigned)((groups_a + StepCfg<1>::WPC - 1) / StepCfg<1>::WPC), StepCfg<1>::WPC * 32, 0, s>>>(p);
    // launch 2: observe, persistent, chunk = ceil(n / resident warps) envs per warp (<= 32)
    const int64_t resident_ctas = (int64_t)ctx->num_sms * IGLU_OBS_CTAS;
    const int64_t warp_slots = resident_ctas * StepCfg<2>::WPC;
    int64_t chunk = (n + warp_slots - 1) / warp_slots;
    // dynamic hand-out pays when every env has ring work in this launch (measured: profiles/r01_launch_policy.txt);
    // with on_change reward and no full obs rewrite only ~6 % of the envs do, and big static chunks are cheaper
    const bool dynamic = p.st.sched != nullptr && !(p.cfg.reserved & 8) &&
                         (p.cfg.reward_mode == IGLU_REWARD_EVERY_STEP || (p.cfg.write_grid_obs == 1 && p.obs.grid));
    p.dynamic = dynamic;
    if (dynamic) {  // small groups handed out dynamically: >= ~6 per warp, 2..IGLU_OBS_DYN_GROUP envs each (measured)
        chunk /= 6;
        chunk = chunk < 2 ? 2 : (chunk > IGLU_OBS_DYN_GROUP ? IGLU_OBS_DYN_GROUP : chunk);
    }
    if ((p.cfg.reserved >> 8) & 63) chunk = (p.cfg.reserved >> 8) & 63;     // tuning override
    p.group = (int)(chunk < 1 ? 1 : (chunk > 32 ? 32 : chunk));
    const int64_t groups_b = (n + p.group - 1) / p.group;
    const int64_t ctas_b = (groups_b + StepCfg<2>::WPC - 1) / StepCfg<2>::WPC;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)(ctas_b < resident_ctas ? ctas_b : resident_ctas));
    lc.blockDim = dim3(StepCfg<2>::WPC * 32);
    lc.dynamicSmemBytes = step_smem_bytes<2>();
    lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    int rc;
    if (dynamic && !(p.cfg.write_grid_obs == 1 && p.obs.grid) && !(p.cfg.reserved & 32)) {   // measured: profiles/
        // tiled observe launch: one tile per resident CTA when the batch is small, OBS_TILE-env tiles otherwise
        // static part: `share` / 64 of each CTA's chunks, in tiles of <= OBS_TILE envs; the rest is the dynamic tail
        const int64_t chunks = n / OBS_CHUNK;
        const int64_t ctas = chunks < resident_ctas ? (chunks > 0 ? chunks : 1) : resident_ctas;
        const int share = ((p.cfg.reserved >> 8) & 63) ? ((p.cfg.reserved >> 8) & 63) : 64;
        int64_t k = (chunks / ctas) * share / 64;        // static chunks per CTA
        int64_t tiles = (k + OBS_TILE / OBS_CHUNK - 1) / (OBS_TILE / OBS_CHUNK);
        if (tiles > 0) k /= tiles;                       // chunks per tile (<= 32)
        p.group = (int)(k * OBS_CHUNK);
        p.tiles = (int)tiles;
        p.tail_begin = tiles * k * ctas * OBS_CHUNK;
        lc.gridDim = dim3((unsigned)ctas);
        lc.dynamicSmemBytes = observe_smem_bytes();
        rc = check(cudaLaunchKernelEx(&lc, k_observe, p));
    } else {
        rc = check(cudaLaunchKernelEx(&lc, k_step<KIND, 2>, p));
    }
    ctx->last_launches = 2;
    return rc ? rc : check(cudaPeekAtLastError());
}

extern "C" int iglu_step_walking(iglu_ctx *ctx, const iglu_state *st, const int32_t *actions,
                                 const iglu_obs *obs, void *stream) {
    if (!ctx || !state_ok(st, obs) || !actions) return IGLU_ERR_INVALID_ARG;
    if (ctx->cfg.action_space != IGLU_WALKING_DISCRETE) return IGLU_ERR_INVALID_ARG;
    StepParams p;
    memset(&p, 0, sizeof(p));
    p.st = *st; p.obs = *obs; p.cfg = ctx->cfg; p.pool = ctx->pool; p.actions = actions;
    return launch_step<IGLU_WALKING_DISCRETE>(ctx, p, stream);
}

extern "C" int iglu_step_flying(iglu_ctx *ctx, const iglu_state *st, const float *movement, const float *camera,
                                const int32_t *inventory, const int32_t *placement, const iglu_obs *obs,
                                void *stream) {
    if (!ctx || !state_ok(st, obs) || !movement || !camera || !inventory || !placement) return IGLU_ERR_INVALID_ARG;
    if (ctx->cfg.action_space != IGLU_FLYING) return IGLU_ERR_INVALID_ARG;
    StepParams p;
    memset(&p, 0, sizeof(p));
    p.st = *st; p.obs = *obs; p.cfg = ctx->cfg; p.pool = ctx->pool;
    p.movement = movement; p.camera = camera; p.inventory = inventory; p.placement = placement;
    return launch_step<IGLU_FLYING>(ctx, p, stream);
}

extern "C" int iglu_step_walking_dict(iglu_ctx *ctx, const iglu_state *st, const int32_t *buttons,
                                      const float *camera, const iglu_obs *obs, void *stream) {
    if (!ctx || !state_ok(st, obs) || !buttons || !camera) return IGLU_ERR_INVALID_ARG;
    if (ctx->cfg.action_space != IGLU_WALKING_DICT) return IGLU_ERR_INVALID_ARG;
    StepParams p;
    memset(&p, 0, sizeof(p));
    p.st = *st; p.obs = *obs; p.cfg = ctx->cfg; p.pool = ctx->pool; p.buttons = buttons; p.camera = camera;
    return launch_step<IGLU_WALKING_DICT>(ctx, p, stream);
}

// Device-visible alias of a pinned (page-locked, UVA-mapped) host pointer, or null for pageable memory.
static void *mapped_alias(const void *host) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

static int step_walking_host(iglu_ctx *ctx, const iglu_state *st, const int32_t *actions_host,
                             int32_t *actions_dev, const iglu_obs *obs, double *reward_host,
                             uint8_t *done_host, void *stream, bool sync) {
    if (!ctx || !state_ok(st, obs) || !actions_host || !actions_dev || !reward_host || !done_host)
        return IGLU_ERR_INVALID_ARG;
    if (ctx->cfg.action_space != IGLU_WALKING_DISCRETE) return IGLU_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    // Pinned host buffers are read / written by the kernels directly over PCIe (zero-copy): the actions are
    // fetched by the physics launch and reward / done are stored to the host by the observe launch as it goes,
    // so no transfer is serialised before or after the step. Pageable buffers take the staged-copy route.
    const int32_t *a_map = static_cast<const int32_t *>(mapped_alias(actions_host));
    // (queried on every call, not cached: a freed pinned buffer's address can come back as pageable memory)
    double *r_map = static_cast<double *>(mapped_alias(reward_host));
    uint8_t *d_map = static_cast<uint8_t *>(mapped_alias(done_host));
    const bool zero_copy = a_map && r_map && d_map && !(ctx->cfg.reserved & 16);
    int rc;
    if (!zero_copy) {
        rc = check(cudaMemcpyAsync(actions_dev, actions_host, sizeof(int32_t) * st->n, cudaMemcpyHostToDevice, s));
        if (rc) return rc;
    }
    StepParams p;
    memset(&p, 0, sizeof(p));
    p.st = *st; p.obs = *obs; p.cfg = ctx->cfg; p.pool = ctx->pool;
    p.actions = zero_copy ? a_map : actions_dev;
    if (zero_copy) { p.reward_host = r_map; p.done_host = d_map; }
    rc = launch_step<IGLU_WALKING_DISCRETE>(ctx, p, stream);
    if (rc) return rc;
    if (!zero_copy) {
        rc = check(cudaMemcpyAsync(reward_host, obs->reward, sizeof(double) * st->n, cudaMemcpyDeviceToHost, s));
        if (rc) return rc;
        rc = check(cudaMemcpyAsync(done_host, obs->done, sizeof(uint8_t) * st->n, cudaMemcpyDeviceToHost, s));
        if (rc) return rc;
    }
    return sync ? check(cudaStreamSynchronize(s)) : IGLU_OK;
}

extern "C" int iglu_step_walking_host(iglu_ctx *ctx, const iglu_state *st, const int32_t *actions_host,
                                      int32_t *actions_dev, const iglu_obs *obs, double *reward_host,
                                      uint8_t *done_host, void *stream) {
    return step_walking_host(ctx, st, actions_host, actions_dev, obs, reward_host, done_host, stream, true);
}

extern "C" int iglu_step_walking_host_async(iglu_ctx *ctx, const iglu_state *st, const int32_t *actions_host,
                                            int32_t *actions_dev, const iglu_obs *obs, double *reward_host,
                                            uint8_t *done_host, void *stream) {
    return step_walking_host(ctx, st, actions_host, actions_dev, obs, reward_host, done_host, stream, false);
}

extern "C" int iglu_wait(void *stream) { return check(cudaStreamSynchronize((cudaStream_t)stream)); }

// ------------------------------------------------------------------ single-env read-back (N = 1 drop-in surface)
namespace iglu {
__global__ void __launch_bounds__(256) k_fetch_env(iglu_state st, iglu_obs obs, int64_t e, iglu_env_snapshot *out) {
    const int64_t n = st.n;
    const int t = threadIdx.x;
    const int8_t *g = st.grid + e * GSTRIDE, *tg = st.target + e * GSTRIDE;
    for (int c = t; c < CELLS; c += blockDim.x) {
        out->grid[c] = (int32_t)g[c];
        out->target_grid[c] = (int32_t)tg[c];
    }
    if (t < 6) out->pose[t] = st.pose[t * n + e];
    if (t < 5) out->agent_pos[t] = obs.agent_pos ? obs.agent_pos[5 * e + t] : 0.f;
    if (t >= 32 && t < 38) out->inventory_obs[t - 32] = obs.inventory ? obs.inventory[6 * e + (t - 32)] : 0.f;
    if (t == 64) {
        const int w0 = st.istate[0 * n + e], w1 = st.istate[1 * n + e], w2 = st.istate[2 * n + e];
        const int w3 = st.istate[3 * n + e];
        out->inventory[0] = lo16(w0); out->inventory[1] = hi16(w0); out->inventory[2] = lo16(w1);
        out->inventory[3] = hi16(w1); out->inventory[4] = lo16(w2); out->inventory[5] = hi16(w2);
        out->active_block = w3 & 0xFF;
        out->time_int_steps = (w3 >> 8) & 0xFF;
        out->error = (w3 >> 24) & 0xFF;
        out->step_no = st.istate[4 * n + e];
        out->grid_size = st.istate[5 * n + e];
        out->max_int = st.istate[6 * n + e];
        out->target_size = st.tmeta[1 * n + e];
        out->done = obs.done[e];
        out->reward = obs.reward[e];
        out->compass = obs.compass ? obs.compass[e] : 0.f;
    }
}
}  // namespace iglu

extern "C" int iglu_fetch_env_host(iglu_ctx *ctx, const iglu_state *st, const iglu_obs *obs, int64_t env,
                                   iglu_env_snapshot *scratch_dev, iglu_env_snapshot *out_host, void *stream) {
    if (!ctx || !state_ok(st, obs) || env < 0 || env >= st->n || !scratch_dev || !out_host) return IGLU_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    k_fetch_env<<<1, 256, 0, s>>>(*st, *obs, env, scratch_dev);
    ctx->last_launches = 1;
    int rc = check(cudaPeekAtLastError());
    if (rc) return rc;
    rc = check(cudaMemcpyAsync(out_host, scratch_dev, sizeof(iglu_env_snapshot), cudaMemcpyDeviceToHost, s));
    if (rc) return rc;
    return check(cudaStreamSynchronize(s));
}

template <bool WANT_KEY>
static int launch_mi(const int8_t *grids, const int8_t *targets, const int8_t *fulls, int64_t gs, int64_t ts, int64_t fs,
                     int invariant, int32_t *out, int32_t *argmax, int64_t count, unsigned grid, size_t smem,
                     cudaStream_t s) {
    int rc = check(cudaFuncSetAttribute(k_max_intersection<WANT_KEY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    if (rc) return rc;
    k_max_intersection<WANT_KEY><<<grid, MI_WARPS * 32, smem, s>>>(grids, targets, fulls, gs, ts, fs, invariant, out,
                                                                  argmax, count);
    return check(cudaPeekAtLastError());
}

extern "C" int iglu_render_pov(iglu_ctx *ctx, const iglu_state *st, uint8_t *pov, int height, int width,
                               const uint8_t *atlas_rgba, void *stream) {
    if (!ctx || !st || st->n <= 0 || !st->grid || !st->pose || !pov || height < 1 || width < 1 || height > 1024 ||
        width > 1024)
        return IGLU_ERR_INVALID_ARG;
    k_render_pov<<<(unsigned)st->n, 256, 0, (cudaStream_t)stream>>>(*st, pov, height, width, atlas_rgba);
    ctx->last_launches = 1;
    return check(cudaPeekAtLastError());
}

extern "C" int iglu_max_intersection_strided(const int8_t *grids, int64_t grid_stride, const int8_t *targets,
                                             int64_t target_stride, const int8_t *fulls, int64_t full_stride,
                                             int invariant, int32_t *out, int32_t *argmax, int64_t count,
                                             void *stream) {
    if (!grids || !targets || !out || count <= 0) return IGLU_ERR_INVALID_ARG;
    if (grid_stride < CELLS || target_stride < CELLS || (fulls && full_stride < CELLS)) return IGLU_ERR_INVALID_ARG;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return IGLU_ERR_CUDA;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return IGLU_ERR_CUDA;
    const size_t smem = TABLE_BYTES + MI_WARPS * sizeof(PairScratch);
    const int64_t ctas = (count + MI_WARPS - 1) / MI_WARPS;
    const int64_t resident = (int64_t)sms * MI_CTAS;
    const unsigned grid = (unsigned)(ctas < resident ? ctas : resident);
    cudaStream_t s = (cudaStream_t)stream;
    return argmax ? launch_mi<true>(grids, targets, fulls, grid_stride, target_stride, full_stride, invariant, out,
                                    argmax, count, grid, smem, s)
                  : launch_mi<false>(grids, targets, fulls, grid_stride, target_stride, full_stride, invariant, out,
                                     argmax, count, grid, smem, s);
}

extern "C" int iglu_max_intersection(const int8_t *grids, const int8_t *targets, const int8_t *fulls,
                                     int invariant, int32_t *out, int32_t *argmax, int64_t count,
                                     void *stream) {
    return iglu_max_intersection_strided(grids, CELLS, targets, CELLS, fulls, CELLS, invariant, out, argmax, count,
                                         stream);
}
