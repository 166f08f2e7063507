// iglu_b200.cu -- sm_100a kernels + C ABI (include/iglu_b200.h) of the IGLU Gridworld step path.
//
// One warp owns 32 consecutive envs:
//   phase A (one thread per env): auto-reset, action parse, World.step in fp64 (iglu_physics.cuh),
//            small observations (agentPos / inventory / compass);
//   phase B (warp-cooperative, one env at a time): stage the env's int8 grid in shared memory with
//            16-byte loads, widen it into the int32 obs['grid'] with coalesced stores, and evaluate
//            Task.maximal_intersection (iglu_reward.cuh) when required;
//   phase C (one thread per env): Task.step_intersection bookkeeping, reward, done (task.py:103-119,
//            env.py:290-296).
// No collective, no host round trip; every entry point only enqueues work on the caller's stream.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/iglu_b200.h"
#include "iglu_physics.cuh"
#include "iglu_reward.cuh"

namespace iglu {

constexpr int WARPS_PER_CTA = 8;
constexpr int CTA_THREADS = WARPS_PER_CTA * 32;
constexpr size_t SMEM_BYTES = SPREAD_N * sizeof(unsigned long long) + WARPS_PER_CTA * sizeof(WarpScratch);

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
};

__device__ __forceinline__ int16_t lo16(int w) { return (int16_t)(w & 0xFFFF); }
__device__ __forceinline__ int16_t hi16(int w) { return (int16_t)((unsigned)w >> 16); }
__device__ __forceinline__ int pack16(int a, int b) { return (a & 0xFFFF) | (b << 16); }

// Cooperative 16-byte copy of one env's padded grid (69 x int4) global -> shared.
__device__ __forceinline__ void load_grid_smem(uint32_t *dst, const int8_t *src, int lane) {
    const int4 *s = reinterpret_cast<const int4 *>(src);
    int4 *d = reinterpret_cast<int4 *>(dst);
    for (int q = lane; q < GSTRIDE / 16; q += 32) d[q] = s[q];
}

template <int KIND>
__global__ void __launch_bounds__(CTA_THREADS, 2) k_step(const StepParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *tbl = reinterpret_cast<unsigned long long *>(smem_raw);
    WarpScratch *scratch = reinterpret_cast<WarpScratch *>(smem_raw + SPREAD_N * sizeof(unsigned long long));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpScratch &ws = scratch[wid];
    const bool need_table = p.cfg.reward_mode == IGLU_REWARD_EVERY_STEP;
    // In ON_CHANGE mode few envs evaluate; the table is still needed, build it once per CTA.
    build_spread_table(tbl, threadIdx.x, CTA_THREADS);
    (void)need_table;
    __syncthreads();

    const int64_t n = p.st.n;
    const int64_t ngroups = (n + 31) / 32;
    const bool flying = KIND == IGLU_FLYING;

    for (int64_t grp = (int64_t)blockIdx.x * WARPS_PER_CTA + wid; grp < ngroups;
         grp += (int64_t)gridDim.x * WARPS_PER_CTA) {
        const int64_t i = grp * 32 + lane;
        const bool valid = i < n;
        const int64_t ii = valid ? i : n - 1;  // clamp: invalid lanes read a real env, write nothing

        // ------------------------------------------------------------- phase A
        int w[8], tm[8];
#pragma unroll
        for (int c = 0; c < 8; c++) w[c] = p.st.istate[c * n + ii];
#pragma unroll
        for (int c = 0; c < 6; c++) tm[c] = p.st.tmeta[c * n + ii];
        const int nblk = tm[0], target_size = tm[1];
        const bool has_start = tm[5] != 0;
        Agent a;
        a.x = p.st.pose[0 * n + ii]; a.y = p.st.pose[1 * n + ii]; a.z = p.st.pose[2 * n + ii];
        a.yaw = p.st.pose[3 * n + ii]; a.pitch = p.st.pose[4 * n + ii]; a.dy = p.st.pose[5 * n + ii];
        a.inv[0] = lo16(w[0]); a.inv[1] = hi16(w[0]); a.inv[2] = lo16(w[1]);
        a.inv[3] = hi16(w[1]); a.inv[4] = lo16(w[2]); a.inv[5] = hi16(w[2]);
        a.active = w[3] & 0xFF; a.tis = (w[3] >> 8) & 0xFF; a.err = (w[3] >> 24) & 0xFF;
        const bool done_prev = ((w[3] >> 16) & 0xFF) != 0;
        int step_no = w[4], dsize = w[5], max_int = w[6];

        // auto-reset (env.py:206-261): grid <- starting grid, cooperatively, 16 bytes per lane
        const bool do_reset = valid && p.cfg.auto_reset && done_prev;
        unsigned rm = __ballot_sync(0xffffffffu, do_reset);
        while (rm) {
            int j = __ffs(rm) - 1;
            rm &= rm - 1;
            int64_t ej = grp * 32 + j;
            bool hs = __shfl_sync(0xffffffffu, (int)has_start, j) != 0;
            int4 *dst = reinterpret_cast<int4 *>(p.st.grid + ej * GSTRIDE);
            const int4 *src = reinterpret_cast<const int4 *>(p.st.init + ej * GSTRIDE);
            for (int q = lane; q < GSTRIDE / 16; q += 32) dst[q] = hs ? src[q] : make_int4(0, 0, 0, 0);
        }
        __syncwarp();
        if (do_reset) {
            a.x = p.st.init_pose[0 * n + i]; a.y = p.st.init_pose[1 * n + i]; a.z = p.st.init_pose[2 * n + i];
            a.yaw = p.st.init_pose[3 * n + i]; a.pitch = p.st.init_pose[4 * n + i];
            a.inv[0] = lo16(tm[2]); a.inv[1] = hi16(tm[2]); a.inv[2] = lo16(tm[3]);
            a.inv[3] = hi16(tm[3]); a.inv[4] = lo16(tm[4]); a.inv[5] = hi16(tm[4]);
            step_no = 0; dsize = 0; max_int = 0;  // dy / tis / active survive (reference quirk)
        }

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

        const int dsize_old = dsize;
        if (valid) {
            step_no += 1;  // env.py:276
            world_step(p.st.grid + i * GSTRIDE, p.st.init + i * GSTRIDE, has_start, a, act, flying,
                       p.cfg.select_and_place != 0, dsize);
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
#pragma unroll
                for (int c = 0; c < 6; c++) iv[c] = (float)a.inv[c];
            }
            if (p.obs.compass) p.obs.compass[i] = (float)(a.yaw - 180.0);
        }
        __syncwarp();  // grid edits of phase A are visible to the whole warp below

        // ------------------------------------------------------------- phase B
        const bool changed = dsize != dsize_old;  // task.py:111-112 gate (wrong_placement != 0)
        const bool want_eval = valid && nblk > 0 &&
                               (p.cfg.reward_mode == IGLU_REWARD_EVERY_STEP ? true : changed);
        const bool want_grid = valid && p.cfg.write_grid_obs && p.obs.grid != nullptr;
        int mi_eval = 0;
        unsigned cm = __ballot_sync(0xffffffffu, want_eval || want_grid);
        const unsigned em = __ballot_sync(0xffffffffu, want_eval);
        const unsigned hm = __ballot_sync(0xffffffffu, has_start);
        while (cm) {
            const int j = __ffs(cm) - 1;
            cm &= cm - 1;
            const int64_t ej = grp * 32 + j;
            const bool ev = (em >> j) & 1u, hs = (hm >> j) & 1u;
            load_grid_smem(ws.grid, p.st.grid + ej * GSTRIDE, lane);
            if (ev && hs) load_grid_smem(ws.init, p.st.init + ej * GSTRIDE, lane);
            __syncwarp();
            if (p.cfg.write_grid_obs && p.obs.grid != nullptr) {
                int32_t *og = p.obs.grid + ej * CELLS;
                const int8_t *sg = reinterpret_cast<const int8_t *>(ws.grid);
                for (int c = lane; c < CELLS; c += 32) og[c] = (int32_t)sg[c];
            }
            if (ev) {
                build_rows(ws, hs, lane);
                const int nb = __shfl_sync(0xffffffffu, nblk, j);
                int r = eval_max_intersection<false>(ws, tbl, p.st.tlist + ej * p.cfg.tlist_cap, nb,
                                                     p.st.trect + ej * 16, 4, lane, nullptr);
                if (lane == j) mi_eval = r;
            }
            __syncwarp();
        }

        // ------------------------------------------------------------- phase C
        if (valid) {
            const int wrong = dsize_old - dsize;                  // task.py:111
            const int mi = wrong != 0 ? mi_eval : max_int;        // task.py:112
            bool done = mi == target_size;                        // task.py:113
            const int right = mi - max_int;                       // task.py:115
            done = done || (step_no == p.cfg.max_steps);          // env.py:292
            const double reward = right == 0 ? (double)wrong * p.cfg.wrong_placement_scale
                                             : (double)right * p.cfg.right_placement_scale;
            p.obs.reward[i] = reward;
            p.obs.done[i] = (uint8_t)done;
            p.st.istate[0 * n + i] = pack16(a.inv[0], a.inv[1]);
            p.st.istate[1 * n + i] = pack16(a.inv[2], a.inv[3]);
            p.st.istate[2 * n + i] = pack16(a.inv[4], a.inv[5]);
            p.st.istate[3 * n + i] = (a.active & 0xFF) | ((a.tis & 0xFF) << 8) | ((int)done << 16) | ((a.err & 0xFF) << 24);
            p.st.istate[4 * n + i] = step_no;
            p.st.istate[5 * n + i] = dsize;
            p.st.istate[6 * n + i] = mi;
        }
    }
}

// ------------------------------------------------------------------ task preparation
// Task.__init__ of the synthetic task (env.py:227-232, task.py:9-72) + Tasks.to_dense'd starting grid.
// One warp per row.
__global__ void __launch_bounds__(256) k_task_prep(iglu_state st, iglu_obs obs, iglu_config cfg,
                                                   const int8_t *__restrict__ target,
                                                   const int8_t *__restrict__ start,
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
    int8_t *gt = st.target + e * GSTRIDE, *gi = st.init + e * GSTRIDE;
    uint16_t *tl = st.tlist + e * cfg.tlist_cap;
    int nout = 0, xmin = 10, xmax = 0, zmin = 10, zmax = 0, any_start = 0;
    int cnt[6] = {0, 0, 0, 0, 0, 0};
    for (int c0 = 0; c0 < GSTRIDE; c0 += 32) {
        const int c = c0 + lane;
        int tv = 0, sv = 0;
        if (c < CELLS) {
            tv = t[c];
            sv = s ? (int)s[c] : 0;
        }
        if (c < GSTRIDE) { gt[c] = (int8_t)tv; gi[c] = (int8_t)sv; }
        if (c < CELLS && obs.target_grid) obs.target_grid[e * CELLS + c] = tv;
        const int d = tv - sv;  // env.py:230
        const unsigned m = __ballot_sync(0xffffffffu, d != 0);
        any_start |= __any_sync(0xffffffffu, sv != 0);
#pragma unroll
        for (int k = 0; k < 6; k++) cnt[k] += __popc(__ballot_sync(0xffffffffu, sv == k + 1));
        if (d != 0) {
            const int y = c / 121, r = c - 121 * y, x = r / 11, z = r - 11 * x;
            const int pos = nout + __popc(m & ((1u << lane) - 1u));
            if (pos < cfg.tlist_cap) tl[pos] = (uint16_t)(y | (x << 4) | (z << 8) | ((d & 15) << 12));
            xmin = min(xmin, x); xmax = max(xmax, x); zmin = min(zmin, z); zmax = max(zmax, z);
        }
        nout += __popc(m);
    }
    xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
    zmin = __reduce_min_sync(0xffffffffu, zmin); zmax = __reduce_max_sync(0xffffffffu, zmax);
    if (lane == 0) {
        int8_t rc[16];
        rects_from_bbox(xmin, xmax, zmin, zmax, rc);
#pragma unroll
        for (int k = 0; k < 16; k++) st.trect[e * 16 + k] = rc[k];
        st.tmeta[0 * n + e] = min(nout, cfg.tlist_cap);
        st.tmeta[1 * n + e] = nout;  // target_size = nnz(target - init)
        st.tmeta[2 * n + e] = pack16(20 - cnt[0], 20 - cnt[1]);
        st.tmeta[3 * n + e] = pack16(20 - cnt[2], 20 - cnt[3]);
        st.tmeta[4 * n + e] = pack16(20 - cnt[4], 20 - cnt[5]);
        st.tmeta[5 * n + e] = any_start;
        st.tmeta[6 * n + e] = 1;
        st.tmeta[7 * n + e] = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) st.init_pose[k * n + e] = init_pose ? init_pose[row * 5 + k] : 0.0;
    }
}

// ------------------------------------------------------------------ reset
// GridWorld.reset (env.py:206-261). One warp per env.
__global__ void __launch_bounds__(256) k_reset(iglu_state st, iglu_obs obs, const uint8_t *__restrict__ mask,
                                               int hard) {
    const int lane = threadIdx.x & 31;
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n = st.n;
    if (e >= n) return;
    if (mask && !mask[e]) return;
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

// ------------------------------------------------------------------ stateless reward sweep
// Task(target, full_grid, invariant).maximal_intersection(grid) per pair (BASELINE config 4).
struct __align__(16) PairScratch {
    uint16_t list[IGLU_TLIST_MAX];
};

template <bool CLIP>
__global__ void __launch_bounds__(CTA_THREADS, 2) k_max_intersection(const int8_t *__restrict__ grids,
                                                                     const int8_t *__restrict__ targets,
                                                                     const int8_t *__restrict__ fulls,
                                                                     int invariant, int32_t *__restrict__ out,
                                                                     int32_t *__restrict__ argmax, int64_t count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *tbl = reinterpret_cast<unsigned long long *>(smem_raw);
    WarpScratch *scratch = reinterpret_cast<WarpScratch *>(smem_raw + SPREAD_N * sizeof(unsigned long long));
    PairScratch *pscratch = reinterpret_cast<PairScratch *>(scratch + WARPS_PER_CTA);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpScratch &ws = scratch[wid];
    uint16_t *list = pscratch[wid].list;
    build_spread_table(tbl, threadIdx.x, CTA_THREADS);
    __syncthreads();
    int8_t *sg = reinterpret_cast<int8_t *>(ws.grid);
    for (int64_t pr = (int64_t)blockIdx.x * WARPS_PER_CTA + wid; pr < count;
         pr += (int64_t)gridDim.x * WARPS_PER_CTA) {
        const int8_t *g = grids + pr * CELLS, *t = targets + pr * CELLS;
        const int8_t *f = fulls ? fulls + pr * CELLS : nullptr;
        int nout = 0, xmin = 10, xmax = 0, zmin = 10, zmax = 0;
        for (int c0 = 0; c0 < GSTRIDE; c0 += 32) {
            const int c = c0 + lane;
            int tv = 0, bv = 0;
            if (c < CELLS) {
                tv = t[c];
                bv = f ? (int)f[c] : tv;
                sg[c] = g[c];
            } else if (c < GSTRIDE) {
                sg[c] = 0;
            }
            const unsigned m = __ballot_sync(0xffffffffu, tv != 0);
            if (c < CELLS) {
                const int y = c / 121, r = c - 121 * y, x = r / 11, z = r - 11 * x;
                if (tv != 0) {
                    const int pos = nout + __popc(m & ((1u << lane) - 1u));
                    list[pos] = (uint16_t)(y | (x << 4) | (z << 8) | ((tv & 15) << 12));
                }
                if (bv != 0) { xmin = min(xmin, x); xmax = max(xmax, x); zmin = min(zmin, z); zmax = max(zmax, z); }
            }
            nout += __popc(m);
        }
        xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
        zmin = __reduce_min_sync(0xffffffffu, zmin); zmax = __reduce_max_sync(0xffffffffu, zmax);
        int8_t rc[16];
        int n_rot = 4;
        if (invariant) {
            rects_from_bbox(xmin, xmax, zmin, zmax, rc);
        } else {  // task.py:59-60 admissible = [[(0, 0)]]
#pragma unroll
            for (int k = 0; k < 16; k++) rc[k] = 0;
            n_rot = 1;
        }
        __syncwarp();
        build_rows(ws, false, lane);
        int key = -1;
        int best = eval_max_intersection<CLIP>(ws, tbl, list, nout, rc, n_rot, lane, argmax ? &key : nullptr);
        if (lane == 0) {
            out[pr] = best;
            if (argmax) {
                int k = 0, dx = 0, dz = 0;
                if (key >= 0) { k = key / 441; int r = key - k * 441; dx = r / 21 - 10; dz = r % 21 - 10; }
                argmax[3 * pr + 0] = dx; argmax[3 * pr + 1] = dz; argmax[3 * pr + 2] = k;
            }
        }
        __syncwarp();
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
    if (cfg->tlist_cap <= 0 || cfg->tlist_cap > IGLU_TLIST_MAX || (cfg->tlist_cap & 7)) return IGLU_ERR_INVALID_ARG;
    iglu_ctx *c = new iglu_ctx();
    c->cfg = *cfg;
    c->last_launches = 0;
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
    if (!rc) rc = check(cudaFuncSetAttribute(k_step<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    if (!rc) rc = check(cudaFuncSetAttribute(k_step<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    if (!rc) rc = check(cudaFuncSetAttribute(k_step<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    if (rc) { delete c; return rc; }
    *out = c;
    return IGLU_OK;
}

extern "C" void iglu_ctx_destroy(iglu_ctx *ctx) { delete ctx; }

extern "C" int iglu_last_launch_count(const iglu_ctx *ctx) { return ctx ? ctx->last_launches : 0; }

static bool state_ok(const iglu_state *st, const iglu_obs *obs) {
    return st && obs && st->n > 0 && st->grid && st->init && st->target && st->pose && st->init_pose &&
           st->istate && st->tlist && st->trect && st->tmeta && obs->reward && obs->done;
}

extern "C" int iglu_set_tasks(iglu_ctx *ctx, const iglu_state *st, const iglu_obs *obs, const int8_t *target,
                              const int8_t *start, const double *init_pose, const int32_t *env_idx,
                              int64_t count, void *stream) {
    if (!ctx || !state_ok(st, obs) || !target || count <= 0) return IGLU_ERR_INVALID_ARG;
    const int threads = 256;
    const int64_t blocks = (count * 32 + threads - 1) / threads;
    k_task_prep<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(*st, *obs, ctx->cfg, target, start,
                                                                         init_pose, env_idx, count);
    ctx->last_launches = 1;
    return check(cudaPeekAtLastError());
}

extern "C" int iglu_reset(iglu_ctx *ctx, const iglu_state *st, const iglu_obs *obs, const uint8_t *mask,
                          int hard, void *stream) {
    if (!ctx || !state_ok(st, obs)) return IGLU_ERR_INVALID_ARG;
    const int threads = 256;
    const int64_t blocks = (st->n * 32 + threads - 1) / threads;
    k_reset<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(*st, *obs, mask, hard);
    ctx->last_launches = 1;
    return check(cudaPeekAtLastError());
}

static unsigned step_grid(const iglu_ctx *ctx, int64_t n) {
    const int64_t groups = (n + 31) / 32;
    const int64_t ctas = (groups + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const int64_t resident = (int64_t)ctx->num_sms * 2;  // __launch_bounds__(256, 2): persistent CTAs
    return (unsigned)(ctas < resident ? ctas : resident);
}

template <int KIND>
static int launch_step(iglu_ctx *ctx, const StepParams &p, void *stream) {
    k_step<KIND><<<step_grid(ctx, p.st.n), CTA_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(p);
    ctx->last_launches = 1;
    return check(cudaPeekAtLastError());
}

extern "C" int iglu_step_walking(iglu_ctx *ctx, const iglu_state *st, const int32_t *actions,
                                 const iglu_obs *obs, void *stream) {
    if (!ctx || !state_ok(st, obs) || !actions) return IGLU_ERR_INVALID_ARG;
    if (ctx->cfg.action_space != IGLU_WALKING_DISCRETE) return IGLU_ERR_INVALID_ARG;
    StepParams p;
    memset(&p, 0, sizeof(p));
    p.st = *st; p.obs = *obs; p.cfg = ctx->cfg; p.actions = actions;
    return launch_step<IGLU_WALKING_DISCRETE>(ctx, p, stream);
}

extern "C" int iglu_step_flying(iglu_ctx *ctx, const iglu_state *st, const float *movement, const float *camera,
                                const int32_t *inventory, const int32_t *placement, const iglu_obs *obs,
                                void *stream) {
    if (!ctx || !state_ok(st, obs) || !movement || !camera || !inventory || !placement) return IGLU_ERR_INVALID_ARG;
    if (ctx->cfg.action_space != IGLU_FLYING) return IGLU_ERR_INVALID_ARG;
    StepParams p;
    memset(&p, 0, sizeof(p));
    p.st = *st; p.obs = *obs; p.cfg = ctx->cfg;
    p.movement = movement; p.camera = camera; p.inventory = inventory; p.placement = placement;
    return launch_step<IGLU_FLYING>(ctx, p, stream);
}

extern "C" int iglu_step_walking_dict(iglu_ctx *ctx, const iglu_state *st, const int32_t *buttons,
                                      const float *camera, const iglu_obs *obs, void *stream) {
    if (!ctx || !state_ok(st, obs) || !buttons || !camera) return IGLU_ERR_INVALID_ARG;
    if (ctx->cfg.action_space != IGLU_WALKING_DICT) return IGLU_ERR_INVALID_ARG;
    StepParams p;
    memset(&p, 0, sizeof(p));
    p.st = *st; p.obs = *obs; p.cfg = ctx->cfg; p.buttons = buttons; p.camera = camera;
    return launch_step<IGLU_WALKING_DICT>(ctx, p, stream);
}

extern "C" int iglu_step_walking_host(iglu_ctx *ctx, const iglu_state *st, const int32_t *actions_host,
                                      int32_t *actions_dev, const iglu_obs *obs, double *reward_host,
                                      uint8_t *done_host, void *stream) {
    if (!ctx || !state_ok(st, obs) || !actions_host || !actions_dev || !reward_host || !done_host)
        return IGLU_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = check(cudaMemcpyAsync(actions_dev, actions_host, sizeof(int32_t) * st->n, cudaMemcpyHostToDevice, s));
    if (rc) return rc;
    rc = iglu_step_walking(ctx, st, actions_dev, obs, stream);
    if (rc) return rc;
    rc = check(cudaMemcpyAsync(reward_host, obs->reward, sizeof(double) * st->n, cudaMemcpyDeviceToHost, s));
    if (rc) return rc;
    rc = check(cudaMemcpyAsync(done_host, obs->done, sizeof(uint8_t) * st->n, cudaMemcpyDeviceToHost, s));
    if (rc) return rc;
    return check(cudaStreamSynchronize(s));
}

extern "C" int iglu_max_intersection(const int8_t *grids, const int8_t *targets, const int8_t *fulls,
                                     int invariant, int32_t *out, int32_t *argmax, int64_t count,
                                     void *stream) {
    if (!grids || !targets || !out || count <= 0) return IGLU_ERR_INVALID_ARG;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return IGLU_ERR_CUDA;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return IGLU_ERR_CUDA;
    const size_t smem = SMEM_BYTES + WARPS_PER_CTA * sizeof(PairScratch);
    const int64_t ctas = (count + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const int64_t resident = (int64_t)sms * 2;
    const unsigned grid = (unsigned)(ctas < resident ? ctas : resident);
    int rc;
    if (fulls && invariant) {
        rc = check(cudaFuncSetAttribute(k_max_intersection<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (rc) return rc;
        k_max_intersection<true><<<grid, CTA_THREADS, smem, (cudaStream_t)stream>>>(grids, targets, fulls, invariant,
                                                                                    out, argmax, count);
    } else {
        rc = check(cudaFuncSetAttribute(k_max_intersection<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (rc) return rc;
        k_max_intersection<false><<<grid, CTA_THREADS, smem, (cudaStream_t)stream>>>(grids, targets, fulls, invariant,
                                                                                     out, argmax, count);
    }
    return check(cudaPeekAtLastError());
}
