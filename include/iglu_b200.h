/*
 * iglu_b200.h -- C ABI of the B200-native IGLU Gridworld step path.
 *
 * The reference (iglu-contest/gridworld) is pure Python and has no FFI: its boundary for
 * this path is the gym surface GridWorld.set_task / reset / step (reference
 * gridworld/env.py:155-166, 206-261, 268-303) which calls World.step
 * (gridworld/core/world.py:434-456) and Task.step_intersection / maximal_intersection
 * (gridworld/tasks/task.py:103-119, 147-161).  This header is the C boundary a maintainer
 * binds with ctypes directly under that surface; each entry point
 * names the reference interface it replaces.
 *
 * Conventions
 *  - Every pointer in iglu_state / iglu_obs and every array argument is DEVICE memory owned by
 *    the caller (PyTorch hands the tensors out) unless the name ends in _host.  The library
 *    allocates nothing persistent except the small iglu_ctx (config copy + trig tables); the only
 *    pointers it retains between calls are those of an attached iglu_task_pool.
 *  - All work is enqueued on the caller's stream (a cudaStream_t passed as void*); device-
 *    buffer entry points never synchronise, so they are CUDA-graph capturable.  *_host entry
 *    points move their host buffers on that stream (pinned buffers: directly from the kernels
 *    over PCIe; pageable ones: staged copies) and synchronise it before returning.
 *  - Return value: IGLU_OK (0) or a negative iglu_status; iglu_strerror() describes it.
 *  - Grids are int8, cell (y, x, z) of env i at  i * IGLU_GRID_STRIDE + y*121 + x*11 + z
 *    (dense index = [y+1, x+5, z+5] of the world block (x, y, z), reference env.py:136-142);
 *    IGLU_GRID_STRIDE = 1104 pads the 1089 cells to a 16-byte multiple (pad bytes are zero).
 *  - Arrays declared [k][n] are structure-of-arrays: component c of env i at  c * n + i.
 */
#ifndef IGLU_B200_H
#define IGLU_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IGLU_ABI_VERSION 1
#define IGLU_CELLS 1089       /* 9 * 11 * 11, reference tasks/task.py:3-5 */
#define IGLU_GRID_STRIDE 1104 /* bytes between consecutive envs in int8 grid state */
#define IGLU_TLIST_MAX 1096   /* max target blocks (uint16 entries) per env */
#define IGLU_TLIST_HDR 8      /* uint16 slots at the start of a tlist row holding the 16-byte rectangle header */
#define IGLU_TLIST_ROW_MAX (IGLU_TLIST_HDR + IGLU_TLIST_MAX) /* 1104 uint16 = 2208 bytes per row */

typedef enum iglu_status {
    IGLU_OK = 0,
    IGLU_ERR_INVALID_ARG = -1,
    IGLU_ERR_CUDA = -2,
    IGLU_ERR_NO_TASK = -3,   /* reference env.py:199-201, 271-275 ValueError('Task is not initialized! ...') */
    IGLU_ERR_UNSUPPORTED = -4
} iglu_status;

/* reference env.py:56-78 action spaces */
typedef enum iglu_action_space {
    IGLU_WALKING_DISCRETE = 0, /* Discrete(18), world.py:360-394 */
    IGLU_FLYING = 1,           /* Dict(movement, camera, inventory, placement), world.py:416-432 */
    IGLU_WALKING_DICT = 2      /* Dict of buttons + camera (discretize=False), world.py:396-414 */
} iglu_action_space;

typedef enum iglu_reward_mode {
    IGLU_REWARD_ON_CHANGE = 0, /* evaluate maximal_intersection only when the block count changed
                                  (the reference's gate, task.py:110-112) */
    IGLU_REWARD_EVERY_STEP = 1 /* evaluate it on every step (BASELINE config 2); the result is still
                                  committed only when the gate fires, so outputs are identical */
} iglu_reward_mode;

/* kwargs of the reference's create_env / GridWorld.__init__ (env.py:27-31, 333-338) that reach
 * the step path, plus the batched additions (auto_reset, reward_mode, write_grid_obs). */
typedef struct iglu_config {
    int32_t action_space;       /* iglu_action_space */
    int32_t select_and_place;   /* env.py:28 / world.py:444-446 */
    int32_t max_steps;          /* env.py:28, 292 */
    int32_t auto_reset;         /* 1: an env whose previous step returned done is reset (env.py:206-261)
                                   at the start of the next step call, then stepped */
    int32_t reward_mode;        /* iglu_reward_mode */
    int32_t write_grid_obs;     /* obs['grid'] int32[n,9,11,11] (env.py:288): 0 = not produced, 1 = rewritten in full
                                   every step, 2 = kept up to date incrementally (the cell a step edits + whole
                                   rows on reset); the buffer contents are identical for 1 and 2 */
    int32_t tlist_cap;          /* uint16 slots per row of iglu_state.tlist: multiple of 8, >= 72,
                                   <= IGLU_TLIST_ROW_MAX; holds IGLU_TLIST_HDR + max target blocks */
    int32_t reserved;           /* scheduling switches (results identical): bit 0 = 1 disables the observation-first /
                                   physics-first warp alternation of the fused launch; bit 1 = 1 forces the fused
                                   launch where the library would use two launches (physics, then observe);
                                   bit 2 = 1 forces the two-launch form; bit 3 = 1 deals the observe launch's env
                                   groups statically even when iglu_state.sched is given; bit 4 = 1 makes
                                   iglu_step_walking_host stage pinned buffers through copies too; bit 5 = 1 uses the
                                   group-wise observe launch where the library would use the tiled one; bit 6 = 1 turns
                                   the overlapped step (widen next to the physics, then commit) off, bit 7 = 1 uses
                                   it for small batches too, bit 15 = 1 gives its commit launch a 3-CTA grid (tests: many
                                   records per warp); bits 8-13:
                                   envs per group of the group-wise observe launch (0 = automatic) */
    double right_placement_scale; /* env.py:29, 296 */
    double wrong_placement_scale; /* env.py:29, 294 */
} iglu_config;

/* Per-env state resident in HBM (reference: Agent world.py:8-29, GridWorld env.py:32-55,
 * the synthetic Task env.py:227-232 / task.py:34-45). */
typedef struct iglu_state {
    int64_t n;
    int8_t *grid;      /* [n][IGLU_GRID_STRIDE]  env.grid */
    int8_t *init;      /* [n][IGLU_GRID_STRIDE]  dense starting grid (_synthetic_init_grid) */
    int8_t *target;    /* [n][IGLU_GRID_STRIDE]  task.target_grid */
    double *pose;      /* [6][n]  x, y, z, yaw, pitch, dy  (agent.position, rotation, dy) */
    double *init_pose; /* [5][n]  x, y, z, yaw, pitch      (initial_position / initial_rotation) */
    int32_t *istate;   /* [8][n]  w0..w2: inventory 6 x int16; w3: active | tis<<8 | done_prev<<16 | err<<24;
                                  w4: step_no; w5: nnz(grid - init) (task.prev_grid_size); w6: max_int;
                                  w7: wrong_placement of the last physics launch (task.py:111) */
    uint16_t *tlist;   /* [n][tlist_cap]  row = 16-byte header + block list of the synthetic task.
                                  header: 16 bytes, per rotation r: dx_lo (int8), number of dx values (bits 0-4),
                                  dz_lo (int8), number of dz values -- the admissible shifts (task.py:57-72;
                                  always a rectangle); bits 5-6 of byte 1 of rotation 0 hold the
                                  evaluation's lane-mapping shift (written by the library, opaque to callers);
                                  then the non-zero cells of (target - init), rotation 0, ascending cell index:
                                  y | x<<4 | z<<8 | (value & 15)<<12 */
    int32_t *tmeta;    /* [8][n]  w0: number of tlist entries; w1: target_size (task.py:35);
                                  w2..w4: inventory after reset, 6 x int16 (env.py:243-246);
                                  w5: 1 if the starting grid is non-empty; w6: 1 once a task is set; w7: reserved */
    int32_t *sched;    /* [8] launch-to-launch scratch, or NULL = groups of envs are dealt statically. Word 0: work
                                  counter of the observe launch; words 2-3 (one uint64): running count of the
                                  maximal_intersection evaluations the step calls performed (a caller may zero it);
                                  words 4-6: counters of the overlapped step (left at zero by every step) */
    int32_t *worklist; /* [8][n] scratch of the overlapped step (a 32-byte record per env a step edited or reset;
                                  16-byte aligned), or NULL = steps use the two-launch / fused forms only */
} iglu_state;

/* Observation / result buffers (reference env.py:283-303). They persist between steps and are
 * overwritten in place; a policy can read them directly. target_grid is written by
 * iglu_set_tasks (it never changes between steps). Any pointer may be NULL = not wanted,
 * except reward and done. */
typedef struct iglu_obs {
    int32_t *grid;        /* [n][1089] int32   obs['grid'] */
    int32_t *target_grid; /* [n][1089] int32   obs['target_grid'] (target_in_obs=True) */
    float *agent_pos;     /* [n][5]  x, y, z, pitch, yaw   obs['agentPos'] */
    float *inventory;     /* [n][6]                        obs['inventory'] */
    float *compass;       /* [n][1]  yaw - 180             obs['compass'] */
    double *reward;       /* [n]     env.py:293-296 (fp64 so that +-0.1 * k is exact as in Python) */
    uint8_t *done;        /* [n]     env.py:292 */
} iglu_obs;

typedef struct iglu_ctx iglu_ctx;

int iglu_abi_version(void);
const char *iglu_strerror(int status);

/* GridWorld.__init__ (env.py:27-115): copies the config, builds the libm trig tables. */
int iglu_ctx_create(const iglu_config *cfg, iglu_ctx **out);
void iglu_ctx_destroy(iglu_ctx *ctx);

/* GridWorld.set_task / initialize_world (env.py:155-193) + Task.__init__ of the synthetic task
 * (env.py:227-232, task.py:9-72) for `count` envs.
 *   target, start : device int8 [count][1089] dense (start may be NULL = empty starting grid [])
 *   init_pose     : device double [count][5] (x, y, z, yaw, pitch) or NULL = (0,0,0,0,0)
 *   env_idx       : device int32 [count] destination env of each row, or NULL = rows 0..count-1
 * Fills state.target/init/tlist/tmeta (+ obs.target_grid). Does NOT reset; call iglu_reset. */
int iglu_set_tasks(iglu_ctx *ctx, const iglu_state *st, const iglu_obs *obs, const int8_t *target,
                   const int8_t *start, const double *init_pose, const int32_t *env_idx, int64_t count,
                   void *stream);

/* Device-resident task generator (reference: GridWorld.set_task_generator env.py:166-172 and the
 * `self._task = self._task_generator.reset()` of GridWorld.reset env.py:213-216; the generator modelled is a
 * finite set of tasks drawn uniformly, i.e. RandomTasks(max_cache=m).reset task_set.py:120-127 or any
 * Tasks-like object whose tasks were enumerated into the pool). While a pool is attached, every reset of an
 * env -- iglu_reset and the auto-reset inside the step calls -- first draws a pool index j and copies
 * pool row j into the env's own target/init/tlist/tmeta rows (+ obs.target_grid), then resets as usual,
 * with no host involvement. The draw of env e for its k-th episode (k = cursor[1][e], then incremented) is
 *     x = seed + 0x9E3779B97F4A7C15 * (env_offset + e + 1) + 0xD1B54A32D192ED03 * k     (mod 2^64)
 *     x ^= x >> 30; x *= 0xBF58476D1CE4E5B9; x ^= x >> 27; x *= 0x94D049BB133111EB; x ^= x >> 31
 *     j = ((x >> 32) * m) >> 32
 * (the reference draws from numpy's global Mersenne Twister, which is neither per-env nor reproducible
 * across shards; env_offset makes sharded runs reproduce the single-device sequence). If next_task is not
 * NULL and 0 <= next_task[e] < m, that index is used instead and next_task[e] is set to -1 (consumed):
 * a caller can impose the exact sequence a reference generator produced. */
typedef struct iglu_task_pool {
    int64_t m;          /* tasks in the pool (0 = no pool) */
    int8_t *target;     /* [m][IGLU_GRID_STRIDE] */
    int8_t *init;       /* [m][IGLU_GRID_STRIDE]  dense starting grids */
    uint16_t *tlist;    /* [m][tlist_cap]         same row format as iglu_state.tlist */
    int32_t *tmeta;     /* [8][m]                 same words as iglu_state.tmeta */
    int32_t *cursor;    /* [2][n] per env: pool index of the current task (-1 before the first draw),
                                  number of draws so far */
    int32_t *next_task; /* [n] or NULL */
    uint64_t seed;
    int64_t env_offset; /* global index of env 0 of this shard */
} iglu_task_pool;

/* Task.__init__ of the synthetic task (env.py:227-232, task.py:9-72) for every pool entry:
 * target, start device int8 [m][1089] dense (start may be NULL) -> pool->target/init/tlist/tmeta. */
int iglu_task_pool_prepare(iglu_ctx *ctx, const iglu_task_pool *pool, const int8_t *target, const int8_t *start,
                           void *stream);

/* RandomTasks.sample_task (task_set.py:135-157) on the device: m random target structures, dense int8 [m][1089] in
 * `targets` (device) -- the input format of iglu_set_tasks / iglu_task_pool_prepare, so a pool can be refilled without
 * the host. Per height level one uniformly placed first block, then max_blocks - 1 more, each at an offset in
 * [-max_dist, max_dist]^2 from the FIRST block of the level, redrawn together with its colour until it is not (0, 0),
 * inside the zone and free -- the reference's control flow. Its random numbers come from numpy's global Mersenne
 * Twister; here the i-th draw (i = 0, 1, ...) of task t (global index first_task + row) with range r is
 *     x = seed + 0x9E3779B97F4A7C15 * (t + 1) + 0xD1B54A32D192ED03 * i   (mod 2^64), the finaliser of the pool draw,
 *     value = ((x >> 32) * r) >> 32
 * in the order block_x, block_z, colour, then (ddx, ddz, colour) per attempt; a block that finds no free cell
 * within 4096 attempts is skipped (the reference would loop forever). */
typedef struct iglu_random_task_spec {
    int32_t max_blocks;     /* blocks per height level (task_set.py:75) */
    int32_t height_levels;  /* 1..9 */
    int32_t max_dist;       /* Chebyshev distance from the level's first block */
    int32_t num_colors;     /* 1..6 */
} iglu_random_task_spec;

int iglu_sample_tasks(int8_t *targets, int64_t m, const iglu_random_task_spec *spec, uint64_t seed,
                      int64_t first_task, void *stream);

/* Multi-turn structures with progressive goals -- reference Subtasks (tasks/task.py:208-308): structure_seq[i] is the
 * structure after turn i; a task (turn_start, turn_goal) starts from structure_seq[turn_start] and has target
 * structure_seq[min(turn_goal, L - 1)] with full_grid = structure_seq[L - 1] (create_task, 256-285). GridWorld.step
 * only ever scores the episode's synthetic task (env.py:291); Subtasks.step_intersection (288-298) is what a
 * training loop calls on the env's grid to get the staged (right, wrong, done), and when the current task is
 * complete it moves the goal on to the next turn. iglu_staged_step does that for all n envs after a step call.
 * Stage rows: one per turn of every structure plus, per structure, one more row that repeats the last turn (the
 * reference advances once past the end, len(structure_seq) > task_goal, and min() then repeats the last target). */
typedef struct iglu_staged_tasks {
    int64_t m;            /* stage rows */
    uint16_t *tlist;      /* [m][tlist_cap]  same row format as iglu_state.tlist: rectangles from the FULL structure,
                                             then the non-zero cells of the stage's target */
    int32_t *tmeta;       /* [8][m]          w0: list entries, w1: target_size */
    int32_t *next_stage;  /* [m]             row of the next turn of the same structure, -1 after the repeated row */
    int32_t *stage_of;    /* [rows of the task pool] goal stage row of task row r (its turn_goal), or NULL = identity */
    int32_t progressive;  /* Subtasks(progressive=...) */
    int32_t reserved;
    int32_t *state;       /* [4][n] per env: goal stage row, prev_grid_size, max_int (of the current staged task), - */
    int32_t *out;         /* [4][n] per env and step: right_placement, wrong_placement, done, goal stage row */
} iglu_staged_tasks;

/* Task.__init__ of every stage row (task.py:9-72 with full_grid): target, full device int8 [m][1089] dense. */
int iglu_staged_prepare(iglu_ctx *ctx, const iglu_staged_tasks *staged, const int8_t *target, const int8_t *full,
                        void *stream);

/* Subtasks.step_intersection(env.grid) for all n envs; call after a step call. An env whose step_no is 1 starts a new
 * episode: its staged task is (re)created first -- Subtasks.reset() inside GridWorld.reset (env.py:219) -- with goal
 * stage_of[task_row[e]] (task_row NULL = e; with a task pool pass iglu_task_pool.cursor, whose first n words are the
 * pool row of each env's current task). */
int iglu_staged_step(iglu_ctx *ctx, const iglu_state *st, const iglu_staged_tasks *staged, const int32_t *task_row,
                     void *stream);

/* Attach (pool != NULL, m > 0) or detach (NULL) the generator. The ctx keeps a COPY of the struct, i.e. it
 * retains the device pointers until detached or destroyed; the caller keeps the buffers alive. */
int iglu_task_pool_attach(iglu_ctx *ctx, const iglu_task_pool *pool);

/* GridWorld.reset (env.py:206-261) for the envs with mask[i] != 0 (mask NULL = all): grid <- start,
 * pose <- initial pose, inventory, step_no, task bookkeeping; writes the reset observation
 * (agentPos and compass are zeros there, env.py:249-254). agent.dy / time_int_steps / active_block
 * are NOT reset (reference behaviour); pass hard != 0 to also reset those (fresh-env state). */
int iglu_reset(iglu_ctx *ctx, const iglu_state *st, const iglu_obs *obs, const uint8_t *mask, int hard,
               void *stream);

/* GridWorld.step (env.py:268-303) for all n envs, Discrete(18) actions (device int32 [n]). */
int iglu_step_walking(iglu_ctx *ctx, const iglu_state *st, const int32_t *actions, const iglu_obs *obs,
                      void *stream);

/* GridWorld.step with the flying Dict action (env.py:71-78): movement f32 [n][3], camera f32 [n][2],
 * inventory int32 [n] (0 = none), placement int32 [n] (0 none, 1 place, 2 break). */
int iglu_step_flying(iglu_ctx *ctx, const iglu_state *st, const float *movement, const float *camera,
                     const int32_t *inventory, const int32_t *placement, const iglu_obs *obs, void *stream);

/* GridWorld.step with the walking Dict action (env.py:60-70, discretize=False): buttons int32 [n][8] =
 * forward, back, left, right, jump, attack, use, hotbar; camera f32 [n][2]. */
int iglu_step_walking_dict(iglu_ctx *ctx, const iglu_state *st, const int32_t *buttons, const float *camera,
                           const iglu_obs *obs, void *stream);

/* Host-buffer form of iglu_step_walking: actions_host int32 [n] in, reward_host f64 [n] / done_host u8 [n]
 * out, stream synchronised on return. This is the call bench.py's e2e number times.
 * Pinned (page-locked) buffers are accessed by the kernels directly over PCIe -- the physics launch reads
 * the actions from host memory and the observe launch stores reward/done to host memory (as well as to
 * obs->reward / obs->done) -- so no copy is serialised around the step; pageable buffers are staged through
 * `actions_dev` (device int32 [n] scratch) and copied back after the step. config.reserved bit 4 = 1 forces
 * the staged route. */
int iglu_step_walking_host(iglu_ctx *ctx, const iglu_state *st, const int32_t *actions_host,
                           int32_t *actions_dev, const iglu_obs *obs, double *reward_host, uint8_t *done_host,
                           void *stream);

/* The same call without the final synchronisation: it returns as soon as the work is enqueued, so a host loop
 * can keep a second env group in flight while it consumes the results of the first (EnvPool-style pipelining;
 * reference loop: examples/run_env.py:16-27). reward_host / done_host are valid after iglu_wait(stream). */
int iglu_step_walking_host_async(iglu_ctx *ctx, const iglu_state *st, const int32_t *actions_host,
                                 int32_t *actions_dev, const iglu_obs *obs, double *reward_host,
                                 uint8_t *done_host, void *stream);

/* Blocks the calling host thread until everything enqueued on `stream` has finished. */
int iglu_wait(void *stream);

/* Everything GridWorld.step / reset hand back for ONE env (reference env.py:247-261, 283-303) plus the agent
 * state its attributes expose, gathered by one small kernel into `scratch_dev` (device, sizeof(iglu_env_snapshot))
 * and brought to `out_host` with one copy; the stream is synchronised on return. This is what the N = 1 gym
 * drop-in (gridworld_b200/env.py) reads after every step instead of one transfer per observation key. */
typedef struct iglu_env_snapshot {
    double pose[6];          /* x, y, z, yaw, pitch, dy */
    double reward;
    float agent_pos[5];      /* obs['agentPos'] */
    float inventory_obs[6];  /* obs['inventory'] */
    float compass;           /* obs['compass'] */
    int32_t inventory[6];    /* agent.inventory */
    int32_t active_block, time_int_steps, error, step_no, grid_size, max_int, target_size, done;
    int32_t grid[IGLU_CELLS];        /* obs['grid'] */
    int32_t target_grid[IGLU_CELLS]; /* obs['target_grid'] */
} iglu_env_snapshot;

int iglu_fetch_env_host(iglu_ctx *ctx, const iglu_state *st, const iglu_obs *obs, int64_t env,
                        iglu_env_snapshot *scratch_dev, iglu_env_snapshot *out_host, void *stream);

/* Stateless Task(target, full_grid=full, invariant=invariant).maximal_intersection(grid)
 * (task.py:9-72, 147-161) for `count` independent pairs (BASELINE config 4).
 *   grids, targets, fulls : device int8 [count][1089] dense (fulls may be NULL)
 *   out                   : device int32 [count]
 *   argmax                : device int32 [count][3] = (dx, dz, rot) of task.py:121-136, or NULL */
int iglu_max_intersection(const int8_t *grids, const int8_t *targets, const int8_t *fulls, int invariant,
                          int32_t *out, int32_t *argmax, int64_t count, void *stream);

/* The same for rows that are not densely packed: row i of `grids` starts at grids + i * grid_stride bytes (stride >=
 * 1089), likewise targets / fulls. Rows that are 16-byte aligned with a stride that is a multiple of 16 and >= 1104 --
 * the layout of iglu_state.grid / target / init -- are staged as they are (no repack of env state is needed to
 * evaluate it); bytes of a row past cell 1088 are ignored. iglu_max_intersection == strides of 1089. */
int iglu_max_intersection_strided(const int8_t *grids, int64_t grid_stride, const int8_t *targets,
                                  int64_t target_stride, const int8_t *fulls, int64_t full_stride, int invariant,
                                  int32_t *out, int32_t *argmax, int64_t count, void *stream);

/* obs['pov'] (reference: gridworld/render.py set_3d 94-113 / add_block 159-170, env.py:257-258, 299-300: the RGB
 * frame of the OpenGL renderer, 64 x 64 by default): a voxel ray caster over every env's grid and pose -- 90 degree
 * vertical field of view, near 0.1 / far 30, cubes seen from outside only (GL_CULL_FACE), top-only floor, nearest texel,
 * clear colour (0.5, 0.69, 1.0); the restatement is spelled out in csrc/iglu_render.cuh. pov: uint8 [n][height][width][3]
 * on the device, rows top-down. atlas_rgba: the 128 x 128 RGBA texture atlas (gridworld/texture.png as an array
 * [128][128][4], rows top-down) on the device, or NULL = flat colours per block id. Enqueues one kernel. */
int iglu_render_pov(iglu_ctx *ctx, const iglu_state *st, uint8_t *pov, int height, int width, const uint8_t *atlas_rgba,
                    void *stream);

/* Number of kernels the last step call on this ctx launched (bench.py's gpu_launches claim). */
int iglu_last_launch_count(const iglu_ctx *ctx);

/* Tuning aid, no reference counterpart: with IGLU_B200_TRACE=1 in the environment when the ctx was created, the
 * overlapped step records timing events around its launches; this call waits for the last step and returns, in
 * milliseconds after the start of the widen launch: out5[0..2] = end of widen + physics (the physics launch is the
 * widen launch's programmatic dependent: nothing may be recorded between them), out5[3] = commit start, [4] = commit end. */
int iglu_debug_overlap_trace(iglu_ctx *ctx, float *out5);

#ifdef __cplusplus
}
#endif
#endif /* IGLU_B200_H */
