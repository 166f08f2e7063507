/*
 * iglu_oracle.c -- CPU restatement of the IGLU Gridworld step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle ("port") for the CUDA
 * path in gridworld_b200/csrc.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it; the product path
 * never does (it fails loudly when the CUDA library is missing).
 *
 * Parity status: the reference (iglu-contest/gridworld) has NO tests of its own, so
 * this oracle is pinned against the EXECUTED reference instead: oracle/make_golden.py
 * imports the unmodified Python reference from /root/reference and records per-step
 * traces + reward known answers into the .npz files under tests/golden; tests/test_oracle_golden.py
 * replays them through this file bit-for-bit (fp64 pose compared exactly).
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference root).  Arithmetic is fp64 in the reference's operation order, libm
 * sin/cos/atan2 (CPython's math module calls the same libm), round-half-even
 * (numba round == rint), no FMA contraction (build with -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GY 9
#define GX 11
#define GZ 11
#define CELLS (GY * GX * GZ) /* tasks/task.py:3-5 BUILD_ZONE_SIZE = 9, 11, 11 */

/* utils.py:9-24 */
#define WALKING_SPEED 5
#define FLYING_SPEED 15
#define GRAVITY 20.0
#define TERMINAL_VELOCITY 50
#define PLAYER_HEIGHT 2
#define AGENT_PAD 0.25 /* core/world.py:9 */
/* utils.py:125-127 */
#define WHITE (-1)
#define GREY 0

typedef struct {
    /* tasks/task.py:9-72 Task fields the path needs */
    int32_t grids[4][CELLS]; /* target_grids: four y-rotations */
    int n_rot;               /* 4 if invariant else 1 (task.py:59-60) */
    int n_adm[4];
    int8_t adm[4][441][2];   /* admissible (dx, dz) per rotation, reference order */
    int target_size;
    int max_int, prev_grid_size, right_placement, wrong_placement;
} orc_task;

typedef struct {
    /* core/world.py:8-29 Agent */
    int flying;
    double strafe[2];
    double pos[3];
    double rot[2]; /* (yaw, pitch) */
    double dy;
    int time_int_steps;
    int inventory[6];
    int active_block;
    /* env.py:26-55 GridWorld */
    int32_t grid[CELLS];      /* env.grid */
    int32_t init[CELLS];      /* _synthetic_init_grid = to_dense(starting_grid) */
    int32_t target[CELLS];    /* _task.target_grid */
    orc_task syn;             /* _synthetic_task */
    int step_no;
    double init_pose[5];      /* initial_position(3) + initial_rotation(2) */
    int done_prev;            /* batched auto-reset bookkeeping (not in the reference) */
    int err;                  /* 1 = bad inventory id seen (world.py:354-355 raises) */
} orc_env;

typedef struct {
    int n;
    int flying, select_and_place, max_steps;
    double right_scale, wrong_scale;
    orc_env *e;
    /* task generator (env.py:166-172, 213-216): a finite pool drawn at every reset */
    int64_t pool_m;
    int8_t *pool_target, *pool_start; /* [m][CELLS] (pool_start may be NULL) */
    uint64_t pool_seed;
    int64_t pool_off;
    int32_t *next_task; /* [n] imposed index for the next reset, -1 = draw */
    int32_t *task_id;   /* [n] current pool index */
    int32_t *episode;   /* [n] draws so far */
} orc_batch;

static double JUMP_SPEED; /* utils.py:21 sqrt(2*GRAVITY*MAX_JUMP_HEIGHT) */

/* CPython mathmodule.c: radians(x) = x * (pi/180), degrees(x) = x * (180/pi) */
static inline double py_radians(double x) { return x * (M_PI / 180.0); }
static inline double py_degrees(double x) { return x * (180.0 / M_PI); }

/* utils.py:57-73 normalize(): int(round(v)); numba round is round-half-even */
static inline long norm1(double v) { return (long)rint(v); }

/* ------------------------------------------------------------------ Task ---- */

/* tasks/task.py:9-72 Task.__init__ (literal: rotations by the triple loop, admissible
 * shifts by counting the non-zero cells of every slice). `full` may be NULL. */
void orc_task_init(orc_task *t, const int32_t *target, const int32_t *full, int invariant) {
    memset(t, 0, sizeof(*t));
    int tsize = 0, fsize;
    for (int i = 0; i < CELLS; i++) tsize += target[i] != 0;
    fsize = tsize;
    if (full) {
        fsize = 0;
        for (int i = 0; i < CELLS; i++) fsize += full[i] != 0;
    }
    t->target_size = tsize;
    memcpy(t->grids[0], target, sizeof(int32_t) * CELLS);
    int32_t fg[4][CELLS];
    if (full) memcpy(fg[0], full, sizeof(int32_t) * CELLS);
    for (int r = 1; r < 4; r++) { /* task.py:47-56 */
        for (int y = 0; y < GY; y++)
            for (int x = 0; x < GX; x++)
                for (int z = 0; z < GZ; z++) {
                    t->grids[r][y * 121 + z * 11 + (GX - x - 1)] = t->grids[r - 1][y * 121 + x * 11 + z];
                    if (full) fg[r][y * 121 + z * 11 + (GX - x - 1)] = fg[r - 1][y * 121 + x * 11 + z];
                }
    }
    if (!invariant) { /* task.py:59-60 */
        t->n_rot = 1;
        t->n_adm[0] = 1;
        t->adm[0][0][0] = 0;
        t->adm[0][0][1] = 0;
        return;
    }
    t->n_rot = 4;
    for (int r = 0; r < 4; r++) { /* task.py:62-72 */
        const int32_t *g = full ? fg[r] : t->grids[r];
        int n = 0;
        for (int dx = -GX + 1; dx < GX; dx++)
            for (int dz = -GZ + 1; dz < GZ; dz++) {
                int x0 = dx > 0 ? dx : 0, x1 = GX + (dx < 0 ? dx : 0);
                int z0 = dz > 0 ? dz : 0, z1 = GZ + (dz < 0 ? dz : 0);
                int cnt = 0;
                for (int y = 0; y < GY; y++)
                    for (int x = x0; x < x1; x++)
                        for (int z = z0; z < z1; z++) cnt += g[y * 121 + x * 11 + z] != 0;
                if (cnt == fsize) {
                    t->adm[r][n][0] = (int8_t)dx;
                    t->adm[r][n][1] = (int8_t)dz;
                    n++;
                }
            }
        t->n_adm[r] = n;
    }
}

/* tasks/task.py:138-145 get_intersection */
static int task_intersection(const orc_task *t, const int32_t *grid, int dx, int dz, int r) {
    int tx0 = dx > 0 ? dx : 0, tx1 = GX + (dx < 0 ? dx : 0);
    int tz0 = dz > 0 ? dz : 0, tz1 = GZ + (dz < 0 ? dz : 0);
    int gx0 = -dx > 0 ? -dx : 0, gz0 = -dz > 0 ? -dz : 0;
    const int32_t *tg = t->grids[r];
    int cnt = 0;
    for (int y = 0; y < GY; y++)
        for (int i = 0; i < tx1 - tx0; i++)
            for (int j = 0; j < tz1 - tz0; j++) {
                int32_t tv = tg[y * 121 + (tx0 + i) * 11 + (tz0 + j)];
                int32_t gv = grid[y * 121 + (gx0 + i) * 11 + (gz0 + j)];
                cnt += (tv == gv) & (tv != 0);
            }
    return cnt;
}

/* tasks/task.py:147-161 maximal_intersection; if arg != NULL also returns
 * tasks/task.py:121-136 argmax_intersection as (dx, dz, rot). */
int orc_task_max_intersection(const orc_task *t, const int32_t *grid, int *arg) {
    int best = 0;
    if (arg) arg[0] = arg[1] = arg[2] = 0;
    for (int r = 0; r < t->n_rot; r++)
        for (int k = 0; k < t->n_adm[r]; k++) {
            int v = task_intersection(t, grid, t->adm[r][k][0], t->adm[r][k][1], r);
            if (v > best) {
                best = v;
                if (arg) { arg[0] = t->adm[r][k][0]; arg[1] = t->adm[r][k][1]; arg[2] = r; }
            }
        }
    return best;
}

/* tasks/task.py:103-119 step_intersection */
static void task_step_intersection(orc_task *t, const int32_t *grid, int *right, int *wrong, int *done) {
    int size = 0;
    for (int i = 0; i < CELLS; i++) size += grid[i] != 0;
    int wr = t->prev_grid_size - size;
    int mi = wr != 0 ? orc_task_max_intersection(t, grid, NULL) : t->max_int;
    *done = mi == t->target_size;
    t->prev_grid_size = size;
    *right = mi - t->max_int;
    t->max_int = mi;
    t->right_placement = *right;
    t->wrong_placement = wr;
    *wrong = wr;
}

/* ----------------------------------------------------------------- World ---- */

/* core/world.py:57-58 build_zone */
static inline int build_zone_i(long x, long y, long z, int pad) {
    return -5 - pad <= x && x <= 5 + pad && -5 - pad <= z && z <= 5 + pad && -1 - pad <= y && y < 8 + pad;
}
static inline int build_zone_d(double x, double y, double z, int pad) {
    return -5 - pad <= x && x <= 5 + pad && -5 - pad <= z && z <= 5 + pad && -1 - pad <= y && y < 8 + pad;
}

/* `key in self.world` / `self.world[key]`: the dict holds the 37x37 floor at y=-2
 * (core/world.py:60-71; WHITE inside the zone footprint, GREY outside) plus every
 * placed block, which env._add_block mirrors into grid[y+1, x+5, z+5] (env.py:136-142).
 * Blocks are only ever placed inside the build zone (world.py:318). Returns 1 and the
 * texture if present. */
static inline int world_get(const orc_env *e, long x, long y, long z, int *tex) {
    if (y == -2) {
        if (labs(x) <= 18 && labs(z) <= 18) {
            *tex = build_zone_i(x, 0, z, 0) ? WHITE : GREY;
            return 1;
        }
        return 0;
    }
    if (build_zone_i(x, y, z, 0)) {
        int32_t c = e->grid[(y + 1) * 121 + (x + 5) * 11 + (z + 5)];
        if (c != 0) { *tex = c; return 1; }
    }
    return 0;
}

/* core/world.py:73-99 hit_test. Returns 1 on hit; *has_prev = 0 when previous is None. */
static int hit_test(const orc_env *e, const double pos[3], const double vec[3], long block[3], long prev[3],
                    int *has_prev) {
    const int m = 5, max_distance = 8;
    double x = pos[0], y = pos[1], z = pos[2];
    double dx = vec[0], dy = vec[1], dz = vec[2];
    long p[3] = {0, 0, 0};
    int hp = 0, tex;
    for (int i = 0; i < max_distance * m; i++) {
        long k[3] = {norm1(x), norm1(y), norm1(z)};
        int differs = !hp || k[0] != p[0] || k[1] != p[1] || k[2] != p[2];
        if (differs && world_get(e, k[0], k[1], k[2], &tex)) {
            memcpy(block, k, sizeof(k));
            memcpy(prev, p, sizeof(p));
            *has_prev = hp;
            return 1;
        }
        memcpy(p, k, sizeof(k));
        hp = 1;
        x = x + dx / m; y = y + dy / m; z = z + dz / m;
    }
    return 0;
}

/* core/world.py:145-161 get_sight_vector */
static void sight_vector(const orc_env *e, double v[3]) {
    double x = e->rot[0], y = e->rot[1];
    double m = cos(py_radians(y));
    v[1] = sin(py_radians(y));
    v[0] = cos(py_radians(x - 90)) * m;
    v[2] = sin(py_radians(x - 90)) * m;
}

/* core/world.py:163-201 get_motion_vector */
static void motion_vector(const orc_env *e, double v[3]) {
    if (e->strafe[0] != 0 || e->strafe[1] != 0) {
        double x = e->rot[0], y = e->rot[1];
        double strafe = py_degrees(atan2(e->strafe[0], e->strafe[1]));
        double y_angle = py_radians(y);
        double x_angle = py_radians(x + strafe);
        if (e->flying) {
            double m = cos(y_angle);
            double dy = sin(y_angle);
            if (e->strafe[1] != 0) { dy = 0.0; m = 1; }
            if (e->strafe[0] > 0) dy *= -1;
            v[0] = cos(x_angle) * m;
            v[1] = dy;
            v[2] = sin(x_angle) * m;
        } else {
            v[1] = 0.0;
            v[0] = cos(x_angle);
            v[2] = sin(x_angle);
        }
    } else {
        v[0] = v[1] = v[2] = 0.0;
    }
}

/* core/world.py:264-310 collide; utils.py:156-163 FACES order */
static void collide(orc_env *e, double p[3]) {
    static const int FACES[6][3] = {{0, 1, 0}, {0, -1, 0}, {-1, 0, 0}, {1, 0, 0}, {0, 0, 1}, {0, 0, -1}};
    const double pad = AGENT_PAD;
    long np[3] = {norm1(p[0]), norm1(p[1]), norm1(p[2])};
    int tex;
    for (int f = 0; f < 6; f++)
        for (int i = 0; i < 3; i++) {
            if (!FACES[f][i]) continue;
            double d = (p[i] - np[i]) * FACES[f][i];
            if (d < pad) continue;
            for (int h = 0; h < PLAYER_HEIGHT; h++) {
                long op[3] = {np[0], np[1], np[2]};
                op[1] -= h;
                op[i] += FACES[f][i];
                if (!world_get(e, op[0], op[1], op[2], &tex)) continue;
                p[i] -= (d - pad) * FACES[f][i];
                if (f < 2) e->dy = 0; /* ground or ceiling */
                break;
            }
        }
}

/* core/world.py:222-262 _update */
static void update_sub(orc_env *e, double dt) {
    int speed = e->flying ? FLYING_SPEED : WALKING_SPEED;
    double d = dt * speed;
    double mv[3];
    motion_vector(e, mv);
    double dx = mv[0] * d, dy = mv[1] * d, dz = mv[2] * d;
    if (!e->flying) {
        e->dy -= dt * GRAVITY;
        if (e->dy < -14) e->time_int_steps = 12;
        else if (e->dy < -10) e->time_int_steps = 8;
        else if (e->dy < -5) e->time_int_steps = 4;
        else e->time_int_steps = 2;
        e->dy = e->dy > -TERMINAL_VELOCITY ? e->dy : -TERMINAL_VELOCITY; /* max(dy, -50) */
    }
    dy += e->dy * dt;
    double x = e->pos[0], y = e->pos[1], z = e->pos[2];
    double cand[3] = {x + dx, y + dy, z + dz};
    if (build_zone_d(cand[0], cand[1], cand[2], 2)) {
        collide(e, cand);
        memcpy(e->pos, cand, sizeof(cand));
    } else if (!e->flying) {
        double c2[3] = {x, y + dy, z};
        collide(e, c2);
        memcpy(e->pos, c2, sizeof(c2));
    }
}

/* core/world.py:203-220 update */
static void update(orc_env *e, double dt) {
    int m = e->time_int_steps;
    dt = dt < 0.2 ? dt : 0.2;
    for (int i = 0; i < m; i++) update_sub(e, dt / m);
    e->strafe[0] = e->strafe[1] = 0; /* sustain=False (env.py:32) */
    if (e->flying) e->dy = 0;
}

/* core/world.py:312-332 place_or_remove_block (+ add_block/remove_block 101-141 and the
 * env callbacks env.py:136-153) */
static void place_or_remove(orc_env *e, int remove, int place) {
    if ((place && remove) || (!place && !remove)) return;
    double vec[3];
    long block[3], prev[3];
    int has_prev = 0;
    sight_vector(e, vec);
    int hit = hit_test(e, e->pos, vec, block, prev, &has_prev);
    if (place) {
        if (hit && has_prev) {
            if (e->inventory[e->active_block - 1] > 0 && build_zone_i(prev[0], prev[1], prev[2], 0)) {
                double x = e->pos[0], y = e->pos[1], z = e->pos[2];
                y = y - (PLAYER_HEIGHT - 1) + AGENT_PAD;
                double bx = prev[0], by = prev[1], bz = prev[2];
                bx -= 0.5;
                bz -= 0.5;
                int inside = bx <= x && x <= bx + 1 && bz <= z && z <= bz + 1 &&
                             ((by <= y && y <= by + 1) || (by <= (y + 1) && (y + 1) <= by + 1));
                if (!inside) {
                    e->grid[(prev[1] + 1) * 121 + (prev[0] + 5) * 11 + (prev[2] + 5)] = e->active_block;
                    e->inventory[e->active_block - 1] -= 1;
                }
            }
        }
    }
    if (remove && hit) {
        int tex = 0;
        world_get(e, block[0], block[1], block[2], &tex);
        if (tex != GREY && tex != WHITE) {
            e->grid[(block[1] + 1) * 121 + (block[0] + 5) * 11 + (block[2] + 5)] = 0;
            e->inventory[tex - 1] += 1;
        }
    }
}

/* Parsed action = the tuple every parse_* returns (core/world.py:394,414,432). */
typedef struct {
    double strafe[2], dy, camera[2];
    int inventory; /* 0 = None */
    int remove, add;
} orc_action;

/* core/world.py:360-394 parse_walking_discrete_action */
static void parse_walking_discrete(int a, orc_action *o) {
    memset(o, 0, sizeof(*o));
    if (a == 1) o->strafe[0] += -1;
    else if (a == 2) o->strafe[0] += 1;
    else if (a == 3) o->strafe[1] += -1;
    else if (a == 4) o->strafe[1] += 1;
    else if (a == 5) o->dy = 1;
    else if (6 <= a && a <= 11) o->inventory = a - 5;
    else if (a == 12) o->camera[0] = -5;
    else if (a == 13) o->camera[0] = 5;
    else if (a == 14) o->camera[1] = -5;
    else if (a == 15) o->camera[1] = 5;
    else if (a == 16) o->remove = 1;
    else if (a == 17) o->add = 1;
}

/* core/world.py:434-456 World.step + 338-356 move_camera/movement */
static void world_step(orc_env *e, orc_action *a, int select_and_place) {
    if (select_and_place && a->inventory != 0) { a->add = 1; a->remove = 0; }
    /* movement (344-356) */
    e->strafe[0] += a->strafe[0];
    e->strafe[1] += a->strafe[1];
    if (a->dy != 0 && e->dy == 0) e->dy = JUMP_SPEED * a->dy;
    if (e->flying && a->dy == 0) e->dy = 0;
    if (a->inventory != 0) {
        if (a->inventory < 1 || a->inventory > 6) e->err = 1; /* reference raises ValueError */
        else e->active_block = a->inventory;
    }
    /* move_camera (338-342) */
    double yaw = e->rot[0] + a->camera[0], pitch = e->rot[1] + a->camera[1];
    double mn = pitch < 90 ? pitch : 90; /* min(90, y) */
    pitch = mn > -90 ? mn : -90;          /* max(-90, .) */
    e->rot[0] = yaw;
    e->rot[1] = pitch;
    place_or_remove(e, a->remove, a->add);
    update(e, 1 / 20.);
    yaw = e->rot[0];
    while (yaw > 360.) yaw -= 360.;
    while (yaw < 0.0) yaw += 360.0;
    e->rot[0] = yaw;
}

/* ------------------------------------------------------------------- Env ---- */

/* env.py:206-261 GridWorld.reset (render=False). The synthetic task is rebuilt by
 * orc_set_task because target/start do not change between resets of one task. */
static void env_reset(orc_env *e) {
    e->step_no = 0;
    e->syn.max_int = 0; /* _synthetic_task.reset() with starting_grid None (task.py:79-85) */
    e->syn.prev_grid_size = 0;
    e->syn.right_placement = e->syn.wrong_placement = 0;
    memcpy(e->grid, e->init, sizeof(e->grid)); /* env.py:234-238 */
    memcpy(e->pos, e->init_pose, 3 * sizeof(double));
    memcpy(e->rot, e->init_pose + 3, 2 * sizeof(double));
    for (int c = 0; c < 6; c++) e->inventory[c] = 20; /* env.py:243-246 */
    for (int i = 0; i < CELLS; i++) {
        int c = e->init[i];
        if (c >= 1 && c <= 6) e->inventory[c - 1] -= 1;
    }
    e->done_prev = 0;
    /* NOT reset (reference quirk): dy, time_int_steps, active_block, strafe */
}

/* env.set_task (env.py:155-165): store target / starting grid, build the synthetic task (env.py:227-232). */
static void env_assign_task(orc_env *e, const int8_t *target, const int8_t *start) {
    int32_t diff[CELLS];
    for (int k = 0; k < CELLS; k++) {
        e->target[k] = target[k];
        e->init[k] = start ? start[k] : 0;
        diff[k] = e->target[k] - e->init[k]; /* always invariant, no full_grid */
    }
    orc_task_init(&e->syn, diff, NULL, 1);
}

/* The batched build's generator draw (include/iglu_b200.h iglu_task_pool). The reference draws with
 * np.random.choice from numpy's global stream (task_set.py:120-127), which has no batched meaning; the
 * draw is therefore part of the new interface and is restated here from its published formula only. */
static int64_t pool_draw(uint64_t seed, int64_t env, uint32_t episode, int64_t m) {
    uint64_t x = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(env + 1) + 0xD1B54A32D192ED03ull * (uint64_t)episode;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (int64_t)(((x >> 32) * (uint64_t)m) >> 32);
}

/* GridWorld.reset with a task generator: `self._task = self._task_generator.reset()` first (env.py:213-216) */
static void batch_reset_env(orc_batch *b, int64_t i) {
    orc_env *e = &b->e[i];
    if (b->pool_m > 0) {
        int64_t j = -1;
        if (b->next_task[i] >= 0 && b->next_task[i] < b->pool_m) { j = b->next_task[i]; b->next_task[i] = -1; }
        if (j < 0) j = pool_draw(b->pool_seed, b->pool_off + i, (uint32_t)b->episode[i], b->pool_m);
        b->task_id[i] = (int32_t)j;
        b->episode[i] += 1;
        env_assign_task(e, b->pool_target + j * CELLS, b->pool_start ? b->pool_start + j * CELLS : NULL);
    }
    env_reset(e);
}

/* env.py:268-303 GridWorld.step for one env with an already parsed action */
static void env_step(const orc_batch *b, orc_env *e, orc_action *a, double *reward, uint8_t *done) {
    e->step_no += 1;
    world_step(e, a, b->select_and_place);
    int32_t diff[CELLS];
    for (int i = 0; i < CELLS; i++) diff[i] = e->grid[i] - e->init[i]; /* env.py:290 */
    int right, wrong, dn;
    task_step_intersection(&e->syn, diff, &right, &wrong, &dn);
    dn = dn || (e->step_no == b->max_steps);
    *reward = right == 0 ? wrong * b->wrong_scale : right * b->right_scale; /* env.py:293-296 */
    *done = (uint8_t)dn;
    e->done_prev = dn;
}


/* ------------------------------------------------ pthread parallel-for ---- */
/* (libgomp is not installed in this image, so no OpenMP.) */
#include <pthread.h>
typedef void (*orc_range_fn)(void *ctx, int64_t lo, int64_t hi);
typedef struct { orc_range_fn fn; void *ctx; int64_t lo, hi; } orc_job;
static void *orc_job_main(void *p) {
    orc_job *j = (orc_job *)p;
    j->fn(j->ctx, j->lo, j->hi);
    return NULL;
}
static void orc_parallel_for(orc_range_fn fn, void *ctx, int64_t n, int threads) {
    if (threads <= 1 || n < 2) { fn(ctx, 0, n); return; }
    if (threads > 256) threads = 256;
    if (threads > n) threads = (int)n;
    pthread_t th[256];
    orc_job jobs[256];
    for (int t = 0; t < threads; t++) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].lo = n * t / threads; jobs[t].hi = n * (t + 1) / threads;
        pthread_create(&th[t], NULL, orc_job_main, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------- batch API ---- */

orc_batch *orc_create(int n, int flying, int select_and_place, int max_steps, double right_scale,
                      double wrong_scale) {
    JUMP_SPEED = sqrt(2 * GRAVITY * 1.2);
    orc_batch *b = (orc_batch *)calloc(1, sizeof(orc_batch));
    b->n = n;
    b->flying = flying;
    b->select_and_place = select_and_place;
    b->max_steps = max_steps;
    b->right_scale = right_scale;
    b->wrong_scale = wrong_scale;
    b->e = (orc_env *)calloc((size_t)n, sizeof(orc_env));
    for (int i = 0; i < n; i++) {
        orc_env *e = &b->e[i];
        e->flying = flying; /* env.py:78 */
        e->time_int_steps = 2;
        e->active_block = 1; /* BLUE */
        for (int c = 0; c < 6; c++) e->inventory[c] = 20;
    }
    return b;
}

void orc_destroy(orc_batch *b) {
    if (!b) return;
    free(b->pool_target); free(b->pool_start);
    free(b->next_task); free(b->task_id); free(b->episode);
    free(b->e);
    free(b);
}

/* env.set_task + initialize_world (env.py:155-193) followed by reset (206-261).
 * target/start are dense int8[9*11*11] (start may be NULL = empty list []);
 * pose = (x, y, z, yaw, pitch) or NULL for the default (0,0,0,0,0). */
void orc_set_task(orc_batch *b, int i, const int8_t *target, const int8_t *start, const double *pose) {
    orc_env *e = &b->e[i];
    env_assign_task(e, target, start);
    for (int k = 0; k < 5; k++) e->init_pose[k] = pose ? pose[k] : 0.0;
    env_reset(e);
}

/* env.set_task_generator (env.py:166-172) with a pool of m tasks; does not reset. */
void orc_set_pool(orc_batch *b, const int8_t *targets, const int8_t *starts, int64_t m, uint64_t seed,
                  int64_t env_offset) {
    free(b->pool_target); free(b->pool_start);
    b->pool_target = (int8_t *)malloc((size_t)m * CELLS);
    memcpy(b->pool_target, targets, (size_t)m * CELLS);
    b->pool_start = NULL;
    if (starts) {
        b->pool_start = (int8_t *)malloc((size_t)m * CELLS);
        memcpy(b->pool_start, starts, (size_t)m * CELLS);
    }
    b->pool_m = m; b->pool_seed = seed; b->pool_off = env_offset;
    if (!b->next_task) {
        b->next_task = (int32_t *)malloc(sizeof(int32_t) * (size_t)b->n);
        b->task_id = (int32_t *)malloc(sizeof(int32_t) * (size_t)b->n);
        b->episode = (int32_t *)calloc((size_t)b->n, sizeof(int32_t));
    }
    for (int i = 0; i < b->n; i++) { b->next_task[i] = -1; b->task_id[i] = -1; b->episode[i] = 0; }
}
int32_t *orc_pool_next_task(orc_batch *b) { return b->next_task; }
int32_t *orc_pool_task_id(orc_batch *b) { return b->task_id; }
void orc_get_target(const orc_batch *b, int32_t *target) {
    for (int i = 0; i < b->n; i++) memcpy(target + (size_t)i * CELLS, b->e[i].target, sizeof(int32_t) * CELLS);
}

void orc_reset(orc_batch *b, const uint8_t *mask) {
    for (int i = 0; i < b->n; i++)
        if (!mask || mask[i]) batch_reset_env(b, i);
}

typedef struct {
    orc_batch *b;
    const int32_t *actions, *inventory, *placement, *buttons;
    const float *movement, *camera;
    double *reward;
    uint8_t *done;
    int auto_reset;
} orc_step_args;

static void step_walking_range(void *ctx, int64_t lo, int64_t hi) {
    orc_step_args *s = (orc_step_args *)ctx;
    for (int64_t i = lo; i < hi; i++) {
        orc_env *e = &s->b->e[i];
        orc_action a;
        if (s->auto_reset && e->done_prev) batch_reset_env(s->b, i);
        parse_walking_discrete(s->actions[i], &a);
        env_step(s->b, e, &a, &s->reward[i], &s->done[i]);
    }
}
void orc_step_walking(orc_batch *b, const int32_t *actions, double *reward, uint8_t *done, int auto_reset,
                      int threads) {
    orc_step_args s = {b, actions, NULL, NULL, NULL, NULL, NULL, reward, done, auto_reset};
    orc_parallel_for(step_walking_range, &s, b->n, threads);
}

/* core/world.py:416-432 parse_flying_action; float32 components widened to fp64
 * (the pinned numpy 1.21 promotion). */
static void step_flying_range(void *ctx, int64_t lo, int64_t hi) {
    orc_step_args *s = (orc_step_args *)ctx;
    for (int64_t i = lo; i < hi; i++) {
        orc_env *e = &s->b->e[i];
        orc_action a;
        if (s->auto_reset && e->done_prev) batch_reset_env(s->b, i);
        memset(&a, 0, sizeof(a));
        a.strafe[0] = s->movement[3 * i + 0];
        a.strafe[1] = s->movement[3 * i + 1];
        a.dy = s->movement[3 * i + 2];
        a.camera[0] = s->camera[2 * i + 0];
        a.camera[1] = s->camera[2 * i + 1];
        a.inventory = s->inventory[i];
        a.add = s->placement[i] == 1;
        a.remove = s->placement[i] == 2;
        env_step(s->b, e, &a, &s->reward[i], &s->done[i]);
    }
}
void orc_step_flying(orc_batch *b, const float *movement, const float *camera, const int32_t *inventory,
                     const int32_t *placement, double *reward, uint8_t *done, int auto_reset, int threads) {
    orc_step_args s = {b, NULL, inventory, placement, NULL, movement, camera, reward, done, auto_reset};
    orc_parallel_for(step_flying_range, &s, b->n, threads);
}

/* core/world.py:396-414 parse_walking_action (Dict, discretize=False).
 * buttons: int32[n,8] = forward, back, left, right, jump, attack, use, hotbar. */
static void step_walking_dict_range(void *ctx, int64_t lo, int64_t hi) {
    orc_step_args *s = (orc_step_args *)ctx;
    orc_batch *b = s->b;
    const float *camera = s->camera;
    double *reward = s->reward;
    uint8_t *done = s->done;
    for (int64_t i = lo; i < hi; i++) {
        orc_env *e = &b->e[i];
        const int32_t *k = s->buttons + 8 * i;
        orc_action a;
        if (s->auto_reset && e->done_prev) batch_reset_env(s->b, i);
        memset(&a, 0, sizeof(a));
        if (k[0]) a.strafe[0] += -1;
        if (k[1]) a.strafe[0] += 1;
        if (k[2]) a.strafe[1] += -1;
        if (k[3]) a.strafe[1] += 1;
        a.dy = k[4];
        a.remove = k[5] != 0;
        a.add = k[6] != 0;
        a.inventory = k[7];
        a.camera[0] = camera[2 * i + 0];
        a.camera[1] = camera[2 * i + 1];
        env_step(b, e, &a, &reward[i], &done[i]);
    }
}
void orc_step_walking_dict(orc_batch *b, const int32_t *buttons, const float *camera, double *reward,
                           uint8_t *done, int auto_reset, int threads) {
    orc_step_args s = {b, NULL, NULL, NULL, buttons, NULL, camera, reward, done, auto_reset};
    orc_parallel_for(step_walking_dict_range, &s, b->n, threads);
}

/* State read-back for the parity tests: all arrays are [n, ...] row-major. */
void orc_get_state(const orc_batch *b, int8_t *grid, double *pos, double *rot, double *dy, int32_t *tis,
                   int32_t *inventory, int32_t *active, int32_t *step_no, int32_t *max_int, int32_t *err) {
    for (int i = 0; i < b->n; i++) {
        const orc_env *e = &b->e[i];
        if (grid) for (int k = 0; k < CELLS; k++) grid[(size_t)i * CELLS + k] = (int8_t)e->grid[k];
        if (pos) memcpy(pos + 3 * i, e->pos, 3 * sizeof(double));
        if (rot) memcpy(rot + 2 * i, e->rot, 2 * sizeof(double));
        if (dy) dy[i] = e->dy;
        if (tis) tis[i] = e->time_int_steps;
        if (inventory) for (int c = 0; c < 6; c++) inventory[6 * i + c] = e->inventory[c];
        if (active) active[i] = e->active_block;
        if (step_no) step_no[i] = e->step_no;
        if (max_int) max_int[i] = e->syn.max_int;
        if (err) err[i] = e->err;
    }
}

/* Overwrite the carried agent state (tests start GPU and oracle from the same point). */
void orc_set_agent(orc_batch *b, int i, const double *pos, const double *rot, double dy, int tis, int active) {
    orc_env *e = &b->e[i];
    if (pos) memcpy(e->pos, pos, 3 * sizeof(double));
    if (rot) memcpy(e->rot, rot, 2 * sizeof(double));
    e->dy = dy;
    e->time_int_steps = tis;
    e->active_block = active;
}

/* env.py:283-289 observation arrays (float32 casts are round-to-nearest-even). */
void orc_get_obs(const orc_batch *b, float *inventory, float *compass, float *agent_pos, int32_t *grid) {
    for (int i = 0; i < b->n; i++) {
        const orc_env *e = &b->e[i];
        for (int c = 0; c < 6; c++) inventory[6 * i + c] = (float)e->inventory[c];
        compass[i] = (float)(e->rot[0] - 180.);
        agent_pos[5 * i + 0] = (float)e->pos[0];
        agent_pos[5 * i + 1] = (float)e->pos[1];
        agent_pos[5 * i + 2] = (float)e->pos[2];
        agent_pos[5 * i + 3] = (float)e->rot[1];
        agent_pos[5 * i + 4] = (float)e->rot[0];
        if (grid) memcpy(grid + (size_t)i * CELLS, e->grid, sizeof(e->grid));
    }
}

/* Stateless reward sweep (BASELINE config 4): out[i] = Task(target_i, full_i, invariant)
 * .maximal_intersection(grid_i); arg (optional, int32[n,3]) = argmax_intersection.
 * include_init != 0 also runs Task.__init__ per pair (what the reference pays per reset). */
typedef struct {
    const int8_t *grids, *targets, *fulls;
    int invariant;
    int32_t *out, *arg;
    orc_task *ts;
} orc_mi_args;

static void max_intersection_range(void *ctx, int64_t lo, int64_t hi) {
    orc_mi_args *m = (orc_mi_args *)ctx;
    orc_task *t = (orc_task *)malloc(sizeof(orc_task));
    int32_t g[CELLS], tg[CELLS], fg[CELLS];
    for (int64_t i = lo; i < hi; i++) {
        for (int k = 0; k < CELLS; k++) {
            g[k] = m->grids[i * CELLS + k];
            tg[k] = m->targets[i * CELLS + k];
            if (m->fulls) fg[k] = m->fulls[i * CELLS + k];
        }
        orc_task_init(t, tg, m->fulls ? fg : NULL, m->invariant);
        int a[3];
        m->out[i] = orc_task_max_intersection(t, g, a);
        if (m->arg) { m->arg[3 * i] = a[0]; m->arg[3 * i + 1] = a[1]; m->arg[3 * i + 2] = a[2]; }
    }
    free(t);
}
void orc_max_intersection(const int8_t *grids, const int8_t *targets, const int8_t *fulls, int invariant,
                          int32_t *out, int32_t *arg, int64_t n, int threads) {
    orc_mi_args m = {grids, targets, fulls, invariant, out, arg, NULL};
    orc_parallel_for(max_intersection_range, &m, n, threads);
}

/* Task.__init__ products for the task-preparation parity test: per rotation the number of
 * admissible shifts and their bounding rectangle (dx_lo, dx_hi, dz_lo, dz_hi), target_size. */
void orc_task_describe(const int8_t *target, const int8_t *full, int invariant, int32_t *n_adm, int32_t *rect,
                       int32_t *target_size) {
    orc_task *t = (orc_task *)malloc(sizeof(orc_task));
    int32_t tg[CELLS], fg[CELLS];
    for (int k = 0; k < CELLS; k++) {
        tg[k] = target[k];
        if (full) fg[k] = full[k];
    }
    orc_task_init(t, tg, full ? fg : NULL, invariant);
    for (int r = 0; r < 4; r++) {
        int lo0 = 127, hi0 = -127, lo1 = 127, hi1 = -127;
        n_adm[r] = r < t->n_rot ? t->n_adm[r] : 0;
        for (int k = 0; k < n_adm[r]; k++) {
            int dx = t->adm[r][k][0], dz = t->adm[r][k][1];
            if (dx < lo0) lo0 = dx;
            if (dx > hi0) hi0 = dx;
            if (dz < lo1) lo1 = dz;
            if (dz > hi1) hi1 = dz;
        }
        rect[4 * r + 0] = lo0; rect[4 * r + 1] = hi0; rect[4 * r + 2] = lo1; rect[4 * r + 3] = hi1;
    }
    *target_size = t->target_size;
    free(t);
}

/* Timing helper for bench.py's cpu_baseline: only the maximal_intersection calls are timed
 * by the caller; tasks are prepared once here. Returns a handle array. */
static void tasks_prepare_range(void *ctx, int64_t lo, int64_t hi) {
    orc_mi_args *m = (orc_mi_args *)ctx;
    for (int64_t i = lo; i < hi; i++) {
        int32_t tg[CELLS];
        for (int k = 0; k < CELLS; k++) tg[k] = m->targets[i * CELLS + k];
        orc_task_init(&m->ts[i], tg, NULL, 1);
    }
}
orc_task *orc_tasks_prepare(const int8_t *targets, int64_t n, int threads) {
    orc_task *ts = (orc_task *)malloc(sizeof(orc_task) * (size_t)n);
    orc_mi_args m = {NULL, targets, NULL, 1, NULL, NULL, ts};
    orc_parallel_for(tasks_prepare_range, &m, n, threads);
    return ts;
}
static void tasks_eval_range(void *ctx, int64_t lo, int64_t hi) {
    orc_mi_args *m = (orc_mi_args *)ctx;
    for (int64_t i = lo; i < hi; i++) {
        int32_t g[CELLS];
        for (int k = 0; k < CELLS; k++) g[k] = m->grids[i * CELLS + k];
        m->out[i] = orc_task_max_intersection(&m->ts[i], g, NULL);
    }
}
void orc_tasks_eval(const orc_task *ts, const int8_t *grids, int32_t *out, int64_t n, int threads) {
    orc_mi_args m = {grids, NULL, NULL, 1, out, NULL, (orc_task *)ts};
    orc_parallel_for(tasks_eval_range, &m, n, threads);
}
void orc_tasks_free(orc_task *ts) { free(ts); }

/* tasks/task_set.py:135-157 RandomTasks.sample_task for a batch of tasks. The control flow is the reference's:
 * per height level one uniformly placed first block, then max_blocks - 1 more, each at an offset
 * (ddx, ddz) in [-max_dist, max_dist]^2 from the FIRST block of the level (block_x / block_z never move), redrawn
 * -- together with a colour -- until it is not (0, 0), inside the zone and free. The reference draws from numpy's
 * global Mersenne Twister, which has no batched meaning; like the task-pool draw above, the batched build defines
 * a counter-based stream instead (include/iglu_b200.h iglu_sample_tasks): the i-th draw of task t with range r is
 *     x = seed + 0x9E3779B97F4A7C15 * (t + 1) + 0xD1B54A32D192ED03 * i   (then the splitmix64 finaliser)
 *     value = ((x >> 32) * r) >> 32
 * and a block that finds no free cell within 4096 redraws is skipped (the reference would loop forever). */
static uint32_t sample_draw(uint64_t seed, int64_t task, uint32_t *counter, uint32_t range) {
    uint64_t x = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(task + 1) + 0xD1B54A32D192ED03ull * (uint64_t)(*counter)++;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (uint32_t)(((x >> 32) * (uint64_t)range) >> 32);
}

void orc_sample_tasks(int8_t *out, int64_t m, int max_blocks, int height_levels, int max_dist, int num_colors,
                      uint64_t seed, int64_t first_task) {
    for (int64_t t = 0; t < m; t++) {
        int8_t *g = out + t * CELLS;
        memset(g, 0, CELLS);
        uint32_t ctr = 0;
        for (int h = 0; h < height_levels && h < GY; h++) {
            const int bx = (int)sample_draw(seed, first_task + t, &ctr, GX);
            const int bz = (int)sample_draw(seed, first_task + t, &ctr, GZ);
            int color = (int)sample_draw(seed, first_task + t, &ctr, (uint32_t)num_colors) + 1;
            g[h * GX * GZ + bx * GZ + bz] = (int8_t)color;
            for (int b = 1; b < max_blocks; b++) {
                int ddx = 0, ddz = 0, tries = 0, ok = 0;
                while (tries++ < 4096) {
                    ddx = (int)sample_draw(seed, first_task + t, &ctr, (uint32_t)(2 * max_dist + 1)) - max_dist;
                    ddz = (int)sample_draw(seed, first_task + t, &ctr, (uint32_t)(2 * max_dist + 1)) - max_dist;
                    color = (int)sample_draw(seed, first_task + t, &ctr, (uint32_t)num_colors) + 1;
                    const int x = bx + ddx, z = bz + ddz;
                    if ((ddx == 0 && ddz == 0) || x >= GX || z >= GZ || x < 0 || z < 0) continue;
                    if (g[h * GX * GZ + x * GZ + z] != 0) continue;
                    ok = 1;
                    break;
                }
                if (ok) g[h * GX * GZ + (bx + ddx) * GZ + (bz + ddz)] = (int8_t)color;
            }
        }
    }
}

/* tasks/task.py:208-308 Subtasks: a multi-turn structure whose goal moves on to the next turn's target when the
 * current one is complete (step_intersection, 288-298). One handle = one structure + its current task. */
typedef struct {
    int L, progressive;
    int32_t *seq;         /* [L][CELLS] dense structure after every turn (Tasks.to_dense of structure_seq[i]) */
    int task_start, task_goal;
    int32_t start[CELLS]; /* dense starting grid of the current task */
    orc_task cur;
} orc_staged;

orc_staged *orc_staged_create(const int8_t *stages, int L, int progressive) {
    orc_staged *s = (orc_staged *)calloc(1, sizeof(orc_staged));
    s->L = L;
    s->progressive = progressive;
    s->seq = (int32_t *)malloc(sizeof(int32_t) * (size_t)L * CELLS);
    for (int64_t i = 0; i < (int64_t)L * CELLS; i++) s->seq[i] = stages[i];
    return s;
}
void orc_staged_destroy(orc_staged *s) { free(s->seq); free(s); }

/* task.py:256-285 create_task(turn_start, turn_goal) + the task.reset() at its end (74-86) */
static void staged_create_task(orc_staged *s, int turn_start, int turn_goal) {
    const int tid = turn_goal < s->L - 1 ? turn_goal : s->L - 1;   /* min(turn_goal, len - 1); self.full is never set */
    for (int k = 0; k < CELLS; k++) s->start[k] = turn_start == -1 ? 0 : s->seq[(int64_t)turn_start * CELLS + k];
    orc_task_init(&s->cur, s->seq + (int64_t)tid * CELLS, s->seq + (int64_t)(s->L - 1) * CELLS, 1);
    int nstart = 0;
    for (int k = 0; k < CELLS; k++) nstart += s->start[k] != 0;
    s->cur.max_int = orc_task_max_intersection(&s->cur, s->start, NULL);  /* starting_grid is a list, never None */
    s->cur.prev_grid_size = nstart;
    s->cur.right_placement = s->cur.wrong_placement = 0;
}

/* task.py:225-243 reset with self.next = turn */
void orc_staged_reset(orc_staged *s, int turn) {
    s->task_start = turn;
    s->task_goal = turn + 1;
    staged_create_task(s, s->task_start, s->task_goal);
}

/* task.py:288-298 */
void orc_staged_step(orc_staged *s, const int8_t *grid, int32_t *out /* right, wrong, done, task_goal, max_int */) {
    int32_t g[CELLS];
    for (int k = 0; k < CELLS; k++) g[k] = grid[k];
    int right, wrong, done;
    task_step_intersection(&s->cur, g, &right, &wrong, &done);
    if (done && s->L > s->task_goal && s->progressive) {
        s->task_goal += 1;
        staged_create_task(s, s->task_start, s->task_goal);
        s->cur.prev_grid_size = 0;
        int r2, w2;
        task_step_intersection(&s->cur, g, &r2, &w2, &done);
    }
    out[0] = right; out[1] = wrong; out[2] = done; out[3] = s->task_goal; out[4] = s->cur.max_int;
}

/* ------------------------------------------------------------------ obs['pov']
 * Restatement of what the reference's OpenGL renderer draws (gridworld/render.py: set_3d 94-113 -- gluPerspective(90,
 * w / h, 0.1, 30), glRotatef(yaw, 0,1,0), glRotatef(-pitch, cos yaw, 0, sin yaw), glTranslatef(-pos); setup 42-50 --
 * clear colour (0.5, 0.69, 1.0), GL_CULL_FACE, GL_NEAREST; add_block 159-170 -- cubes with all six faces, floor blocks
 * top-only; gridworld/utils.py: cube_vertices 26-41, tex_coord / tex_coords 75-118, id2texture 130-139) as one ray per
 * pixel centre. PARITY UNPINNED against OpenGL itself: no GL context can be created in the build container, so this
 * function is checked by known answers (tests/test_oracle_golden.py) and serves as the bit-exact twin of the CUDA ray
 * caster (csrc/iglu_render.cuh): fp32, same operation order, no FMA contraction.
 * grid: int8 [CELLS]; pose: x, y, z, yaw, pitch; atlas: [128][128][4] RGBA rows top-down, or NULL = flat palette;
 * out: uint8 [height][width][3], rows top-down (render()[::-1][..., :-1], render.py:139-146, env.py:258). */
static const uint8_t POV_PALETTE[8][3] = {{245, 245, 245}, {70, 72, 76}, {0, 90, 128}, {17, 101, 36},
                                          {133, 16, 17},   {146, 71, 22}, {92, 37, 96}, {145, 137, 0}};

static void pov_texel(const uint8_t *atlas, int id, int qx, int qy, float u, float v, uint8_t *rgb) {
    if (!atlas) { memcpy(rgb, POV_PALETTE[id + 1], 3); return; }
    static const int TX[8] = {0, 1, 2, 3, 0, 1, 2, 3}, TY[8] = {0, 0, 0, 0, 1, 1, 1, 1}; /* utils.py:130-139, index id + 1 */
    const int res = id <= 0 ? 32 : 16; /* whole tile on the floor (split=False), a quarter per face otherwise */
    int tu = (int)(u * (float)res), tv = (int)(v * (float)res);
    tu = tu < 0 ? 0 : (tu > res - 1 ? res - 1 : tu);
    tv = tv < 0 ? 0 : (tv > res - 1 ? res - 1 : tv);
    const int tx = 32 * TX[id + 1] + (id <= 0 ? 0 : qx) + tu, ty = 32 * TY[id + 1] + (id <= 0 ? 0 : qy) + tv;
    memcpy(rgb, atlas + ((127 - ty) * 128 + tx) * 4, 3); /* texture origin is bottom-left */
}

void orc_render_pov(const int8_t *grid, const double *pose, int height, int width, const uint8_t *atlas, uint8_t *out) {
    const double yaw = pose[3], pitch = pose[4];
    const double sy_ = sin(py_radians(yaw)), cy = cos(py_radians(yaw)), sp = sin(py_radians(pitch)), cp = cos(py_radians(pitch));
    const float o[3] = {(float)pose[0], (float)pose[1], (float)pose[2]};
    const float f[3] = {(float)(sy_ * cp), (float)sp, (float)(-(cy * cp))}; /* sight vector, world.py:145-161 */
    const float rx = (float)cy, rz = (float)sy_;                            /* camera x axis after the yaw rotation */
    const float u3[3] = {-(rz * f[1]), rz * f[0] - rx * f[2], rx * f[1]};   /* up = right x forward */
    const float kx = 2.0f / (float)width, ky = 2.0f / (float)height, aspect = (float)width / (float)height;
    /* the walk through the cells is restricted to the bounding box of the blocks (cell indices; empty: b0 > b1) */
    int b0[3] = {16, 16, 16}, b1[3] = {-1, -1, -1};
    for (int c = 0; c < CELLS; c++)
        if (grid[c] != 0) {
            const int q[3] = {(c % 121) / 11, c / 121, c % 11};
            for (int a = 0; a < 3; a++) { if (q[a] < b0[a]) b0[a] = q[a]; if (q[a] > b1[a]) b1[a] = q[a]; }
        }
    const float lo[3] = {(float)b0[0] - 5.5f, (float)b0[1] - 1.5f, (float)b0[2] - 5.5f};
    const int dim[3] = {b1[0] - b0[0] + 1, b1[1] - b0[1] + 1, b1[2] - b0[2] + 1};
    for (int j = 0; j < height; j++)
        for (int i = 0; i < width; i++) {
            uint8_t *px_out = out + 3 * (j * width + i);
            const float sx = ((float)i + 0.5f) * kx - 1.0f, sy = 1.0f - ((float)j + 0.5f) * ky, sxa = sx * aspect;
            const float d[3] = {(f[0] + sxa * rx) + sy * u3[0], f[1] + sy * u3[1], (f[2] + sxa * rz) + sy * u3[2]};
            const float inv[3] = {d[0] != 0.0f ? 1.0f / d[0] : 0.0f, d[1] != 0.0f ? 1.0f / d[1] : 0.0f,
                                  d[2] != 0.0f ? 1.0f / d[2] : 0.0f}; /* plane distances are (plane - origin) * inv */
            float best_t = 30.0f;
            px_out[0] = 128; px_out[1] = 176; px_out[2] = 255; /* glClearColor(0.5, 0.69, 1.0) */
            if (d[1] < 0.0f && o[1] > -1.5f) { /* floor: top faces at y = -2 + 0.5, seen from above only */
                const float t = (-1.5f - o[1]) * inv[1];
                const float px = o[0] + d[0] * t, pz = o[2] + d[2] * t;
                if (t >= 0.1f && t <= best_t && px >= -18.5f && px <= 18.5f && pz >= -18.5f && pz <= 18.5f) {
                    const float cx = floorf(px + 0.5f), cz = floorf(pz + 0.5f);
                    const int id = (cx >= -5.0f && cx <= 5.0f && cz >= -5.0f && cz <= 5.0f) ? WHITE : GREY; /* world.py:60-71 */
                    pov_texel(atlas, id, 0, 0, pz - (cz - 0.5f), px - (cx - 0.5f), px_out);
                    best_t = t;
                }
            }
            if (b0[0] > b1[0]) continue;
            float t_in = 0.0f, t_out = best_t;
            int axis_in = -1, miss = 0;
            for (int a = 0; a < 3; a++) {
                const float hi = lo[a] + (float)dim[a];
                if (d[a] != 0.0f) {
                    const float t1 = (lo[a] - o[a]) * inv[a], t2 = (hi - o[a]) * inv[a];
                    const float tn = fminf(t1, t2), tf = fmaxf(t1, t2);
                    if (tn > t_in) { t_in = tn; axis_in = a; }
                    t_out = fminf(t_out, tf);
                } else if (o[a] < lo[a] || o[a] > hi) {
                    miss = 1;
                }
            }
            if (miss || t_in > t_out) continue;
            int cell[3], step[3];
            float t_max[3], t_delta[3];
            for (int a = 0; a < 3; a++) {
                const float p = (o[a] + d[a] * t_in) - lo[a];
                int ci = (int)floorf(p);
                ci = ci < 0 ? 0 : (ci > dim[a] - 1 ? dim[a] - 1 : ci);
                if (a == axis_in) ci = d[a] > 0.0f ? 0 : dim[a] - 1;
                cell[a] = ci;
                step[a] = d[a] > 0.0f ? 1 : -1;
                if (d[a] != 0.0f) {
                    const float bound = lo[a] + (float)(ci + (d[a] > 0.0f ? 1 : 0));
                    t_max[a] = (bound - o[a]) * inv[a];
                    t_delta[a] = fabsf(inv[a]);
                } else {
                    t_max[a] = 3.0e38f;
                    t_delta[a] = 0.0f;
                }
            }
            int axis = axis_in; /* face the current cell was entered through; -1: the eye is inside the zone */
            float t = t_in;
            for (int it = 0; it < 40; it++) {
                if (axis >= 0 && t >= 0.1f) { /* a cube is seen from outside, in front of the near plane */
                    if (t > best_t) break;
                    const int gx = cell[0] + b0[0], gy = cell[1] + b0[1], gz = cell[2] + b0[2];
                    const int v = grid[gy * 121 + gx * 11 + gz];
                    if (v != 0) {
                        const int id = v < -1 ? 0 : (v > 6 ? 6 : v);
                        const float px = o[0] + d[0] * t, py = o[1] + d[1] * t, pz = o[2] + d[2] * t;
                        const float X = (float)(gx - 5), Y = (float)(gy - 1), Z = (float)(gz - 5);
                        float u, w;
                        int qx, qy;
                        if (axis == 1) { /* cube_vertices: top / bottom */
                            if (step[1] < 0) { u = pz - (Z - 0.5f); w = px - (X - 0.5f); qx = 0; qy = 16; }
                            else { u = px - (X - 0.5f); w = pz - (Z - 0.5f); qx = 16; qy = 0; }
                        } else if (axis == 0) { /* left (-x) / right (+x) */
                            if (step[0] > 0) { u = pz - (Z - 0.5f); w = py - (Y - 0.5f); }
                            else { u = (Z + 0.5f) - pz; w = py - (Y - 0.5f); }
                            qx = 0; qy = 0;
                        } else { /* front (+z) / back (-z) */
                            if (step[2] < 0) { u = px - (X - 0.5f); w = py - (Y - 0.5f); }
                            else { u = (X + 0.5f) - px; w = py - (Y - 0.5f); }
                            qx = 16; qy = 16;
                        }
                        pov_texel(atlas, id, qx, qy, u, w, px_out);
                        break;
                    }
                }
                axis = (t_max[0] <= t_max[1] && t_max[0] <= t_max[2]) ? 0 : (t_max[1] <= t_max[2] ? 1 : 2);
                t = t_max[axis];
                if (t > t_out) break;
                cell[axis] += step[axis];
                if (cell[axis] < 0 || cell[axis] >= dim[axis]) break;
                t_max[axis] += t_delta[axis];
            }
        }
}
