// iglu_physics.cuh -- per-env (one thread) World/Agent update in fp64.
//
// Restates, in the reference's operation order, gridworld/core/world.py:73-99 (hit_test),
// 145-201 (sight / motion vectors), 203-262 (update/_update), 264-310 (collide), 312-332
// (place_or_remove_block), 338-356 (move_camera / movement), 434-456 (step) and the env
// callbacks gridworld/env.py:136-153.  Compiled with -fmad=false: every mul/add is a separate
// IEEE operation exactly as CPython evaluates it.  round() is half-even (numba round == rint).
//
// Trigonometry: CPython's math.sin/cos/atan2 are glibc's.  The device libm is not bit-identical,
// so every angle that is an exact multiple of 5 degrees inside [-180, 720] (all angles the
// Discrete(18) action space can produce from the default pose) is served from a table the HOST
// filled with glibc at iglu_ctx_create time, keyed by the un-reduced angle.  Other angles take
// the device sincos/atan2 (<= 2 ulp, pose parity 1e-6 as BASELINE.json states for flying).
//
// Memory-level parallelism: the voxel lookups of the ray cast and of collide() do not depend on
// each other's results (sample positions and the 12 neighbour cells are known up front), only the
// decisions do.  They are therefore issued as batches of independent predicated loads and the
// reference's sequential logic then runs on the gathered bits -- same results, one memory latency
// per batch instead of one per lookup.
#pragma once
#include <cstdint>

namespace iglu {

constexpr int CELLS = 1089;
constexpr int GSTRIDE = 1104;
constexpr int TRIG_LO = -36;  // table index 0 <-> -180 degrees
constexpr int TRIG_N = 181;   // ... index 180 <-> 720 degrees

// Filled by iglu_ctx_create (host glibc). g_trig[i] = (sin, cos)(radians(5 * (i + TRIG_LO))).
__device__ double2 g_trig[TRIG_N];
// degrees(atan2(s0, s1)) for s0, s1 in {-1, 0, 1}; index (s0 + 1) * 3 + (s1 + 1).
__device__ double g_strafe_deg[9];
__device__ double g_jump_speed;  // utils.py:21 sqrt(2 * GRAVITY * MAX_JUMP_HEIGHT)

constexpr double DEG2RAD = 3.14159265358979323846 / 180.0;  // CPython: x * (pi / 180)
constexpr double RAD2DEG = 180.0 / 3.14159265358979323846;  // CPython: x * (180 / pi)

struct Agent {
    double x, y, z, yaw, pitch, dy;
    int tis, active, err;
    int inv01, inv23, inv45;  // inventory, six int16 packed as stored in istate w0..w2
};

__device__ __forceinline__ int inv_get(const Agent &a, int i) {
    int w = i < 2 ? a.inv01 : (i < 4 ? a.inv23 : a.inv45);
    return (int)(int16_t)((unsigned)w >> (16 * (i & 1)));
}
__device__ __forceinline__ void inv_add(Agent &a, int i, int delta) {
    int w = i < 2 ? a.inv01 : (i < 4 ? a.inv23 : a.inv45);
    const int sh = 16 * (i & 1);
    const int v = ((int)(int16_t)((unsigned)w >> sh) + delta) & 0xFFFF;
    w = (w & ~(0xFFFF << sh)) | (v << sh);
    if (i < 2) a.inv01 = w;
    else if (i < 4) a.inv23 = w;
    else a.inv45 = w;
}

// The tuple every parse_* returns (world.py:394, 414, 432).
struct Action {
    double s0, s1, jump, cam0, cam1;
    int hotbar;  // 0 = None
    bool remove, add;
};

__device__ __forceinline__ void trig_deg(double deg, double &s, double &c) {
    double k = rint(deg * 0.2);
    if (k * 5.0 == deg && k >= (double)TRIG_LO && k <= (double)(TRIG_LO + TRIG_N - 1)) {
        double2 v = g_trig[(int)k - TRIG_LO];
        s = v.x;
        c = v.y;
    } else {
        sincos(deg * DEG2RAD, &s, &c);
    }
}

__device__ __forceinline__ double strafe_degrees(double s0, double s1) {
    if ((s0 == -1.0 || s0 == 0.0 || s0 == 1.0) && (s1 == -1.0 || s1 == 0.0 || s1 == 1.0))
        return g_strafe_deg[((int)s0 + 1) * 3 + ((int)s1 + 1)];
    return atan2(s0, s1) * RAD2DEG;
}

// `key in self.world`: the 37x37 floor at y = -2 (world.py:60-71) plus the placed blocks mirrored
// in grid[y+1, x+5, z+5] (env.py:136-142).  A lookup is split in two so that a batch issues all its
// loads before any result is consumed: in_zone()/cell_of() say whether a byte has to be fetched at
// all (a scattered 1-byte load costs a full L1 wavefront per lane, so keys outside the zone, floor
// keys and keys equal to the previous sample are predicated off), world_has() combines the byte
// with the bounds / floor tests.
__device__ __forceinline__ bool in_zone(int x, int y, int z) {
    return (unsigned)(x + 5) <= 10u && (unsigned)(z + 5) <= 10u && (unsigned)(y + 1) <= 8u;
}
__device__ __forceinline__ int cell_of(int x, int y, int z) { return (y + 1) * 121 + (x + 5) * 11 + (z + 5); }

// Predicated signed-byte load (0 when pred is false; no memory access then). Inline PTX so that the
// compiler neither turns the predicate into a branch nor waits on one load before issuing the next.
__device__ __forceinline__ int ld_s8_if(const int8_t *p, bool pred) {
    int v;
    asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n mov.b32 %0, 0;\n @q ld.global.s8 %0, [%1];\n}"
                 : "=r"(v) : "l"(p), "r"((int)pred) : "memory");
    return v;
}
__device__ __forceinline__ bool world_has(int cell_value, int x, int y, int z) {
    const bool floor_hit = y == -2 && abs(x) <= 18 && abs(z) <= 18;
    return floor_hit || cell_value != 0;  // cell_value is 0 for keys outside the zone
}

__device__ __forceinline__ bool build_zone_i(int x, int y, int z) {
    return -5 <= x && x <= 5 && -5 <= z && z <= 5 && -1 <= y && y < 8;
}

// Neighbour cells already fetched around one centre voxel. The grid does not change between the sub-steps of one
// update() (the step's edit happens before it), so a sub-step that stays in the centre cell of the one before only
// fetches the neighbours it still lacks -- usually none: one memory round trip per step instead of one per sub-step.
struct CollideCache {
    uint32_t key;     // packed centre voxel (0 = nothing cached: real keys have bit 9 of every field set or more)
    unsigned have;    // bit c: neighbour cell c (of the 12) was fetched
    unsigned filled;  // bit c: ... and holds a block
};

// world.py:264-310 collide (FACES order utils.py:156-163: +y, -y, -x, +x, +z, -z)
__device__ __forceinline__ void collide(const int8_t *__restrict__ g, double &px, double &py, double &pz,
                                        double &agent_dy, CollideCache &cc) {
    const double pad = 0.25;
    const int nx = (int)rint(px), ny = (int)rint(py), nz = (int)rint(pz);
    // A face can only act when the overlap along its axis is >= pad, and the two faces of one axis
    // exclude each other (after +axis pushes the position back the -axis overlap is -pad), so which
    // faces need their two neighbour cells is known from the incoming position: fetch only those,
    // all loads issued before any is used.
    const double ex = px - (double)nx, ey = py - (double)ny, ez = pz - (double)nz;
    const bool on[6] = {ey >= pad, -ey >= pad, -ex >= pad, ex >= pad, ez >= pad, -ez >= pad};
    // The 12 neighbour cells use 3 x, 4 y and 3 z coordinates: zone / floor tests once per coordinate,
    // cell indices as offsets from the centre cell.
    const bool xz[3] = {(unsigned)(nx + 4) <= 10u, (unsigned)(nx + 5) <= 10u, (unsigned)(nx + 6) <= 10u};  // nx-1, nx, nx+1
    const bool zz[3] = {(unsigned)(nz + 4) <= 10u, (unsigned)(nz + 5) <= 10u, (unsigned)(nz + 6) <= 10u};
    const bool yz[4] = {(unsigned)(ny - 1) <= 8u, (unsigned)ny <= 8u, (unsigned)(ny + 1) <= 8u,
                        (unsigned)(ny + 2) <= 8u};                                                       // ny-2 .. ny+1
    const bool xf[3] = {(unsigned)(nx + 17) <= 36u, (unsigned)(nx + 18) <= 36u, (unsigned)(nx + 19) <= 36u};  // |x| <= 18
    const bool zf[3] = {(unsigned)(nz + 17) <= 36u, (unsigned)(nz + 18) <= 36u, (unsigned)(nz + 19) <= 36u};
    const bool yf[4] = {ny == 0, ny == -1, ny == -2, ny == -3};                                          // y == -2
    const int c0 = cell_of(nx, ny, nz);
    const uint32_t ckey = (uint32_t)(nx + 512) + ((uint32_t)(ny + 512) << 10) + ((uint32_t)(nz + 512) << 20);
    if (ckey != cc.key) { cc.key = ckey; cc.have = 0u; cc.filled = 0u; }
    // per (face, height): index into the x / y / z coordinate tables and the cell offset
    //                      +y        -y        -x        +x        +z        -z
    constexpr int XI[12] = {1, 1,     1, 1,     0, 0,     2, 2,     1, 1,     1, 1};
    constexpr int YI[12] = {3, 2,     1, 0,     2, 1,     2, 1,     2, 1,     2, 1};
    constexpr int ZI[12] = {1, 1,     1, 1,     1, 1,     1, 1,     2, 2,     0, 0};
    int cv[12];
    unsigned fetched = 0u;
#pragma unroll
    for (int c = 0; c < 12; c++) {
        const bool need = on[c >> 1] && xz[XI[c]] && yz[YI[c]] && zz[ZI[c]] && !((cc.have >> c) & 1u);
        const int off = (XI[c] - 1) * 11 + (YI[c] - 2) * 121 + (ZI[c] - 1);
        cv[c] = ld_s8_if(g + (need ? c0 + off : 0), need);
        fetched |= (unsigned)need << c;
    }
    unsigned solid = 0;
#pragma unroll
    for (int c = 0; c < 12; c++) {
        const bool floor_hit = yf[YI[c]] && xf[XI[c]] && zf[ZI[c]];
        solid |= (unsigned)floor_hit << c;
    }
    if (fetched) {  // nothing to wait for when every neighbour this sub-step needs is cached
#pragma unroll
        for (int c = 0; c < 12; c++) cc.filled |= (unsigned)(cv[c] != 0) << c;
        cc.have |= fetched;
    }
    solid |= cc.filled;  // cells never fetched are outside the zone (empty) or belong to faces that cannot act
#pragma unroll
    for (int f = 0; f < 6; f++) {
        const int axis = (f < 2) ? 1 : (f < 4 ? 0 : 2);
        const int sgn = (f == 0 || f == 3 || f == 4) ? 1 : -1;
        double &p = (axis == 0) ? px : (axis == 1 ? py : pz);
        const int npi = (axis == 0) ? nx : (axis == 1 ? ny : nz);
        const double d = (p - (double)npi) * (double)sgn;
        if (d < pad) continue;
        if ((solid >> (2 * f)) & 3u) {  // either height hits: same adjustment, then break
            p -= (d - pad) * (double)sgn;
            if (f < 2) agent_dy = 0.0;
        }
    }
}

// world.py:222-262 _update
__device__ __forceinline__ void update_sub(const int8_t *__restrict__ g, Agent &a, bool flying, double mvx,
                                           double mvy, double mvz, double dt, CollideCache &cc) {
    const double speed = flying ? 15.0 : 5.0;
    const double d = dt * speed;
    double dx = mvx * d, dy = mvy * d, dz = mvz * d;
    if (!flying) {
        a.dy -= dt * 20.0;
        if (a.dy < -14.0) a.tis = 12;
        else if (a.dy < -10.0) a.tis = 8;
        else if (a.dy < -5.0) a.tis = 4;
        else a.tis = 2;
        a.dy = a.dy > -50.0 ? a.dy : -50.0;
    }
    dy += a.dy * dt;
    const double cx = a.x + dx, cy = a.y + dy, cz = a.z + dz;
    const bool inside = -7.0 <= cx && cx <= 7.0 && -7.0 <= cz && cz <= 7.0 && -3.0 <= cy && cy < 10.0;
    if (inside || !flying) {  // world.py:258-261: collide(cand) or, when walking, collide((x, y + dy, z))
        double ux = inside ? cx : a.x, uy = inside ? cy : a.y + dy, uz = inside ? cz : a.z;
        collide(g, ux, uy, uz, a.dy, cc);
        a.x = ux; a.y = uy; a.z = uz;
    }
}

// world.py:434-456 World.step for one env. `g` = this env's int8 grid (global memory),
// `init` = its starting grid (only read when has_start, for the block-count bookkeeping),
// `dsize` = nnz(grid - init), kept incrementally (task.py:110 recomputes it from scratch).
// edit_cell / edit_val report the single cell this step changed (-1 = none).
__device__ __forceinline__ void world_step(int8_t *__restrict__ g, const int8_t *__restrict__ init, bool has_start,
                                           Agent &a, Action act, bool flying, bool select_and_place, int &dsize,
                                           int &edit_cell, int &edit_val) {
    edit_cell = -1;
    edit_val = 0;
    if (select_and_place && act.hotbar != 0) { act.add = true; act.remove = false; }
    // movement (world.py:344-356); agent.strafe is [0, 0] on entry because sustain=False
    const double s0 = act.s0, s1 = act.s1;
    if (act.jump != 0.0 && a.dy == 0.0) a.dy = g_jump_speed * act.jump;
    if (flying && act.jump == 0.0) a.dy = 0.0;
    if (act.hotbar != 0) {
        if (act.hotbar < 1 || act.hotbar > 6) a.err = 1;  // the reference raises ValueError here
        else a.active = act.hotbar;
    }
    // move_camera (world.py:338-342)
    a.yaw = a.yaw + act.cam0;
    {
        double p = a.pitch + act.cam1;
        double mn = p < 90.0 ? p : 90.0;
        a.pitch = mn > -90.0 ? mn : -90.0;
    }
    // place_or_remove_block (world.py:312-332)
    if (act.add != act.remove) {
        double sp, cp, sy, cy;
        trig_deg(a.pitch, sp, cp);
        trig_deg(a.yaw - 90.0, sy, cy);
        const double vx = cy * cp, vy = sp, vz = sy * cp;  // get_sight_vector (145-161)
        // hit_test (73-99): 40 samples, step 1/5; lookups gathered 8 at a time
        const double sx = vx / 5.0, syy = vy / 5.0, sz = vz / 5.0;
        double px = a.x, py = a.y, pz = a.z;
        // Keys are compared in packed form (x + 1024 y + 2^20 z, offset 512 each): consecutive samples differ by at
        // most one cell per axis, so equality of packed keys is exact whatever the magnitude; only keys of a hit
        // (which lie in the zone or on the floor) and of the sample before it are unpacked.
        // `previous` (world.py:97) is always the key of the sample before, so "key != previous" is `fresh`.
        uint32_t kkey = 0, qkey = 0, last = 0;
        int hit_val = 0;  // grid byte of the hit cell (0 for floor hits): self.world[block] without a second load
        bool has_prev = false, hit = false, lvalid = false;
        for (int i0 = 0; i0 < 40 && !hit; i0 += 8) {
            uint32_t key[8];
            int cv[8];
            unsigned fresh = 0, solid = 0;
            const uint32_t last_in = last;
            const bool lvalid_in = lvalid;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int bx = (int)rint(px), by = (int)rint(py), bz = (int)rint(pz);
                key[j] = (uint32_t)(bx + 512) + ((uint32_t)(by + 512) << 10) + ((uint32_t)(bz + 512) << 20);
                // a sample whose key equals the previous sample's key is never looked up (world.py:95)
                const bool fr = !lvalid || key[j] != last;
                const bool need = fr && in_zone(bx, by, bz);
                cv[j] = ld_s8_if(g + (need ? cell_of(bx, by, bz) : 0), need);
                const bool floor_hit = by == -2 && (unsigned)(bx + 18) <= 36u && (unsigned)(bz + 18) <= 36u;
                fresh |= (unsigned)fr << j;
                solid |= (unsigned)floor_hit << j;
                last = key[j];
                lvalid = true;
                px = px + sx; py = py + syy; pz = pz + sz;
            }
#pragma unroll
            for (int j = 0; j < 8; j++) solid |= (unsigned)(cv[j] != 0) << j;  // cv is 0 for keys outside the zone
            const unsigned hm = fresh & solid;
            if (hm) {  // first fresh sample whose key is in the world: (block, previous) = (its key, the key before)
                hit = true;
                const int jh = __ffs(hm) - 1;
                kkey = key[0]; qkey = last_in; has_prev = lvalid_in; hit_val = cv[0];
#pragma unroll
                for (int j = 1; j < 8; j++)
                    if (jh == j) { kkey = key[j]; qkey = key[j - 1]; has_prev = true; hit_val = cv[j]; }
            }
        }
        const int kx = (int)(kkey & 1023u) - 512, ky = (int)((kkey >> 10) & 1023u) - 512, kz = (int)(kkey >> 20) - 512;
        const int qx = (int)(qkey & 1023u) - 512, qy = (int)((qkey >> 10) & 1023u) - 512, qz = (int)(qkey >> 20) - 512;
        int cell = -1, newv = 0, oldv = 0;
        if (act.add) {
            if (hit && has_prev) {
                if (inv_get(a, a.active - 1) > 0 && build_zone_i(qx, qy, qz)) {
                    double y = a.y - 1.0 + 0.25;
                    double bx = (double)qx - 0.5, by = (double)qy, bz = (double)qz - 0.5;
                    bool inside = bx <= a.x && a.x <= bx + 1.0 && bz <= a.z && a.z <= bz + 1.0 &&
                                  ((by <= y && y <= by + 1.0) || (by <= (y + 1.0) && (y + 1.0) <= by + 1.0));
                    if (!inside) {
                        cell = (qy + 1) * 121 + (qx + 5) * 11 + (qz + 5);
                        newv = a.active;
                        inv_add(a, a.active - 1, -1);
                    }
                }
            }
        } else if (hit) {
            // self.world[block]: the floor's WHITE / GREY or the byte the hit test already fetched
            const int tex = ky == -2 ? ((abs(kx) <= 5 && abs(kz) <= 5) ? -1 : 0) : hit_val;
            if (tex != 0 && tex != -1) {  // not GREY / WHITE floor
                cell = (ky + 1) * 121 + (kx + 5) * 11 + (kz + 5);
                newv = 0;
                oldv = tex;
                if (tex >= 1 && tex <= 6) inv_add(a, tex - 1, 1);
            }
        }
        if (cell >= 0) {
            // old value of the edited cell: a removed block's colour; 0 for a placed one (`previous` is a sample
            // whose key was looked up and found absent, world.py:95-97, and build_zone puts it inside the grid)
            int iv = has_start ? (int)init[cell] : 0;
            dsize += (int)(newv != iv) - (int)(oldv != iv);
            g[cell] = (int8_t)newv;
            edit_cell = cell;
            edit_val = newv;
        }
    }
    // update (world.py:203-220): m = time_int_steps read once, dt = min(1/20, 0.2)
    {
        double mvx = 0.0, mvy = 0.0, mvz = 0.0;  // get_motion_vector (163-201)
        if (s0 != 0.0 || s1 != 0.0) {
            double sa, ca;
            trig_deg(a.yaw + strafe_degrees(s0, s1), sa, ca);
            if (flying) {
                double sp, cp;
                trig_deg(a.pitch, sp, cp);
                double m = cp, dyv = sp;
                if (s1 != 0.0) { dyv = 0.0; m = 1.0; }
                if (s0 > 0.0) dyv *= -1.0;
                mvx = ca * m; mvy = dyv; mvz = sa * m;
            } else {
                mvx = ca; mvy = 0.0; mvz = sa;
            }
        }
        const int m = a.tis;
        const double dt = 0.05 / (double)m;
        CollideCache cc = {0u, 0u, 0u};
        for (int i = 0; i < m; i++) update_sub(g, a, flying, mvx, mvy, mvz, dt, cc);
        if (flying) a.dy = 0.0;
    }
    // world.py:451-456. The iteration cap only matters for |yaw| so large that +-360 no longer
    // changes it (the reference would spin forever there); it keeps the kernel from hanging.
    double yaw = a.yaw;
    for (int it = 0; it < (1 << 16) && yaw > 360.0; it++) yaw -= 360.0;
    for (int it = 0; it < (1 << 16) && yaw < 0.0; it++) yaw += 360.0;
    a.yaw = yaw;
}

}  // namespace iglu
