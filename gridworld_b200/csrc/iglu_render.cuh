// iglu_render.cuh -- obs['pov']: a voxel ray caster that stands in for the reference's OpenGL renderer
// (gridworld/render.py: set_3d 94-113, add_block 159-170; gridworld/utils.py: cube_vertices 26-41, tex_coord /
// tex_coords 75-118, id2texture 130-139; env.py:257-258, 299-300 obs['pov'] = render()[..., :-1]).
//
// What the reference draws, restated as a per-pixel ray query:
//   * camera: gluPerspective(90, w / h, 0.1, 30) at agent.position, glRotatef(yaw, 0,1,0) then
//     glRotatef(-pitch, cos yaw, 0, sin yaw): forward f = (sin yaw cos pitch, sin pitch, -cos yaw cos pitch) (the sight
//     vector of world.py:145-161), right r = (cos yaw, 0, sin yaw), up u = r x f; the ray through the centre of pixel
//     (column i, row j from the top) is f + sx * (w / h) * r + sy * u with sx = 2 (i + .5) / w - 1, sy = 1 - 2 (j + .5) / h,
//     so its parameter t is the eye-space depth and the near / far planes are t = 0.1 and t = 30;
//   * geometry: every placed block is a unit cube centred on its integer cell, drawn with all six faces; the 37 x 37
//     floor (y = -2) is drawn as TOP faces only (top_only=True); GL_CULL_FACE is on, so a face is seen from outside only
//     -- a ray that starts inside a cube (or enters one before the near plane) does not see that cube;
//   * colour: no lighting, no fog (setup_fog is commented out), GL_NEAREST texel of a 4 x 4 atlas of 32 x 32 tiles
//     (texture.png, origin bottom-left): WHITE (0,0) and GREY (1,0) use the whole tile on the floor; the six colours
//     BLUE (2,0) GREEN (3,0) RED (0,1) ORANGE (1,1) PURPLE (2,1) YELLOW (3,1) are "split" tiles, a 16 x 16 quarter per
//     face kind: sides (-x, +x) quarter (0,0), bottom (16,0), top (0,16), front / back (+z, -z) (16,16); texture
//     coordinates run over a face as cube_vertices / tex_coord pair them up (u, v below); the clear colour is
//     (0.5, 0.69, 1.0).
// The atlas is an input ([128][128][4] uint8, rows top-down as PIL / the PNG hold them); without one a flat palette close
// to the tile means is used. The arithmetic is fp32 in a fixed operation order (-fmad=false), mirrored by the test
// suite's C restatement (orc_render_pov), so kernel and checker agree bit for bit; agreement with the OpenGL
// rasteriser itself cannot be pinned here (no GL in this image): pixels on face edges may differ from it.
#pragma once
#include <cstdint>

namespace iglu {

constexpr float POV_NEAR = 0.1f, POV_FAR = 30.0f;

struct PovCamera {
    float ox, oy, oz;      // eye
    float fx, fy, fz;      // forward
    float rx, rz;          // right (ry = 0)
    float ux, uy, uz;      // up
};

__device__ __forceinline__ PovCamera pov_camera(double x, double y, double z, double yaw, double pitch) {
    double sy, cy, sp, cp;
    trig_deg(yaw, sy, cy);
    trig_deg(pitch, sp, cp);
    PovCamera c;
    c.ox = (float)x; c.oy = (float)y; c.oz = (float)z;
    c.fx = (float)(sy * cp); c.fy = (float)sp; c.fz = (float)(-(cy * cp));
    c.rx = (float)cy; c.rz = (float)sy;
    c.ux = -(c.rz * c.fy);
    c.uy = c.rz * c.fx - c.rx * c.fz;
    c.uz = c.rx * c.fy;
    return c;
}

// flat stand-in palette (WHITE -1 .. YELLOW 6), close to the tile means of the reference atlas, packed r | g << 8 | b << 16
__device__ __forceinline__ uint32_t pov_palette(int id) {
    switch (id) {
        case -1: return 245u | (245u << 8) | (245u << 16);
        case 0: return 70u | (72u << 8) | (76u << 16);
        case 1: return 0u | (90u << 8) | (128u << 16);
        case 2: return 17u | (101u << 8) | (36u << 16);
        case 3: return 133u | (16u << 8) | (17u << 16);
        case 4: return 146u | (71u << 8) | (22u << 16);
        case 5: return 92u | (37u << 8) | (96u << 16);
        default: return 145u | (137u << 8) | (0u << 16);
    }
}

// texel of block id `id` (-1..6), face quarter (qx, qy in texels, split tiles), face coordinates u, v in [0, 1]
__device__ __forceinline__ uchar3 pov_texel(const uint8_t *__restrict__ atlas, int id, int qx, int qy, float u, float v) {
    if (!atlas) {
        const uint32_t c = pov_palette(id);
        return make_uchar3((uint8_t)c, (uint8_t)(c >> 8), (uint8_t)(c >> 16));
    }
    // tile of the id: utils.py:130-139
    const int tile_x = id == -1 ? 0 : (id == 0 ? 1 : (id == 1 ? 2 : (id == 2 ? 3 : (id == 3 ? 0 : (id == 4 ? 1 : (id == 5 ? 2 : 3))))));
    const int tile_y = id >= 3 ? 1 : 0;
    const int res = id <= 0 ? 32 : 16;   // floor tiles use the whole tile, colours a quarter per face
    int tu = (int)(u * (float)res), tv = (int)(v * (float)res);
    tu = tu < 0 ? 0 : (tu > res - 1 ? res - 1 : tu);
    tv = tv < 0 ? 0 : (tv > res - 1 ? res - 1 : tv);
    const int tx = 32 * tile_x + (id <= 0 ? 0 : qx) + tu, ty = 32 * tile_y + (id <= 0 ? 0 : qy) + tv;   // from bottom-left
    const uint8_t *p = atlas + ((127 - ty) * 128 + tx) * 4;
    return make_uchar3(p[0], p[1], p[2]);
}

// Bounding box of the env's blocks in cell indices (x: 0..10, y: 0..8, z: 0..10), lo > hi when there is none. The walk
// through the cells is restricted to it: most worlds hold a few dozen blocks, so most rays never enter the box.
struct PovBox {
    int x0, x1, y0, y1, z0, z1;
};

// Colour of the ray o + t d, d = f + sxa r + sy u. g8 = the env's int8 grid (cell (y+1) * 121 + (x+5) * 11 + (z+5)).
__device__ __forceinline__ uchar3 pov_trace(const int8_t *g8, const PovBox &bb, const uint8_t *__restrict__ atlas,
                                            const PovCamera &c, float sxa, float sy) {
    const float o[3] = {c.ox, c.oy, c.oz};
    const float d[3] = {(c.fx + sxa * c.rx) + sy * c.ux, c.fy + sy * c.uy, (c.fz + sxa * c.rz) + sy * c.uz};
    // one IEEE reciprocal per axis; every plane distance below is (plane - origin) * inv (mirrored by the restatement)
    const float inv[3] = {d[0] != 0.0f ? 1.0f / d[0] : 0.0f, d[1] != 0.0f ? 1.0f / d[1] : 0.0f, d[2] != 0.0f ? 1.0f / d[2] : 0.0f};
    float best_t = POV_FAR;   // nearest accepted hit so far (hits beyond the far plane are clipped)
    uchar3 col = make_uchar3(128, 176, 255);   // glClearColor(0.5, 0.69, 1.0)
    // ---- floor: top faces at y = -1.5 over |x|, |z| <= 18.5, seen from above only
    if (d[1] < 0.0f && o[1] > -1.5f) {
        const float t = (-1.5f - o[1]) * inv[1];
        const float px = o[0] + d[0] * t, pz = o[2] + d[2] * t;
        if (t >= POV_NEAR && t <= best_t && px >= -18.5f && px <= 18.5f && pz >= -18.5f && pz <= 18.5f) {
            const float cx = floorf(px + 0.5f), cz = floorf(pz + 0.5f);
            const int id = (cx >= -5.0f && cx <= 5.0f && cz >= -5.0f && cz <= 5.0f) ? -1 : 0;   // world.py:60-71
            // top face: u runs with z, v with x (cube_vertices top: (x-n, z-n), (x-n, z+n), (x+n, z+n), (x+n, z-n))
            col = pov_texel(atlas, id, 0, 0, pz - (cz - 0.5f), px - (cx - 0.5f));
            best_t = t;
        }
    }
    // ---- blocks: DDA through the cells of the blocks' bounding box (cell (i, j, k) of the zone is the unit cube with
    // lower corner (i - 5.5, j - 1.5, k - 5.5))
    if (bb.x0 > bb.x1) return col;
    const float lo[3] = {(float)bb.x0 - 5.5f, (float)bb.y0 - 1.5f, (float)bb.z0 - 5.5f};
    const int dim[3] = {bb.x1 - bb.x0 + 1, bb.y1 - bb.y0 + 1, bb.z1 - bb.z0 + 1};
    float t_in = 0.0f, t_out = best_t;
    int axis_in = -1;
    bool miss = false;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float hi = lo[a] + (float)dim[a];
        if (d[a] != 0.0f) {
            const float t1 = (lo[a] - o[a]) * inv[a], t2 = (hi - o[a]) * inv[a];
            const float tn = fminf(t1, t2), tf = fmaxf(t1, t2);
            if (tn > t_in) { t_in = tn; axis_in = a; }
            t_out = fminf(t_out, tf);
        } else if (o[a] < lo[a] || o[a] > hi) {
            miss = true;
        }
    }
    if (miss || t_in > t_out) return col;
    // starting cell: the one holding the eye, or the boundary cell the ray enters the zone through
    // (scalars per axis instead of arrays indexed by the stepping axis: no local memory)
    int c0, c1, c2;
    float m0, m1, m2, e0, e1, e2;   // t of the next boundary per axis, t between boundaries per axis
    auto setup = [&](int a, int &ci, float &tm, float &td) {
        const float p = (o[a] + d[a] * t_in) - lo[a];
        ci = (int)floorf(p);
        ci = ci < 0 ? 0 : (ci > dim[a] - 1 ? dim[a] - 1 : ci);
        if (a == axis_in) ci = d[a] > 0.0f ? 0 : dim[a] - 1;
        if (d[a] != 0.0f) {
            const float bound = lo[a] + (float)(ci + (d[a] > 0.0f ? 1 : 0));
            tm = (bound - o[a]) * inv[a];
            td = fabsf(inv[a]);
        } else {
            tm = 3.0e38f;
            td = 0.0f;
        }
    };
    setup(0, c0, m0, e0);
    setup(1, c1, m1, e1);
    setup(2, c2, m2, e2);
    const int s0 = d[0] > 0.0f ? 1 : -1, s1 = d[1] > 0.0f ? 1 : -1, s2 = d[2] > 0.0f ? 1 : -1;
    int axis = axis_in;   // the face the current cell was entered through (-1: the eye is inside the zone)
    float t = t_in;
    for (int it = 0; it < 40; it++) {
        if (axis >= 0 && t >= POV_NEAR) {   // an entry event in front of the near plane
            if (t > best_t) break;
            const int gx = c0 + bb.x0, gy = c1 + bb.y0, gz = c2 + bb.z0;   // cell in the zone
            const int v = g8[gy * 121 + gx * 11 + gz];
            if (v != 0) {
                const int id = v < -1 ? 0 : (v > 6 ? 6 : v);
                const float px = o[0] + d[0] * t, py = o[1] + d[1] * t, pz = o[2] + d[2] * t;
                const float X = (float)(gx - 5), Y = (float)(gy - 1), Z = (float)(gz - 5);
                float u, w;
                int qx, qy;
                if (axis == 1) {
                    if (s1 < 0) { u = pz - (Z - 0.5f); w = px - (X - 0.5f); qx = 0; qy = 16; }        // top
                    else { u = px - (X - 0.5f); w = pz - (Z - 0.5f); qx = 16; qy = 0; }                // bottom
                } else if (axis == 0) {
                    if (s0 > 0) { u = pz - (Z - 0.5f); w = py - (Y - 0.5f); }                         // left (-x)
                    else { u = (Z + 0.5f) - pz; w = py - (Y - 0.5f); }                                // right (+x)
                    qx = 0; qy = 0;
                } else {
                    if (s2 < 0) { u = px - (X - 0.5f); w = py - (Y - 0.5f); }                         // front (+z)
                    else { u = (X + 0.5f) - px; w = py - (Y - 0.5f); }                                // back (-z)
                    qx = 16; qy = 16;
                }
                return pov_texel(atlas, id, qx, qy, u, w);
            }
        }
        // next cell boundary (ties: x before y before z)
        if (m0 <= m1 && m0 <= m2) {
            axis = 0; t = m0;
            if (t > t_out) break;
            c0 += s0;
            if (c0 < 0 || c0 >= dim[0]) break;
            m0 += e0;
        } else if (m1 <= m2) {
            axis = 1; t = m1;
            if (t > t_out) break;
            c1 += s1;
            if (c1 < 0 || c1 >= dim[1]) break;
            m1 += e1;
        } else {
            axis = 2; t = m2;
            if (t > t_out) break;
            c2 += s2;
            if (c2 < 0 || c2 >= dim[2]) break;
            m2 += e2;
        }
    }
    return col;
}

}  // namespace iglu
