"""CPU model of the packed-column comparison the stateless sweep uses for crowded zone-spanning targets
(gridworld_b200/csrc/iglu_reward.cuh: pack_columns, nz_nibbles, eval_zero_shift_packed), checked against the oracle.

The model restates the kernel's integer arithmetic with numpy uint32 (wrap-around included) so that the identities it
relies on are pinned without a GPU:
  * lo = 0x88888888 + sum_y v_y << 4y (mod 2^32) holds nibble y = v_y + 8 for v_y in -8..7 (borrows of negative values
    cancel against the bias), hi = v_8 + 8;
  * a level matches iff the nibble of t ^ g is zero and the target nibble is not 8;
  * rotation k of the target compares column 11 x + z with the grid column 11 z + 10 - x (k = 1), 120 - column (k = 2),
    11 (10 - z) + x (k = 3)  (reference tasks/task.py:47-56), the first rotation reaching the maximum wins (task.py:127-136).
"""
import numpy as np

from oracle import oracle as O

M32 = np.uint64(0xFFFFFFFF)


def _pack(rows):
    """rows int8 [n, 9, 121] -> lo uint32 [n, 121], hi uint32 [n, 121] as pack_columns builds them."""
    v = rows.astype(np.int64)
    lo = np.full(v.shape[::2], 0x88888888, np.uint64)
    for y in range(8):
        lo = (lo + ((v[:, y, :] << (4 * y)).astype(np.uint64) & M32)) & M32   # two's complement, mod 2^32
    hi = (np.uint64(8) + (v[:, 8, :].astype(np.uint64) & M32)) & M32
    return lo.astype(np.uint32), hi.astype(np.uint32)


def _nz_nibbles(x):
    a = x | (x >> np.uint32(1))
    return (a | (a >> np.uint32(2))) & np.uint32(0x11111111)


def _popc(x):
    return np.unpackbits(np.ascontiguousarray(x).view(np.uint8).reshape(x.shape + (4,)), axis=-1).sum(axis=-1)


def _model(G, T, n_rot=4):
    n = G.shape[0]
    tl, th = _pack(T.reshape(n, 9, 121))
    gl, gh = _pack(G.reshape(n, 9, 121))
    col = np.arange(121)
    x, z = col // 11, col % 11
    maps = [col, 11 * z + 10 - x, 120 - col, 11 * (10 - z) + x][:n_rot]
    on = _nz_nibbles(tl ^ np.uint32(0x88888888))
    on8 = th != 8
    counts = []
    for m in maps:
        a = tl ^ gl[:, m]
        b = a | (a >> np.uint32(1))
        eq = on & ~(b | (b >> np.uint32(2)))
        counts.append((_popc(eq) + (on8 & (gh[:, m] == th))).sum(axis=1))
    counts = np.stack(counts, axis=1)
    best = counts.max(axis=1)
    kbest = counts.argmax(axis=1)           # first maximum
    return best.astype(np.int32), kbest.astype(np.int32)


def _pairs(seed, n, rho, diff):
    rng = np.random.RandomState(seed)
    draw = lambda r: np.where(rng.random_sample((n, 9, 11, 11)) < r, rng.randint(1, 7, (n, 9, 11, 11)), 0).astype(np.int8)  # noqa: E731
    T, G = draw(rho), draw(rho)
    for i in range(0, n, 2):
        g = np.rot90(T[i], k=-rng.randint(4), axes=(1, 2)).copy()
        keep = rng.random_sample(g.shape) < 0.9
        G[i] = np.where(keep, g, G[i])
    if diff:                                 # difference grids: the classes -6..-1 appear too (env.py:230, 290)
        S = draw(rho / 2)
        T = (T.astype(np.int16) - S).astype(np.int8)
        G = (G.astype(np.int16) - S).astype(np.int8)
    return G, T


def test_nibbles_of_packed_columns():
    rng = np.random.RandomState(1)
    rows = rng.randint(-8, 8, (50, 9, 121)).astype(np.int8)       # everything a nibble can hold
    lo, hi = _pack(rows)
    for y in range(8):
        assert np.array_equal((lo >> np.uint32(4 * y)) & np.uint32(15), (rows[:, y, :].astype(np.int32) + 8).astype(np.uint32))
    assert np.array_equal(hi, (rows[:, 8, :].astype(np.int32) + 8).astype(np.uint32))


def test_packed_columns_model_matches_oracle():
    for seed, rho, diff in ((3, 0.2, False), (4, 0.5, False), (5, 0.3, True)):
        G, T = _pairs(seed, 400, rho, diff)
        occ = (T != 0).any(axis=1)
        spans = occ.any(axis=2)[:, 0] & occ.any(axis=2)[:, 10] & occ.any(axis=1)[:, 0] & occ.any(axis=1)[:, 10]
        assert spans.all()                                        # crowded random targets span the zone
        want, warg = O.max_intersection(G, T, threads=4, return_argmax=True)
        best, kbest = _model(G, T)
        assert np.array_equal(best, want)
        hit = want > 0
        assert np.array_equal(kbest[hit], warg[hit, 2]) and not warg[hit, :2].any()
        want1 = O.max_intersection(G, T, invariant=False, threads=4)
        assert np.array_equal(_model(G, T, n_rot=1)[0], want1)
