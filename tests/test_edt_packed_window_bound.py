"""The dense-image EDT row pass (edt_row_window16, csrc/preprocess.cu) runs on packed 16-bit lanes: column distances are clamped
to EDT16_G = 170 before squaring, both row copies carry EDT16_PAD = 192 halfwords of EDT16_BIG = 30000, and a row is redone
with the 32-bit routine when any result reaches 170^2.  This CPU test replays that arithmetic in numpy on adversarial rows and
checks the two claims the kernel relies on: (1) no 16-bit sum can wrap, (2) every result below 170^2 equals the exact
d2(x) = min_x' (g(x')^2 + (x - x')^2) (the quantity cv::distanceTransform PRECISE yields, src/SolveDVO.cpp:1709,1771)."""
import numpy as np

G, PAD, BIG = 170, 192, 30000


def exact_row(g):
    x = np.arange(len(g))
    g2 = g.astype(np.int64) ** 2
    return (g2[None, :] + (x[:, None] - x[None, :]) ** 2).min(axis=1)


def packed_row(g):
    w = len(g)
    c = np.minimum(g.astype(np.int64), G) ** 2
    row = np.concatenate([np.full(PAD, BIG), c, np.full(PAD, BIG)])
    best = c.copy()
    worst = 0
    kmax = int(np.ceil(np.sqrt(best.max()))) + 3            # the kernel stops at k^2 >= best for all lanes, checked every 4 steps
    assert kmax + 1 <= PAD
    for k in range(1, kmax + 1):
        cand = np.minimum(row[PAD - k:PAD - k + w], row[PAD + k:PAD + k + w]) + k * k
        worst = max(worst, int(np.maximum(row[PAD - k:PAD - k + w], row[PAD + k:PAD + k + w]).max()) + k * k)
        best = np.minimum(best, cand)
    return best, worst


def test_packed_window_is_exact_below_the_clamp_and_never_wraps():
    rng = np.random.default_rng(9)
    rows = []
    for w in (80, 160, 640, 1280):
        rows.append(np.full(w, 16384, np.int64))                               # no edge in any column
        r = np.full(w, 16384, np.int64); r[3] = 0; rows.append(r)               # one edge pixel near the border
        r = np.full(w, 16384, np.int64); r[w // 2] = 169; rows.append(r)        # just below / at / above the clamp
        r = np.full(w, 16384, np.int64); r[w // 2] = 170; rows.append(r)
        r = np.full(w, 16384, np.int64); r[w // 2] = 171; rows.append(r)
        rows.append(rng.integers(0, 400, w))                                    # dense, far larger than the clamp in places
        rows.append(rng.integers(0, 30, w))                                     # dense, typical
        r = rng.integers(150, 200, w); r[::97] = rng.integers(0, 5, len(r[::97])); rows.append(r)
    for g in rows:
        got, worst = packed_row(g)
        assert worst < 65536, worst
        ex = exact_row(g)
        ok = got < G * G
        assert np.array_equal(got[ok], ex[ok])
        assert np.all(ex[~ok] >= G * G)                                         # the rows the kernel redoes really are far from every edge
