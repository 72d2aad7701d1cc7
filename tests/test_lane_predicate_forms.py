"""The step kernel evaluates pedestrian.is_in_front / is_crossing_in_front (SC:463-476) without the branch on the walking
direction (env_core.cuh lane_pred).  This checks, in IEEE double arithmetic (numpy), that the folded form is bit-for-bit the
reference's two-sided form - also exactly on the thresholds and one ulp either side of them."""
import numpy as np


def _reference(Spy, d, cross, L, line, nl, pl):
    Lf = float(L); W = Lf * cross; Hn = (-W) / 2.0; Hp = W / 2.0
    inf_neg = Spy >= (Hp - cross * ((Lf - 0.5 * nl) - line)) - 0.001
    inf_pos = Spy <= (Hn + cross * ((line - 0.5 * nl) + 1.0)) + 0.001
    cif_neg = Spy < Hp - cross * (((Lf - line) - 1.0) - pl)
    cif_pos = Spy > Hn + cross * (line - pl)
    return np.where(d == -1, inf_neg, inf_pos), np.where(d == -1, cif_neg, cif_pos)


def _folded(Spy, d, cross, L, line, nl, pl):
    Lf = float(L); W = Lf * cross; Hn = (-W) / 2.0
    neg = d == -1
    y = np.where(neg, -Spy, Spy)
    ai = np.where(neg, L - line, line + 1).astype(np.float64)
    return y <= (Hn + cross * (ai - 0.5 * nl)) + 0.001, y > Hn + cross * ((ai - 1.0) - pl)


def test_folded_lane_predicates_equal_the_two_sided_form():
    rng = np.random.default_rng(5)
    n = 200000
    for L in (1, 2, 3, 5, 29):
        for nl, pl in ((0.0, 0.0), (1.0, 0.5)):
            cross = rng.uniform(2.0, 4.5, n)
            line = rng.integers(0, L, n).astype(np.float64)
            d = rng.choice([-1, 0, 1], n)
            Spy = rng.uniform(-1.2, 1.2, n) * L * cross
            # half of the positions sit exactly on a threshold of one of the four forms, or one ulp beside it
            Lf = float(L); W = Lf * cross; Hn = (-W) / 2.0; Hp = W / 2.0
            thr = np.stack([(Hp - cross * ((Lf - 0.5 * nl) - line)) - 0.001, (Hn + cross * ((line - 0.5 * nl) + 1.0)) + 0.001,
                            Hp - cross * (((Lf - line) - 1.0) - pl), Hn + cross * (line - pl)])
            pick = thr[rng.integers(0, 4, n), np.arange(n)]
            k = rng.integers(-1, 2, n)
            near = np.where(k < 0, np.nextafter(pick, -np.inf), np.where(k > 0, np.nextafter(pick, np.inf), pick))
            Spy = np.where(rng.random(n) < 0.5, near, Spy)
            a = _reference(Spy, d, cross, L, line, nl, pl)
            b = _folded(Spy, d, cross, L, line, nl, pl)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (L, nl)
