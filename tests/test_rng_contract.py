"""The project's Philox contract: Random123 known-answer vectors, and the three implementations
(pure Python shim that drives the reference, C oracle, kernel header built for the host) agree."""
import numpy as np

from philox import PhiloxRandom, philox4x32_10

KAT = [  # Random123 kat_vectors, philox4x32-10
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers():
    for ctr, key, want in KAT:
        assert philox4x32_10(ctr, key) == want


def test_oracle_philox_matches_python(oracle_mod):
    rng = np.random.default_rng(0)
    for _ in range(200):
        ctr, env, seed = int(rng.integers(0, 2**32)), int(rng.integers(0, 2**40)), int(rng.integers(0, 2**63))
        want = philox4x32_10((ctr, 0, env & 0xffffffff, env >> 32), (seed & 0xffffffff, seed >> 32))
        assert oracle_mod.philox(ctr, env, seed) == want


def test_stream_consumption_rules():
    r = PhiloxRandom(5, 9)
    r.uniform(0, 1); r.randint(0, 9); r.normalvariate(0, 1)
    assert r.ctr == 3
    x = list(range(5)); r.shuffle(x)
    assert r.ctr == 7 and sorted(x) == list(range(5))
    s = r.sample(range(8), 3)
    assert r.ctr == 10 and len(set(s)) == 3
    vals = [PhiloxRandom(5, 9, k).randint(2, 5) for k in range(500)]
    assert min(vals) == 2 and max(vals) == 5
