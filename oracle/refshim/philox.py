"""Counter-based RNG contract of the project, in pure Python (TEST INFRASTRUCTURE).

The reference draws from CPython's global Mersenne Twister (`random.*`, seeded to
10 at import: Env_hybrid_multi_coop_scalable.py:10).  A GPU env cannot share one
global sequential generator, so the project DEFINES a per-env counter-based
stream and injects it into the unmodified reference modules
(`Environments.Env_X.random = PhiloxRandom(...)`) so both sides consume the same
numbers in the same (data-dependent) order.  The CUDA kernel
(`mh-ppo_b200/csrc/philox.cuh`) and the C oracle (`oracle/mhppo_oracle.c`)
implement exactly this contract:

  block(k)  = Philox4x32-10(counter=(k, 0, env_lo, env_hi), key=(seed_lo, seed_hi))
  every `random.*` call consumes ONE block (k += 1), except shuffle/sample which
  consume one block per swap.
  u         = ((w0 >> 5) * 2^26 + (w1 >> 6)) / 2^53            (53-bit, [0,1))
  uniform(a,b)       = a + (b - a) * u
  randint(a,b)       = a + min(floor(u * (b - a + 1)), b - a)
  normalvariate(m,s) = m + s * sqrt(-2 ln(1 - u)) * cos(2 pi u2),
                       u2 = ((w2 >> 5) * 2^26 + (w3 >> 6)) / 2^53   (same block)
  shuffle(x)         = for i = n-1 .. 1: j = floor(u * (i + 1)); swap(x[i], x[j])
  sample(pop, k)     = partial Fisher-Yates from the front: for i = 0 .. k-1:
                       j = i + floor(u * (n - i)); swap(pool[i], pool[j]); -> pool[:k]

Bit-equality with CPython's Mersenne Twister is neither possible nor required
("RNG control").
"""
import math

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for r in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK
        hi1, lo1 = p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def u53(wa, wb):
    return ((wa >> 5) * 67108864.0 + (wb >> 6)) / 9007199254740992.0


class PhiloxRandom:
    """Drop-in for the `random` module attribute of a reference env module."""

    def __init__(self, seed=0, env_id=0, ctr=0):
        self.seed_ = int(seed)
        self.env_id = int(env_id)
        self.ctr = int(ctr)
        self.log = None  # optional list of (kind, value) for debugging

    # -- stream control -----------------------------------------------------
    def set_stream(self, seed, env_id, ctr=0):
        self.seed_, self.env_id, self.ctr = int(seed), int(env_id), int(ctr)

    def seed(self, *a, **k):  # `random.seed(10)` at module import: ignored
        pass

    def _block(self):
        w = philox4x32_10(
            (self.ctr & MASK, 0, self.env_id & MASK, (self.env_id >> 32) & MASK),
            (self.seed_ & MASK, (self.seed_ >> 32) & MASK),
        )
        self.ctr += 1
        return w

    # -- the `random` API used by the reference ------------------------------
    def random(self):
        w = self._block()
        return u53(w[0], w[1])

    def uniform(self, a, b):
        return a + (b - a) * self.random()

    def randint(self, a, b):
        n = b - a + 1
        return a + min(int(math.floor(self.random() * n)), n - 1)

    def normalvariate(self, mu, sigma):
        w = self._block()
        u1 = u53(w[0], w[1])
        u2 = u53(w[2], w[3])
        return mu + sigma * (math.sqrt(-2.0 * math.log(1.0 - u1)) * math.cos(2.0 * math.pi * u2))

    def shuffle(self, x):
        for i in range(len(x) - 1, 0, -1):
            j = min(int(math.floor(self.random() * (i + 1))), i)
            x[i], x[j] = x[j], x[i]

    def sample(self, population, k):
        pool = list(population)
        n = len(pool)
        for i in range(k):
            j = i + min(int(math.floor(self.random() * (n - i))), n - i - 1)
            pool[i], pool[j] = pool[j], pool[i]
        return pool[:k]

    def choices(self, population, k=1):
        return [population[min(int(self.random() * len(population)), len(population) - 1)] for _ in range(k)]
