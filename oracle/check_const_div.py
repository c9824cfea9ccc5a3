"""Exactness check for pm_div_const (csrc/pm_particles.cu): the drift of integrate.py:95 divides by
(a+da)^2, the same divisor for every particle, and the CUDA kernel replaces the IEEE division by
    r = RN(1/b) (host);  q = RN(x*r);  e = fma(-b, q, x);  q' = fma(e, r, q)
This replays the sequence in exact rational arithmetic (an FMA is one rounding of an exact
product-sum) and compares q' with the correctly rounded x/b.  Test infrastructure only."""
import random
from fractions import Fraction as F

import numpy as np


def div_const(x: float, b: float) -> float:
    r = 1.0 / b
    q = x * r
    e = float(F(x) - F(b) * F(q))      # fma(-b, q, x)
    return float(F(q) + F(e) * F(r))   # fma(e, r, q)


def mismatches(n_divisors: int, n_values: int, seed: int = 1) -> int:
    rng = random.Random(seed)
    bad = 0
    for _ in range(n_divisors):
        a = rng.uniform(0.005, 1.0)
        da = rng.choice([0.99 / 1000, 0.99 / 100, 0.0099, rng.uniform(1e-5, 0.1)])
        b = (a + da) * (a + da)
        for _ in range(n_values):
            v = float(np.float32(rng.gauss(0, 1) * 10 ** rng.uniform(-12, 3)))
            x = da * v
            bad += div_const(x, b) != x / b
    return bad


if __name__ == "__main__":
    print("mismatches", mismatches(400, 500), "of", 400 * 500)
