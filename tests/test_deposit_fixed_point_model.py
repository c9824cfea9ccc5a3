"""CPU model of the tile deposit's exact accumulation (csrc/pm_deposit_tiles.cuh).

* pm_fx_add / pm_fx_add32: a 64-bit sum kept as two 32-bit words, the low word added atomically and the carry
  derived from the RETURNED old value -- exact for any interleaving of the adders and any order of the
  low / high adds (each wrap-around of the low word is seen by exactly one adder).
* the conversion of an ordinary contribution: the integer value of a double c in [0, 2^32) is the low mantissa
  word of c + 1.5 * 2^52 (one DADD, round-to-nearest-even), which is what (uint32)__double2loint(...) reads.
* integer sums do not depend on the order of the terms, float32 running sums (the reference's grid) do --
  the reason the kernel can use tiles, atomics and work items and stay bit-reproducible."""
import random
import struct

import numpy as np
from hypothesis import given, settings, strategies as st

M32 = (1 << 32) - 1


def lo_word_of_magic_add(c):
    x = np.float64(c) + np.float64(6755399441055744.0)           # 1.5 * 2^52
    return struct.unpack("<II", struct.pack("<d", float(x)))[0]


@settings(max_examples=300, deadline=None)
@given(st.floats(min_value=0.0, max_value=float(2 ** 32 - 1), allow_nan=False))
def test_magic_add_gives_the_nearest_even_integer(c):
    want = int(np.rint(np.float64(c)))                           # round half to even, as cvt.rni would
    assert lo_word_of_magic_add(c) == (want & M32)


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(min_value=-(1 << 62), max_value=(1 << 62)), min_size=1, max_size=40), st.randoms())
def test_two_word_accumulation_is_exact_under_any_interleaving(values, rnd):
    """Each adder performs (1) old = fetch_add(lo, vlo), (2) fetch_add(hi, vhi + carry(old)); the two steps of
    different adders interleave arbitrarily.  The final (hi:lo) is the exact sum modulo 2^64."""
    lo = hi = 0
    pending = []                                                 # (vhi + carry) waiting to be added to hi
    todo = list(values)
    while todo or pending:
        if todo and (not pending or rnd.random() < 0.5):
            v = todo.pop(rnd.randrange(len(todo))) & ((1 << 64) - 1)
            vlo, vhi = v & M32, v >> 32
            if vlo:
                old = lo
                lo = (lo + vlo) & M32
                vhi = (vhi + (1 if ((old + vlo) & M32) < old else 0)) & M32
            if vhi:
                pending.append(vhi)
        else:
            hi = (hi + pending.pop(rnd.randrange(len(pending)))) & M32
    assert ((hi << 32) | lo) == (sum(values) & ((1 << 64) - 1))


def test_integer_sums_are_order_independent_and_float32_running_sums_are_not():
    rng = np.random.default_rng(1)
    w = rng.uniform(0, 1, 4000) * 8.0                            # contributions of mass 8
    fx = np.rint(w * 2.0 ** 24).astype(np.int64)                 # 2^-24 mass units
    order = rng.permutation(len(w))
    assert int(fx.sum()) == int(fx[order].sum())
    a = np.float32(0)
    for v in w.astype(np.float32):
        a = np.float32(a + v)
    b = np.float32(0)
    for v in w.astype(np.float32)[order]:
        b = np.float32(b + v)
    assert a != b                                                # the reference's grid depends on particle order
    exact = float(w.sum())
    assert abs(float(np.float32(fx.sum() * 2.0 ** -24)) - exact) <= abs(float(a) - exact) + 1e-3
