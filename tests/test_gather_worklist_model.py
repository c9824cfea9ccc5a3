"""CPU model of the gather's work list (csrc/pm_particles.cu, k_gather_items): the same filing rule restated in
Python, checked on random and adversarial plane loads for the two properties the kernel relies on --
every particle of every (row block, plane) belongs to exactly one item, and the list never outgrows the
bound its buffer and its launch grid are sized with (pm_gather_item_bound: chunks + 6 np / T + 64)."""
import numpy as np
from hypothesis import given, settings, strategies as st


def file_chunk(counts, T):
    """counts: particles per plane of one base chunk.  Returns items (first_plane, planes, beg, end) with
    beg/end a particle range of a single crowded plane (end > beg) or (0, 0); heavy items flagged."""
    H = T // 2
    zc0 = len(counts)
    total = int(sum(counts))
    if total <= T:
        return [("light", 0, zc0, 0, 0)]
    items, z_first, acc = [], 0, 0
    for i, c in enumerate(counts):
        c = int(c)
        if c > T:
            if i > z_first:
                items.append(("heavy", z_first, i - z_first, 0, 0))
            q = 0
            while q < c:
                items.append(("heavy", i, 1, q, c if c - q < H else q + H))
                q += H
            z_first, acc = i + 1, 0
        elif acc + c > T:
            items.append(("heavy", z_first, i - z_first, 0, 0))
            z_first, acc = i, c
        else:
            acc += c
    if zc0 > z_first:
        items.append(("heavy", z_first, zc0 - z_first, 0, 0))
    return items


def check(chunks, T):
    np_total = int(sum(int(sum(c)) for c in chunks))
    n_items = 0
    for counts in chunks:
        items = file_chunk(counts, T)
        n_items += len(items)
        covered = [0] * len(counts)
        for kind, z, nz, b, e in items:
            assert nz >= 1 and 0 <= z and z + nz <= len(counts)
            if e > b:
                assert nz == 1 and e <= counts[z] and e - b <= T // 2 + 0
                covered[z] += e - b
            else:
                for k in range(z, z + nz):
                    covered[k] += int(counts[k])
        assert covered == [int(c) for c in counts]              # every particle in exactly one item
        # a piece of whole planes never exceeds T unless it is a single plane (those are cut into ranges)
        for kind, z, nz, b, e in items:
            if kind == "heavy" and e == b:
                assert sum(int(c) for c in counts[z:z + nz]) <= T
    assert n_items <= len(chunks) + 6 * (np_total // T) + 64, (n_items, len(chunks), np_total, T)
    return n_items


@settings(max_examples=200, deadline=None)
@given(st.lists(st.lists(st.integers(0, 40000), min_size=1, max_size=32), min_size=1, max_size=40),
       st.sampled_from([4096, 5000, 16384]))
def test_random_loads_are_covered_once_and_respect_the_bound(chunks, T):
    check(chunks, T)


def test_adversarial_loads():
    T = 4096
    # every chunk just above the threshold, in the most piece-producing way: planes of T/2 + 1
    chunks = [[T // 2 + 1] * 32 for _ in range(64)]
    n = check(chunks, T)
    assert n > len(chunks)
    # one plane holding everything; alternating empty / crowded planes; all light
    check([[10 ** 6] + [0] * 31], T)
    check([[0, T + 1] * 16 for _ in range(8)], T)
    assert check([[100] * 32 for _ in range(2048)], T) == 2048   # a uniform load files every chunk whole
