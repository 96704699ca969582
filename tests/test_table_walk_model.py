"""Executable statement of the table walk of gp_decide_tables (csrc/gangpack_tables.cuh), on the CPU.

The kernel emits ExecutorNodes of tightly-pack from the per-shape prefix table S (S[i] = sum of the capacities of the nodes
before i) instead of from the capacities: it starts at the shape's first hosting node (one word per (shape, group) from the
table build), takes differences of neighbouring prefix words with the next word already requested, uses cap(d | driver) on
the driver's node and jumps zero-capacity runs with a galloping search.  This model mirrors that control flow line by line
and checks it against the definition (pack_tightly.go:45-61: node n takes min(capacity, remaining), in order) on random
tables -- the argument the CUDA code relies on, kept runnable without a GPU.  (The CUDA kernel itself is checked against the
oracles by the -m gpu tests.)"""
import numpy as np
import pytest


def tab_at(tab, i, ne, total):
    return int(tab[i]) if i < ne else total


def next_above(tab, lo, ne, total, val):
    """smallest p in (lo, ne] with S(p) > val; requires S(ne) = total > val -- galloping, then bisection"""
    a, b, step = lo + 1, ne, 1
    probes = 0
    while a < b:
        t = min(a + step - 1, b - 1)
        probes += 1
        if tab_at(tab, t, ne, total) > val:
            b = t
            break
        a = t + 1
        step <<= 1
    while a < b:
        mid = (a + b) >> 1
        probes += 1
        if tab_at(tab, mid, ne, total) > val:
            b = mid
        else:
            a = mid + 1
    return a, probes


def walk(tab, ne, total, k, dpos, cd, first_host):
    out, pos, prev, placed = [], first_host, 0, 0
    nxt = tab_at(tab, pos + 1, ne, total)
    while placed < k and pos < ne:
        nxt2 = tab_at(tab, pos + 2, ne, total)          # in flight while this node is handled
        c = nxt - prev
        if pos == dpos:
            c = cd
        if c == 0:
            if nxt >= total:
                break
            p, _ = next_above(tab, pos + 1, ne, total, nxt)
            pos, prev = p - 1, nxt
            nxt = tab_at(tab, pos + 1, ne, total)
            continue
        take = min(c, k - placed)
        out += [pos] * take
        placed += take
        prev, nxt, pos = nxt, nxt2, pos + 1
    return out


def definition(caps, k, dpos, cd):
    out = []
    for n, c in enumerate(caps):
        c = cd if n == dpos else int(c)
        take = min(c, k - len(out))
        out += [n] * take
        if len(out) == k:
            break
    return out


@pytest.mark.parametrize("seed", range(4))
def test_walk_equals_definition(seed):
    rng = np.random.default_rng(seed)
    checked = 0
    for _ in range(4000):
        ne = int(rng.integers(1, 80))
        caps = rng.integers(0, 5, ne) * (rng.random(ne) < rng.random())       # zero runs of every length
        if caps.sum() == 0:
            continue
        tab = np.concatenate([[0], np.cumsum(caps)[:-1]])
        total = int(caps.sum())
        dpos = int(rng.integers(0, ne + 2))
        cd = int(rng.integers(0, caps[dpos] + 1)) if dpos < ne else 0
        if dpos >= ne:
            dpos = 0x7fffffff                                                  # the driver's node is no executor candidate
        room = total - ((int(caps[dpos]) - cd) if dpos < ne else 0)
        if room <= 0:
            continue
        k = int(rng.integers(1, room + 1))                                     # feasible: S[ne] - (c0(d) - cap(d|drv)) >= k
        first_host = int(np.nonzero(caps)[0][0])
        assert walk(tab, ne, total, k, dpos, cd, first_host) == definition(caps, k, dpos, cd)
        checked += 1
    assert checked > 2000


def test_galloping_search_is_the_upper_bound():
    rng = np.random.default_rng(9)
    for _ in range(3000):
        ne = int(rng.integers(1, 300))
        caps = rng.integers(0, 3, ne) * (rng.random(ne) < 0.3)
        tab = np.concatenate([[0], np.cumsum(caps)[:-1]])
        total = int(caps.sum())
        lo = int(rng.integers(0, ne))
        val = tab_at(tab, lo, ne, total)
        if val >= total:
            continue
        p, probes = next_above(tab, lo, ne, total, val)
        full = np.concatenate([tab, [total]])
        want = lo + 1 + int(np.argmax(full[lo + 1:] > val))
        assert p == want
        assert probes <= 2 * max(1, int(np.ceil(np.log2(max(2, p - lo)))) + 1)   # O(log distance), not O(log ne)
