"""CPU property test of K1's per-hop pool update against the reference's one-insert-at-a-time NeighborPriorityQueue.

The reference inserts the scored neighbours of a hop one by one (include/efanna2e/neighbor.h:150-183) and pops with
closest_unexpanded (:185-192).  K1 (csrc/rg_search.cu, "merge") applies a whole hop at once: candidates behind the tail of a
full pool are dropped when they are scored, the rest are ranked against the pool by binary search (a candidate equal to a
pool entry is the re-scored entry point and is dropped) and against each other by counting, pool entries behind the first
insertion point shift right by the number of candidates in front of them - in place, top-down in chunks of 4T entries, a
thread moving two pairs of neighbouring entries with one search per pair - and the cursor restarts at
min(first insertion point, position of the expanded entry + 1).

`K1Pool` below restates exactly those steps (same variable names as the kernel) in Python; the test drives it and the
oracle's pool (pinned to the compiled reference by tests/test_oracle_golden.py) with the same random hop sequences - heavy
(distance, id) ties, a re-scored entry point, pools from 1 to 300 entries, chunk sizes that force several shift rounds - and
compares the popped ids and the final pool (ids, distance bits, expanded flags).  The GPU suite checks the CUDA code itself;
this pins the ALGORITHM on a box without a GPU."""
import struct

import numpy as np
import pytest

from conftest import ROOT  # noqa: F401

DUP = (1 << 64) - 1


def ordered(d):
    u = struct.unpack("<I", struct.pack("<f", d))[0]
    return (~u & 0xFFFFFFFF) if u & 0x80000000 else (u | 0x80000000)


def make_key(d, i):          # rg_common.cuh make_key: ordered(distance) << 32 | id << 1 | expanded
    return (ordered(d) << 32) | (i << 1)


def key_id(k):
    return (k >> 1) & 0x7FFFFFFF


def lower_bound_key(P, n, key):
    lo, hi = 0, n
    while lo < hi:
        mid = (lo + hi) >> 1
        if (P[mid] & ~1) < key:
            lo = mid + 1
        else:
            hi = mid
    return lo


class K1Pool:
    def __init__(self, L, T):
        self.L, self.T = L, T
        self.P = [0] * (L + 1)
        self.size, self.cur, self.have_cur = 0, 0, False

    def tail(self):
        return (self.P[self.L - 1] & ~1) if self.size == self.L else DUP

    def hop(self, scored):
        """scored: (distance, id) of the hop's unvisited neighbours in the order the warps appended them"""
        P, L, T, size, cur, have_cur = self.P, self.L, self.T, self.size, self.cur, self.have_cur
        tail = self.tail()
        s_cand = [k for k in (make_key(d, i) for d, i in scored) if k < tail]
        C = len(s_cand)
        if C == 0:
            P[cur] |= 1
            start = cur + 1
        else:
            # (a)
            ndup, minlo, lb = 0, L, []
            for j, key in enumerate(s_cand):
                lo = lower_bound_key(P, size, key)
                lb.append(lo)
                if lo < size and (P[lo] & ~1) == key:
                    s_cand[j] = DUP
                    ndup += 1
                else:
                    minlo = min(minlo, lo)
            Cn = C - ndup
            minlo = min(minlo, size)
            # (b)
            s_sorted, s_pos = [None] * C, [None] * C
            for j, key in enumerate(s_cand):
                if key == DUP:
                    continue
                r = sum(1 for x in s_cand if x < key)
                s_sorted[r] = key
                s_pos[r] = lb[j] + r
            curpos = L
            if have_cur and cur < minlo:
                P[cur] |= 1
                curpos = cur
            # (c)
            hi = size
            while hi > minlo:
                lo_c = hi - 4 * T if hi - minlo > 4 * T else minlo
                moves = []
                for tid in range(T):
                    for u in range(2):
                        i = lo_c + 2 * (tid + u * T)
                        if i >= hi:
                            continue
                        e0 = P[i]
                        a, b = 0, Cn
                        while a < b:
                            mid = (a + b) >> 1
                            if s_pos[mid] - mid <= i:
                                a = mid + 1
                            else:
                                b = mid
                        pos0 = i + a
                        if have_cur and i == cur:
                            e0 |= 1
                            curpos = pos0
                        moves.append((pos0, e0))
                        if i + 1 < hi:
                            e1 = P[i + 1]
                            while a < Cn and s_pos[a] - a <= i + 1:
                                a += 1
                            pos1 = i + 1 + a
                            if have_cur and i + 1 == cur:
                                e1 |= 1
                                curpos = pos1
                            moves.append((pos1, e1))
                for pos, e in moves:        # after the barrier
                    if pos < L:
                        P[pos] = e
                hi = lo_c
            for r in range(Cn):
                if s_pos[r] < L:
                    P[s_pos[r]] = s_sorted[r]
            size = min(L, size + Cn)
            start = min(minlo, curpos + 1) if have_cur else 0
        nxt = size
        for i in range(start, size):
            if (P[i] & 1) == 0:
                nxt = i
                break
        self.size, self.cur = size, nxt
        self.have_cur = True
        return nxt < size            # something left to expand

    def state(self):
        ids = np.array([key_id(k) for k in self.P[:self.size]], np.uint32)
        flags = np.array([k & 1 for k in self.P[:self.size]], np.uint8)
        dbits = np.array([k >> 32 for k in self.P[:self.size]], np.uint64)
        return ids, dbits, flags


@pytest.fixture(scope="module")
def oracle():
    import oracle.binding as ob

    ob.build(ref=False)
    return ob.Oracle()


def dist_of(i, levels, sign):
    # deterministic in the id (a node has ONE distance to the query), few levels -> many (distance, id) ties
    return np.float32(sign * (((i * 2654435761) >> 7) % levels) / 8.0)


@pytest.mark.parametrize("seed", range(12))
def test_hop_merge_equals_sequential_inserts(oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(25):
        L = int(rng.choice([1, 2, 3, 7, 10, 33, 64, 100, 257, 300]))
        T = int(rng.choice([1, 2, 8, 32, 64]))            # 4T entries per shift round: small T forces many rounds
        levels = int(rng.choice([3, 16, 1000]))
        sign = float(rng.choice([-1.0, 1.0]))              # inner-product scores are negative
        max_deg = int(rng.choice([1, 4, 35, 70]))
        ids = rng.permutation(4000).astype(np.uint32)
        ep, fresh = int(ids[0]), list(map(int, ids[1:]))
        pool = K1Pool(L, T)
        kind, sid, sdist, pops = [0], [ep], [dist_of(ep, levels, sign)], []
        pool.hop([(float(dist_of(ep, levels, sign)), ep)])  # entry point: scored and inserted, not marked visited
        ep_again = int(rng.integers(1, 6))                  # ... so it is scored a second time when it shows up as a neighbour
        hop_no = 0
        more = True
        while more and fresh and hop_no < 400:
            hop_no += 1
            pops.append(key_id(pool.P[pool.cur]))
            kind.append(1); sid.append(0); sdist.append(np.float32(0))
            deg = int(rng.integers(0, max_deg + 1))
            nb = [fresh.pop() for _ in range(min(deg, len(fresh)))]
            if hop_no == ep_again:
                nb.insert(int(rng.integers(0, len(nb) + 1)), ep)
            scored = [(float(dist_of(i, levels, sign)), i) for i in nb]
            for d, i in scored:                             # the reference inserts in adjacency order
                kind.append(0); sid.append(i); sdist.append(np.float32(d))
            order = rng.permutation(len(scored))            # K1's warps append in any order
            more = pool.hop([scored[j] for j in order])
        oi, od, of, opop = oracle.pool_script(L, np.array(kind, np.uint8), np.array(sid, np.uint32), np.array(sdist, np.float32))
        ids_k1, dbits_k1, flags_k1 = pool.state()
        tag = f"L={L} T={T} levels={levels} deg<={max_deg} hops={hop_no}"
        assert list(opop) == pops, tag
        assert np.array_equal(oi, ids_k1), tag
        assert np.array_equal(np.array([ordered(float(x)) for x in od], np.uint64), dbits_k1), tag
        assert np.array_equal(of, flags_k1), tag
