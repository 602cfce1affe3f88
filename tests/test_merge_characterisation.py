"""The data-parallel class resolution on the GPU (k_resolve_paths_par, geno_session.cu) does not run the reference's
two-pointer walk (`intersect` without the truncate, src/mapping.rs:283-308); it uses a closed form of what that walk does
to a vector that is no longer sorted:
  * element i of v1 "matches" iff v1[i] is in v2 AND v1[i] is a left-to-right maximum of v1;
  * the r-th match lands at position r;
  * a match position p >= M (M = number of matches) receives the displaced element v1[m], where m starts at the match's rank
    and follows m -> rank(m) while m is itself a match position with rank(m) < m; every other position >= M keeps its element.
This test states both forms in plain Python and compares them on random inputs, including chains of merges in which v1 is
the (permuted) output of earlier merges - the situation the reference's quirk creates (SURVEY.md finding #2)."""
import random


def merge_reference(v1, v2):
    """src/mapping.rs:283-308 as executed: two pointers, swap on equality, NO truncate"""
    v = list(v1)
    f = i = j = 0
    while i < len(v) and j < len(v2):
        if v[i] < v2[j]:
            i += 1
        elif v[i] > v2[j]:
            j += 1
        else:
            v[f], v[i] = v[i], v[f]
            i += 1
            j += 1
            f += 1
    return v


def merge_closed_form(v1, v2):
    """what k_resolve_paths_par computes with a membership bitset, a prefix-maximum scan, a prefix sum and a gather"""
    n = len(v1)
    members = set(v2)
    rank_plus_1 = [0] * n          # rk[] of the kernel: rank + 1 at match positions, 0 elsewhere
    out = list(v1)
    prefix_max = -1
    m_count = 0
    for i, x in enumerate(v1):
        if x in members and x > prefix_max:
            out[m_count] = x        # B[rank] = x
            m_count += 1
            rank_plus_1[i] = m_count
        prefix_max = max(prefix_max, x)
    for i in range(m_count, n):
        if rank_plus_1[i]:
            m = rank_plus_1[i] - 1
            while rank_plus_1[m] and rank_plus_1[m] - 1 < m:
                m = rank_plus_1[m] - 1
            out[i] = v1[m]
        else:
            out[i] = v1[i]
    return out


def test_closed_form_equals_two_pointer_walk():
    rng = random.Random(20191101)
    checked = 0
    for trial in range(4000):
        universe = rng.randint(1, 40)
        base = sorted(rng.sample(range(universe), rng.randint(0, universe)))
        v = list(base)
        for _ in range(rng.randint(1, 8)):                  # a path: the last node's colour merged with every earlier one
            other = sorted(rng.sample(range(universe), rng.randint(0, universe)))
            if not other:
                continue                                    # colours are never empty in the index (every node has an allele)
            want = merge_reference(v, other)
            assert merge_closed_form(v, other) == want, (v, other)
            assert sorted(want) == base                     # a permutation: nothing is dropped without the truncate
            v = want
            checked += 1
    assert checked > 15000


def test_closed_form_on_adversarial_shapes():
    cases = [([5, 1, 2, 3], [1, 2, 3, 5]), ([1, 2, 3, 100], [1, 2, 100]), ([3, 1, 2], [1, 2, 3]), ([2, 1], [1, 2]),
             ([0, 1, 2, 3, 4, 5], [1, 3, 5]), ([5, 4, 3, 2, 1, 0], [0, 1, 2, 3, 4, 5]), ([1, 0, 3, 2, 5, 4], [0, 2, 4]),
             ([], [1]), ([7], [7]), ([7], [8])]
    for v1, v2 in cases:
        assert merge_closed_form(v1, v2) == merge_reference(v1, v2), (v1, v2)


def test_merge_with_the_same_set_twice_is_idempotent():
    """merge(merge(v, C), C) == merge(v, C). After a merge the matches are a sorted front block; an element x of C outside it
    had a larger element e in front of it: if e is in C it matched and now sits in the front block, and if it is not, either a
    later match (> e > x) sits in the front block or no rotation happened after x was queued behind e - so x is never a
    left-to-right maximum afterwards and the second merge finds every match already in place. (Not exploitable on real data:
    consecutive nodes of a path practically never share a colour; measured 0 of 14 293 merges.)"""
    rng = random.Random(7)
    for _ in range(5000):
        universe = rng.randint(1, 30)
        v = sorted(rng.sample(range(universe), rng.randint(0, universe)))
        for _ in range(rng.randint(0, 5)):
            v = merge_reference(v, sorted(rng.sample(range(universe), rng.randint(1, universe))))
        c = sorted(rng.sample(range(universe), rng.randint(1, universe)))
        once = merge_reference(v, c)
        assert merge_reference(once, c) == once


def _merge_with_stretch_hops(A, v2):
    """The r02 form of the data-parallel merge (k_resolve_paths_par): matches to the front, the list of non-match positions
    written from the back of the output vector, and every displaced element located by crossing whole stretches between
    consecutive non-matches with one division instead of following the chain link by link."""
    n = len(A)
    s2 = set(v2)
    rk, B, rank, pm = [0] * n, [None] * n, 0, -1
    for i, x in enumerate(A):
        match = x > pm and x in s2
        pm = max(pm, x)
        if match:
            rk[i] = rank + 1
            B[rank] = x
            rank += 1
        else:
            B[n - 1 - (i - rank)] = i          # position of the (i - rank)-th non-match
    M = rank
    last_pos = max([i + 1 for i in range(n) if rk[i]] or [0])
    if M == 0 or last_pos == M:
        return list(A)
    src = list(range(n))
    for i in range(M, n):
        if rk[i]:
            m = rk[i] - 1
            while rk[m]:
                j = m - (rk[m] - 1)            # non-matches before m = link length inside this stretch
                if j == 0:
                    break
                u = B[n - j]                   # the last non-match before m
                m -= ((m - u - 1) // j + 1) * j
            src[i] = m
    for i in range(M, n):
        B[i] = A[src[i]]
    return B


def test_stretch_hops_equal_the_sequential_walk():
    import random
    rng = random.Random(7)

    def seq(v1, v2):
        v1 = list(v1)
        f = i = j = 0
        while i < len(v1) and j < len(v2):
            if v1[i] < v2[j]:
                i += 1
            elif v1[i] > v2[j]:
                j += 1
            else:
                v1[i], v1[f] = v1[f], v1[i]
                i += 1
                j += 1
                f += 1
        return v1
    for _ in range(6000):
        n = rng.randint(1, 80)
        A = sorted(rng.sample(range(300), n))
        for _ in range(rng.randint(1, 8)):
            p = rng.choice([0.05, 0.5, 0.9, 0.98])
            v2 = sorted(x for x in range(300) if rng.random() < p)
            want = seq(A, v2)
            assert _merge_with_stretch_hops(A, v2) == want
            A = want
    # the case the hops are for: one allele missing at the front of a long sorted vector (a chain of length n - 1)
    A = list(range(2000))
    assert _merge_with_stretch_hops(A, A[1:]) == seq(A, A[1:]) == A[1:] + [0]
