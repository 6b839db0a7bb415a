"""CPU property tests (hypothesis) of the oracle: size-independent invariants of the reference algorithm that the GPU
parity tests rely on (the oracle is the checker, so it gets checked from a second angle besides the golden vectors)."""

import collections

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

keys_st = st.lists(st.integers(min_value=-5, max_value=40), min_size=0, max_size=300)


@settings(max_examples=60, deadline=None)
@given(keys=keys_st, batch=st.integers(min_value=1, max_value=64), seed=st.integers(0, 2**31 - 1))
def test_groupby_invariants(oracle, keys, batch, seed):
    rng = np.random.default_rng(seed)
    k = np.array(keys, dtype=np.int64)
    v = rng.integers(-(2**62), 2**62, len(k)).astype(np.int64)
    r = oracle.groupby(k, None, ["sum", "count", "min", "max", "size"], [v] * 5, batch_size=batch)
    # one output row per distinct key, in first-appearance order
    assert list(r["keys"]) == list(dict.fromkeys(keys))
    sums, counts = r["cols"][0][0], r["cols"][1][0]
    assert int(counts.sum()) == len(k)
    assert int(np.sum(sums.astype(np.uint64))) % (1 << 64) == int(np.sum(v.astype(np.uint64))) % (1 << 64)  # wraps like int64
    for key, mn, mx, c in zip(r["keys"], r["cols"][2][0], r["cols"][3][0], r["cols"][4][0]):
        sel = v[k == key]
        assert mn == sel.min() and mx == sel.max() and c == len(sel)
    # the batch size (how the stream is cut) never changes the result
    r2 = oracle.groupby(k, None, ["sum", "count"], [v] * 2, batch_size=max(1, len(k)))
    assert dict(zip(r["keys"], sums)) == dict(zip(r2["keys"], r2["cols"][0][0]))


@settings(max_examples=40, deadline=None)
@given(keys=keys_st, n_pes=st.integers(min_value=1, max_value=9))
def test_sharded_groupby_is_a_partition_of_the_groups(oracle, keys, n_pes):
    k = np.array(keys, dtype=np.int64)
    v = np.arange(len(k), dtype=np.int64)
    full = oracle.groupby(k, None, ["sum"], [v])
    seen = {}
    for rank in range(n_pes):
        part = oracle.groupby(k, None, ["sum"], [v], n_pes=n_pes, rank=rank)
        assert (oracle.hash_to_rank(part["keys"], None, n_pes) == rank).all()
        for kk, s in zip(part["keys"], part["cols"][0][0]):
            assert kk not in seen
            seen[int(kk)] = int(s)
    assert seen == {int(kk): int(s) for kk, s in zip(full["keys"], full["cols"][0][0])}


@settings(max_examples=60, deadline=None)
@given(keys=keys_st, n_pes=st.integers(min_value=1, max_value=16), null_every=st.integers(min_value=0, max_value=5))
def test_shuffle_partition_is_a_stable_permutation(oracle, keys, n_pes, null_every):
    k = np.array(keys, dtype=np.int64)
    valid = None if null_every == 0 else (np.arange(len(k)) % (null_every + 1) != 0)
    counts, perm = oracle.shuffle_partition(k, valid, n_pes)
    assert sorted(perm.tolist()) == list(range(len(k)))
    dest = oracle.hash_to_rank(k, valid, n_pes)
    off = 0
    for d, c in enumerate(counts):
        seg = perm[off:off + c]
        assert (dest[seg] == d).all()
        assert (np.diff(seg) > 0).all()  # input order kept inside a destination (fill_send_array is stable)
        off += c
    assert off == len(k)


@settings(max_examples=60, deadline=None)
@given(bk=keys_st, pk=keys_st, bo=st.booleans(), po=st.booleans())
def test_hash_join_row_counts(oracle, bk, pk, bo, po):
    b = np.array(bk, dtype=np.int64)
    p = np.array(pk, dtype=np.int64)
    bi, pi = oracle.hash_join(b, None, p, None, bo, po, True)
    cb, cp = collections.Counter(bk), collections.Counter(pk)
    inner = sum(cb[x] * cp[x] for x in cb)
    expect = inner + (sum(c for x, c in cp.items() if x not in cb) if po else 0) + (sum(c for x, c in cb.items() if x not in cp) if bo else 0)
    assert len(bi) == expect
    m = (bi >= 0) & (pi >= 0)
    assert int(m.sum()) == inner and (b[bi[m]] == p[pi[m]]).all()
