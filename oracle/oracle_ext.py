"""Oracle (TEST INFRASTRUCTURE — never imported by bodo_b200/): CPU restatement of the reference's streaming groupby for the
aggregates added in round 2 — var / std / var_pop / std_pop, skew, first / last, nunique.

Plain Python / numpy in the reference's own structure (per-batch update -> combine into the running values -> eval), meant for
small inputs (thousands of rows).  Each function cites what it restates:

  update   var_agg  (Welford: count, mean, m2)                 bodo/libs/groupby/_groupby_agg_funcs.h:694-716
           skew_agg (count, sum, sum of squares, sum of cubes)  bodo/libs/groupby/_groupby_agg_funcs.h:723-745
           aggfunc<first> / aggfunc<last> (skip NA)             bodo/libs/groupby/_groupby_agg_funcs.h:594-611
  combine  var_combine (Chan's pairwise merge)                  bodo/libs/groupby/_groupby_update.cpp:1144-1195
           skew_combine (sums add)                              bodo/libs/groupby/_groupby_update.cpp:1239-1285
  eval     var_eval / std_eval / skew_eval                      bodo/libs/groupby/_groupby_eval.h:71-137
  nunique  distinct non-NA values of a group                    bodo/libs/groupby/_groupby_col_set.cpp:1771-1810 (nunique_computation)

Pinned by tests/test_oracle_ext.py against pandas (the reference's own oracle, bodo/tests/utils.py:157-186) on the reference's
fixture shapes.  The device accumulates POWER SUMS for var / std too (atomics cannot run Welford's recurrence): the same value
up to rounding unless |mean| >> spread; tests/test_gpu_groupby.py states that bound.
"""

from __future__ import annotations

import math

import numpy as np


def _isna(v, valid_i) -> bool:
    return (not valid_i) or (isinstance(v, float) and math.isnan(v))


def _batches(n, batch_size):
    for r0 in range(0, n, batch_size):
        yield r0, min(n, r0 + batch_size)


def _key_of(keys, key_valid, i, dropna):
    if key_valid is not None and not key_valid[i]:
        return None if dropna else ("NA",)
    return int(keys[i])


def groupby_moments(keys, key_valid, vals, val_valid, dropna=True, batch_size=32768):
    """Returns {key: dict(count, mean, m2, s1, s2, s3)} after the reference's update / combine sequence over the batches."""
    run = {}
    vals = np.asarray(vals)
    for r0, r1 in _batches(len(keys), batch_size):
        upd = {}
        for i in range(r0, r1):  # update table of this batch (var_agg / skew_agg per row, in row order)
            k = _key_of(keys, key_valid, i, dropna)
            if k is None:
                continue
            st = upd.setdefault(k, dict(count=0, mean=0.0, m2=0.0, s1=0.0, s2=0.0, s3=0.0))
            v = vals[i].item()
            if _isna(v, val_valid is None or val_valid[i]):
                continue
            x = float(v)
            st["count"] += 1
            delta = x - st["mean"]
            st["mean"] += delta / st["count"]
            st["m2"] += delta * (x - st["mean"])
            st["s1"] += x; st["s2"] += x * x; st["s3"] += x * x * x
        for k, b in upd.items():  # combine into the running values (var_combine / skew_combine)
            a = run.setdefault(k, dict(count=0, mean=0.0, m2=0.0, s1=0.0, s2=0.0, s3=0.0))
            if b["count"] == 0:
                continue
            count = a["count"] + b["count"]
            delta = b["mean"] - a["mean"]
            a["m2"] = a["m2"] + b["m2"] + delta * delta * a["count"] * b["count"] / count
            a["mean"] = (a["count"] * a["mean"] + b["count"] * b["mean"]) / count
            a["count"] = count
            a["s1"] += b["s1"]; a["s2"] += b["s2"]; a["s3"] += b["s3"]
    return run


def var_eval(st, pop=False):
    n = st["count"]
    if pop:
        return st["m2"] / n if n >= 1 else math.nan
    return st["m2"] / (n - 1) if n > 1 else math.nan


def std_eval(st, pop=False):
    v = var_eval(st, pop)
    return math.sqrt(v) if not math.isnan(v) else v


def skew_eval(st):
    n = st["count"]
    if n < 3:
        return math.nan
    m1, m2, m3 = st["s1"], st["s2"], st["s3"]
    s = m1 / n
    num = m3 - 3.0 * m2 * s + 2.0 * n * s ** 3
    base = m2 - s * m1
    den = base ** 1.5 if base >= 0 else math.nan
    if num == 0.0 or math.isnan(den) or abs(den) < 1e-14 or math.log2(abs(den)) - math.log2(abs(num)) < -20:
        return 0.0
    return ((n * (n - 1) ** 1.5 / (n - 2)) * num / den) / (n - 1)


def groupby_first_last(keys, key_valid, vals, val_valid, dropna=True, batch_size=32768):
    """{key: (first non-NA value or None, last non-NA value or None)} — update per batch, then combine in batch order."""
    run = {}
    vals = np.asarray(vals)
    for r0, r1 in _batches(len(keys), batch_size):
        upd = {}
        for i in range(r0, r1):
            k = _key_of(keys, key_valid, i, dropna)
            if k is None:
                continue
            f, l = upd.get(k, (None, None))
            v = vals[i].item()
            if not _isna(v, val_valid is None or val_valid[i]):
                f = v if f is None else f   # aggfunc<first>: keep the first non-NA value
                l = v                       # aggfunc<last>: the latest non-NA value wins
            upd[k] = (f, l)
        for k, (f, l) in upd.items():
            rf, rl = run.get(k, (None, None))
            run[k] = (rf if rf is not None else f, l if l is not None else rl)
    return run


def groupby_nunique(keys, key_valid, vals, val_valid, dropna=True):
    """{key: number of distinct non-NA values}."""
    sets = {}
    vals = np.asarray(vals)
    for i in range(len(keys)):
        k = _key_of(keys, key_valid, i, dropna)
        if k is None:
            continue
        s = sets.setdefault(k, set())
        v = vals[i].item()
        if not _isna(v, val_valid is None or val_valid[i]):
            s.add(v)
    return {k: len(s) for k, s in sets.items()}
