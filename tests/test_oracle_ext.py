"""The round-2 oracle restatements (oracle/oracle_ext.py: var / std / skew / first / last / nunique in the reference's
update -> combine -> eval structure) pinned against pandas, the reference's own oracle, on its fixture shapes
(bodo/tests/test_df_lib/test_gpu/test_gpu_end_to_end.py:68-110 aggregation matrix; test_streaming/test_groupby.py:83-96)."""

import math

import numpy as np
import pandas as pd
import pytest

from oracle import oracle_ext as X


def _frame(seed, n=4000, ng=37, nullable=True):
    rng = np.random.default_rng(seed)
    k = rng.integers(0, ng, n)
    v = rng.normal(50.0, 7.0, n)
    i = rng.integers(-1000, 1000, n)
    df = pd.DataFrame({"k": pd.array(k, dtype="Int64"), "v": v, "i": pd.array(i, dtype="Int64")})
    if nullable:
        df.loc[rng.random(n) < 0.05, "k"] = pd.NA
        df.loc[rng.random(n) < 0.2, "v"] = np.nan
        df.loc[rng.random(n) < 0.2, "i"] = pd.NA
        df.loc[df.k == 3, "i"] = pd.NA   # a group without values
        df.loc[(df.k == 5) & (np.arange(n) > 40), "v"] = np.nan  # a group with very few values
    return df


def _cols(df, c):
    kv = (~df.k.isna()).to_numpy()
    keys = df.k.fillna(0).to_numpy(dtype=np.int64)
    if df[c].dtype.kind == "f":
        return keys, kv, df[c].to_numpy(), None
    vv = (~df[c].isna()).to_numpy()
    return keys, kv, df[c].fillna(0).to_numpy(dtype=np.int64), vv


@pytest.mark.parametrize("batch_size", [97, 32768])
@pytest.mark.parametrize("col", ["v", "i"])
def test_var_std_skew_follow_pandas(batch_size, col):
    df = _frame(1)
    keys, kv, vals, vv = _cols(df, col)
    st = X.groupby_moments(keys, kv, vals, vv, True, batch_size)
    g = df.groupby("k")[col]
    exp = pd.DataFrame({"var": g.var(), "std": g.std(), "var_pop": g.var(ddof=0), "std_pop": g.std(ddof=0), "skew": g.skew(), "count": g.count()})
    assert set(st) == set(int(x) for x in exp.index)
    for k, row in exp.iterrows():
        s = st[int(k)]
        assert s["count"] == row["count"]
        for name, got in (("var", X.var_eval(s)), ("std", X.std_eval(s)), ("var_pop", X.var_eval(s, True)), ("std_pop", X.std_eval(s, True)),
                          ("skew", X.skew_eval(s))):
            e = row[name]
            e = float("nan") if e is pd.NA else float(e)
            if math.isnan(e):
                assert math.isnan(got), (k, name, got)
            else:
                assert got == pytest.approx(e, rel=1e-9, abs=1e-9), (k, name)


def test_welford_and_power_sums_agree_when_the_spread_is_not_tiny():
    """The device forms M2 = sum x^2 - (sum x)^2 / n; the reference carries Welford's M2.  Same value up to rounding
    for data like the fixtures' (|mean| / spread ~ 7)."""
    df = _frame(2, nullable=False)
    keys, kv, vals, vv = _cols(df, "v")
    for s in X.groupby_moments(keys, kv, vals, vv).values():
        if s["count"] > 1:
            m2_power = s["s2"] - s["s1"] * s["s1"] / s["count"]
            assert m2_power == pytest.approx(s["m2"], rel=1e-9)


@pytest.mark.parametrize("dropna", [True, False])
@pytest.mark.parametrize("batch_size", [61, 32768])
def test_first_last_nunique_follow_pandas(dropna, batch_size):
    df = _frame(3)
    keys, kv, vals, vv = _cols(df, "i")
    fl = X.groupby_first_last(keys, kv, vals, vv, dropna, batch_size)
    nu = X.groupby_nunique(keys, kv, vals, vv, dropna)
    g = df.groupby("k", dropna=dropna)["i"]
    exp = pd.DataFrame({"first": g.first(), "last": g.last(), "nunique": g.nunique()})
    assert len(fl) == len(exp) == len(nu)
    for k, row in exp.iterrows():
        kk = ("NA",) if k is pd.NA or (isinstance(k, float) and math.isnan(k)) else int(k)
        f, l = fl[kk]
        assert (f is None) == (row["first"] is pd.NA) and (f is None or f == row["first"])
        assert (l is None) == (row["last"] is pd.NA) and (l is None or l == row["last"])
        assert nu[kk] == row["nunique"]
    # float column: NaN counts as NA
    keys, kv, vals, vv = _cols(df, "v")
    flv = X.groupby_first_last(keys, kv, vals, vv, True, batch_size)
    gv = df.groupby("k")["v"]
    for k, e in gv.first().items():
        f = flv[int(k)][0]
        assert (f is None and math.isnan(e)) or f == e
