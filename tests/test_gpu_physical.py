"""GPU tests of the physical-operator layer (the reference's pipeline protocol) against pandas — the shape of the
reference's DataFrame-library end-to-end tests (bodo/tests/test_df_lib/test_end_to_end.py: _test_equal(bdf, pdf,
sort_output=True, reset_index=True))."""

import numpy as np
import pandas as pd
import pytest

from bodo_b200.physical import groupby_agg, groupby_agg_parquet, merge

pytestmark = pytest.mark.gpu


def _sorted(df):
    d = df.copy()
    d.columns = [f"c{i}" for i in range(d.shape[1])]
    for c in d.columns:
        d[c] = d[c].to_numpy(dtype="float64", na_value=np.nan)
    return d.sort_values(list(d.columns), na_position="last").reset_index(drop=True)


def test_readme_example_groupby_sum(gpu_lib):
    # README.md:105-112 / BASELINE.json configs[0] shape: A = arange(N) % 30, B = arange(N); df.groupby("A")["B"].sum()
    n = 2_000_000
    df = pd.DataFrame({"A": np.arange(n) % 30, "B": np.arange(n)})
    got = groupby_agg(df, "A", [("B", "B", "sum")], batch_size=1 << 18)
    exp = df.groupby("A", as_index=False)["B"].sum()
    np.testing.assert_array_equal(_sorted(got).to_numpy(), _sorted(exp).to_numpy())
    # closed form: sum_g = sum of k with k % 30 == g
    g = got.sort_values("A")["B"].to_numpy()
    k = np.arange(n)
    assert int(g[7]) == int(k[k % 30 == 7].sum())


def test_groupby_agg_matrix_vs_pandas(gpu_lib):
    rng = np.random.default_rng(4)
    n = 100_000
    df = pd.DataFrame({"k": rng.integers(0, 97, n), "x": rng.integers(-50, 50, n), "y": rng.random(n), "unused": rng.random(n)})
    got = groupby_agg(df, "k", [("sx", "x", "sum"), ("cx", "x", "count"), ("my", "y", "mean"), ("mn", "y", "min"), ("mx", "x", "max"), ("n", None, "size")])
    exp = df.groupby("k", as_index=False).agg(sx=("x", "sum"), cx=("x", "count"), my=("y", "mean"), mn=("y", "min"), mx=("x", "max"), n=("x", "size"))
    g, e = _sorted(got), _sorted(exp)
    np.testing.assert_allclose(g.to_numpy(), e.to_numpy(), rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("how", ["inner", "left", "right", "outer"])
def test_merge_vs_pandas(gpu_lib, how):
    rng = np.random.default_rng(8)
    left = pd.DataFrame({"a": rng.integers(0, 500, 20_000), "lv": rng.random(20_000)})
    right = pd.DataFrame({"b": rng.integers(250, 750, 3_000), "rv": rng.integers(0, 9, 3_000)})
    got = merge(left, right, "a", "b", how=how, batch_size=4096)
    exp = left.merge(right, left_on="a", right_on="b", how=how)[["b", "rv", "a", "lv"]]
    assert len(got) == len(exp)
    np.testing.assert_allclose(_sorted(got).to_numpy(), _sorted(exp).to_numpy(), rtol=0, atol=0, equal_nan=True)


def test_groupby_from_parquet_vs_pandas(gpu_lib, tmp_path):
    # read_parquet -> groupby -> agg (the shape of bodo/tests/test_df_lib/test_gpu/test_gpu_end_to_end.py:68-110):
    # PhysicalReadParquet streams Arrow record batches (nullable columns, several row groups) into PhysicalAggregate
    import pyarrow as pa
    import pyarrow.parquet as pq

    rng = np.random.default_rng(12)
    n = 50_000
    x = rng.integers(-1000, 1000, n).astype("float64")
    x[rng.random(n) < 0.05] = np.nan
    df = pd.DataFrame({"k": rng.integers(0, 300, n), "x": pd.array(x).astype("Int64"), "y": rng.random(n), "unused": rng.random(n)})
    path = str(tmp_path / "t.parquet")
    pq.write_table(pa.Table.from_pandas(df, preserve_index=False), path, row_group_size=7000)
    got = groupby_agg_parquet(path, "k", [("sx", "x", "sum"), ("cx", "x", "count"), ("my", "y", "mean"), ("n", None, "size")], batch_size=4096)
    exp = df.groupby("k", as_index=False).agg(sx=("x", "sum"), cx=("x", "count"), my=("y", "mean"), n=("y", "size"))
    np.testing.assert_allclose(_sorted(got).to_numpy(), _sorted(exp).to_numpy(), rtol=1e-5, atol=1e-8)
