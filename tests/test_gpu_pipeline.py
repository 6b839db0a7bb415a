"""GPU parity tests of the pipeline neighbours: fused filter + projection, dictionary-encoded string
keys, and TPC-H Q1 through the whole operator chain on the reference's own test data
(bodo/tests/test_df_lib/test_tpch.py, benchmarks/tpch/bodo/dataframe_queries.py:96-119) against pandas."""

import datetime
import os

import numpy as np
import pandas as pd
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from bodo_b200.dictionary import DictionaryBuilder
from bodo_b200.expr import col, lit
from bodo_b200.physical import (OperatorResult, PhysicalAggregate, PhysicalFilterProject, PhysicalReadArrowDevice, ResultCollector,
                                run_pipeline)
from bodo_b200.table import Table
from tests.helpers import table_to_device

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run_fp(df, predicate, outputs):
    op = PhysicalFilterProject(predicate, outputs)
    out, _ = op.ProcessBatch(table_to_device(Table.from_pandas(df)), OperatorResult.FINISHED)
    return out.to_pandas()


def test_filter_project_arithmetic_comparisons_nulls(gpu_lib):
    rng = np.random.default_rng(2)
    n = 100_003
    df = pd.DataFrame({"a": rng.integers(-50, 50, n).astype(np.int64), "b": rng.random(n) * 10,
                       "c": pd.array(rng.integers(0, 9, n), dtype="Int32"), "d": pd.array(rng.random(n), dtype="Float64"),
                       "day": pd.to_datetime("1998-01-01").to_numpy().astype("datetime64[D]") + rng.integers(0, 700, n)})
    df.loc[rng.random(n) < 0.1, "c"] = pd.NA
    df.loc[rng.random(n) < 0.1, "d"] = pd.NA
    df["day"] = df["day"].dt.date if hasattr(df["day"], "dt") else df["day"]
    t = pa.table({"a": df.a.to_numpy(), "b": df.b.to_numpy(), "c": pa.array(df.c), "d": pa.array(df.d),
                  "day": pa.array(pd.to_datetime(df.day).dt.date, type=pa.date32())})
    pred = ((col("a") > -10) & (col("b") * 2.0 <= 15.0) & (col("day") <= lit(datetime.date(1999, 3, 1)))) | (col("c") == 7)
    outs = [("a", col("a")), ("e", col("b") * (lit(1.0) - col("d"))), ("f", col("a") * 3 + col("c")), ("g", col("a") / 4), ("c", col("c")),
            ("isn", col("d").isnull())]
    op = PhysicalFilterProject(pred, outs)
    from bodo_b200.streaming.dist_join import to_device
    got_t, _ = op.ProcessBatch(to_device(Table.from_arrow(t), 0), OperatorResult.FINISHED)
    got = got_t.to_pandas()
    days = pd.to_datetime(df.day)
    m = (((df.a > -10) & (df.b * 2.0 <= 15.0) & (days <= pd.Timestamp("1999-03-01"))) | (df.c == 7).fillna(False)).to_numpy(dtype=bool)
    e = df[m]
    exp = pd.DataFrame({"a": e.a.to_numpy(), "e": (e.b * (1.0 - e.d)).to_numpy(dtype="float64", na_value=np.nan),
                        "f": (e.a * 3 + e.c).to_numpy(dtype="float64", na_value=np.nan), "g": (e.a / 4).to_numpy(),
                        "c": e.c.to_numpy(dtype="float64", na_value=np.nan), "isn": e.d.isna().to_numpy().astype(np.float64)})
    assert len(got) == len(exp)
    g = pd.DataFrame({c: got[c].to_numpy(dtype="float64", na_value=np.nan) for c in got.columns})
    key = list(exp.columns)
    pd.testing.assert_frame_equal(g.sort_values(key).reset_index(drop=True), exp.sort_values(key).reset_index(drop=True), check_dtype=False, rtol=1e-12)


def test_dictionary_unification_across_batches(gpu_lib):
    b = DictionaryBuilder()
    batches = [pa.array(["x", "y", None, "x"]), pa.array(["z", "y", "y"]).dictionary_encode(), pa.array(["w", None, "x"], type=pa.large_string())]
    ids = []
    for arr in batches:
        c = b.unify(arr, 0)
        v = c.values_numpy()
        mask = c.valid_mask_numpy()
        ids.append((v, mask))
    assert b.values == ["x", "y", "z", "w"]  # first-appearance order across batches (DictionaryBuilder::InsertIfNotExists)
    dec = [list(b.decode(v, m)) for v, m in ids]
    assert dec == [["x", "y", None, "x"], ["z", "y", "y"], ["w", None, "x"]]


def _tpch_q1_pandas(lineitem: pd.DataFrame) -> pd.DataFrame:
    # benchmarks/tpch/bodo/dataframe_queries.py:96-119, verbatim semantics in pandas
    filt = lineitem[lineitem["L_SHIPDATE"] <= datetime.date(1998, 9, 2)].copy()
    filt["DISC_PRICE"] = filt.L_EXTENDEDPRICE * (1.0 - filt.L_DISCOUNT)
    filt["CHARGE"] = filt.L_EXTENDEDPRICE * (1.0 - filt.L_DISCOUNT) * (1.0 + filt.L_TAX)
    gb = filt.groupby(["L_RETURNFLAG", "L_LINESTATUS"], as_index=False)
    agg = gb.agg(SUM_QTY=pd.NamedAgg(column="L_QUANTITY", aggfunc="sum"), SUM_BASE_PRICE=pd.NamedAgg(column="L_EXTENDEDPRICE", aggfunc="sum"),
                 SUM_DISC_PRICE=pd.NamedAgg(column="DISC_PRICE", aggfunc="sum"), SUM_CHARGE=pd.NamedAgg(column="CHARGE", aggfunc="sum"),
                 AVG_QTY=pd.NamedAgg(column="L_QUANTITY", aggfunc="mean"), AVG_PRICE=pd.NamedAgg(column="L_EXTENDEDPRICE", aggfunc="mean"),
                 AVG_DISC=pd.NamedAgg(column="L_DISCOUNT", aggfunc="mean"), COUNT_ORDER=pd.NamedAgg(column="L_ORDERKEY", aggfunc="size"))
    return agg.sort_values(["L_RETURNFLAG", "L_LINESTATUS"]).reset_index(drop=True)


@pytest.mark.parametrize("batch_size", [4096, 1 << 20])
def test_tpch_q1_on_the_reference_fixture(gpu_lib, batch_size):
    """read (Arrow) -> filter L_SHIPDATE <= 1998-09-02 + DISC_PRICE / CHARGE projections (one kernel) -> groupby on two
    dictionary-encoded string keys with 4 sums, 3 means, size -> decode; equals pandas on the same rows (rtol 1e-9)."""
    at = pq.read_table(os.path.join(GOLDEN, "tpch_q1_lineitem.parquet"))
    builders = {"L_RETURNFLAG": DictionaryBuilder(), "L_LINESTATUS": DictionaryBuilder()}
    src = PhysicalReadArrowDevice(at, batch_size, 0, builders)
    price, disc, tax = col("L_EXTENDEDPRICE"), col("L_DISCOUNT"), col("L_TAX")
    fp = PhysicalFilterProject(col("L_SHIPDATE") <= lit(datetime.date(1998, 9, 2)),
                               [("L_RETURNFLAG", col("L_RETURNFLAG")), ("L_LINESTATUS", col("L_LINESTATUS")), ("L_QUANTITY", col("L_QUANTITY")),
                                ("L_EXTENDEDPRICE", price), ("DISC_PRICE", price * (lit(1.0) - disc)),
                                ("CHARGE", price * (lit(1.0) - disc) * (lit(1.0) + tax)), ("L_DISCOUNT", disc), ("L_ORDERKEY", col("L_ORDERKEY"))])
    aggs = [("sum", 2), ("sum", 3), ("sum", 4), ("sum", 5), ("mean", 2), ("mean", 3), ("mean", 6), ("size", None)]
    agg = PhysicalAggregate((0, 1), aggs)
    run_pipeline(src, [fp], agg)
    coll = ResultCollector()
    run_pipeline(agg, [], coll)
    agg.Finalize()
    got = coll.result()
    got.columns = ["L_RETURNFLAG", "L_LINESTATUS", "SUM_QTY", "SUM_BASE_PRICE", "SUM_DISC_PRICE", "SUM_CHARGE", "AVG_QTY", "AVG_PRICE", "AVG_DISC", "COUNT_ORDER"]
    for nm in ("L_RETURNFLAG", "L_LINESTATUS"):
        got[nm] = builders[nm].decode(got[nm].to_numpy(dtype="int64"))
    got = got.sort_values(["L_RETURNFLAG", "L_LINESTATUS"]).reset_index(drop=True)
    exp = _tpch_q1_pandas(at.to_pandas())
    assert list(got["L_RETURNFLAG"]) == list(exp["L_RETURNFLAG"]) and list(got["L_LINESTATUS"]) == list(exp["L_LINESTATUS"])
    assert (got["COUNT_ORDER"].to_numpy(dtype="int64") == exp["COUNT_ORDER"].to_numpy()).all()
    for c in exp.columns[2:-1]:
        np.testing.assert_allclose(got[c].to_numpy(dtype="float64"), exp[c].to_numpy(dtype="float64"), rtol=1e-9, err_msg=c)
