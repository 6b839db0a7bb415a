"""Host-side source operators (bodo_b200/physical.py) and the Arrow -> Table boundary: no GPU needed.
Reference shapes: PhysicalReadPandas (bodo/pandas/physical/read_pandas.h:13-120), PhysicalReadParquet
(physical/read_parquet.h:31-210), Bodo_CTypes codes for temporal columns (bodo/libs/_bodo_common.h:331-359)."""
import numpy as np
import pandas as pd
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from bodo_b200.physical import OperatorResult, PhysicalReadArrow, PhysicalReadPandas, PhysicalReadParquet
from bodo_b200.table import ArrTypes, CTypes, Table


def _drain(src):
    frames, n_batches = [], 0
    while True:
        batch, res = src.ProduceBatch()
        n_batches += 1
        frames.append(batch.to_pandas())
        assert res in (OperatorResult.HAVE_MORE_OUTPUT, OperatorResult.FINISHED)
        if res == OperatorResult.FINISHED:
            return pd.concat(frames, ignore_index=True), n_batches


def _arrow_table(n):
    rng = np.random.default_rng(3)
    k = pa.array([None if i % 97 == 5 else int(v) for i, v in enumerate(rng.integers(0, 50, n))], type=pa.int64())
    return pa.table({"k": k, "v": rng.random(n), "w": rng.integers(-9, 9, n).astype("int32"),
                     "d": pa.array(rng.integers(0, 20000, n).astype("int32")).cast(pa.date32()),
                     "ts": pa.array(rng.integers(0, 2**40, n), type=pa.timestamp("us"))})


def test_read_parquet_batches_match_the_file(tmp_path):
    t = _arrow_table(10_000)
    path = str(tmp_path / "a.parquet")
    pq.write_table(t, path, row_group_size=3000)
    got, nb = _drain(PhysicalReadParquet(path, batch_size=1024))
    assert nb >= 10 and len(got) == 10_000
    exp = t.to_pandas()
    assert got["k"].isna().sum() == exp["k"].isna().sum()
    np.testing.assert_array_equal(got["k"].to_numpy(dtype="float64", na_value=np.nan), exp["k"].to_numpy(dtype="float64", na_value=np.nan))
    np.testing.assert_array_equal(got["v"].to_numpy(dtype="float64"), exp["v"].to_numpy())
    np.testing.assert_array_equal(got["w"].to_numpy(dtype="int64"), exp["w"].to_numpy())
    np.testing.assert_array_equal(got["ts"].to_numpy(), exp["ts"].to_numpy().astype("datetime64[ns]"))
    np.testing.assert_array_equal(got["d"].to_numpy().astype("datetime64[D]"), pd.to_datetime(exp["d"]).to_numpy().astype("datetime64[D]"))


def test_read_parquet_prunes_columns_and_keeps_their_order(tmp_path):
    t = _arrow_table(2_000)
    path = str(tmp_path / "a.parquet")
    pq.write_table(t, path)
    src = PhysicalReadParquet(path, columns=["w", "k"], batch_size=512)
    batch, _ = src.ProduceBatch()
    assert batch.names == ["w", "k"] and batch.n_rows <= 512
    assert [c.c_type for c in batch.columns] == [CTypes.INT32, CTypes.INT64]
    assert all(c.arr_type == ArrTypes.NULLABLE_INT_BOOL for c in batch.columns)   # Arrow columns are nullable by construction


def test_read_parquet_empty_file_yields_one_finished_empty_batch(tmp_path):
    t = _arrow_table(0)
    path = str(tmp_path / "empty.parquet")
    pq.write_table(t, path)
    batch, res = PhysicalReadParquet(path, columns=["k", "v"]).ProduceBatch()
    assert res == OperatorResult.FINISHED and batch.n_rows == 0 and batch.names == ["k", "v"]


def test_read_parquet_dataset_directory(tmp_path):
    t = _arrow_table(3_000)
    for i in range(3):
        pq.write_table(t.slice(i * 1000, 1000), str(tmp_path / f"part-{i}.parquet"))
    got, _ = _drain(PhysicalReadParquet(str(tmp_path), columns=["w"], batch_size=400))
    assert len(got) == 3000 and int(got["w"].to_numpy(dtype="int64").sum()) == int(t["w"].to_numpy().sum())


@pytest.mark.parametrize("n,bs", [(10, 3), (9, 3), (1, 32768), (0, 8)])
def test_read_arrow_and_pandas_sources_agree(n, bs):
    t = _arrow_table(n).select(["k", "v", "w"])
    a, na = _drain(PhysicalReadArrow(t, bs))
    p, npd = _drain(PhysicalReadPandas(t.to_pandas(types_mapper={pa.int64(): pd.Int64Dtype()}.get), bs))
    assert na == npd == max(1, -(-n // bs))
    assert len(a) == len(p) == n
    np.testing.assert_array_equal(a["k"].to_numpy(dtype="float64", na_value=np.nan), p["k"].to_numpy(dtype="float64", na_value=np.nan))
    np.testing.assert_array_equal(a["v"].to_numpy(dtype="float64"), p["v"].to_numpy(dtype="float64"))


def test_temporal_arrow_columns_keep_reference_dtype_codes():
    t = pa.table({"d": pa.array([1, None, 3], type=pa.int32()).cast(pa.date32()),
                  "ts": pa.array([1, 2, None], type=pa.timestamp("ms")),
                  "td": pa.array([5, None, 7], type=pa.duration("s"))})
    tab = Table.from_arrow(t)
    assert [c.c_type for c in tab.columns] == [CTypes.DATE, CTypes.DATETIME, CTypes.TIMEDELTA]
    assert tab.columns[0].data.dtype == np.int32 and tab.columns[1].data.dtype == np.int64
    assert int(tab.columns[1].data[1]) == 2_000_000 and int(tab.columns[2].data[0]) == 5_000_000_000   # brought to ns
    back = tab.to_pandas()
    assert back["d"].isna().tolist() == [False, True, False] and back["ts"].isna().tolist() == [False, False, True]
    assert back["td"].to_numpy()[0] == np.timedelta64(5, "s")


def test_strings_are_rejected_loudly():
    with pytest.raises(TypeError, match="unsupported Arrow type"):
        Table.from_arrow(pa.table({"s": ["a", "b"]}))
