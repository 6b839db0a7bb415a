"""CPU tests (no GPU): the C-ABI library loads and exports every declared symbol, refuses to compute without a device,
and the host-side table / operator plumbing behaves."""

import ctypes
import os

import numpy as np
import pandas as pd
import pytest

from bodo_b200 import B200Error, _lib
from bodo_b200.table import ArrTypes, CTable, CTypes, Table


def test_library_exports_every_declared_symbol():
    syms = _lib.declared_symbols()
    assert len(syms) >= 25
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert _lib.lib().b200_abi_version() == 1


def test_header_cites_reference_for_every_entry_point():
    src = open(_lib.HEADER).read()
    for needle in ("_groupby.cpp:4917-4970", "_groupby.cpp:4663-4676", "_groupby.cpp:4772-4783", "_join.cpp:4087-4136",
                   "_join.cpp:4149-4185", "_join.cpp:4205-4260", "_shuffle.cpp:94-163"):
        assert needle in src, needle


def test_no_cpu_fallback_without_gpu():
    L = _lib.lib()
    if L.b200_device_count() > 0:
        pytest.skip("a GPU is visible")
    from bodo_b200.streaming.groupby import groupby_build_consume_batch, init_groupby_state
    st = init_groupby_state(-1, (0,), ("sum",), (0, 1), (1,))
    t = Table.from_pandas(pd.DataFrame({"a": [1, 2], "b": [3, 4]}))
    with pytest.raises(B200Error, match="no CUDA device|no CPU fallback|CUDA-only"):
        groupby_build_consume_batch(st, t, True, True)
    from bodo_b200.streaming.join import init_join_state, join_build_consume_batch, join_probe_consume_batch
    js = init_join_state(-1, (0,), (0,), ("a", "b"), ("a", "b"), False, False)
    with pytest.raises(B200Error, match="no CUDA device|no CPU fallback|CUDA-only"):
        join_build_consume_batch(js, t, True)  # build batches go straight to the device: the first one already fails
    with pytest.raises(B200Error):
        join_probe_consume_batch(js, t, True)


def test_product_package_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "bodo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in txt and "import oracle" not in txt and "libbodo_oracle" not in txt, f


def test_temporal_columns_are_nanoseconds_whatever_unit_pandas_hands_out():
    # pandas >= 3 defaults to datetime64[us]; DATETIME / TIMEDELTA columns are int64 ns (Bodo_CTypes)
    for unit in ("s", "ms", "us", "ns"):
        ts = pd.Series(np.array(["2020-01-01T00:00:01", "1999-12-31T23:59:59", "NaT"], dtype=f"datetime64[{unit}]"))
        td = pd.Series(np.array([1, -5, 86400], dtype=f"timedelta64[{unit}]"))
        t = Table.from_pandas(pd.DataFrame({"ts": ts, "td": td}))
        assert [c.c_type for c in t.columns] == [CTypes.DATETIME, CTypes.TIMEDELTA]
        assert t.columns[0].data[0] == np.datetime64("2020-01-01T00:00:01", "ns").astype("int64")
        back = t.to_pandas()
        assert back["ts"][0] == pd.Timestamp("2020-01-01T00:00:01") and back["ts"][1] == pd.Timestamp("1999-12-31T23:59:59")
        assert pd.isna(back["ts"][2])
        assert back["td"][2] == pd.Timedelta(86400, unit=unit)
    from bodo_b200.table import Column
    c = Column(np.array(["2021-06-01"], dtype="datetime64[D]"))
    assert c.c_type == CTypes.DATETIME and c.data.dtype == np.int64 and c.data[0] == np.datetime64("2021-06-01", "ns").astype("int64")


def test_table_from_pandas_nullable_and_numpy():
    df = pd.DataFrame({"a": pd.array([1, None, 3], dtype="Int64"), "b": np.array([1.0, 2.0, 3.0]), "c": np.array([1, 2, 3], dtype=np.int32)})
    t = Table.from_pandas(df)
    assert [c.c_type for c in t.columns] == [CTypes.INT64, CTypes.FLOAT64, CTypes.INT32]
    assert [c.arr_type for c in t.columns] == [ArrTypes.NULLABLE_INT_BOOL, ArrTypes.NUMPY, ArrTypes.NUMPY]
    assert t.columns[0].valid_mask_numpy().tolist() == [True, False, True]
    s = t.slice(1, 3)
    assert s.n_rows == 2 and s.columns[0].valid_mask_numpy().tolist() == [False, True]
    rt = t.to_pandas()
    assert rt["a"].isna().tolist() == [False, True, False] and rt["c"].tolist() == [1, 2, 3]
    ct = CTable(t)
    assert ct.ptr.n_rows == 3 and ct.ptr.n_cols == 3 and ct.ptr.device == -1


def test_table_from_arrow_zero_copy():
    import pyarrow as pa
    tbl = pa.table({"k": pa.array([1, 2, None, 4], type=pa.int64()), "v": pa.array([1.0, 2.0, 3.0, 4.0])})
    t = Table.from_arrow(tbl)
    assert t.columns[0].valid_mask_numpy().tolist() == [True, True, False, True]
    assert t.columns[1].validity is None and t.columns[1].data.tolist() == [1.0, 2.0, 3.0, 4.0]


def test_unsupported_inputs_fail_loudly():
    from bodo_b200.streaming.groupby import init_groupby_state
    with pytest.raises(B200Error, match="unsupported aggregate function"):
        init_groupby_state(-1, (0,), ("median",), (0, 1), (1,))
    with pytest.raises(B200Error, match="min_row_number_filter"):
        init_groupby_state(-1, (0,), ("sum",), (0, 1), (1,), mrnf_sort_col_inds=(1,))
    with pytest.raises(TypeError, match="object dtype|unsupported"):
        Table.from_pandas(pd.DataFrame({"s": ["a", "b"]}))


def test_expression_programs_and_dictionary_decode_host_logic():
    import datetime
    import struct

    from bodo_b200.dictionary import DictionaryBuilder
    from bodo_b200.expr import OPS, col, compile_program, lit
    e = (col("a") * (lit(1.0) - col("b")) <= 3) & ~col("d").isnull()
    prog, starts = compile_program([e, col("d"), lit(datetime.date(1970, 1, 11))], {"a": 0, "b": 1, "d": 2})
    assert starts == [0, 12, 14] and prog[-1] == (OPS["end"], 0) and prog[-2] == (OPS["const_i64"], 10)
    ops = [o for o, _ in prog[:12]]
    assert ops == [OPS[x] for x in ("col", "const_f64", "col", "sub", "mul", "const_i64", "le", "col", "is_null", "not", "and", "end")]
    assert prog[1][1] == struct.unpack("<q", struct.pack("<d", 1.0))[0] and e.columns() == {"a", "b", "d"}
    with pytest.raises(KeyError):
        compile_program([col("zz")], {"a": 0})
    b = DictionaryBuilder()
    b.values, b.index = ["N", "R"], {"N": 0, "R": 1}
    assert list(b.decode(np.array([1, 0, 1]), np.array([True, True, False]))) == ["R", "N", None]
