"""Regenerates tests/golden/*.json: the inputs of the reference's own tests for this path and the expected
values those tests assert (the reference's oracle is pandas: check_func(..., py_output=<pandas>),
bodo/tests/utils.py:157-186).  Run in the authoring container (pandas 3.0.2); the GPU box only reads the JSON.

    python tests/golden/make_golden.py
"""

import json
import os

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))


def frame_to_json(df):
    out = {}
    for c in df.columns:
        v = df[c]
        out[str(c)] = [None if pd.isna(x) else (float(x) if isinstance(x, (float, np.floating)) else int(x)) for x in v.tolist()]
    return out


def groupby_cases():
    cases = []
    # test_groupby_basic (bodo/tests/test_streaming/test_groupby.py:83-108)
    groups = [1, 2, 1, 1, 2, 0, 1, 2] * 100
    data = [1, 3, 5, 11, 1, 3, 5, 3] * 100
    df = pd.DataFrame({"A": groups, "B": data})
    for f in ["sum", "mean", "count", "min", "max"]:
        exp = df.groupby("A", as_index=False).agg(f)
        cases.append({"name": f"test_groupby_basic[{f}]", "ref": "bodo/tests/test_streaming/test_groupby.py:83-108", "key": groups,
                      "val": data, "func": f, "dropna": True, "expected": frame_to_json(exp)})
    exp = df.groupby("A", as_index=False).size()
    cases.append({"name": "test_groupby_basic[size]", "ref": "bodo/tests/test_streaming/test_groupby.py:83-108", "key": groups, "val": data,
                  "func": "size", "dropna": True, "expected": frame_to_json(exp)})
    # quickstart example (bodo/tests/test_quickstart_docs.py:29-63): 2000 rows % 30 groups, max
    df = pd.DataFrame({"A": np.arange(2000) % 30, "B": np.arange(2000)})
    cases.append({"name": "test_quickstart[max]", "ref": "bodo/tests/test_quickstart_docs.py:29-63", "key": df.A.tolist(), "val": df.B.tolist(),
                  "func": "max", "dropna": True, "expected": frame_to_json(df.groupby("A", as_index=False).B.max())})
    # test_series_groupby (bodo/tests/test_df_lib/test_end_to_end.py:1606-1626): nullable keys with NA, Float64 values, dropna
    key = [1, 2, None, 1, 2, None, 7, 8]
    val = [1.5, None, 3.0, 4.0, 5.5, 6.0, None, 8.0]
    df = pd.DataFrame({"A": pd.array(key, dtype="Int64"), "B": pd.array(val, dtype="Float64")})
    for dropna in (True, False):
        for f in ["sum", "mean", "count", "min", "max"]:
            exp = df.groupby("A", as_index=False, dropna=dropna).agg(f)
            cases.append({"name": f"test_series_groupby[{f}-dropna={dropna}]", "ref": "bodo/tests/test_df_lib/test_end_to_end.py:1606-1626",
                          "key": key, "val": val, "func": f, "dropna": dropna, "expected": frame_to_json(exp)})
    # test_dataframe_groupby (:1631-1660): Int32 keys incl. INT32_MAX
    key = [2147483647, 1, 2147483647, 3, 1, 3, 3]
    val = [10, 20, 30, 40, 50, 60, 70]
    df = pd.DataFrame({"A": pd.array(key, dtype="Int32"), "B": pd.array(val, dtype="Int64")})
    for f in ["sum", "count", "max"]:
        cases.append({"name": f"test_dataframe_groupby[{f}]", "ref": "bodo/tests/test_df_lib/test_end_to_end.py:1631-1660", "key": key, "val": val,
                      "func": f, "dropna": True, "key_dtype": "int32", "expected": frame_to_json(df.groupby("A", as_index=False).agg(f))})
    # int64 wraparound (2**62 + 2**62 -> -2**63, reference -fwrapv)
    key, val = [0, 0, 1], [2**62, 2**62, 5]
    df = pd.DataFrame({"A": key, "B": val})
    cases.append({"name": "int64_sum_wraps", "ref": "SURVEY.md §8c / CMakeLists.txt:223 (-fwrapv)", "key": key, "val": val, "func": "sum",
                  "dropna": True, "expected": frame_to_json(df.groupby("A", as_index=False).B.sum())})
    return cases


def join_cases():
    cases = []
    # test_hash_join_non_nullable_outer (bodo/tests/test_streaming/test_join.py:922-940)
    b = pd.DataFrame({"A": [1, 2, 3, 4, 5] * 25, "B": [1, 2, 3, 4, 5] * 25})
    p = pd.DataFrame({"C": [2, 6] * 25, "D": [2, 6] * 25})
    for how, bo, po in [("inner", False, False), ("outer", True, True), ("left", True, False), ("right", False, True)]:
        exp = b.merge(p, left_on="A", right_on="C", how=how)
        cases.append({"name": f"test_hash_join_non_nullable_outer[{how}]", "ref": "bodo/tests/test_streaming/test_join.py:922-940",
                      "build_key": b.A.tolist(), "probe_key": p.C.tolist(), "build_outer": bo, "probe_outer": po, "n_rows": len(exp),
                      "key_sum": int(exp.A.fillna(0).sum() + exp.C.fillna(0).sum())})
    # test_merge (bodo/tests/test_df_lib/test_end_to_end.py:1255-1286): nullable keys [2,2,3] vs [2,3,8]
    l = pd.DataFrame({"A": pd.array([2, 2, 3], dtype="Int64")})
    r = pd.DataFrame({"C": pd.array([2, 3, 8], dtype="Int64")})
    for how, bo, po in [("inner", False, False), ("left", False, True), ("right", True, False), ("outer", True, True)]:
        exp = l.merge(r, left_on="A", right_on="C", how=how)  # build side = right table (reference convention)
        cases.append({"name": f"test_merge[{how}]", "ref": "bodo/tests/test_df_lib/test_end_to_end.py:1255-1286", "build_key": [2, 3, 8],
                      "probe_key": [2, 2, 3], "build_outer": bo, "probe_outer": po, "n_rows": len(exp),
                      "key_sum": int(exp.A.fillna(0).sum() + exp.C.fillna(0).sum())})
    # test_shuffle_batching (:4803-4816): 60 000-row 1:1 join
    cases.append({"name": "test_shuffle_batching", "ref": "bodo/tests/test_streaming/test_join.py:4803-4816", "build_key": "arange:60000",
                  "probe_key": "arange:60000", "build_outer": False, "probe_outer": False, "n_rows": 60000, "key_sum": 2 * (59999 * 60000 // 2)})
    return cases


def hash_cases():
    """Known-answer vectors of the reference's hash_inner_32 (vendored xxHash) for placement parity."""
    import ctypes as C
    so = os.path.join(HERE, "..", "..", "oracle", "_ref", "libref_xxh3.so")
    R = C.CDLL(so)
    R.ref_hash_inner_32_i64.restype = C.c_uint32
    R.ref_hash_inner_32_i64.argtypes = [C.c_int64, C.c_uint32]
    R.ref_hash_inner_32_i32.restype = C.c_uint32
    R.ref_hash_inner_32_i32.argtypes = [C.c_int32, C.c_uint32]
    rng = np.random.default_rng(42)
    keys64 = [0, 1, -1, 2**63 - 1, -(2**63), 42, 1000000] + [int(x) for x in rng.integers(-(2**63), 2**63 - 1, 200)]
    keys32 = [0, 1, -1, 2**31 - 1, -(2**31), 42] + [int(x) for x in rng.integers(-(2**31), 2**31 - 1, 100)]
    out = []
    for seed in (0xB0D01289, 0xB0D01286, 0xB0D01285):
        out.append({"seed": seed, "keys64": keys64, "hash64": [int(R.ref_hash_inner_32_i64(k, seed)) for k in keys64],
                    "keys32": keys32, "hash32": [int(R.ref_hash_inner_32_i32(k, seed)) for k in keys32]})
    return out


if __name__ == "__main__":
    json.dump({"generated_with": f"pandas {pd.__version__}", "cases": groupby_cases()}, open(os.path.join(HERE, "groupby.json"), "w"))
    json.dump({"generated_with": f"pandas {pd.__version__}", "cases": join_cases()}, open(os.path.join(HERE, "join.json"), "w"))
    json.dump({"source": "bodo/libs/vendored/xxhash.h compiled in place (oracle/_ref/libref_xxh3.so)", "vectors": hash_cases()},
              open(os.path.join(HERE, "xxh3_hash_inner_32.json"), "w"))
    print("golden fixtures written")
