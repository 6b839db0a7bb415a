"""Shared helpers for the parity tests (streaming loops shaped like the reference's tests)."""

import numpy as np
import pandas as pd

from bodo_b200.streaming.groupby import (delete_groupby_state, groupby_build_consume_batch,
                                         groupby_produce_output_batch, init_groupby_state)
from bodo_b200.table import Table


def stream_groupby(table: Table, key_inds, fnames, f_in_offsets, f_in_cols, batch_size=None, to_device=False, **kw):
    """The reference's streaming test loop (bodo/tests/test_streaming/test_groupby.py:51-81): init, feed
    batch_size-row slices until is_last, pop output batches, delete the state; returns one pandas frame."""
    state = init_groupby_state(-1, key_inds, fnames, f_in_offsets, f_in_cols, **kw)
    n = table.n_rows
    bs = batch_size or max(n, 1)
    is_last = False
    it = 0
    while not is_last:
        batch = table.slice(it * bs, (it + 1) * bs)
        if to_device:
            batch = table_to_device(batch)
        is_last = (it + 1) * bs >= n
        it += 1
        is_last, _ = groupby_build_consume_batch(state, batch, is_last, True)
    outs = []
    out_last = False
    while not out_last:
        out, out_last = groupby_produce_output_batch(state, True)
        outs.append(out.to_pandas())
    delete_groupby_state(state)
    return pd.concat(outs, ignore_index=True)


def table_to_device(t: Table, device=0) -> Table:
    import torch

    from bodo_b200.table import Column

    cols = []
    for c in t.columns:
        d = torch.from_numpy(np.ascontiguousarray(c.data)).to(f"cuda:{device}")
        v = None
        if c.validity is not None:
            # pad to a multiple of 8 bytes so device-side word reads stay in bounds
            vb = np.zeros((len(c.validity) + 7) // 8 * 8, dtype=np.uint8)
            vb[: len(c.validity)] = c.validity
            v = torch.from_numpy(vb).to(f"cuda:{device}")
        cols.append(Column(d, v, c.c_type, c.arr_type, c.length))
    return Table(cols, list(t.names))


def positional(df: pd.DataFrame) -> pd.DataFrame:
    """Rename columns positionally (key, f0, f1, ...) so frames from different sources line up."""
    df = df.copy()
    df.columns = ["key"] + [f"f{j}" for j in range(df.shape[1] - 1)]
    return df


def oracle_groupby_frame(O, table: Table, key_ind, fnames, in_cols, dropna=True, batch_size=32768, n_pes=1, rank=0):
    """Run the CPU oracle on the same host table; returns a positional pandas frame (key, f0, f1, ...)."""
    kc = table.columns[key_ind]
    kvalid = kc.valid_mask_numpy()
    vals, vvalids = [], []
    for f, ci in zip(fnames, in_cols):
        c = table.columns[ci if ci is not None else key_ind]
        vals.append(c.data)
        vvalids.append(c.valid_mask_numpy() if ci is not None else None)
    keys = kc.data.astype(np.int64) if kc.data.dtype != np.int64 else kc.data
    r = O.groupby(keys, kvalid, list(fnames), vals, vvalids, dropna=dropna, batch_size=batch_size, n_pes=n_pes, rank=rank)
    out = {}
    if kvalid is not None:
        out["key"] = pd.array(pd.arrays.IntegerArray(r["keys"].copy(), ~r["key_valid"]))
    else:
        out["key"] = r["keys"]
    for j in range(len(fnames)):
        data, valid = r["cols"][j]
        if valid.all():
            out[f"f{j}"] = data
        elif data.dtype.kind == "f":
            d = data.copy()
            d[~valid] = np.nan
            out[f"f{j}"] = d
        else:
            out[f"f{j}"] = pd.array(pd.arrays.IntegerArray(data.copy(), ~valid))
    return pd.DataFrame(out)


def sort_frame(df: pd.DataFrame) -> pd.DataFrame:
    """sort_output=True, reset_index=True of the reference's check_func (bodo/tests/utils.py:157-186)."""
    return df.sort_values(list(df.columns), na_position="last").reset_index(drop=True)


def assert_frames_equal(got: pd.DataFrame, exp: pd.DataFrame, rtol=1e-5, atol=1e-8):
    """_test_equal with check_dtype=False: integer columns bit-exact, float columns within
    rtol=1e-5 / atol=1e-8 (the reference's defaults, bodo/tests/utils.py:179-180)."""
    assert list(got.columns) == list(exp.columns), (list(got.columns), list(exp.columns))
    assert len(got) == len(exp), (len(got), len(exp))
    g, e = sort_frame(got), sort_frame(exp)
    for c in g.columns:
        if g[c].dtype.kind in "iu" and e[c].dtype.kind in "iu" and (g[c].isna().any() or e[c].isna().any()):
            # nullable integer columns: same NA positions, the other values bit-exact
            np.testing.assert_array_equal(g[c].isna().to_numpy(), e[c].isna().to_numpy(), err_msg=f"column {c} (NA mask)")
            np.testing.assert_array_equal(g[c].fillna(0).to_numpy(dtype=np.int64), e[c].fillna(0).to_numpy(dtype=np.int64), err_msg=f"column {c}")
            continue
        gv = g[c].to_numpy(dtype="float64", na_value=np.nan) if g[c].dtype.kind not in "iu" or g[c].isna().any() else g[c].to_numpy()
        ev = e[c].to_numpy(dtype="float64", na_value=np.nan) if e[c].dtype.kind not in "iu" or e[c].isna().any() else e[c].to_numpy()
        if gv.dtype.kind in "iu" and ev.dtype.kind in "iu":
            np.testing.assert_array_equal(gv.astype(np.int64), ev.astype(np.int64), err_msg=f"column {c}")
        else:
            np.testing.assert_allclose(gv.astype(np.float64), ev.astype(np.float64), rtol=rtol, atol=atol, equal_nan=True,
                                       err_msg=f"column {c}")
