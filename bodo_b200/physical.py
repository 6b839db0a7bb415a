"""Physical-operator layer — the Python mirror of the reference's pipeline protocol ("door 2",
bodo/pandas/physical/operator.h:46-50,247-458; aggregate.h:65-365; join.h:58-744; _pipeline.cpp:389-466).

    OperatorResult.{NEED_MORE_INPUT, HAVE_MORE_OUTPUT, FINISHED}
    PhysicalReadPandas / PhysicalReadArrow / PhysicalReadParquet : sources (batches of host columns)
    PhysicalAggregate : sink of one pipeline (ConsumeBatch) and source of the next (ProduceBatch)
    PhysicalJoin      : sink for the build side, ProcessBatch for the probe side
    Pipeline          : while not finished: batch = source.ProduceBatch(); ... sink.ConsumeBatch(batch)

plus two helpers that run those pipelines over pandas frames the way bodo.pandas does for
`df.groupby(keys)[cols].agg(...)` and `left.merge(right, on=...)` (PhysicalReadPandas slices the frame into
STREAMING_BATCH_SIZE-row Arrow batches, bodo/pandas/physical/read_pandas.h:13-120).  All computation happens in
libbodo_b200.so; this module only moves batches between operators.
"""

from __future__ import annotations

import enum
from typing import Iterable, Sequence

from .streaming import groupby as G
from .streaming import join as J
from .table import Table

STREAMING_BATCH_SIZE = 32768  # bodo/libs/streaming/_shuffle.h:27-31


class OperatorResult(enum.Enum):
    NEED_MORE_INPUT = 0
    HAVE_MORE_OUTPUT = 1
    FINISHED = 2


class PhysicalReadPandas:
    """Source: slices a pandas DataFrame into batch_size-row batches (read_pandas.h:13-120)."""

    def __init__(self, df, batch_size: int = STREAMING_BATCH_SIZE):
        self.table = Table.from_pandas(df)
        self.batch_size = batch_size
        self.cur = 0

    def ProduceBatch(self):
        n = self.table.n_rows
        batch = self.table.slice(self.cur, self.cur + self.batch_size)
        self.cur += self.batch_size
        return batch, (OperatorResult.FINISHED if self.cur >= n else OperatorResult.HAVE_MORE_OUTPUT)


class PhysicalReadArrow:
    """Source over an in-memory pyarrow Table: batch_size-row zero-copy slices (the Arrow half of
    PhysicalReadPandas, read_pandas.h:13-120: the reference converts every pandas slice to Arrow first)."""

    def __init__(self, table, batch_size: int = STREAMING_BATCH_SIZE):
        self.arrow = table.combine_chunks()
        self.batch_size = batch_size
        self.cur = 0

    def ProduceBatch(self):
        n = self.arrow.num_rows
        batch = Table.from_arrow(self.arrow.slice(self.cur, self.batch_size))
        self.cur += self.batch_size
        return batch, (OperatorResult.FINISHED if self.cur >= n else OperatorResult.HAVE_MORE_OUTPUT)


class PhysicalReadParquet:
    """Source over a Parquet file or dataset directory (host side of physical/read_parquet.h:31-210 — the reference
    streams Arrow record batches out of its ParquetReader; here pyarrow's reader produces them).  Only the selected
    columns are decoded; every batch has at most batch_size rows; an empty dataset yields one empty, FINISHED batch."""

    def __init__(self, path: str, columns: Sequence[str] | None = None, batch_size: int = STREAMING_BATCH_SIZE):
        import pyarrow.dataset as ds

        self.dataset = ds.dataset(path, format="parquet")
        self.columns = list(columns) if columns is not None else None
        self.schema = self.dataset.schema
        self._it = iter(self.dataset.to_batches(columns=self.columns, batch_size=batch_size))
        self._next = self._pull()

    def _pull(self):
        for rb in self._it:
            if rb.num_rows:
                return rb
        return None

    def ProduceBatch(self):
        import pyarrow as pa

        cur = self._next
        if cur is None:  # empty dataset
            names = self.columns if self.columns is not None else list(self.schema.names)
            empty = pa.table({n: pa.array([], type=self.schema.field(n).type) for n in names})
            return Table.from_arrow(empty), OperatorResult.FINISHED
        self._next = self._pull()
        return Table.from_arrow(cur), (OperatorResult.FINISHED if self._next is None else OperatorResult.HAVE_MORE_OUTPUT)


class PhysicalAggregate:
    """Groupby sink/source (aggregate.h:65-365). `aggs` = [(func_name, input_column_index or None for size)]."""

    def __init__(self, key_inds: Sequence[int], aggs: Sequence[tuple], dropna: bool = True, parallel: bool = False, **kw):
        fnames = tuple(f for f, _ in aggs)
        f_in_offsets, f_in_cols = [0], []
        for _, c in aggs:
            if c is not None:
                f_in_cols.append(c)
            f_in_offsets.append(len(f_in_cols))
        self.state = G.init_groupby_state(-1, tuple(key_inds), fnames, tuple(f_in_offsets), tuple(f_in_cols), parallel=parallel,
                                          dropna=dropna, **kw)
        self.finished_build = False

    def ConsumeBatch(self, batch: Table, prev: OperatorResult) -> OperatorResult:
        is_last = prev == OperatorResult.FINISHED
        global_last, _ = G.groupby_build_consume_batch(self.state, batch, is_last, True)
        self.finished_build = global_last
        return OperatorResult.FINISHED if global_last else OperatorResult.NEED_MORE_INPUT

    def ProduceBatch(self):
        out, last = G.groupby_produce_output_batch(self.state, True)
        return out, (OperatorResult.FINISHED if last else OperatorResult.HAVE_MORE_OUTPUT)

    def Finalize(self):
        G.delete_groupby_state(self.state)


class PhysicalJoin:
    """Hash join: sink for build batches, ProcessBatch for probe batches (join.h:58-744)."""

    def __init__(self, build_key: int, probe_key: int, build_names, probe_names, how: str = "inner", **kw):
        build_outer = how in ("right", "outer")   # the build side is the RIGHT table (reference convention)
        probe_outer = how in ("left", "outer")
        kw.setdefault("is_na_equal", True)  # pandas merge semantics: NA joins NA (bodo/pandas/physical/join.h:267)
        self.state = J.init_join_state(-1, (build_key,), (probe_key,), tuple(build_names), tuple(probe_names), build_outer, probe_outer, **kw)

    def ConsumeBatch(self, batch: Table, prev: OperatorResult) -> OperatorResult:
        is_last = prev == OperatorResult.FINISHED
        J.join_build_consume_batch(self.state, batch, is_last)
        return OperatorResult.FINISHED if is_last else OperatorResult.NEED_MORE_INPUT

    def ProcessBatch(self, batch: Table, prev: OperatorResult):
        is_last = prev == OperatorResult.FINISHED
        out, out_last, _ = J.join_probe_consume_batch(self.state, batch, is_last, True)
        return out, (OperatorResult.FINISHED if out_last else OperatorResult.NEED_MORE_INPUT)

    def Finalize(self):
        J.delete_join_state(self.state)


class ResultCollector:
    """PhysicalResultCollector: concatenates output batches into one pandas frame."""

    def __init__(self):
        self.frames = []

    def ConsumeBatch(self, batch: Table, prev: OperatorResult) -> OperatorResult:
        self.frames.append(batch.to_pandas())
        return OperatorResult.FINISHED if prev == OperatorResult.FINISHED else OperatorResult.NEED_MORE_INPUT

    def result(self):
        import pandas as pd

        return pd.concat(self.frames, ignore_index=True) if self.frames else pd.DataFrame()


def run_pipeline(source, between: Iterable, sink) -> None:
    """Pipeline::Execute (bodo/pandas/_pipeline.cpp:389-466): push batches source -> between ops -> sink."""
    finished = False
    while not finished:
        batch, res = source.ProduceBatch()
        for op in between:
            batch, res2 = op.ProcessBatch(batch, res)
            if res == OperatorResult.FINISHED and res2 != OperatorResult.FINISHED:
                res2 = OperatorResult.FINISHED
            res = res2
        sink.ConsumeBatch(batch, res)
        finished = res == OperatorResult.FINISHED


def groupby_agg(df, by, aggs: Sequence[tuple], dropna: bool = True, batch_size: int = STREAMING_BATCH_SIZE, **kw):
    """df.groupby(by, as_index=False, dropna=dropna).agg(...) through the streaming operators.

    aggs: [(out_name, column, func)] with func in {'sum','count','mean','min','max','size'}.
    Returns a pandas DataFrame (group order unspecified, as in the reference)."""
    by = [by] if isinstance(by, str) else list(by)
    cols = list(df.columns)
    used = list(by)
    for _, c, _ in aggs:
        if c is not None and c not in used:
            used.append(c)
    sub = df[used]
    key_inds = [used.index(k) for k in by]
    agg_spec = [(f, None if f == "size" or c is None else used.index(c)) for _, c, f in aggs]
    op = PhysicalAggregate(key_inds, agg_spec, dropna=dropna, **kw)
    run_pipeline(PhysicalReadPandas(sub, batch_size), [], op)
    coll = ResultCollector()
    run_pipeline(op, [], coll)
    op.Finalize()
    out = coll.result()
    out.columns = by + [name for name, _, _ in aggs]
    return out


def groupby_agg_parquet(path: str, by, aggs: Sequence[tuple], dropna: bool = True, batch_size: int = STREAMING_BATCH_SIZE, **kw):
    """bodo.pandas.read_parquet(path).groupby(by, as_index=False, dropna=dropna).agg(...): PhysicalReadParquet feeding
    PhysicalAggregate.  Only the key and aggregated columns are decoded (column pruning, as the reference's planner does
    for ReadParquet under an aggregate).  Same `aggs` format and result shape as groupby_agg."""
    by = [by] if isinstance(by, str) else list(by)
    used = list(by)
    for _, c, _ in aggs:
        if c is not None and c not in used:
            used.append(c)
    key_inds = [used.index(k) for k in by]
    agg_spec = [(f, None if f == "size" or c is None else used.index(c)) for _, c, f in aggs]
    op = PhysicalAggregate(key_inds, agg_spec, dropna=dropna, **kw)
    run_pipeline(PhysicalReadParquet(path, used, batch_size), [], op)
    coll = ResultCollector()
    run_pipeline(op, [], coll)
    op.Finalize()
    out = coll.result()
    out.columns = by + [name for name, _, _ in aggs]
    return out


def merge(left, right, left_on: str, right_on: str, how: str = "inner", batch_size: int = STREAMING_BATCH_SIZE, **kw):
    """left.merge(right, left_on=..., right_on=..., how=...) through the streaming join (right = build side).
    Output columns: right's columns then left's columns (the reference's build-then-probe order), renamed on clashes."""
    rcols, lcols = list(right.columns), list(left.columns)
    op = PhysicalJoin(rcols.index(right_on), lcols.index(left_on), rcols, lcols, how=how, **kw)
    run_pipeline(PhysicalReadPandas(right, batch_size), [], op)
    coll = ResultCollector()
    run_pipeline(PhysicalReadPandas(left, batch_size), [op], coll)
    op.Finalize()
    return coll.result()
