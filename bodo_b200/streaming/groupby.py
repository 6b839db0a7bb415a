"""Streaming groupby operator API — the host-side mirror of bodo/libs/streaming/groupby.py.

Same four verbs, same argument meaning and the same calling protocol as the reference
(init_groupby_state :702-715, groupby_build_consume_batch :1295-1395, groupby_produce_output_batch
:1502-1600, delete_groupby_state), so the reference's streaming test loops
(bodo/tests/test_streaming/test_groupby.py:51-81) run unchanged against this module.  The work happens in
libbodo_b200.so (CUDA, sm_100a); there is no CPU implementation behind these calls.

Differences: the reference types the state at Numba compile time; here the build-table schema is taken
from the first consumed batch.  With parallel=True the state is one shard of a torch.distributed process
group (one process per GPU): the last consume call runs the hash-partition exchange that replaces the
reference's MPI shuffle (streaming/_shuffle.cpp:687-804).
"""

from __future__ import annotations



from .. import _lib
from .._lib import ffi
from ..table import CTable, Table, table_from_ctable

# names must match supported_agg_funcs positions / Bodo_FTypes (groupby/_groupby_ftypes.h:17-110)
FTYPES = {"size": 4, "sum": 6, "count": 7, "nunique": 8, "mean": 14, "min": 15, "max": 16, "first": 18, "last": 19, "var_pop": 22, "std_pop": 23, "var": 24, "std": 25, "skew": 27}


class GroupbyState:
    """Python handle of the C GroupbyState (created lazily at the first consume call)."""

    def __init__(self, operator_id, key_inds, fnames, f_in_offsets, f_in_cols, parallel, dropna, output_batch_size,
                 expected_groups, device, stream, process_group):
        self.operator_id = int(operator_id)
        self.key_inds = tuple(int(k) for k in key_inds)
        self.fnames = tuple(fnames)
        for f in self.fnames:
            if f not in FTYPES:
                raise _lib.B200Error(
                    f"Streaming Groupby: unsupported aggregate function '{f}' (supported: {sorted(FTYPES)})")
        self.f_in_offsets = tuple(int(x) for x in f_in_offsets)
        self.f_in_cols = tuple(int(x) for x in f_in_cols)
        if len(self.f_in_offsets) != len(self.fnames) + 1:
            raise _lib.B200Error("Streaming Groupby: f_in_offsets must have len(fnames) + 1 entries")
        self.parallel = bool(parallel)
        self.dropna = bool(dropna)
        self.output_batch_size = int(output_batch_size)
        self.expected_groups = int(expected_groups)
        self.device = device
        self.stream = int(stream)
        self.process_group = process_group
        self.handle = None
        self.build_indices = None  # physical column order: keys first (as the reference's build_indices)
        self.out_names = None
        self._out_cols = None
        self._out_tab = None
        self.exchanged = False

    # -- lazy C state creation once the input schema is known
    def _ensure(self, table: Table):
        if self.handle is not None:
            return
        L = _lib.lib()
        _lib.require_gpu()
        n = table.n_cols
        others = [i for i in range(n) if i not in self.key_inds]
        self.build_indices = list(self.key_inds) + others
        remap = {logical: phys for phys, logical in enumerate(self.build_indices)}
        phys_f_in_cols = [remap[c] for c in self.f_in_cols]
        cols = [table.columns[i] for i in self.build_indices]
        c_types = ffi.new("int8_t[]", [c.c_type for c in cols])
        a_types = ffi.new("int8_t[]", [c.arr_type for c in cols])
        ftypes = ffi.new("int32_t[]", [FTYPES[f] for f in self.fnames] or [0])
        offs = ffi.new("int32_t[]", list(self.f_in_offsets))
        fcols = ffi.new("int32_t[]", phys_f_in_cols or [0])
        if self.device is None:
            self.device = table.device if table.device >= 0 else _current_device()
        n_pes, rank = 1, 0
        if self.parallel:
            import torch.distributed as dist

            n_pes, rank = dist.get_world_size(self.process_group), dist.get_rank(self.process_group)
        self.n_pes, self.rank = n_pes, rank
        h = L.b200_groupby_state_init(self.operator_id, c_types, a_types, len(cols), ftypes, offs, fcols,
                                      len(self.fnames), len(self.key_inds), self.output_batch_size,
                                      int(self.parallel and n_pes > 1), int(self.dropna), self.device, n_pes, rank,
                                      self.expected_groups, ffi.cast("void*", self.stream))
        self.handle = _lib.check_ptr(h, "init_groupby_state")
        key_names = [table.names[i] for i in self.key_inds]
        fn_names = []
        for j, f in enumerate(self.fnames):
            lo, hi = self.f_in_offsets[j], self.f_in_offsets[j + 1]
            fn_names.append(table.names[self.f_in_cols[lo]] if hi > lo else f)
        # output names must be unique: a column aggregated more than once gets a _<func> suffix
        seen = set(key_names)
        uniq = []
        for f, nm in zip(self.fnames, fn_names):
            cand, k = nm, 0
            while cand in seen:
                cand = f"{nm}_{f}" if k == 0 else f"{nm}_{f}{k}"
                k += 1
            seen.add(cand)
            uniq.append(cand)
        self.out_names = key_names + uniq

    # ---- reduce-or-shuffle (reference: GroupbyIncrementalShuffleState::ShouldShuffleAfterProcessing,
    # bodo/libs/streaming/_groupby.cpp:1655-1711: an HLL estimate of how many NEW groups the pending rows hold decides whether they are
    # pre-reduced locally or shuffled as they are; threshold agg_reduction_threshold = 0.85, :1556) ----
    # Here every batch is aggregated into the local table first (that is the fast kernel), so the uniqueness is measured, not
    # estimated: groups in the tables / rows consumed, summed over the ranks.  While it stays below the threshold the ranks keep
    # pre-aggregating and exchange PARTIAL AGGREGATES once, at the end.  Above it (groups ~ rows: pre-aggregation buys nothing, the
    # table of every rank would grow towards its share of ALL rows and the one final exchange would move them all at once) the
    # state switches to the raw-row form: every further batch is hash-partitioned (b200_shuffle_partition) and exchanged right away
    # (all-to-all-v), each rank aggregates only rows of groups it owns, and the final exchange only carries what was aggregated
    # before the switch.  Collective: decided once, from the first >= B200_SHUFFLE_DECISION_ROWS rows, identically on every rank.
    raw_row_mode = False
    shuffle_decided = False
    raw_rows_shuffled = 0

    def _decide_reduce_or_shuffle(self):
        import os

        import torch
        import torch.distributed as dist

        L = _lib.lib()
        dev = torch.device("cuda", self.device)
        t = torch.tensor([int(L.b200_groupby_get_metric(self.handle, 13)), int(L.b200_groupby_get_metric(self.handle, 2))], dtype=torch.int64, device=dev)
        dist.all_reduce(t, group=self.process_group)
        groups, rows = (int(x) for x in t.tolist())
        if rows < int(os.environ.get("B200_SHUFFLE_DECISION_ROWS", 1 << 22)):
            return
        thr = min(1.0, max(0.0, float(os.environ.get("BODO_STREAM_GROUPBY_AGG_REDUCTION_THRESHOLD", os.environ.get("B200_AGG_REDUCTION_THRESHOLD", 0.85)))))
        self.shuffle_decided = True
        self.raw_row_mode = rows > 0 and groups / rows >= thr
        self.local_uniqueness = groups / max(rows, 1)

    def _shuffle_rows(self, phys: Table) -> Table:
        import torch

        from ..shuffle import shuffle_table
        from ..table import to_device

        dev_tab = to_device(phys, self.device)
        stream = torch.cuda.ExternalStream(self.stream) if self.stream else torch.cuda.default_stream(self.device)
        with torch.cuda.stream(stream):
            out = shuffle_table(dev_tab, len(self.key_inds), True, group=self.process_group, stream=self.stream)
        self.raw_rows_shuffled += phys.n_rows
        return out

    def _exchange(self):
        """Hash-partition exchange of the partial aggregates after the last local batch, then finalize.

        Fused form (default): ONE kernel packs every partial row another rank owns straight into that rank's receive slab over
        NVLink (symmetric memory, streaming/exchange.py), a device-side barrier, one combine kernel — no count exchange, no
        send buffer, no host synchronisation before finalize.  Falls back to the NCCL all-to-all-v (`_exchange_nccl`) when
        symmetric memory is unavailable or a rank's share did not fit its slab segment (finalize reports that: -2)."""
        import os
        import time

        import torch

        from . import exchange as X

        L = _lib.lib()
        h = self.handle
        trace = os.environ.get("B200_TRACE") and self.rank == 0
        t0 = time.perf_counter()
        slabs = X.get_slabs(self.process_group, self.device)
        done = False
        # nunique: the nested (key, value) distinct states are exchanged first (fused form only: their partial rows are multi-key)
        for i in range(int(L.b200_groupby_num_inner_states(h))):
            hi = _lib.check_ptr(L.b200_groupby_inner_state(h, i), "groupby nunique")
            if slabs is None:
                raise _lib.B200Error("groupby nunique on the sharded path needs the fused exchange (torch symmetric memory)")
            row_bytes = int(L.b200_groupby_exchange_row_bytes(hi))
            cap_rows = (slabs.slab_bytes - X.HDR_BYTES) // (self.n_pes * row_bytes)
            peers_dev, my_slab, hdl = slabs.next()
            stream = torch.cuda.ExternalStream(self.stream) if self.stream else torch.cuda.default_stream(self.device)
            with torch.cuda.stream(stream):
                _lib.check(L.b200_groupby_exchange_fused_pack(hi, ffi.cast("void* const*", peers_dev), cap_rows), "groupby nunique exchange (pack)")
                hdl.barrier(channel=0)
                _lib.check(L.b200_groupby_exchange_fused_combine(hi, ffi.cast("void*", my_slab), cap_rows), "groupby nunique exchange (combine)")
            if int(L.b200_groupby_finalize(hi)) < 0:
                raise _lib.B200Error("groupby nunique: the distinct (key, value) pairs did not fit the exchange slab "
                                     "(raise B200_XCHG_SLAB_BYTES)")
        if slabs is not None:
            row_bytes = int(L.b200_groupby_exchange_row_bytes(h))
            cap_rows = (slabs.slab_bytes - X.HDR_BYTES) // (self.n_pes * row_bytes)
            peers_dev, my_slab, hdl = slabs.next()
            stream = torch.cuda.ExternalStream(self.stream) if self.stream else torch.cuda.default_stream(self.device)
            with torch.cuda.stream(stream):
                _lib.check(L.b200_groupby_exchange_fused_pack(h, ffi.cast("void* const*", peers_dev), cap_rows), "groupby fused exchange (pack)")
                hdl.barrier(channel=0)
                _lib.check(L.b200_groupby_exchange_fused_combine(h, ffi.cast("void*", my_slab), cap_rows), "groupby fused exchange (combine)")
            rc = int(L.b200_groupby_finalize(h))
            if rc == -1:
                _lib.check(-1, "groupby finalize")
            done = rc != -2
            self.shuffle_bytes = None
        if not done:
            self._exchange_nccl()
            _lib.check(int(L.b200_groupby_finalize(h)), "groupby finalize")
        self.exchanged = True
        self.exchange_path = "fused" if done else "nccl"
        if trace:
            torch.cuda.synchronize()
            print(f"[b200 exchange ms] {'fused' if done else 'nccl'} exchange + finalize = {(time.perf_counter() - t0) * 1e3:.3f}", flush=True)

    def _exchange_nccl(self):
        """The same exchange through NCCL: count all-gather + one all-to-all-v of the packed partial rows, then combine."""
        import os
        import time

        import torch
        import torch.distributed as dist

        trace = os.environ.get("B200_TRACE") and self.rank == 0
        t = [time.perf_counter()]

        def mark():
            if trace:
                torch.cuda.synchronize()
                t.append(time.perf_counter())

        L = _lib.lib()
        h = self.handle
        counts = ffi.new("int64_t[]", self.n_pes)
        row_bytes = _lib.check(L.b200_groupby_shuffle_prepare(h, counts), "groupby shuffle prepare")
        mark()
        send_counts = [int(counts[i]) for i in range(self.n_pes)]
        dev = torch.device("cuda", self.device)
        words = row_bytes // 8
        n_send = sum(send_counts)
        # counts travel as one small all-gather of the n_pes x n_pes matrix (mpi_comm_info's MPI_Alltoall,
        # _shuffle.cpp:210-213), issued asynchronously so it overlaps the pack kernel
        sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
        allc = torch.empty(self.n_pes * self.n_pes, dtype=torch.int64, device=dev)
        work = dist.all_gather_into_tensor(allc, sc, group=self.process_group, async_op=True)
        send = torch.empty((max(n_send, 1), words), dtype=torch.int64, device=dev)
        _lib.check(L.b200_groupby_shuffle_pack(h, ffi.cast("void*", send.data_ptr())), "groupby shuffle pack")
        mark()
        work.wait()
        recv_counts = allc.view(self.n_pes, self.n_pes)[:, self.rank].tolist()
        mark()
        n_recv = sum(recv_counts)
        recv = torch.empty((max(n_recv, 1), words), dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv[:n_recv], send[:n_send], output_split_sizes=recv_counts,
                               input_split_sizes=send_counts, group=self.process_group)
        torch.cuda.current_stream(dev).synchronize()
        mark()
        _lib.check(L.b200_groupby_shuffle_combine(h, ffi.cast("void*", recv.data_ptr()), n_recv),
                   "groupby shuffle combine")
        L.b200_stream_synchronize(ffi.cast("void*", self.stream))
        mark()
        self.shuffle_bytes = n_send * row_bytes
        if trace:
            names = ["prepare", "pack", "counts", "alltoall", "combine"]
            print("[b200 exchange ms] " + " ".join(f"{n}={(b - a) * 1e3:.3f}" for n, a, b in zip(names, t, t[1:])), flush=True)


def _current_device() -> int:
    import torch

    return torch.cuda.current_device()


def init_groupby_state(operator_id, key_inds, fnames, f_in_offsets, f_in_cols, mrnf_sort_col_inds=None,
                       mrnf_sort_col_asc=None, mrnf_sort_col_na=None, mrnf_col_inds_keep=None, op_pool_size_bytes=-1,
                       parallel=False, *, dropna=True, output_batch_size=32768, expected_groups=0, device=None,
                       stream=0, process_group=None) -> GroupbyState:
    """Mirror of bodo.libs.streaming.groupby.init_groupby_state (groupby.py:702-715).

    key_inds / f_in_cols index the logical input table; fnames are names from supported_agg_funcs.
    The MRNF arguments must be None (min_row_number_filter is out of scope) and op_pool_size_bytes is
    ignored (the table is sized in HBM, there is no host operator pool).
    Keyword-only extras: dropna (pandas_drop_na of the C++ ctor), output_batch_size, expected_groups
    (sizing hint), device, stream (cudaStream_t as int), process_group (torch.distributed).
    """
    if any(x is not None for x in (mrnf_sort_col_inds, mrnf_sort_col_asc, mrnf_sort_col_na, mrnf_col_inds_keep)):
        raise _lib.B200Error("Streaming Groupby: min_row_number_filter is not supported by bodo_b200")
    key_inds = getattr(key_inds, "meta", key_inds)
    fnames = getattr(fnames, "meta", fnames)
    f_in_offsets = getattr(f_in_offsets, "meta", f_in_offsets)
    f_in_cols = getattr(f_in_cols, "meta", f_in_cols)
    return GroupbyState(operator_id, key_inds, fnames, f_in_offsets, f_in_cols, parallel, dropna, output_batch_size,
                        expected_groups, device, stream, process_group)


def groupby_build_consume_batch(groupby_state: GroupbyState, table: Table, is_last: bool, is_final_pipeline: bool = True):
    """Mirror of groupby_build_consume_batch (groupby.py:1295-1395): returns (is_last, request_input).

    Collective when the state is parallel: every rank must pass is_last=True in the same call (ranks that
    ran out of input keep calling with empty batches, as in the reference's pipeline loop,
    bodo/pandas/_pipeline.cpp:453-457)."""
    st = groupby_state
    st._ensure(table)
    L = _lib.lib()
    phys = table.select(st.build_indices)
    sharded = st.parallel and st.n_pes > 1
    if sharded and st.raw_row_mode:
        phys = st._shuffle_rows(phys)  # this rank's share of every rank's batch: all of its rows are owned here
    ct = CTable(phys)
    req = ffi.new("int32_t*")
    rc = _lib.check(L.b200_groupby_build_consume_batch(st.handle, ct.ptr, int(bool(is_last)), int(bool(is_final_pipeline)), req),
                    "groupby_build_consume_batch")
    if sharded and not is_last and not st.shuffle_decided:
        st._decide_reduce_or_shuffle()
    if is_last and sharded and not st.exchanged:
        st._exchange()
    return bool(rc), bool(req[0])


def groupby_produce_output_batch(groupby_state: GroupbyState, produce_output: bool = True):
    """Mirror of groupby_produce_output_batch (groupby.py:1502-1600): returns (out_table, is_last).
    The returned Table wraps library-owned device columns that stay valid until the next produce call."""
    st = groupby_state
    if st.handle is None:
        raise _lib.B200Error("groupby_produce_output_batch called before any build batch was consumed")
    L = _lib.lib()
    ncols = len(st.key_inds) + len(st.fnames)
    st._out_cols = ffi.new("b200_column[]", ncols)
    st._out_tab = ffi.new("b200_table*")
    st._out_tab.cols = st._out_cols
    last = ffi.new("int32_t*")
    _lib.check(L.b200_groupby_produce_output_batch(st.handle, st._out_tab, last, int(bool(produce_output))),
               "groupby_produce_output_batch")
    out = table_from_ctable(st._out_tab, ncols, st.out_names, owner=st)
    return out, bool(last[0])


def delete_groupby_state(groupby_state: GroupbyState) -> None:
    if groupby_state.handle is not None:
        _lib.lib().b200_delete_groupby_state(groupby_state.handle)
        groupby_state.handle = None


def get_metric(groupby_state: GroupbyState, which: int) -> int:
    return int(_lib.lib().b200_groupby_get_metric(groupby_state.handle, which))
