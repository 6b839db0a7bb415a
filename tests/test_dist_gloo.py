"""world_size-2 gloo tests (CPU): the N>1 host logic — count exchange, all-to-all-v of destination-grouped buffers,
per-destination bitmap segments — with the oracle's partition standing in for the CUDA kernel (tests only)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_partition(table, n_keys, n_pes):
    """Test-only stand-in for b200_shuffle_partition built on the oracle (same output contract)."""
    from bodo_b200.table import Column, Table
    from oracle import oracle as O
    kc = table.columns[0]
    keys = kc.data.numpy() if hasattr(kc.data, "numpy") else kc.data
    kvalid = kc.valid_mask_numpy()
    counts, perm = O.shuffle_partition(keys, kvalid, n_pes)
    cols = []
    for c in table.columns:
        d = c.data.numpy() if hasattr(c.data, "numpy") else c.data
        data = torch.from_numpy(np.ascontiguousarray(d[perm]))
        v = None
        if c.validity is not None:
            mask = c.valid_mask_numpy()[perm]
            segs, off = [], 0
            for cnt in counts:
                segs.append(np.packbits(mask[off:off + cnt], bitorder="little"))
                off += cnt
            v = torch.from_numpy(np.concatenate(segs + [np.zeros(8, dtype=np.uint8)]))
        cols.append(Column(data, v, c.c_type, c.arr_type, c.length))
    return Table(cols, list(table.names)), [int(x) for x in counts]


def _numpy_merge_bitmaps(bitmap, counts):
    """Test-only stand-in for b200_merge_segment_bitmaps (per-source byte-padded segments -> one Arrow bitmap)."""
    bits, off, b = [], 0, bitmap.cpu().numpy()
    for c in counts:
        nb = (c + 7) // 8
        bits.append(np.unpackbits(b[off:off + nb], bitorder="little")[:c])
        off += nb
    allbits = np.concatenate(bits) if bits else np.zeros(0, dtype=np.uint8)
    packed = np.packbits(allbits, bitorder="little")
    pad = np.zeros((len(packed) + 7) // 8 * 8 + 8, dtype=np.uint8)
    pad[:len(packed)] = packed
    return torch.from_numpy(pad)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pandas as pd
        from bodo_b200.shuffle import exchange_table, with_schema_validity
        from bodo_b200.table import Table
        from oracle import oracle as O
        rng = np.random.default_rng(100 + rank)
        n = 5000 + 37 * rank
        keys = rng.integers(0, 400, n).astype(np.int64)
        df = pd.DataFrame({"k": keys, "v": rng.integers(-9, 9, n).astype(np.int64),
                           "w": pd.array(rng.integers(0, 50, n), dtype="Int64")})
        if rank == 0:
            df.loc[rng.random(n) < 0.2, "w"] = pd.NA  # rank 1 has NO nulls in its nullable column: its bitmap must still travel
        t = Table.from_pandas(df)
        if rank == 1:  # what an Arrow reader hands over for a chunk without nulls: nullable column, no bitmap
            from bodo_b200.table import Column
            t.columns[2] = Column(t.columns[2].data, None, t.columns[2].c_type, t.columns[2].arr_type, t.columns[2].length)
        part, send_counts = _oracle_partition(with_schema_validity(t), 1, world)
        out = exchange_table(part, send_counts, merge_bitmaps=_numpy_merge_bitmaps)
        odf = out.to_pandas()
        # 1. every received key belongs to this rank under the reference's hash_to_rank
        dest = O.hash_to_rank(odf["k"].to_numpy(), None, world)
        ok_owner = bool((dest == rank).all())
        # 2. nothing lost or duplicated: global multiset of rows is preserved
        mine = torch.tensor([len(odf), int(odf["v"].sum()), int(odf["w"].fillna(0).sum()), int(odf["w"].isna().sum()),
                             n, int(df["v"].sum()), int(df["w"].fillna(0).sum()), int(df["w"].isna().sum())], dtype=torch.int64)
        dist.all_reduce(mine)
        ok_conserved = mine[:4].tolist() == mine[4:].tolist()
        # 3. a local groupby of the shuffled rows equals the oracle's sharded groupby of the global table
        gathered = [None] * world
        dist.all_gather_object(gathered, df)
        gdf = pd.concat(gathered, ignore_index=True)
        exp = O.groupby(gdf["k"].to_numpy(), None, ["sum", "count"], [gdf["v"].to_numpy()] * 2, n_pes=world, rank=rank)
        got = odf.groupby("k").v.agg(["sum", "count"])
        e = pd.DataFrame({"k": exp["keys"], "sum": exp["cols"][0][0], "count": exp["cols"][1][0]}).set_index("k").sort_index()
        ok_groupby = bool((got["sum"].to_numpy() == e["sum"].to_numpy()).all() and (got["count"].to_numpy() == e["count"].to_numpy()).all()
                          and (got.index.to_numpy() == e.index.to_numpy()).all())
        q.put((rank, ok_owner, ok_conserved, ok_groupby))
    except Exception as ex:  # surface the failure in the parent instead of hanging it
        import traceback
        q.put((rank, False, False, traceback.format_exc() or str(ex)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_shuffle_table_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    for r in sorted(res):
        assert r[1] and r[2] and r[3] is True, r
