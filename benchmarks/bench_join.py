#!/usr/bin/env python
"""Secondary benchmark (not the bench.py contract): BASELINE.json configs[2] — hash inner join 1B x 100M int64 key,
4 payload columns (2 per side) on 1 x B200.  Reports build and probe time, probe rows/s and achieved GB/s against the
compulsory stream |B|*row_B + |P|*row_P + |Out|*row_Out (2.4 + 24 + 40 = 66.4 GB).

    python benchmarks/bench_join.py [--build-rows 100000000] [--probe-rows 1000000000] [--batch 250000000]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build-rows", type=int, default=100_000_000)
    ap.add_argument("--probe-rows", type=int, default=1_000_000_000)
    ap.add_argument("--batch", type=int, default=250_000_000)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch

    from bodo_b200 import synth
    from bodo_b200.streaming import join as J
    from bodo_b200.table import Column, Table

    dev = torch.device("cuda", 0)
    nb, npr = args.build_rows, args.probe_rows
    # build: unique keys = a permutation-like bijection of [0, nb) (odd multiplier modulo 2^k is not onto [0, nb), so use
    # arange and let the hash table see them in a scrambled order via the payload-independent key column itself)
    bk = torch.arange(nb, dtype=torch.int64, device=dev)
    perm_mul = 0x9E3779B97F4A7C15 % nb | 1
    bk = (bk * perm_mul) % nb if nb & (nb - 1) == 0 else bk[torch.randperm(nb, device=dev)]
    b1 = torch.empty(nb, dtype=torch.int64, device=dev)
    b2 = torch.empty(nb, dtype=torch.float64, device=dev)
    synth.device_fill(None, b1, 0, 1, 31)
    synth.device_fill(None, b2, 0, 1, 32)
    pk = torch.empty(npr, dtype=torch.int64, device=dev)
    p1 = torch.empty(npr, dtype=torch.int64, device=dev)
    p2 = torch.empty(npr, dtype=torch.float64, device=dev)
    synth.device_fill(pk, p1, 0, nb, 41)   # every probe key in [0, nb): exactly one match per probe row
    synth.device_fill(None, p2, 0, 1, 42)
    torch.cuda.synchronize()
    build = Table([Column(bk), Column(b1), Column(b2)], ["k", "b1", "b2"])
    res = []
    for rep in range(args.reps):
        st = J.init_join_state(-1, (0,), (0,), ("k", "b1", "b2"), ("k", "p1", "p2"), False, False, expected_build_rows=nb, device=0)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        J.join_build_consume_batch(st, build, True)
        # the C state is created at the first probe call (it needs both schemas); force it with an empty batch
        empty = Table([Column(pk[:0]), Column(p1[:0]), Column(p2[:0])], ["k", "p1", "p2"])
        J.join_probe_consume_batch(st, empty, False, True, ([0, 1, 2], [1, 2]))
        e[1].record()
        out_rows, chk = 0, 0
        for r0 in range(0, npr, args.batch):
            r1 = min(npr, r0 + args.batch)
            t = Table([Column(pk[r0:r1]), Column(p1[r0:r1]), Column(p2[r0:r1])], ["k", "p1", "p2"])
            out, last, _ = J.join_probe_consume_batch(st, t, r1 == npr, True, ([0, 1, 2], [1, 2]))
            out_rows += out.n_rows
        e[2].record()
        torch.cuda.synchronize()
        # spot check of the last batch: output key column equals the probe keys as a multiset (1 match per row)
        ok = out.n_rows == (r1 - r0)
        okeys = torch.as_tensor(out.columns[0].data, device=dev)
        ok = ok and int(okeys.sum().item()) == int(pk[r0:r1].sum().item())
        J.delete_join_state(st)
        res.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), out_rows, ok))
    bms = min(r[0] for r in res)
    pms = min(r[1] for r in res)
    stream_gb = (nb * 24 + npr * 24 + res[-1][2] * 40) / 1e9
    peak = 6574.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    print(json.dumps({
        "workload": f"hash inner join {npr} x {nb} int64 key, 2 payload cols per side (BASELINE.json configs[2])",
        "build_ms": bms, "probe_ms": pms, "out_rows": res[-1][2], "check": "ok" if all(r[3] for r in res) else "MISMATCH",
        "probe_rows_per_s": npr / (pms * 1e-3), "compulsory_stream_gb": stream_gb,
        "achieved_gbs": stream_gb / ((bms + pms) * 1e-3), "frac_of_hbm_peak": stream_gb / ((bms + pms) * 1e-3) / peak, "peak_gbs": peak,
    }))


if __name__ == "__main__":
    main()
