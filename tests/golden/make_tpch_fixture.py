"""Generates tests/golden/tpch_q1_lineitem.parquet from the reference's own TPC-H test data
(bodo/tests/data/tpch-test_data/parquet/lineitem.pq, the fixture of bodo/tests/test_df_lib/test_tpch.py): the seven columns
TPC-H Q1 touches, every 4th row (30 k rows) to keep the committed file small.  Run in the authoring container only:

    python tests/golden/make_tpch_fixture.py
"""
import os

import pyarrow as pa
import pyarrow.parquet as pq

SRC = "/root/reference/bodo/tests/data/tpch-test_data/parquet"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    cols = ["L_ORDERKEY", "L_QUANTITY", "L_EXTENDEDPRICE", "L_DISCOUNT", "L_TAX", "L_RETURNFLAG", "L_LINESTATUS", "L_SHIPDATE"]
    t = pq.read_table(os.path.join(SRC, "lineitem.pq"), columns=cols)
    t = t.take(pa.array(range(0, t.num_rows, 4)))
    out = os.path.join(HERE, "tpch_q1_lineitem.parquet")
    pq.write_table(t, out, compression="zstd")
    print(out, t.num_rows, os.path.getsize(out))
