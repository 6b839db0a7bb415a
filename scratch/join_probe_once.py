import sys
import torch
sys.path.insert(0, ".")
from bodo_b200.streaming import join as J
from bodo_b200.table import Column, Table
dev = torch.device("cuda", 0)
nb, npr = 100_000_000, 250_000_000
g = torch.Generator(device=dev); g.manual_seed(1)
bk = torch.randperm(nb, device=dev, generator=g)
b1 = torch.arange(nb, device=dev); b2 = torch.rand(nb, device=dev, dtype=torch.float64)
st = J.init_join_state(-1, (0,), (0,), ("k", "b1", "b2"), ("k", "p1", "p2"), False, False, output_batch_size=1 << 40)
J.join_build_consume_batch(st, Table([Column(bk), Column(b1), Column(b2)], ["k", "b1", "b2"]), True)
p1 = torch.arange(npr, device=dev); p2 = torch.rand(npr, device=dev, dtype=torch.float64)
pk = torch.randint(0, nb, (npr,), device=dev, generator=g)
tab = Table([Column(pk), Column(p1), Column(p2)], ["k", "p1", "p2"])
for it in range(2):
    out, _, _ = J.join_probe_consume_batch(st, tab, False, True, ([0, 1, 2], [1, 2]))
torch.cuda.synchronize()
print(out.n_rows)
