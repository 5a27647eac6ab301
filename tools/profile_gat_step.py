"""One GAT / GCN / SAGE-max training step on the arxiv shape, repeated a few times, for an ncu launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_gat_step.py [gat|gcn|sage]
Shows where a whole step (cuBLAS GEMMs + sparse kernels + autograd glue) spends its time."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cogdl_b200  # noqa: E402
from cogdl_b200 import synth  # noqa: E402
from cogdl_b200.layers import GAT, GCN, SAGE  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "gat"
dev = torch.device("cuda")
n, e = synth.SHAPES["arxiv"]
rp, col = synth.powerlaw_csr(n, e, seed=0)
g = cogdl_b200.Graph(x=torch.randn(n, 128), row_ptr=rp, col=col, num_nodes=n).to(dev)
model = {"gat": lambda: GAT(128, 16, 40, nhead=8, last_nhead=1), "gcn": lambda: GCN(128, 128, 40, dropout=0.0),
         "sage": lambda: SAGE(128, 128, 40, aggr="max")}[which]().to(dev)
opt = torch.optim.SGD(model.parameters(), lr=0.01)
y = torch.randint(0, 40, (n,), device=dev)
for _ in range(4):
    opt.zero_grad(set_to_none=True)
    loss = F.cross_entropy(model(g), y)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("done", which, float(loss))
