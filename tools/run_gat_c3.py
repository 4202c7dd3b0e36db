"""Development: a few forward+backward calls of GATConv(300 -> 8 x 64) on the C3-shaped graph (ncu target)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import gmlm_b200 as G
from gmlm_b200 import synth

dev = torch.device("cuda:0")
w = synth.WORKLOADS["c3"]
ei = synth.make_graph(w, device=dev)
x = synth.make_features(w.num_nodes, w.feat, device=dev).requires_grad_(True)
gat = G.GATConv(w.feat, 64, heads=8).to(dev)
for _ in range(4):
    y = gat(x, ei)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
