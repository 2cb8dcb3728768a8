"""Per-round trace of the full-size cfg3 / cfg5 graph:  peel_trace_big.py cfg3 out.csv"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1])); sys.path.insert(0, str(Path(__file__).resolve().parent))
os.environ["KOMBGPU_DEBUG"] = "1"
import torch, komb_b200
from scale_probe import rmat_device, ramp_device
ctx = komb_b200.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
if sys.argv[1] == "cfg3":
    n = 50_000_000; u, v = rmat_device(26, 540_000_000, n, 42)
else:
    u, v, n = ramp_device(5000, 40, 9_800_000, 24, 40_000_000, 7)
g = ctx.graph_from_edges(u, v, n)
g.coreness(copy=False, again=True)
os.environ["KOMBGPU_TRACE"] = sys.argv[2]
g.coreness(copy=False, again=True)
print(g.stats()["ms_peel_kernel"])
