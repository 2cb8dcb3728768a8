"""Per-round trace (KOMBGPU_TRACE) of one peel for a graph shape and a kernel mode:  peel_trace.py cfg2 warp out.csv"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
which, mode, out = sys.argv[1], sys.argv[2], sys.argv[3]
os.environ["KOMBGPU_PEEL_MODE"] = mode
os.environ["KOMBGPU_DEBUG"] = "1"
import komb_b200
from peel_ab import make_graph
ctx = komb_b200.Context(0)
g = make_graph(ctx, which)
g.coreness(copy=False, again=True)          # warm-up
os.environ["KOMBGPU_TRACE"] = out
g.coreness(copy=False, again=True)
print(which, mode, g.stats()["ms_peel_kernel"])
