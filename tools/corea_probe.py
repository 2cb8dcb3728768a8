import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import komb_b200
from komb_b200 import synth
ctx = komb_b200.Context(0)
m1, m2 = synth.metagenome_hits(1_000_000, 5_000_000, seed=11)
rk = torch.from_numpy(np.concatenate([m1.read_key, m2.read_key]).view(np.int32)).cuda()
ut = torch.from_numpy(np.concatenate([m1.unitig, m2.unitig]).view(np.int32)).cuda()
for mode in ("lut", "sort", "lut", "sort"):
    if mode == "sort": os.environ["KOMBGPU_COREA_SORT"] = "1"
    else: os.environ.pop("KOMBGPU_COREA_SORT", None)
    g = ctx.build_graph(rk, ut, 1_000_000); g.analyse(0); st = g.stats()
    l0 = ctx.launches(); 
    print(mode, "ms_corea", st["ms_corea"], "max_deg", st["max_degree"], "kmax", st["max_coreness"], "launches", st["kernel_launches"])
    g.close()
