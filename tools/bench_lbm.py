"""Throughput of the lattice-Boltzmann generator kernel against its HBM roofline: B simulations on the reference's
384 x 128 lattice, CUDA events over a block of steps.  Algorithmic bytes per cell and step: 18 populations (9 read,
9 written) + 1 solid byte."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "steady-state-flow-with-neural-nets_b200")); sys.path.insert(0, ROOT)
from fluid_generator import lbm
from utils.synthetic import synth_boundary
peak = 6558.1
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak)
except Exception:
    pass
for dtype, B in ((torch.float32, 256), (torch.float32, 32), (torch.float64, 128)):
    solid = lbm.domain_from_boundary(synth_boundary(B, 128, 256, seed=0))
    sim = lbm.ChannelFlow(solid, dtype=dtype)
    sim.step(20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record(); sim.step(n); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    cells = B * 128 * 384
    nbytes = cells * (18 * (4 if dtype == torch.float32 else 8) + 1)
    print(json.dumps({"dtype": str(dtype), "B": B, "us_per_step": round(ms * 1e3, 1), "Mcell_updates_per_s": round(cells / ms / 1e3, 1),
                      "GBps": round(nbytes / ms / 1e6, 1), "frac_hbm": round(nbytes / ms / 1e6 / peak, 3),
                      "samples_per_s_at_10000_steps": round(B / (ms * 10), 2)}), flush=True)
