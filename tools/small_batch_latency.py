import os, sys, json, torch
ROOT="/root/repo"
sys.path.insert(0, os.path.join(ROOT, "steady-state-flow-with-neural-nets_b200")); sys.path.insert(0, ROOT)
import model.flow_architecture as arch, model.flow_net as flow_net
from utils.synthetic import synth_boundary
for B in (1, 8, 16):
    arch.reset_variables(1234)
    x = torch.from_numpy(synth_boundary(B, 128, 256, seed=0)).cuda()
    net = flow_net.CompiledInference(B, (128, 256)); net(x)
    for _ in range(5): net.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): net.replay()
    e1.record(); torch.cuda.synchronize()
    print("PDL", os.environ.get("FLOWNET_PDL"), "B", B, "ms/step %.4f" % (e0.elapsed_time(e1)/50), flush=True)
