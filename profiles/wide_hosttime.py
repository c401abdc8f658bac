"""Host-side cost of enqueuing one wide-regime train step (167 launches + tensor-map encodes): config 4 at a batch small
enough that the GPU is never the bottleneck, so wall time per step = CPU time per step.  usage: python profiles/wide_hosttime.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from torch.nn import CrossEntropyLoss
from oracle.spec_io import config_spec, CONFIGS
from model_utils import model_from_spec
from multimodn_b200 import FusedAdam
dev = torch.device("cuda"); c = CONFIGS["c4_wide"]; feats = c["features"]
model = model_from_spec(config_spec("c4_wide", 3), 1.0, 0.3, dev, "row", precision="bf16")
opt = FusedAdam(model, lr=1e-4)
g = torch.Generator(device=dev).manual_seed(3)
for B in (128, 8192):
    b = ([torch.randn((B, F), device=dev, generator=g) for F in feats], (torch.rand((B, 2), device=dev, generator=g) < 0.3).long())
    for i in range(3): model.train_epoch([b], opt, CrossEntropyLoss())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(20): model.train_epoch([b], opt, CrossEntropyLoss())
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"B={B}: host enqueue {1e3*(t1-t0)/20:.2f} ms/step, wall {1e3*(t2-t0)/20:.2f} ms/step")
