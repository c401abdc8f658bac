import os, sys, time
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from torch.nn import CrossEntropyLoss
from oracle.spec_io import config_spec, CONFIGS
from model_utils import model_from_spec
from multimodn_b200 import FusedAdam
dev = torch.device("cuda"); c = CONFIGS["c4_wide"]; feats = c["features"]; B = 8192
model = model_from_spec(config_spec("c4_wide", 3), 1.0, 0.3, dev, "row", precision="bf16")
opt = FusedAdam(model, lr=1e-4)
g = torch.Generator(device=dev).manual_seed(3)
b = ([torch.randn((B, F), device=dev, generator=g) for F in feats], (torch.rand((B, 2), device=dev, generator=g) < 0.3).long())
for i in range(3): model.train_epoch([b], opt, CrossEntropyLoss())
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10): model.train_epoch([b], opt, CrossEntropyLoss())
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/10:.2f} ms/step, total {1e3*(t2-t0)/10:.2f} ms/step")
