import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from torch.nn import CrossEntropyLoss
import bench
from model_utils import model_from_spec
from multimodn_b200 import FusedAdam
dev = torch.device("cuda")
w = bench.WORKLOAD
model = model_from_spec(bench.make_spec(), 1.0, 0.3, dev, "row")
opt = FusedAdam(model, lr=1e-3)
xs, y = bench.make_batch(np.random.default_rng(0), 65536, device=dev)
for i in range(3):
    model.train_epoch([(xs, y)], opt, CrossEntropyLoss())
torch.cuda.synchronize()
os.environ["MMN_DEBUG_TIMERS"] = "1"
model.train_epoch([(xs, y)], opt, CrossEntropyLoss())
torch.cuda.synchronize()
