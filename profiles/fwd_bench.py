"""Forward-only (test()/predict() path) timing of the step kernel on config C2, per engine.
usage: MMN_ENGINE=fma|tc|tc2 python profiles/fwd_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bench
from model_utils import model_from_spec
from oracle import multimodn_oracle as O

dev = torch.device("cuda")
B = 65536
spec = bench.make_spec()
model = model_from_spec(spec, 1.0, 0.3, dev, "row")
rt = model.runtime()
xs, y = bench.make_batch(np.random.default_rng(0), B, device=dev)
seq = [(i, i) for i in range(3)]
metrics = rt.new_metrics()
preds = torch.zeros((4, 2, B), dtype=torch.uint8, device=dev)
def step():
    mb, keep, n = rt.prepare_batch(xs, y, seq, "row", None)
    rt.forward(mb, n, metrics=metrics, predictions=preds)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
# parity of a 512-row sample vs the oracle
idx = np.arange(0, B, B // 512)[:512]
sub = [x[idx].cpu().numpy() for x in xs]
ofwd = O.forward(O.cast_spec(spec, np.float32), sub, y[idx].cpu().numpy(), None, "row")
mis = (preds[:, :, idx].cpu().numpy() != ofwd["predictions"]).mean()
print(f"engine {rt.lib.dll.mmn_plan_engine(rt.plan)}: forward {ms:.3f} ms/launch, {B / ms / 1e3:.1f} M rows/s, prediction mismatch on sample {mis:.4f}")
