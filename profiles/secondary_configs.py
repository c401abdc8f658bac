"""Secondary BASELINE.json configurations on one B200 (not the bench headline): C1 Titanic model at a scaled batch
(N = B = 2^20), C3 MNAR-stress train step (E=8, D=6, S=256, B=65536, 30 % MNAR) and C5 inference sweep (predict over
permuted encoding sequences).  C4 (wide regime): profiles/wide_config4.py / bench.py --workload c4_wide.
Inputs are generated on the device.  usage: python profiles/secondary_configs.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from torch.nn import CrossEntropyLoss
from oracle.spec_io import config_spec, CONFIGS
from model_utils import model_from_spec
from multimodn_b200 import FusedAdam

dev = torch.device("cuda")
feats = CONFIGS["c3_mnar"]["features"]
spec = config_spec("c3_mnar", 2)
model = model_from_spec(spec, 1.0, 0.3, dev, "row")
opt = FusedAdam(model, lr=1e-3)
g = torch.Generator(device=dev).manual_seed(2)

def batch(B, mnar=True):
    y = (torch.rand((B, 6), device=dev, generator=g) < 0.5).long()
    xs = [torch.randn((B, F), device=dev, generator=g) for F in feats]
    if mnar:
        p = torch.where(y[:, 0] == 1, 0.5, 0.1)
        for x in xs:
            x[torch.rand(B, device=dev, generator=g) < p] = float("nan")
    return xs, y

def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

# ---- C1: Titanic MLP model (S = 1, 6 features, (5, 5) hidden, 1 logistic decoder), scaled throughput case ----
spec1 = config_spec("c1_titanic", 0)
m1 = model_from_spec(spec1, 0.7, 0.3, dev, "row")
o1 = FusedAdam(m1, lr=1e-2)
B1 = 1 << 20
b1 = [([torch.randn((B1, 6), device=dev, generator=g)], (torch.rand((B1, 1), device=dev, generator=g) < 0.4).long()) for _ in range(2)]
s1 = lambda i: m1.train_epoch([b1[i % 2]], o1, CrossEntropyLoss())
for i in range(3): s1(i)
ms1 = timed(s1, 10)
print(f"C1 train (scaled): B={B1}: {ms1:.3f} ms/step, {B1 / ms1 / 1e3:.1f} M samples/s, {B1 * (2 * 4 * 6 + 8) / ms1 / 1e6:.0f} GB/s algorithmic "
      f"(56 B/sample; HBM-bound regime: 65 MAC/sample)")
del m1, o1, b1

B = 65536
bs = [batch(B) for _ in range(2)]
miss = float(np.mean([torch.isnan(x[:, 0]).float().mean().item() for x in bs[0][0]]))
step = lambda i: model.train_epoch([bs[i % 2]], opt, CrossEntropyLoss())
for i in range(3): step(i)
ms = timed(step, 8)
macs = sum((F + 256) * 32 + 32 * 32 + 32 * 256 for F in feats) + (256 * 32 + 32 * 32 + 32 * 2) * 6 * 9
print(f"C3 train: B={B}, missing cells {miss:.3f}: {ms:.2f} ms/step, {B / ms / 1e3:.2f} M samples/s, "
      f"{6 * macs * B / ms / 1e9:.1f} TFLOP/s fp32 (algorithmic), {B * (2 * 4 * sum(feats) + 48) / ms / 1e6:.0f} GB/s algorithmic")

N = 1 << 18
xs, _ = batch(N, mnar=True)
perms = [np.roll(np.arange(8), s) for s in range(4)] + [np.random.default_rng(4).permutation(8) for _ in range(4)]
def pred(i):
    p = perms[i % len(perms)]
    model.predict([xs[e] for e in p], p)
pred(0)
t0 = time.perf_counter()
for i in range(8): pred(i)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 8
print(f"C5 predict sweep: N={N} rows x 8 sequence orders, 30 % MNAR: {dt * 1e3:.1f} ms per predict() incl. D2H of the (9,6,N) class ids, "
      f"{N / dt / 1e6:.1f} M rows/s")
