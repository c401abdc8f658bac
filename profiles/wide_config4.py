"""BASELINE.json configs[3] — the wide regime (state 1024, encoder hidden 2048, bf16) — on one B200:
train step at 8192 rows per GPU (the per-GPU share of the 8-GPU data-parallel configuration).
Inputs are generated on the device.  usage: python profiles/wide_config4.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from torch.nn import CrossEntropyLoss
from oracle.spec_io import config_spec, CONFIGS
from model_utils import model_from_spec
from multimodn_b200 import FusedAdam, _lib

dev = torch.device("cuda")
c = CONFIGS["c4_wide"]
feats, S = c["features"], c["S"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
spec = config_spec("c4_wide", 3)
model = model_from_spec(spec, 1.0, 0.3, dev, "row", precision="bf16")
opt = FusedAdam(model, lr=1e-4)
g = torch.Generator(device=dev).manual_seed(3)
bs = [([torch.randn((B, F), device=dev, generator=g) for F in feats], (torch.rand((B, 2), device=dev, generator=g) < 0.3).long())
      for _ in range(2)]
lib = _lib.get_lib()

def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

step = lambda i: model.train_epoch([bs[i % 2]], opt, CrossEntropyLoss())
for i in range(3): step(i)
l0 = lib.dll.mmn_wide_launch_count()
ms = timed(step, 10)
launches = (lib.dll.mmn_wide_launch_count() - l0) / 10
h = c["enc_hidden"]
macs = sum((F + S) * h[0] + h[0] * h[1] + h[1] * S for F in feats) + (S * c["dec_hidden"][0] + c["dec_hidden"][0] * 2) * 2 * (len(feats) + 1)
peaks = {}
try:
    import json
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
tf = 6 * macs * B / ms / 1e9
print(f"C4 train (bf16 wide regime): B={B}: {ms:.3f} ms/step, {B / ms / 1e3:.3f} M samples/s, {tf:.0f} TFLOP/s algorithmic "
      f"({macs / 1e6:.1f} M MAC/sample x 6), {tf / peaks.get('bf16_tflops_sustained', 1375.5):.3f} of the measured sustained bf16 peak, "
      f"{launches:.0f} kernel launches/step, workspace {model.runtime()._ws.numel() / 2**30:.2f} GiB")
rt = model.runtime()
fwd = lambda i: model.predict(bs[i % 2][0])
fwd(0)
ms_f = timed(fwd, 5)
print(f"C4 predict: {ms_f:.3f} ms per predict() of {B} rows incl. D2H of the class ids, {B / ms_f / 1e3:.3f} M rows/s")
if os.environ.get("MMN_WIDE_TIMERS_ONCE"):
    os.environ["MMN_WIDE_TIMERS"] = "1"
    step(0)
    torch.cuda.synchronize()
