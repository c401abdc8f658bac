import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from torch.nn import CrossEntropyLoss
from oracle.spec_io import config_spec, CONFIGS
from model_utils import model_from_spec
from multimodn_b200 import FusedAdam
dev = torch.device("cuda")
for name, B in (("c3_mnar", 65536), ("c2_mimic", 65536)):
    feats = CONFIGS[name]["features"]; D = CONFIGS[name]["n_decoders"]
    model = model_from_spec(config_spec(name, 2), 1.0, 0.3, dev, "row", precision="bf16")
    opt = FusedAdam(model, lr=1e-3)
    g = torch.Generator(device=dev).manual_seed(2)
    y = (torch.rand((B, D), device=dev, generator=g) < 0.5).long()
    xs = [torch.randn((B, F), device=dev, generator=g) for F in feats]
    if name == "c3_mnar":
        p = torch.where(y[:, 0] == 1, 0.5, 0.1)
        for x in xs: x[torch.rand(B, device=dev, generator=g) < p] = float("nan")
    step = lambda: model.train_epoch([(xs, y)], opt, CrossEntropyLoss())
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name} bf16 layer-wise: B={B}: {ms:.2f} ms/step, {B/ms/1e3:.2f} M samples/s, workspace {model.runtime()._ws.numel()/2**30:.2f} GiB")
    del model, opt, xs
    torch.cuda.empty_cache()
