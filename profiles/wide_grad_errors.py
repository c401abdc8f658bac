import sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from torch.nn import CrossEntropyLoss
from oracle import multimodn_oracle as O
from oracle.spec_io import random_spec, synthetic_batch
from multimodn_b200 import MultiModNHistory
from model_utils import model_from_spec, GradTap
import test_gpu_wide as T
for name in sys.argv[1:]:
    S, feats, kind, eh, D, dh, C, B, mnar, p = T.CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    spec = random_spec(rng, S, feats, enc_kind=kind, enc_hidden=eh, dropout=p, n_decoders=D, dec_hidden=dh, n_classes=C)
    data, y = synthetic_batch(rng, feats, D, B, mnar=mnar, n_classes=C)
    model = model_from_spec(spec, 0.8, 0.6, "cuda", "row", precision="bf16")
    tap = GradTap(model.parameters())
    loader = [([torch.from_numpy(x).cuda() for x in data], torch.from_numpy(y).cuda())]
    rt = model.runtime(); rt.dropout_base_seed, rt.step_counter = 77, 0
    seed = (77 * 0x9E3779B1 + 1 * 0x85EBCA77) & 0xFFFFFFFF
    model.train_epoch(loader, tap, CrossEntropyLoss(), MultiModNHistory([str(i) for i in range(D)]))
    fwd, loss, grads, _ = O.train_step(O.cast_spec(spec, np.float32), data, y, 0.8, 0.006, missing_mode="row", dropout_seed=seed)
    def cmp(tag, got, exp):
        got = got.detach().cpu().numpy().astype(np.float64); exp = np.asarray(exp, np.float64)
        print(f"  {tag:16s} max|exp| {np.abs(exp).max():.3e}  maxerr/max {np.abs(got-exp).max()/max(np.abs(exp).max(),1e-30):.3e}  relL2 {np.linalg.norm(got-exp)/max(np.linalg.norm(exp),1e-30):.3e}")
    print(name)
    g = lambda p_: tap.acc[p_]
    cmp("init", g(model.init_state.state_value).reshape(-1), grads["init_state"])
    for e, enc in enumerate(model.encoders):
        lins = [m for m in enc.layers if isinstance(m, torch.nn.Linear)]
        for j, l in enumerate(lins):
            cmp(f"enc{e}.W{j}", g(l.weight), grads["encoders"][e][j][0]); cmp(f"enc{e}.b{j}", g(l.bias), grads["encoders"][e][j][1])
    for d, dec in enumerate(model.decoders):
        lins = [dec.fc] if hasattr(dec, "fc") else list(dec.layers)
        for j, l in enumerate(lins):
            cmp(f"dec{d}.W{j}", g(l.weight), grads["decoders"][d][j][0]); cmp(f"dec{d}.b{j}", g(l.bias), grads["decoders"][d][j][1])
