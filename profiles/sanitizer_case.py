import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from torch.nn import CrossEntropyLoss
from oracle.spec_io import random_spec, synthetic_batch
from model_utils import model_from_spec, GradTap
from multimodn_b200 import MultiModNHistory
which = sys.argv[1]
rng = np.random.default_rng(1)
if which == "wide":
    feats = [72, 40]; spec = random_spec(rng, 64, feats, enc_hidden=(96, 80), n_decoders=2, dec_hidden=(48,), dropout=0.2)
    prec = "bf16"; B = 300
else:
    feats = [6, 19, 40]; spec = random_spec(rng, 16, feats, enc_hidden=(8, 8), n_decoders=2, dec_hidden=(8, 8), dropout=0.2)
    prec = "fp32"; B = 300
data, y = synthetic_batch(rng, feats, 2, B, mnar=True)
model = model_from_spec(spec, 1.0, 0.3, "cuda", "row", precision=prec)
tap = GradTap(model.parameters())
loader = [([torch.from_numpy(x).cuda() for x in data], torch.from_numpy(y).cuda())]
model.train_epoch(loader, tap, CrossEntropyLoss(), MultiModNHistory(["a", "b"]))
model.predict([torch.from_numpy(x) for x in data])
torch.cuda.synchronize()
print("done", which)
