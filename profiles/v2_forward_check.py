import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from multimodn_b200 import MultiModN, MultiModNHistory, _lib
from oracle import multimodn_oracle as O
from oracle.spec_io import random_spec, synthetic_batch
from model_utils import model_from_spec
from helpers import assert_close
from torch.nn import CrossEntropyLoss
lib=_lib.get_lib()
def run(S, feats, kind, eh, D, dh, C, B, mnar, seed=0, mode="row"):
    rng=np.random.default_rng(seed)
    spec=random_spec(rng,S,feats,enc_kind=kind,enc_hidden=eh,n_decoders=D,dec_hidden=dh,n_classes=C)
    data,y=synthetic_batch(rng,feats,D,B,mnar=mnar,n_classes=C)
    model=model_from_spec(spec,0.8,0.6,"cuda",mode)
    print("fwd engine", lib.dll.mmn_plan_forward_engine(model.runtime().plan), flush=True)
    ofwd=O.forward(O.cast_spec(spec,np.float32),data,y,None,mode)
    pred=model.predict([torch.from_numpy(x) for x in data])
    torch.cuda.synchronize()
    mis=(pred!=ofwd["predictions"]).mean()
    loader=[([torch.from_numpy(x).cuda() for x in data], torch.from_numpy(y).cuda())]
    st=torch.stack(model.get_states(loader)).cpu().numpy()
    assert_close(st, ofwd["final_state"], rtol=1e-5, what="states")
    hist=MultiModNHistory([str(i) for i in range(D)])
    model.test(loader, CrossEntropyLoss(), hist, tag="val")
    acc=O.EpochAccumulator(len(feats),D); acc.add(ofwd); fin=acc.finalize()
    assert_close(hist.loss["val"][0], fin["loss"], rtol=1e-5, what="val loss")
    print("ok", S, feats, kind, mode, "pred mismatch", mis, flush=True)
run(16,[6,19,40],"mimic",(8,8),2,(8,8),2,16,False)
run(16,[6,19,40],"mimic",(8,8),2,(8,8),2,16,False,mode="batch")
run(16,[5,12,33],"mimic",(8,8),2,(8,),2,300,False)
run(24,[6,9,17,4],"mimic",(16,),3,(8,8),2,517,True)
run(8,[7,40],"mlp",(40,12),2,(),2,300,True)
run(40,[9],"mimic",(4,),1,(),3,70,True)
