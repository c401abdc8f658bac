"""Fused Adam on the packed parameter buffer.

``torch.optim.Adam`` works with ``MultiModN.train_epoch`` unchanged (parameters are ordinary
``nn.Parameter``s whose ``.grad`` the step fills).  ``FusedAdam`` is the same update
(torch.optim.Adam defaults: no weight decay, no amsgrad) as ONE launch over the flat buffer, with
the rule "an encoder that took no row this step keeps parameters, moments and step count
untouched" (``.grad is None`` in the reference, multimodn.py:137,168-169) decided on the device
from the gradient buffer's tail, so ``train_epoch`` needs no host synchronisation per batch.
"""
import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    _mmn_fused = True

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.model = model
        rt = model.runtime()
        super().__init__(list(model.parameters()), dict(lr=lr, betas=betas, eps=eps))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam keeps one parameter group")
        self.exp_avg = torch.zeros(rt.packed.n_params, dtype=torch.float32, device=rt.device)
        self.exp_avg_sq = torch.zeros_like(self.exp_avg)
        self.step_count = torch.zeros(1 + rt.E, dtype=torch.int32, device=rt.device)

    def zero_grad(self, set_to_none: bool = True):
        pass                                    # mmn_train_step overwrites the packed gradient

    @torch.no_grad()
    def step(self, closure=None):
        rt = self.model.runtime()
        rt.ensure_packed()
        g = self.param_groups[0]
        rt.lib.check(rt.lib.dll.mmn_adam_step(rt.plan, rt.flat.data_ptr(), rt.gflat.data_ptr(), self.exp_avg.data_ptr(),
                                              self.exp_avg_sq.data_ptr(), self.step_count.data_ptr(), float(g["lr"]),
                                              float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), rt.stream()))
