"""CPU, build container only (needs /root/reference; skipped elsewhere): the drop-in driver accepts the
reference's OWN encoder / decoder / init-state instances and reproduces the reference driver on them."""
import itertools
import os
import sys

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


def test_reference_module_instances(emu):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "oracle", "ref_shims"))
    sys.path.insert(0, REF)
    try:
        torch._utils._accumulate = itertools.accumulate
        import torch.nn.functional as F
        from multimodn.multimodn import MultiModN as RefModN
        from multimodn.encoders import MIMIC_MLPEncoder, MLPEncoder
        from multimodn.decoders import MLPDecoder, LogisticDecoder
        from multimodn_b200 import MultiModN
        torch.manual_seed(0)
        encs = [MIMIC_MLPEncoder(8, 5, (6,), dropout=0.0), MLPEncoder(8, 4, (7, 3), F.relu)]
        decs = [MLPDecoder(8, (5,), 2), LogisticDecoder(8)]
        ref = RefModN(8, encs, decs, 1.0, 0.3, device=torch.device("cpu"))
        x = [torch.randn(40, 5), torch.randn(40, 4)]
        want = ref.predict(x)
        mine = MultiModN(8, encs, decs, 1.0, 0.3, device=torch.device("cpu"), init_state=ref.init_state)
        assert (mine.predict(x) == want).all()
        assert sorted(mine.state_dict()) == sorted(ref.state_dict())
    finally:
        sys.path.remove(REF)
        for name in [m for m in sys.modules if m == "multimodn" or m.startswith("multimodn.")]:
            del sys.modules[name]
