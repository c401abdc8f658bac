"""GPU: tcgen05.mma kind::tf32 with the engine's SWIZZLE_128B operand images in all three operand
configurations (K-major x K-major, K-major x MN-major, MN-major x MN-major), 3xTF32 -> fp32 accuracy."""
import pytest

pytestmark = pytest.mark.gpu

from tc_selftest_common import run_selftest  # noqa: E402


@pytest.mark.parametrize("mode,n", [(0, 32), (0, 64), (1, 32), (1, 64), (2, 32), (3, 32), (3, 64), (4, 32), (4, 64)])
def test_umma_hardware(mode, n):
    from multimodn_b200 import _lib
    err = run_selftest(_lib.get_lib(), "cuda", mode, n)
    assert err < 2e-6, f"mode {mode} n {n}: max err / max|ref| = {err:.3e}"
