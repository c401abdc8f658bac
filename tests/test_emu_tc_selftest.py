"""CPU: the tensor-core engine's staging / descriptor arithmetic against the functional tcgen05 model
of the emulator build (mmn_tc.cuh, #ifdef MMN_EMU).  The real instruction is checked by
tests/test_gpu_tc_selftest.py."""
import pytest

from tc_selftest_common import run_selftest


@pytest.mark.parametrize("mode,n", [(0, 32), (0, 64), (1, 32), (1, 64), (2, 32), (3, 32), (3, 64), (4, 32), (4, 64)])
def test_umma_model(emu, mode, n):
    assert run_selftest(emu, "cpu", mode, n) < 2e-6
