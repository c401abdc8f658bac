"""CPU: product path on the emulated kernels vs the oracle — multi-tile batches, ragged tiles,
all tile heights, MNAR missingness, dropout (tests/parity_cases.py)."""
import pytest

from parity_cases import CASES, run_parity_case


@pytest.mark.parametrize("name", sorted(CASES))
def test_parity_case(emu, name):
    run_parity_case(name, "cpu")


@pytest.mark.parametrize("name", ["multi_tile_ragged", "mnar_rows", "dropout_mnar", "mlp_kind_wide_hidden"])
def test_parity_case_tc_engine(emu, monkeypatch, name):
    """the tcgen05 engine's staging / command protocol against the emulator's functional UMMA model"""
    monkeypatch.setenv("MMN_ENGINE", "tc")
    model = run_parity_case(name, "cpu")
    assert emu.dll.mmn_plan_engine(model.runtime().plan) == 1


@pytest.mark.parametrize("name", ["multi_tile_ragged", "mnar_rows", "dropout_mnar", "mlp_kind_wide_hidden"])
def test_parity_case_tmem_resident_engine(emu, monkeypatch, name):
    """training through the TMEM-resident kernel (forward + backward) on the emulator's tcgen05 / TMEM model"""
    monkeypatch.setenv("MMN_ENGINE", "tc2")
    model = run_parity_case(name, "cpu")
    assert emu.dll.mmn_plan_engine(model.runtime().plan) == 2
