"""CPU: product path on the emulated kernels vs the oracle — multi-tile batches, ragged tiles,
all tile heights, MNAR missingness, dropout (tests/parity_cases.py)."""
import pytest

from parity_cases import CASES, run_parity_case


@pytest.mark.parametrize("name", sorted(CASES))
def test_parity_case(emu, name):
    run_parity_case(name, "cpu")


@pytest.mark.parametrize("name", ["multi_tile_ragged", "mnar_rows", "mlp_kind_wide_hidden"])
def test_forward_engine_is_tmem_resident_where_the_model_qualifies(emu, name):
    """run_parity_case's predict / get_states / test legs go through the TMEM-resident forward kernel on the emulator's
    tcgen05 / TMEM model"""
    model = run_parity_case(name, "cpu")
    assert emu.dll.mmn_plan_engine(model.runtime().plan) == 0
    assert emu.dll.mmn_plan_forward_engine(model.runtime().plan) == 2
