"""CPU: product path on the emulated kernels vs the oracle — multi-tile batches, ragged tiles,
all tile heights, MNAR missingness, dropout (tests/parity_cases.py)."""
import pytest

from parity_cases import CASES, run_parity_case


@pytest.mark.parametrize("name", sorted(CASES))
def test_parity_case(emu, name):
    run_parity_case(name, "cpu")
