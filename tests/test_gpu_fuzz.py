"""A short randomised parity campaign on the device (tests/fuzz_cases.py): 150 random graph / sampler / tuning combinations,
every one bit-identical to the CPU oracle.  (tools/fuzz_parity.py runs thousands; 1500 cases were clean at the end of round 1.)"""
import pytest

import fuzz_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_cases_are_bit_exact(gpu, seed):
    bad, _ = fuzz_cases.run_cases(gpu, 50, seed)
    assert bad == 0
