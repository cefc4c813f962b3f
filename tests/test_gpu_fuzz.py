"""A short randomised parity campaign on the device (tests/fuzz_cases.py): 150 random graph / sampler / tuning combinations,
every one bit-identical to the CPU oracle.  (tools/fuzz_parity.py runs thousands; 1500 cases were clean at the end of round 1.)"""
import pytest

import fuzz_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_cases_are_bit_exact(gpu, seed):
    bad, _ = fuzz_cases.run_cases(gpu, 50, seed)
    assert bad == 0


@pytest.mark.parametrize("seed", [21, 22])
def test_random_cases_sequential_chains(gpu, seed):
    """The same campaign through the sequential-chain schedule (zz_seq.cuh): plain ZigZag on random lattices / sparse graphs, some
    of them cut into scattered independent chains; 1, 2 or 4 warps per chain."""
    bad, _ = fuzz_cases.run_cases(gpu, 40, seed, seq=True)
    assert bad == 0
