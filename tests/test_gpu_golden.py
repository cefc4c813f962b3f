"""The CUDA path (through the C-ABI) against the committed fixtures of tests/golden/ -- no oracle in the loop."""
import json
import os

import pytest

import golden_cases as GC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GC.CASES)
def test_device_reproduces_fixture(gpu, name):
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", name + ".json")))
    assert GC.run_device(gpu, GC.case_inputs(gpu, name)) == g
