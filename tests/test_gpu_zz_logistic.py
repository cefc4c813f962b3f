"""Config 3 of BASELINE.json on the device: sparse logistic regression, subsampled ZigZag (scripts/logistic.jl) through
zzb_problem_create_logistic + zzb_spdmp_run, bit for bit against the CPU oracle (mode ctr|lazy)."""
import numpy as np
import pytest

import logistic_cases as LC
import oracle_lib as O

# Named test_gpu_zz_* so that it runs after the other GPU files (a d = 442 problem with a complete dependency graph is the
# slowest thing in the suite).  Seen green on a B200 in round 1 (profiles/r01f_logit_tests.log).
pytestmark = pytest.mark.gpu


# schedule 2 = sequential chains (zz_seq.cuh, the default for this target), 1 = the windowed relaxation (zz_run_kernel_csr_logit)
@pytest.mark.parametrize("schedule", [2, 1])
@pytest.mark.parametrize("case", LC.SMALL)
def test_small_designs_bit_exact(gpu, case, schedule):
    *design, T = case
    cfg = LC.make(gpu, *design)
    ref = LC.run_oracle(O, cfg, T)
    got, Xi = LC.run_device(gpu, cfg, T, tune=dict(schedule=schedule))
    assert (Xi.stats["windows"] == 0) == (schedule == 2)
    O.assert_same_run(ref, got)
    assert (got.c != cfg["c"]).any()                       # adapt = true multiplied some bounds by `factor`
    m1, m2 = Xi.moments
    assert np.allclose(m1, ref.m1, rtol=1e-12, atol=1e-300) and np.allclose(m2, ref.m2, rtol=1e-12, atol=1e-300)


def test_window_policy_and_tag_rebase_do_not_change_results(gpu):
    *design, T = LC.SMALL[1]
    cfg = LC.make(gpu, *design)
    ref = LC.run_oracle(O, cfg, T)
    for tune in (dict(delta0=1e-3, target_frac=0.1), dict(delta0=2.0, target_frac=4.0, tag_limit=40), dict(grid=3)):
        got, _ = LC.run_device(gpu, cfg, T, tune=dict(schedule=1, **tune))
        O.assert_same_run(ref, got)


def test_bound_violation_error(gpu):
    """adapt = false with the script's c = 0.01: error("Tuning parameter `c` too small."), sfact.jl:124."""
    cfg = LC.make(gpu, *LC.SMALL[0][:4])
    for schedule in (2, 1):
        with pytest.raises(gpu.BoundError, match="Tuning parameter `c` too small"):
            LC.run_device(gpu, cfg, 20.0, adapt=False, tune=dict(schedule=schedule))


def test_full_size_config3_bit_exact(gpu):
    """n = 8840, p = 442, k = 10, c = 0.01, adapt = true, factor = 5 (scripts/logistic.jl:21,148,167; README.md:50)."""
    cfg = LC.make(gpu, *LC.FULL)
    T = 25.0
    ref = LC.run_oracle(O, cfg, T)
    for schedule in (2, 1):
        got, Xi = LC.run_device(gpu, cfg, T, tune=dict(schedule=schedule))
        O.assert_same_run(ref, got)
        assert len(got.events) > 5000
        print(f"config 3, schedule {schedule}: {len(got.events)} events, {got.num} proposals in {Xi.device_ms:.1f} ms on the device "
              f"({ref.loop_seconds * 1e3:.1f} ms in the oracle); stats {Xi.stats}")


def test_refuses_unsupported_combinations(gpu):
    cfg = LC.make(gpu, *LC.SMALL[0][:4])
    lg = cfg["logistic"]
    grad = gpu.LogisticSubsampled(lg["A"], lg["At"], lg["y"], lg["ny"], lg["mu"], lg["gamma0"], lg["k"])
    Z = gpu.ZigZag(cfg["Gamma_drop"], cfg["mu"])
    with pytest.raises(NotImplementedError):
        gpu.spdmp(grad, 0.0, cfg["x0"], cfg["theta0"], 1.0, gpu.LocalBound(cfg["c"]), Z)
    prob = gpu.Problem(grad, Z)
    with pytest.raises(gpu.ZZBError):
        gpu.Run(prob, kappa=np.ones(cfg["p"]))
    bad = gpu.LogisticSubsampled(lg["A"], lg["A"], lg["y"], lg["ny"], lg["mu"], lg["gamma0"], lg["k"])   # At is not A'
    with pytest.raises(gpu.ZZBError):
        gpu.Problem(bad, Z)


def test_replicas_bit_exact(gpu):
    """R independent chains as one block-diagonal problem (config 3 = plumbing + replicas): bit-exact against the oracle on
    the joint problem; the device time per event falls with R because the passes of the chains overlap."""
    cfg = LC.make(gpu, *LC.FULL)
    R, T = 8, 4.0
    big = LC.replicas(gpu, cfg, R)
    ref = LC.run_oracle(O, big, T)
    for schedule in (2, 1):
        got, Xi = LC.run_device(gpu, big, T, tune=dict(schedule=schedule))
        O.assert_same_run(ref, got)
        one, Xi1 = LC.run_device(gpu, cfg, T, tune=dict(schedule=schedule))
        print(f"config 3 replicas, schedule {schedule}: R = {R}: {len(got.events)} events in {Xi.device_ms:.1f} ms; R = 1: {len(one.events)} "
              f"events in {Xi1.device_ms:.1f} ms; oracle (R = {R}) {ref.loop_seconds * 1e3:.1f} ms")


def test_random_designs_on_the_device(gpu):
    """tools/fuzz_logistic.py on the CUDA path: 25 random cases, bit-exact or a bound error on both sides."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("fuzz_logistic", os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools", "fuzz_logistic.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    n_ok, n_err = fz.run(gpu, 11, 25, gpu=True, verbose=False, schedule=2)     # sequential chains (the default for this target)
    assert n_ok >= 12
    n_ok, n_err = fz.run(gpu, 12, 12, gpu=True, verbose=False, schedule=1)     # windowed relaxation
    assert n_ok >= 5


@pytest.mark.auto_schedule
def test_device_reproduces_logistic_fixture(gpu):
    """tests/golden/logistic26.json (written from the oracle) through the CUDA path -- no oracle in the loop."""
    import json
    import os

    import golden_cases as GC
    for name in GC.LOGISTIC_CASES:
        g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", name + ".json")))
        assert GC.run_device(gpu, GC.case_inputs(gpu, name)) == g
