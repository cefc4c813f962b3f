"""The scalar primitives the kernels and the oracle share (csrc/zz_math.h: log, exp, sincos, poisson_time, the counter-based
uniforms) evaluated ON THE DEVICE (zzb_math_probe), pinned (i) bit for bit against the host build of the same header and
(ii) within 1 ulp against the platform's libm (numpy).  Without this, an error in a shared primitive would cancel in every
device == oracle comparison."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
N = 10_000_000


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def device(zzb, kind, x, y=None, z=None, two=False):
    from zzb200 import _capi
    o1, o2 = np.empty(len(x) if kind != 4 else N), (np.empty(len(x)) if two else None)
    _capi.check(_capi.lib().zzb_math_probe(kind, len(o1), ptr(x), ptr(y), ptr(z), ptr(o1), ptr(o2)))
    return o1, o2


def host(kind, x, y=None, z=None, two=False):
    L = O.lib()
    L.zzo_math_probe.restype = None
    L.zzo_math_probe.argtypes = [C.c_int, C.c_int64] + [C.c_void_p] * 5
    o1, o2 = np.empty(len(x) if kind != 4 else N), (np.empty(len(x)) if two else None)
    L.zzo_math_probe(kind, len(o1), ptr(x), ptr(y), ptr(z), ptr(o1), ptr(o2))
    return o1, o2


def ulps(a, b):
    """|a - b| in units of the spacing of b."""
    return np.abs(a - b) / np.spacing(np.abs(b))


def same_bits(a, b):
    return np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_log_device_vs_host_vs_libm(gpu):
    rng = np.random.default_rng(1)
    # the arguments the sampler uses are uniforms in (0, 1); add wide-range positives and values next to 1
    x = np.concatenate([rng.random(N // 2) * (1 - 2 ** -52) + 2 ** -53, np.exp(rng.uniform(-700, 700, N // 4)),
                        1.0 + rng.uniform(-1e-6, 1e-6, N // 4)])
    d, _ = device(gpu, 0, x)
    h, _ = host(0, x)
    assert same_bits(d, h)
    ref = np.log(x)
    ok = ref != 0
    assert ulps(d[ok], ref[ok]).max() <= 1.0


def test_exp_device_vs_host_vs_libm(gpu):
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(-40, 40, N // 2), rng.uniform(-745, 709, N // 4), rng.uniform(-1e-7, 1e-7, N // 4)])
    d, _ = device(gpu, 1, x)
    h, _ = host(1, x)
    assert same_bits(d, h)
    ref = np.exp(x)
    ok = ref > 1e-300   # (subnormal results: scaled by 2^-1000 in two roundings)
    assert ulps(d[ok], ref[ok]).max() <= 1.0


def test_sincos_device_vs_host_vs_libm(gpu):
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-50, 50, N // 2), rng.uniform(-1e4, 1e4, N // 2)])
    (ds, dc), (hs, hc) = device(gpu, 2, x, two=True), host(2, x, two=True)
    assert same_bits(ds, hs) and same_bits(dc, hc)
    # absolute error relative to the spacing at 1 (near the zeros of sin / cos a 2-term Cody-Waite reduction cannot do better)
    assert np.abs(ds - np.sin(x)).max() <= 4e-16 and np.abs(dc - np.cos(x)).max() <= 4e-16


def test_poisson_time_device_vs_host(gpu):
    """poisson_time(a, b, u) (src/poissontime.jl:8-30) over all six branches incl. b == 0 and the infinite cases."""
    rng = np.random.default_rng(4)
    n = N // 4
    a = rng.standard_normal(n) * 3
    b = rng.standard_normal(n) * 2
    b[::7] = 0.0
    a[::11] = 0.0
    u = rng.random(n) * (1 - 2 ** -52) + 2 ** -53
    d, _ = device(gpu, 3, a, b, u)
    h, _ = host(3, a, b, u)
    assert same_bits(d, h)
    fin = np.isfinite(d)
    assert 0.3 < fin.mean() < 0.9 and np.all(d[fin] >= 0)
    # integral identity of test/poisson.jl:20-50: int_0^tau (a + b t)^+ dt == -log(u) on the finite branch
    t = d[fin]; aa, bb, uu = a[fin], b[fin], u[fin]
    t0 = np.where((aa < 0) & (bb > 0), -aa / np.where(bb != 0, bb, 1), 0.0)   # rate is zero before t0
    integ = np.where(bb != 0, aa * (t - t0) + 0.5 * bb * (t * t - t0 * t0), aa * t)
    well = (bb == 0) | (np.abs(bb) > 1e-2)     # (the closed form cancels catastrophically for tiny |b|; the sampler never divides there)
    assert np.allclose(integ[well], -np.log(uu[well]), rtol=1e-7, atol=1e-9)


def test_counter_uniforms_device_vs_host(gpu):
    seed = np.array([0x1234567, 0x89abcdef], dtype=np.uint64).view(np.float64)
    d, _ = device(gpu, 4, seed[:1].copy(), seed[1:].copy())
    h, _ = host(4, seed[:1].copy(), seed[1:].copy())
    assert same_bits(d, h)
    assert d.min() >= 2 ** -53 and d.max() <= 1 - 2 ** -53 and abs(d.mean() - 0.5) < 1e-3
