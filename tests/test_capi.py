"""CPU-side checks of the boundary: the C-ABI library loads without a GPU, exports every symbol the header declares,
refuses to compute without CUDA (no fallback), and the host graph builder rejects malformed matrices."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(zzb):
    hdr = open(os.path.join(ROOT, "include", "zzb200.h")).read()
    declared = sorted(set(re.findall(r"\bint32_t\s+(zzb_\w+)\s*\(", hdr)))
    assert declared and sorted(zzb._capi.SYMBOLS) == declared
    lib = ctypes.CDLL(zzb._capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert os.path.getsize(os.path.join(ROOT, "zigzagboomerang.jl_b200", "zzb200_kernels.cubin")) > 10000


def test_kernel_image_is_sm100a_and_uses_no_fma_contraction():
    import subprocess
    cubin = os.path.join(ROOT, "zigzagboomerang.jl_b200", "zzb200_kernels.cubin")
    out = subprocess.run(["cuobjdump", "-elf", cubin], capture_output=True, text=True).stdout
    assert "sm_100a" in out or "sm_100" in out
    names = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    for k in ("zz_run_kernel_grid", "zz_run_kernel_csr", "zz_run_kernel_grid_multi_lb", "zz_run_kernel_grid_boom", "zz_run_kernel_csr_boom", "zz_init_kernel", "zz_setup_kernel", "zz_export_kernel"):
        assert k in names


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback(zzb):
    with pytest.raises(zzb.ZZBError, match="no CPU fallback|not found"):
        zzb._capi._inited = False
        zzb.init(0)
    G = zzb.grid_precision(4)
    with pytest.raises(zzb.ZZBError):
        zzb.spdmp(zzb.GaussianPotential(G), 0.0, np.zeros(16), np.ones(16), 1.0, np.ones(16), zzb.ZigZag(G, np.zeros(16)))


def test_spdmp_rejects_closures(zzb):
    G = zzb.grid_precision(4)
    with pytest.raises(TypeError, match="descriptor"):
        zzb.spdmp(lambda x, i: 0.0, 0.0, np.zeros(16), np.ones(16), 1.0, np.ones(16), zzb.ZigZag(G, np.zeros(16)))


def test_gaussian_potential_is_callable_like_the_reference_closure(zzb):
    G = zzb.grid_precision(4)
    x = np.arange(16.0)
    g = zzb.GaussianPotential(G)
    dense = G.to_scipy().toarray()
    for i in (1, 6, 16):
        assert g(x, i) == pytest.approx(dense[:, i - 1] @ x)


def test_grid_precision_matches_the_reference_construction(zzb):
    """scripts/gridlaplace.jl:4-21 restated densely."""
    m, n = 4, 3
    S = np.zeros((m * n, m * n))
    lin = lambda i, j: i + j * m
    for i in range(m):
        for j in range(n):
            for i2, j2 in ((i + 1, j), (i, j + 1)):
                if i2 < m and j2 < n:
                    S[lin(i, j), lin(i2, j2)] -= 1
                    S[lin(i2, j2), lin(i, j)] -= 1
                    S[lin(i, j), lin(i, j)] += 1
                    S[lin(i2, j2), lin(i2, j2)] += 1
    G = zzb.grid_precision(m, n)
    assert np.array_equal(G.to_scipy().toarray(), 0.01 * np.eye(m * n) + S)
    assert zzb.grid_precision(100).nnz == 49600  # scripts/gaussianrandomfield.jl:19


def test_header_is_plain_c_and_links(tmp_path):
    """include/zzb200.h compiles as C99 with warnings as errors, and a C program binds every entry point of the shared
    library the way a foreign-function interface would (no GPU needed: zzb_init fails cleanly without a driver/device)."""
    import subprocess
    src = tmp_path / "abi.c"
    hdr = open(os.path.join(ROOT, "include", "zzb200.h")).read()
    names = sorted(set(re.findall(r"\bint32_t\s+(zzb_\w+)\s*\(", hdr)))
    body = "\n".join(f"    p[{k}] = (fn_t){n};" for k, n in enumerate(names))
    src.write_text(f"""#include <stdio.h>
#include "zzb200.h"
typedef void (*fn_t)(void);
int main(void) {{
    fn_t p[{len(names)}];
{body}
    char msg[256];
    int32_t st = zzb_init(1, NULL, NULL);
    zzb_last_error(msg, sizeof msg);
    zzb_event e; e.t = 0; e.i = 0; e.x = 0; e.theta = 0; (void)e;
    printf("%d %d %zu\\n", (int)st, (int)(p[0] != (fn_t)0), sizeof(zzb_event));
    return 0;
}}
""")
    exe = tmp_path / "abi"
    pkg = os.path.join(ROOT, "zigzagboomerang.jl_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe), "-L", pkg, "-lzzb200", f"-Wl,-rpath,{pkg}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert out[2] == "32"                         # sizeof(zzb_event) == sizeof(Tuple{Float64,Int64,Float64,Float64})
    if not os.path.exists("/dev/nvidia0"):
        assert out[0] == "2"                      # ZZB_E_CUDA: no driver / device, and no fallback
