"""hypot_cr (stair_step_detector_b200/csrc/ssd_device.cuh) stands in for std::hypot (segmentation.cpp:346-349, :381-385) in the
outline kernels. The function is compiled here for the HOST from the same header (nvcc, no GPU needed) and checked against
exact arithmetic: it must be the correctly rounded square root of x^2 + y^2 on the operand classes of the path -- integer
line coefficients of images up to 4096 x 3072, components of normalised lines and of their bisectors.
glibc's hypot, which the reference calls, is itself NOT correctly rounded (about 0.6 % of operand pairs come out one ulp off,
and the value depends on whether the CPU dispatches to the FMA variant); the test records how often the two differ so the
claim in the header stays honest. A last-bit difference of a line normalisation is far below the 0.1 mm parity bar."""
import decimal
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def exact_hypot(x, y):
    decimal.getcontext().prec = 80
    return float((decimal.Decimal(float(x)) ** 2 + decimal.Decimal(float(y)) ** 2).sqrt())  # Decimal -> float rounds to nearest even


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not found")
def test_hypot_cr_is_correctly_rounded(tmp_path):
    exe = os.path.join(ROOT, "build", "hypot_check")
    src = os.path.join(ROOT, "tests", "native", "hypot_check.cu")
    hdr = os.path.join(ROOT, "stair_step_detector_b200", "csrc", "ssd_device.cuh")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run([NVCC, "-O2", "-std=c++17", "-Xcompiler", "-ffp-contract=off", "-o", exe, src], check=True, capture_output=True)
    rng = np.random.default_rng(7)
    n = 12000
    ints = np.stack([rng.integers(-3072, 3073, n), rng.integers(-4096, 4097, n)], 1).astype(np.float64)
    unit = rng.uniform(-2.0, 2.0, (n, 2))
    mixed = rng.uniform(-1.0, 1.0, (n, 2)) * np.exp2(rng.integers(-12, 12, (n, 2)))
    spec = np.array([[a, b] for a in (0.0, 1.0, -1.0, 3.0, 4.0, 1e-3, 4096.0, 0.5, 1 / 3) for b in (0.0, 1.0, -1.0, 3.0, 4.0, 1e-3, 4096.0, 0.5, 1 / 3)])
    pairs = np.concatenate([ints, unit, mixed, spec])
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    pairs.tofile(fin)
    subprocess.run([exe, fin, fout], check=True)
    out = np.fromfile(fout, np.float64).reshape(-1, 2)
    exact = np.array([exact_hypot(x, y) for x, y in pairs])
    assert np.array_equal(out[:, 0], exact), f"hypot_cr differs from the correctly rounded value on {(out[:, 0] != exact).sum()} of {len(exact)} pairs"
    # libm on this machine (informational): within one ulp of the exact value, not always equal to it
    ulp = np.spacing(exact)
    assert np.all(np.abs(out[:, 1] - exact) <= ulp)
    print(f"libm hypot differs from the correctly rounded value on {(out[:, 1] != exact).sum()} of {len(exact)} pairs")
