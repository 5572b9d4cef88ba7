"""CPU check of the product's host/device math header against the oracle (no GPU needed)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_math_matches_oracle():
    import oracle

    oracle.lib()  # builds oracle/_build/*.o if needed
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "host_math_check")
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-ffp-contract=off",
                    "-o", exe, os.path.join(ROOT, "tests", "host_math_check.cu"),
                    os.path.join(ROOT, "oracle", "_build", "cvmodels.o"), "-lm"], check=True, cwd=ROOT)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:]


def test_restated_logf_equals_libm():
    """MapPoint::PredictScale takes libm's logf of a distance ratio; glibc's logf is not correctly rounded, so the kernels
    follow its algorithm (dvmslam_b200/csrc/glibc_logf.h).  oracle/check_logf.c compares the restatement with this box's
    libm over every float in [2^-20, 2^20] (335 M inputs, a few seconds)."""
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "check_logf")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "oracle", "check_logf.c"), "-lm"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith(" 0 differ"), r.stdout
