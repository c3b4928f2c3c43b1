"""The drop-in boundary is a C ABI: drive it from a PURE C program (tests/c_abi/c_abi_smoke.c; no Python, no torch in
the process) that links librcbf_b200.so + libcudart and checks known answers from the reference source / the oracle.

The expected numbers in the C file were produced with:
    O.safe_action('Unicycle', st, ac, mu, sg, solver='exact', gamma_b=20.0)   and   O.assemble_unicycle(...)
on the three instances hard-coded there, plus the SURVEY 8(c) KATs of UnicycleEnv.reset()/step([1.0, 0.5])."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pure_c_driver(tmp_path):
    from sac_rcbf_b200 import build
    lib = build.build()
    exe = str(tmp_path / "c_abi_smoke")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["gcc", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                           os.path.join(ROOT, "tests", "c_abi", "c_abi_smoke.c"), "-o", exe,
                           "-L", os.path.dirname(lib), "-lrcbf_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart",
                           "-lm", "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath," + os.path.join(cuda, "lib64")])
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0 and "C ABI SMOKE OK" in r.stdout, r.stdout
