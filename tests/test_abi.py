"""CPU checks of the drop-in boundary: the C-ABI library builds/loads and exports every symbol include/rcbf_b200.h
declares; the ctypes parameter structs match the C structs; the product fails loudly without CUDA; the product package
never imports the oracle."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rcbf_b200.h")


@pytest.fixture(scope="module")
def lib():
    from sac_rcbf_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(rcbf_[a-z0-9_]+)\s*\(", src))
    # the env block is declared once per precision suffix
    return sorted(names)


def test_library_exports_every_declared_symbol(lib):
    from sac_rcbf_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 26
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    # and the Python binding covers them all (no silently unbound entry point)
    bound = set(_lib.SIGNATURES) | {"rcbf_version"}
    assert set(names) == bound, set(names) ^ bound
    assert b"sm_100a" in lib.rcbf_version()


def test_ctypes_structs_match_c_layout(tmp_path):
    from sac_rcbf_b200 import _params as P
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rcbf_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(rcbf_unicycle_params),sizeof(rcbf_cars_params),sizeof(rcbf_unicycle_env_params),'
                    'sizeof(rcbf_cars_env_params),offsetof(rcbf_unicycle_params,p_diag),'
                    'offsetof(rcbf_unicycle_env_params,max_episode_steps),offsetof(rcbf_cars_params,slack_coeff),'
                    'sizeof(rcbf_gp_posterior),offsetof(rcbf_gp_posterior,n_pad),offsetof(rcbf_gp_posterior,min_variance),sizeof(rcbf_replay_ring),offsetof(rcbf_replay_ring,obs_dim));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(P.UnicycleParams), C.sizeof(P.CarsParams), C.sizeof(P.UnicycleEnvParams), C.sizeof(P.CarsEnvParams),
            P.UnicycleParams.p_diag.offset, P.UnicycleEnvParams.max_episode_steps.offset, P.CarsParams.slack_coeff.offset,
            C.sizeof(P.GpPosterior), P.GpPosterior.n_pad.offset, P.GpPosterior.min_variance.offset,
            C.sizeof(P.ReplayRing), P.ReplayRing.obs_dim.offset]
    assert got == want


def test_header_is_plain_c_and_the_pure_c_driver_links(lib, tmp_path):
    """include/rcbf_b200.h must compile as C (no C++/torch types) and every entry the C driver uses must resolve
    against the built library (the driver itself runs in the GPU suite)."""
    from sac_rcbf_b200 import build
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I",
                           os.path.join(cuda, "include"), os.path.join(ROOT, "tests", "c_abi", "c_abi_smoke.c"), "-o",
                           exe, "-L", os.path.dirname(build.LIB), "-lrcbf_b200", "-L", os.path.join(cuda, "lib64"),
                           "-lcudart", "-lm"])
    assert os.path.exists(exe)


def test_params_mirror_reference_constants():
    from sac_rcbf_b200 import _params as P
    p = P.unicycle_params(gamma_b=20.0)
    assert abs(p.collision_radius_sq - (1.2 * 0.6) ** 2) < 1e-7 and list(p.p_diag) == [1.0, pytest.approx(1e-2), 1e5]
    assert [list(r) for r in p.hazards] == [[0, 0], [-1.5, 1.5], [-1.5, -1.5], [1.5, -1.5], [1.5, 1.5]]
    c = P.cars_params(gamma_b=20.0)
    assert c.gamma_2 == 40.0 and c.gamma_sq == 400.0 and c.slack_coeff == 200.0 and c.collision_radius_sq == 12.25
    with pytest.raises(ValueError):
        P.unicycle_params(hazards_locations=[[0, 0]] * 6)
    p3 = P.unicycle_params(hazards_locations=[[0, 0]] * 3)       # padded with inert far-away hazards
    assert p3.hazards[3][0] == P.FAR_HAZARD and p3.hazards[4][1] > P.FAR_HAZARD


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_fails_loudly_without_cuda():
    import types
    import sac_rcbf_b200 as S
    with pytest.raises(S.RcbfLibraryError):
        S.UnicycleEnv()
    env = types.SimpleNamespace(dynamics_mode="Unicycle")
    with pytest.raises(S.RcbfLibraryError):
        S.CBFQPLayer(env, types.SimpleNamespace(cuda=True))
    with pytest.raises(S.RcbfLibraryError):
        S.DynamicsModel(env, types.SimpleNamespace(cuda=True, gp_model_size=10))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sac_rcbf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "hostsim" not in src or f == "rcbf_core.cuh", f
    code = "import sys; import sac_rcbf_b200, sac_rcbf_b200._lib, sac_rcbf_b200._params, sac_rcbf_b200.diff_cbf_qp; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
