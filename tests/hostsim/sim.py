"""TEST INFRASTRUCTURE ONLY: ctypes driver for the host build of csrc/rcbf_core.cuh (see hostsim.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def build(force=False, contract="off"):
    so = os.path.join(HERE, "libhostsim.so")
    src = os.path.join(HERE, "hostsim.cpp")
    csrc = os.path.join(HERE, "..", "..", "sac_rcbf_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("rcbf_core.cuh", "rcbf_backward.cuh", "rcbf_f2.cuh")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-march=native", "-ffp-contract=" + contract, "-fopenmp", "-shared",
                               "-fPIC", "-x", "c++", src, "-o", so])
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def unicycle_safe_action(st, ac, mu, sg, params, mode=0):
    n = st.shape[0]
    st, ac, mu, sg = (np.ascontiguousarray(a, np.float32) for a in (st, ac, mu, sg))
    o = dict(out=np.zeros((n, 2), np.float32), x=np.zeros((n, 3)), lam=np.zeros((n, 9)), s=np.zeros((n, 9)),
             status=np.zeros(n, np.int32), iters=np.zeros(n, np.int32), Gn=np.zeros((n, 9, 3), np.float32),
             hn=np.zeros((n, 9), np.float32), G=np.zeros((n, 9, 3), np.float32), h=np.zeros((n, 9), np.float32))
    lib().hs_unicycle_safe_action(C.c_int64(n), _p(st, C.c_float), _p(ac, C.c_float), _p(mu, C.c_float),
                                  _p(sg, C.c_float), C.byref(params), _p(o["out"], C.c_float), _p(o["x"], C.c_double),
                                  _p(o["lam"], C.c_double), _p(o["s"], C.c_double), _p(o["status"], C.c_int),
                                  _p(o["iters"], C.c_int), _p(o["Gn"], C.c_float), _p(o["hn"], C.c_float),
                                  _p(o["G"], C.c_float), _p(o["h"], C.c_float), C.c_int(mode))
    return o


def cars_safe_action(st, ac, sg, params, mode=0):
    n = st.shape[0]
    st, ac, sg = (np.ascontiguousarray(a, np.float32) for a in (st, ac, sg))
    o = dict(out=np.zeros((n, 1), np.float32), x=np.zeros((n, 2)), lam=np.zeros((n, 4)), s=np.zeros((n, 4)),
             status=np.zeros(n, np.int32), iters=np.zeros(n, np.int32), Gn=np.zeros((n, 4, 2), np.float32),
             hn=np.zeros((n, 4), np.float32), G=np.zeros((n, 4, 2), np.float32), h=np.zeros((n, 4), np.float32))
    lib().hs_cars_safe_action(C.c_int64(n), _p(st, C.c_float), _p(ac, C.c_float), _p(sg, C.c_float), C.byref(params),
                              _p(o["out"], C.c_float), _p(o["x"], C.c_double), _p(o["lam"], C.c_double),
                              _p(o["s"], C.c_double), _p(o["status"], C.c_int), _p(o["iters"], C.c_int),
                              _p(o["Gn"], C.c_float), _p(o["hn"], C.c_float), _p(o["G"], C.c_float),
                              _p(o["h"], C.c_float), C.c_int(mode))
    return o


def unicycle_bwd(st, ac, mu, sg, gout, params):
    """(grad_action of the qpth-clamp form on dense saved tensors, of the exact active-set form on the mask, status)."""
    n = st.shape[0]
    st, ac, mu, sg, gout = (np.ascontiguousarray(a, np.float32) for a in (st, ac, mu, sg, gout))
    gd, ga, status = np.zeros((n, 2), np.float32), np.zeros((n, 2), np.float32), np.zeros(n, np.int32)
    lib().hs_unicycle_bwd(C.c_int64(n), _p(st, C.c_float), _p(ac, C.c_float), _p(mu, C.c_float), _p(sg, C.c_float),
                          _p(gout, C.c_float), C.byref(params), _p(gd, C.c_float), _p(ga, C.c_float), _p(status, C.c_int))
    return gd, ga, status


def cars_bwd(st, ac, sg, gout, params):
    n = st.shape[0]
    st, ac, sg, gout = (np.ascontiguousarray(a, np.float32) for a in (st, ac, sg, gout))
    gd, ga, status = np.zeros((n, 1), np.float32), np.zeros((n, 1), np.float32), np.zeros(n, np.int32)
    lib().hs_cars_bwd(C.c_int64(n), _p(st, C.c_float), _p(ac, C.c_float), _p(sg, C.c_float), _p(gout, C.c_float),
                      C.byref(params), _p(gd, C.c_float), _p(ga, C.c_float), _p(status, C.c_int))
    return gd, ga, status


def unicycle_general(st, ac, mu, sg, gout, params, hazards):
    """Layer on K = len(hazards) (1..12) hazards: (safe action, grad_action of the exact active-set form, status)."""
    n = st.shape[0]
    st, ac, mu, sg, gout = (np.ascontiguousarray(a, np.float32) for a in (st, ac, mu, sg, gout))
    hz = np.ascontiguousarray(hazards, np.float32)
    out, ga, status = np.zeros((n, 2), np.float32), np.zeros((n, 2), np.float32), np.zeros(n, np.int32)
    lib().hs_unicycle_general(C.c_int64(n), _p(st, C.c_float), _p(ac, C.c_float), _p(mu, C.c_float), _p(sg, C.c_float),
                              _p(gout, C.c_float), C.byref(params), _p(hz, C.c_float), C.c_int(hz.shape[0]),
                              _p(out, C.c_float), _p(ga, C.c_float), _p(status, C.c_int))
    return out, ga, status
