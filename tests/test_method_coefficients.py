"""CPU: the integrator coefficient sets compiled into the kernels (csrc/pk_common.cuh) satisfy
their defining conditions — parsed from the header, so a typo in a constant fails here."""
import os
import re
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))


def _parse_methods():
    text = open(os.path.join(ROOT, "phoskintime_b200", "csrc", "pk_common.cuh")).read()
    out = {}
    for name in ("METHOD_RODAS4", "METHOD_ROS5L", "METHOD_ROS6L"):
        body = text[text.index(f"constexpr Method {name}"):]
        body = body[body.index("{") + 1:body.index("};")]
        nsol = int(re.findall(r",\s*(\d+)\s*$", body)[0])                       # last field: solves per step
        body = body.replace("1.0f / 6.0f", repr(1.0 / 6.0))
        nums = [float(x.rstrip("f")) for x in re.findall(r"-?\d+\.\d+(?:e[+-]\d+)?f?", body)]
        assert len(nums) == 16, (name, nums)                                     # gamma, mu[7], eps[7], expo
        assert np.all(np.array(nums[1:8])[nsol:] == 0.0) and np.all(np.array(nums[8:15])[nsol:] == 0.0)
        out[name] = {"gamma": nums[0], "mu": np.array(nums[1:1 + nsol]), "eps": np.array(nums[8:8 + nsol]),
                     "expo": nums[15], "nsol": nsol}
    return out


def _R(mu, gamma, z):
    w = 1.0 / (1.0 - gamma * z)
    return 1.0 + z * sum(m * w ** (k + 1) for k, m in enumerate(mu))


def test_coefficients_match_derivation_scripts():
    m = _parse_methods()
    import derive_ros5l
    mu, eps = derive_ros5l.design()
    assert np.allclose(m["METHOD_ROS5L"]["mu"], [float(x) for x in mu], rtol=0, atol=1e-16)
    assert np.allclose(m["METHOD_ROS5L"]["eps"], [float(x) for x in eps], rtol=0, atol=1e-15)
    assert m["METHOD_ROS5L"]["gamma"] == float(derive_ros5l.GAMMA) and abs(m["METHOD_ROS5L"]["expo"] - 0.2) < 1e-7
    import derive_ros6l
    mu6, eps6, _ = derive_ros6l.design()
    assert m["METHOD_ROS6L"]["nsol"] == 7 and m["METHOD_ROS6L"]["gamma"] == float(derive_ros6l.GAMMA)
    assert np.allclose(m["METHOD_ROS6L"]["mu"], [float(x) for x in mu6], rtol=0, atol=2e-15)
    assert np.allclose(m["METHOD_ROS6L"]["eps"], [float(x) for x in eps6], rtol=0, atol=2e-14)
    assert abs(m["METHOD_ROS6L"]["expo"] - 1.0 / 6.0) < 1e-7
    import derive_rodas4_linear
    mu4, eps4 = derive_rodas4_linear.derive()
    assert np.allclose(m["METHOD_RODAS4"]["mu"], [float(x) for x in mu4[1:]], rtol=0, atol=1e-16)
    assert np.allclose(m["METHOD_RODAS4"]["eps"], [float(x) for x in eps4[1:]], rtol=0, atol=1e-16)


def test_order_and_stability_of_compiled_coefficients():
    for name, order in (("METHOD_RODAS4", 4), ("METHOD_ROS5L", 5), ("METHOD_ROS6L", 6)):
        c = _parse_methods()[name]
        g, mu, eps = c["gamma"], c["mu"], c["eps"]
        assert mu[0] == g and eps[0] == 0.0                        # L-stable main and embedded solutions
        # order: R(z) - exp(z) = O(z^(order+1)), estimator = O(z^order)
        # (ROS6L sits next to the member of its family whose z^7 error coefficient vanishes: C7 = 1.1e-6, so at a
        #  testable z the observed slope is between 7 and 8 — checked as "at least order 6")
        za, zb = (0.02, 0.04) if order < 6 else (0.1, 0.2)
        e1 = abs(_R(mu, g, za) - np.exp(za))
        e2 = abs(_R(mu, g, zb) - np.exp(zb))
        if order < 6:
            assert abs(np.log2(e2 / e1) - (order + 1)) < 0.1
        else:
            assert order + 0.9 < np.log2(e2 / e1) < order + 2.1
        s1 = abs(0.02 * sum(e * (1 / (1 - g * 0.02)) ** (k + 1) for k, e in enumerate(eps)))
        s2 = abs(0.04 * sum(e * (1 / (1 - g * 0.04)) ** (k + 1) for k, e in enumerate(eps)))
        assert abs(np.log2(s2 / s1) - order) < 0.1
        # A-stability on the imaginary axis and damping on the negative real axis
        y = np.concatenate([np.linspace(0, 60, 12001), np.logspace(1.7, 8, 1000)])
        assert np.abs(_R(mu, g, 1j * y)).max() <= 1.0 + 1e-12
        x = -np.logspace(-3, 8, 2000)
        assert np.abs(_R(mu, g, x)).max() < 1.0 and abs(_R(mu, g, -1e8)) < 1e-6


def test_runtime_ros5l_family_matches_design_and_stays_stable():
    """`ros5l_coeffs` (csrc/pk_common.cuh; the dense kernel evaluates it on the device for steps that reuse an
    inverse computed for a larger step) through its host export `pk_ros5l_coeffs`:
    at gamma = 0.19 it reproduces the compiled ROS5L constants; for every gamma' the kernel can request
    (0.19/1.03 .. 0.19*16) the member has order 5 with an order-4 estimate, MU[0] = gamma' (L-stable), is stable on
    the negative real axis and in the 60-degree sector, and a shortened step's local error stays within 1.5x of the
    full step's (and drops quickly for larger gamma')."""
    import ctypes as C
    from phoskintime_b200 import _lib
    lib = _lib.load()

    def coeffs(g):
        mu, eps = (C.c_double * 6)(), (C.c_double * 6)()
        assert lib.pk_ros5l_coeffs(g, mu, eps) == 0
        return np.array(mu[:]), np.array(eps[:])

    ref = _parse_methods()["METHOD_ROS5L"]
    mu0, eps0 = coeffs(0.19)
    assert np.allclose(mu0, ref["mu"], rtol=1e-10, atol=1e-11) and np.allclose(eps0, ref["eps"], rtol=1e-9, atol=1e-10)
    from math import comb, factorial
    c6 = lambda g, mu: 1 / factorial(6) - sum(mu[k] * comb(k + 5, 5) * g ** 5 for k in range(6))
    C60 = abs(c6(0.19, mu0))
    assert abs(C60 - 7.6482963444444444e-5) < 1e-12
    x = -np.logspace(-3, 8, 3000)
    sector = np.logspace(-3, 8, 3000) * np.exp(1j * (np.pi - np.pi / 3))
    for g in (0.19 / 1.03, 0.2, 0.25, 0.3, 0.334, 0.5, 0.76, 1.0, 1.9, 3.04):
        mu, eps = coeffs(g)
        assert mu[0] == g and eps[0] == 0.0
        z1, z2 = 0.01 / max(g, 0.19), 0.02 / max(g, 0.19)
        e1, e2 = abs(_R(mu, g, z1) - np.exp(z1)), abs(_R(mu, g, z2) - np.exp(z2))
        if abs(c6(g, mu)) > 1e-5:                                  # (C6 changes sign near gamma = 0.334)
            assert abs(np.log2(e2 / e1) - 6) < 0.2, (g, np.log2(e2 / e1))
        s1 = abs(z1 * sum(e * (1 / (1 - g * z1)) ** (k + 1) for k, e in enumerate(eps)))
        s2 = abs(z2 * sum(e * (1 / (1 - g * z2)) ** (k + 1) for k, e in enumerate(eps)))
        assert abs(np.log2(s2 / s1) - 5) < 0.1
        assert np.abs(_R(mu, g, x)).max() < 1.0 and abs(_R(mu, g, -1e8)) < 1e-6
        assert np.abs(_R(mu, g, sector)).max() <= 1.0 + 1e-12
        # local error of the shortened step (gamma' = g, step h*0.19/g) relative to the full step at gamma = 0.19:
        # at most 1.5x (gamma' in 0.19..0.27, where C6 grows faster than the step shrinks), far smaller beyond
        rel = abs(c6(g, mu)) / C60 * (0.19 / g) ** 6
        assert rel <= 1.5 and (g < 0.3 or rel < 0.7)
    with np.errstate(all="ignore"):
        assert lib.pk_ros5l_coeffs(-1.0, (C.c_double * 6)(), (C.c_double * 6)()) != 0


def test_runtime_ros6l_family_matches_design_and_stays_stable():
    """`rosl_coeffs<7>` (the dense kernel re-derives ROS6L(gamma') for steps that reuse an inverse) through its host export
    `pk_ros6l_coeffs`: at gamma = 0.205 it reproduces the compiled ROS6L constants; every member the kernel can request
    (0.205/1.03 .. 0.205*16) has MU[0] = gamma' (L-stable), EPS[0] = 0, an order-5 embedded solution (estimate O(z^6)),
    order >= 6, and is stable on the negative real axis and in the 60-degree sector."""
    import ctypes as C
    from phoskintime_b200 import _lib
    lib = _lib.load()

    def coeffs(g):
        mu, eps = (C.c_double * 7)(), (C.c_double * 7)()
        assert lib.pk_ros6l_coeffs(g, mu, eps) == 0
        return np.array(mu[:]), np.array(eps[:])

    ref = _parse_methods()["METHOD_ROS6L"]
    mu0, eps0 = coeffs(0.205)
    assert np.allclose(mu0, ref["mu"], rtol=1e-9, atol=1e-10) and np.allclose(eps0, ref["eps"], rtol=1e-8, atol=1e-9)
    x = -np.logspace(-3, 8, 3000)
    sector = np.logspace(-3, 8, 3000) * np.exp(1j * (np.pi - np.pi / 3))
    for g in (0.205 / 1.03, 0.21, 0.25, 0.3, 0.41, 0.5, 0.82, 1.0, 1.64, 3.28):
        mu, eps = coeffs(g)
        assert mu[0] == g and eps[0] == 0.0
        z1 = 0.05 / max(1.0, g / 0.3)
        z2 = 2.0 * z1
        e1, e2 = abs(_R(mu, g, 4 * z1) - np.exp(4 * z1)), abs(_R(mu, g, 4 * z2) - np.exp(4 * z2))
        assert np.log2(e2 / e1) > 6.7, (g, np.log2(e2 / e1))                  # local error O(z^7) or better
        s1 = abs(z1 * sum(e * (1 / (1 - g * z1)) ** (k + 1) for k, e in enumerate(eps)))
        s2 = abs(z2 * sum(e * (1 / (1 - g * z2)) ** (k + 1) for k, e in enumerate(eps)))
        assert abs(np.log2(s2 / s1) - 6) < 0.3, (g, np.log2(s2 / s1))
        assert np.abs(_R(mu, g, x)).max() < 1.0 and abs(_R(mu, g, -1e8)) < 1e-6
        assert np.abs(_R(mu, g, sector)).max() <= 1.0 + 1e-12
    assert lib.pk_ros6l_coeffs(-1.0, (C.c_double * 7)(), (C.c_double * 7)()) != 0
