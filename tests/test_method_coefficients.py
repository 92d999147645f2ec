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
    for name in ("METHOD_RODAS4", "METHOD_ROS5L"):
        body = text[text.index(f"constexpr Method {name}"):]
        body = body[body.index("{") + 1:body.index("};")]
        nums = [float(x.rstrip("f")) for x in re.findall(r"-?\d+\.\d+(?:e[+-]\d+)?f?", body)]
        assert len(nums) == 14, (name, nums)
        out[name] = {"gamma": nums[0], "mu": np.array(nums[1:7]), "eps": np.array(nums[7:13]), "expo": nums[13]}
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
    import derive_rodas4_linear
    mu4, eps4 = derive_rodas4_linear.derive()
    assert np.allclose(m["METHOD_RODAS4"]["mu"], [float(x) for x in mu4[1:]], rtol=0, atol=1e-16)
    assert np.allclose(m["METHOD_RODAS4"]["eps"], [float(x) for x in eps4[1:]], rtol=0, atol=1e-16)


def test_order_and_stability_of_compiled_coefficients():
    for name, order in (("METHOD_RODAS4", 4), ("METHOD_ROS5L", 5)):
        c = _parse_methods()[name]
        g, mu, eps = c["gamma"], c["mu"], c["eps"]
        assert mu[0] == g and eps[0] == 0.0                        # L-stable main and embedded solutions
        # order: R(z) - exp(z) = O(z^(order+1)), estimator = O(z^order)
        e1 = abs(_R(mu, g, 0.02) - np.exp(0.02))
        e2 = abs(_R(mu, g, 0.04) - np.exp(0.04))
        assert abs(np.log2(e2 / e1) - (order + 1)) < 0.1
        s1 = abs(0.02 * sum(e * (1 / (1 - g * 0.02)) ** (k + 1) for k, e in enumerate(eps)))
        s2 = abs(0.04 * sum(e * (1 / (1 - g * 0.04)) ** (k + 1) for k, e in enumerate(eps)))
        assert abs(np.log2(s2 / s1) - order) < 0.1
        # A-stability on the imaginary axis and damping on the negative real axis
        y = np.concatenate([np.linspace(0, 60, 12001), np.logspace(1.7, 8, 1000)])
        assert np.abs(_R(mu, g, 1j * y)).max() <= 1.0 + 1e-12
        x = -np.logspace(-3, 8, 2000)
        assert np.abs(_R(mu, g, x)).max() < 1.0 and abs(_R(mu, g, -1e8)) < 1e-6
