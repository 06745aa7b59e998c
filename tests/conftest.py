import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_package():
    """The package directory is 'sc-a-loam_b200' (hyphen): load it as module sc_a_loam_b200."""
    if "sc_a_loam_b200" in sys.modules:
        return sys.modules["sc_a_loam_b200"]
    d = os.path.join(ROOT, "sc-a-loam_b200")
    spec = importlib.util.spec_from_file_location("sc_a_loam_b200", os.path.join(d, "__init__.py"),
                                                  submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["sc_a_loam_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def s2m():
    return load_package()


@pytest.fixture(scope="session")
def built():
    """Build every native library once per session (no-op when up to date)."""
    import harness
    import oracle
    harness.build()
    oracle.build()
    spec = importlib.util.spec_from_file_location("s2m_build", os.path.join(ROOT, "sc-a-loam_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build_hostmath()
    if not os.path.exists(b.LIB):
        b.build_cuda()
    return b


def quat_mul(a, b):
    import numpy as np
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def rot_angle(qa, qb):
    """angle (rad) of qa^-1 * qb"""
    import numpy as np
    d = quat_mul(np.array([-qa[0], -qa[1], -qa[2], qa[3]]), qb)
    return 2.0 * np.arcsin(min(1.0, float(np.linalg.norm(d[:3]))))
