import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _make(*targets):
    subprocess.run(["make", "-C", ROOT, *targets], check=True, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def built():
    """All native pieces exist (built by __graft_entry__.build(); rebuilt here if a file is missing)."""
    need = ["raytracert_b200/_build/librt_host.so", "raytracert_b200/_build/librt_b200.so", "oracle/_build/librt_oracle.so"]
    if not all(os.path.exists(os.path.join(ROOT, p)) for p in need):
        _make("host", "cuda", "oracle")
    return True


@pytest.fixture(scope="session")
def port(built):
    from oracle import pyoracle
    return pyoracle.PortOracle()


@pytest.fixture(scope="session")
def ref(built):
    from oracle import pyoracle
    if not os.path.exists(pyoracle.REF_SO):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return pyoracle.RefOracle()


def load_scene(name):
    from raytracert_b200 import host
    return host.Scene.load(os.path.join(GOLDEN, "scenes", name + ".npz"))


def render_cases():
    d = os.path.join(GOLDEN, "renders")
    return sorted(f[:-4] for f in os.listdir(d) if f.endswith(".npz"))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, "renders", name + ".npz"))
    c = {k: z[k] for k in z.files}
    for k in ("W", "H", "pfx", "pfy", "max_lvl", "features"):
        c[k] = int(c[k])
    c["scene"] = str(c["scene"])
    return c


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="session")
def gpu(built):
    """One Renderer on cuda:0 for the whole session.  No GPU => the test FAILS (no CPU fallback exists)."""
    from raytracert_b200 import binding
    R = binding.Renderer(1)
    yield R
    R.shutdown()
