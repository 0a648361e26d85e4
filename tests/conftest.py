import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """The native libraries, built in-tree (no-op when they are up to date)."""
    from raytracingoneweekendapplication_b200 import build

    build.build_cuda()
    build.build_host()
    build.build_oracle()
    return True


@pytest.fixture(scope="session")
def ctx(built):
    """One rt_ctx on cuda:0.  Fails loudly if the CUDA path is unavailable (no fallback)."""
    from raytracingoneweekendapplication_b200 import capi

    c = capi.Context(0)
    yield c
    c.close()


_scene_cache = {}


def reference_inputs_available() -> bool:
    """The `monkey` scene runs on the reference's OWN inputs (monkey.obj, Images/earthmap.jpg): staged under
    scenes/assets/reference/ by the build where /root/reference is mounted (and shipped to the GPU box with the other
    built artefacts), decoded by an stb-enabled libscenes_b200.so.  Elsewhere its tests are skipped."""
    from raytracingoneweekendapplication_b200 import capi

    d = os.path.join(ROOT, "scenes", "assets", "reference")
    if not (os.path.exists(os.path.join(d, "monkey.obj")) and os.path.exists(os.path.join(d, "earthmap.jpg"))):
        return False
    s = capi.load_scenes()
    return hasattr(s, "rtsc_has_stb") and s.rtsc_has_stb() == 1


@pytest.fixture(scope="session")
def scene_of(built):
    from raytracingoneweekendapplication_b200 import capi

    def get(name, seed=1):
        if name == "monkey" and not reference_inputs_available():
            pytest.skip("the reference's own inputs (monkey.obj, earthmap.jpg) are not staged on this box")
        key = (name, seed)
        if key not in _scene_cache:
            _scene_cache[key] = capi.Scene(name, seed)
        return _scene_cache[key]

    return get
