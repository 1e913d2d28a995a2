import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    # GPU tests never silently pass on a box without a GPU: they are skipped unless selected with -m gpu,
    # and when selected they fail loudly if CUDA or the extension is missing (see gpu fixture).
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_cuda = False
    selected = config.getoption("-m") or ""
    if has_cuda or "gpu" in selected.replace("not gpu", ""):
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from tests.oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(ROOT, "tests", "golden", name))
    return load


@pytest.fixture(scope="session")
def fa():
    """The product package with the CUDA library loaded -- raises if it is missing (no fallback)."""
    import torch
    import flashattention_kernel_project_b200 as pkg
    pkg.load_library()
    if not torch.cuda.is_available():
        raise RuntimeError("GPU test selected but no CUDA device is visible")
    return pkg


@pytest.hookimpl(hookwrapper=True)
def pytest_runtest_makereport(item, call):
    """A failing GPU test also reports the library's hang-diagnostic record (which bounded mbarrier wait expired)."""
    outcome = yield
    rep = outcome.get_result()
    if rep.when == "call" and rep.failed and "gpu" in item.keywords:
        try:
            import flashattention_kernel_project_b200 as pkg
            rep.sections.append(("fa_last_debug", repr([hex(w) for w in pkg.last_debug()])))
        except Exception as e:  # noqa: BLE001
            rep.sections.append(("fa_last_debug", f"unavailable: {e}"))
