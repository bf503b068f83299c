"""pytest configuration: the `gpu` marker and shared fixtures."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# ---- measured parity numbers: every GPU parity test that is not bit-exact records what it measured; the session writes
# them to gpurun_out/parity_r02.json (merged back by gpurun) and the committed copy lives in profiles/r02_parity.json
_PARITY = {}


def record_parity(key, value):
    _PARITY[key] = value


@pytest.fixture(scope="session")
def parity_log():
    return record_parity


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY:
        return
    import json

    out_dir = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        path = os.path.join(out_dir, "parity_r02.json")
        data = {}
        if os.path.exists(path):
            try:
                data = json.load(open(path))
            except Exception:
                data = {}
        data.update(_PARITY)
        with open(path, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    except OSError:
        pass


def parity_bound(key, field, default):
    """Asserted bar for a measured (not bit-exact) parity number: 2x what was last measured on a B200 and committed in
    tests/golden/parity_bounds.json (written by tools/update_parity_bounds.py from gpurun_out/parity_r02.json), or
    `default` when the case has not been measured yet.  Never below 1e-4 so that noise-floor flips do not flake."""
    import json

    path = os.path.join(GOLDEN, "parity_bounds.json")
    if os.path.exists(path):
        try:
            v = json.load(open(path)).get(key, {}).get(field)
            if v is not None:
                return max(2.0 * float(v), 1e-4)
        except Exception:
            pass
    return default


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


@pytest.fixture(scope="session")
def golden_rtn():
    return load_golden("rtn_pack.pt")


@pytest.fixture(scope="session")
def golden_gptq():
    return load_golden("gptq_layer.pt")


@pytest.fixture(scope="session")
def golden_awq():
    return load_golden("awq_module.pt")


@pytest.fixture(scope="session")
def golden_config1():
    return load_golden("config1_rtn_linear1024.pt")


@pytest.fixture(scope="session")
def golden_e2e():
    return load_golden("e2e_tiny_llama.pt")


@pytest.fixture(scope="session")
def golden_options():
    return load_golden("options_matrix.pt")


@pytest.fixture(scope="session")
def golden_options_extra():
    return load_golden("options_extra.pt")
