import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device; run with -m gpu")


@pytest.fixture(scope="session")
def dev():
    import torch

    from whisper_nemo_b200 import _cabi

    _cabi.require_device()  # raises (does not skip) when the library / device is missing: no silent fallback
    return torch.device("cuda")


@pytest.fixture(scope="session")
def weights(dev):
    """The fixed-seed random-init TitaNet-L of the benchmark (embedding-BN statistics from the committed CPU-oracle
    fixture), shared by the product and the oracle."""
    from whisper_nemo_b200 import checkpoint

    return checkpoint.seeded()


@pytest.fixture(scope="session")
def oracle_model(weights):
    import torch

    from oracle.titanet import TitaNetL

    torch.set_num_threads(os.cpu_count() or 8)
    model = TitaNetL(compute_logits=False)
    res = model.load_state_dict(weights, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.startswith("preprocessor.") or k.startswith("decoder.final") for k in res.missing_keys), res.missing_keys
    return model.eval()
