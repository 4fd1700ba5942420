import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def model_dirs(tmp_path_factory):
    """Seeded random-init model directories (encoder-/decoder-/joiner- containers + tokens.txt)."""
    from sherpa_vietnamese_asr_b200 import weights
    base = tmp_path_factory.mktemp("models")
    made = {}

    def get(name, seed):
        key = (name, seed)
        if key not in made:
            cfg = weights.CONFIGS[name]()
            d = str(base / f"{name}-{seed}")
            paths = weights.write_model_dir(d, cfg, seed)
            made[key] = (cfg, paths, d)
        return made[key]
    return get
