import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


GOLDEN = os.path.join(ROOT, "tests", "golden")
PAR1999 = os.path.join(ROOT, "desirna_b200", "params", "turner1999_37C.par")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle, build
    build()
    return Oracle(PAR1999)


@pytest.fixture(scope="session")
def engine():
    from desirna_b200 import engine as eng
    eng.init(0)
    eng.params_builtin(1999)
    return eng


def load_golden(tag):
    import json
    with open(os.path.join(GOLDEN, f"{tag}.jsonl")) as f:
        return [json.loads(line) for line in f]
