import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_package():
    """import the product package (directory name has a '-') as module `bwa_mem_gpu_b200`"""
    name = "bwa_mem_gpu_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(ROOT, "bwa-mem_gpu_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    mod = load_package()
    mod.build()
    return mod


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.build_oracle()
    return oracle_py


@pytest.fixture(scope="session")
def small_index(pkg, tmp_path_factory):
    """200 kb genome with repeats, index built by the product's host builder."""
    from tools import synth
    d = tmp_path_factory.mktemp("idx")
    g = synth.make_genome(200_000, repeats=True)
    prefix = str(d / "g")
    pkg.build_index(g, prefix, sa_intv=16, also_stock_layout=True, n_threads=4)
    return g, prefix
