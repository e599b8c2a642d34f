"""pytest configuration: registers the `gpu` marker (tests that need a real B200) and puts the
repository root on sys.path.  `-m "not gpu"` runs here on CPU; `-m gpu` runs on the GPU box."""
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
    """Build every native artefact once per session (CUDA library, oracle, host-emulation test library)."""
    import __graft_entry__ as g
    g.build()
    import _hostemu
    _hostemu.build()
    return True


@pytest.fixture(scope="session")
def cuda_engine_factory(built):
    """Factory for product engines (CUDA library through the C ABI).  Fails loudly without a GPU."""
    from w2dynamics_b200.CTQMC import MCTQMC

    def make(problem, nchains=64, nbins=4, seed=43890, **opts):
        eng = MCTQMC()
        eng.set_option("nchains", nchains)
        eng.set_option("nbins", nbins)
        eng.set_option("device", 0)
        for k, v in opts.items():
            eng.set_option(k, v)
        eng.init_rnd(seed)
        eng.init_paras(problem.pstring())
        eng.init_solver(problem.u_matrix, problem.ftau, problem.muimp, None, 1)
        eng.set_taudiffmax(problem.qmc["TaudiffMax"])
        eng.init_counters()
        return eng
    return make
