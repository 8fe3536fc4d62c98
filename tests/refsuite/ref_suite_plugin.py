"""pytest plugin: run the REFERENCE's own test-suite against the b200 backend (CPU test double of the C ABI).

    cd /root/reference && PYTHONPATH=/root/repo:/root/repo/tests/refsuite PYTHONDONTWRITEBYTECODE=1 \
        python -m pytest -p no:cacheprovider -p ref_suite_plugin -q tests/test_suite.py tests/test_kgtests.py ...

Every `KlongInterpreter()` the reference tests construct becomes `KlongInterpreter(backend='b200', device=<double>)`
with the leaking verbs overridden (`klongpy_b200.verbs.install`) — the same thing `klongpy_b200.interpreter()` does.
This measures how complete the drop-in is at the HOST level (provider + np shim + verb overrides + lowering); the
kernels themselves are covered by the `-m gpu` parity tests.  Test infrastructure only.
"""
import pytest


@pytest.hookimpl(trylast=True)
def pytest_configure(config):
    import os
    import sys
    # appended (not prepended): the reference tests do `from conftest import ...` and must get their own conftest
    sys.path.append(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    dev = os.environ.get("REF_SUITE_DEVICE")  # e.g. cuda:0 on a GPU box: the REAL library instead of the CPU double
    if not dev:
        import fake_lib
        dev = fake_lib.install()
    import klongpy
    import klongpy_b200
    from klongpy_b200 import verbs
    klongpy_b200.register()
    inner = klongpy.KlongInterpreter.__init__

    def init(self, *args, backend=None, device=None, **kwargs):
        if backend in (None, "numpy") and not verbs.BUILDING_HOST_TWIN:
            backend, device = "b200", dev
        inner(self, *args, backend=backend, device=device, **kwargs)
        if backend == "b200":
            verbs.install(self)

    klongpy.KlongInterpreter.__init__ = init
