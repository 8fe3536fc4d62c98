"""klongpy_b200 — a B200-native (sm_100a) array backend for KlongPy.

    import klongpy_b200                       # registers backend 'b200' when klongpy is importable
    klong = klongpy_b200.interpreter()        # KlongInterpreter(backend='b200') + verb overrides installed

The hot path (fused elementwise / reduce / scan kernels, radix-sort grade, group / find / range, grouped
aggregation) lives in libkgb200.so behind the C ABI of include/kgb200.h; this package is the host-side mirror
of KlongPy's backend interface.  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .array import DeviceArray  # noqa: F401
from .backend import B200BackendProvider, HAVE_KLONGPY, UnsupportedDtypeError  # noqa: F401

BACKEND_NAME = "b200"


def register():
    """Register the provider with klongpy's backend registry (klongpy/backends/registry.py:23-25)."""
    from klongpy.backends import register_backend
    register_backend(BACKEND_NAME, B200BackendProvider)


def interpreter(device=None):
    """A KlongInterpreter on the b200 backend with the leaking verbs replaced (see verbs.install)."""
    from klongpy import KlongInterpreter
    from . import verbs
    register()
    return verbs.install(KlongInterpreter(backend=BACKEND_NAME, device=device))


if HAVE_KLONGPY:
    try:
        register()
    except Exception:  # pragma: no cover
        pass
