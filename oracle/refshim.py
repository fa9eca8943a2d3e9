"""Import the UNMODIFIED reference (AMEP 1.3.0) from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used solely by ``oracle/make_golden.py`` (run in the
build container, where /root/reference exists) to produce the golden vectors
committed under ``tests/golden/``, and by the optional ``-m "not gpu"`` test
that re-checks the oracle against the live reference when it is present.  It is
never imported by the product (``amep_b200``), by ``bench.py`` or by anything
that runs on the GPU box -- /root/reference does not exist there.

``import amep`` needs five packages that are absent from this image (h5py,
matplotlib, scikit-image, gsd, chemfiles).  None of them does arithmetic on the
RDF / S(q) path (storage, plotting, field clustering, file readers), so
permissive stub packages are written to a temp directory and put on the module
search path.  They are written to disk (not just ``sys.modules``) because the
reference forces the *spawn* start method (amep/utils.py:49-50) and its pool
workers re-import ``amep`` in a fresh interpreter.  See SURVEY.md section 8(c).

Callers must be guarded by ``if __name__ == "__main__":`` for the same reason.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("AMEP_REFERENCE_ROOT", "/root/reference")

_STUB_PACKAGES = [
    "h5py",
    "matplotlib", "matplotlib.pyplot", "matplotlib.ticker",
    "matplotlib.patches", "matplotlib.collections", "matplotlib.colors",
    "matplotlib.animation", "matplotlib.figure", "matplotlib.axes",
    "matplotlib.cm", "matplotlib.lines", "matplotlib.transforms",
    "mpl_toolkits", "mpl_toolkits.axes_grid1",
    "mpl_toolkits.axes_grid1.inset_locator",
    "skimage", "skimage.segmentation", "skimage.measure",
    "gsd", "gsd.hoomd",
    "chemfiles",
]

_STUB_SOURCE = '''
class _Anything:
    """Tolerates attribute access, calls, subclassing and ``X | Y`` unions."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __or__(self, other):
        return self

    __ror__ = __or__

    def __mro_entries__(self, bases):
        return (object,)

    def __iter__(self):
        return iter(())

    def __getitem__(self, key):
        return _Anything()


def __getattr__(name):
    if name.startswith("__") and name.endswith("__"):
        raise AttributeError(name)
    return _Anything()
'''


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "amep"))


def _write_disk_stubs(root, missing):
    for name in _STUB_PACKAGES:
        if name.split(".")[0] not in missing:
            continue
        d = os.path.join(root, *name.split("."))
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "__init__.py"), "w") as fh:
            fh.write(_STUB_SOURCE)


def load_reference():
    """Return the reference ``amep`` package."""
    mod = sys.modules.get("amep")
    if mod is not None and str(getattr(mod, "__file__", "")).startswith(REFERENCE_ROOT):
        return mod
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    missing = []
    for top in ("h5py", "matplotlib", "mpl_toolkits", "skimage", "gsd", "chemfiles"):
        try:
            __import__(top)
        except Exception:
            missing.append(top)
    extra = [REFERENCE_ROOT]
    if missing:
        stub_root = os.path.join(os.environ.get("TMPDIR", "/tmp"), "amep_ref_stubs")
        _write_disk_stubs(stub_root, missing)
        extra.append(stub_root)
    # children (spawned pool workers) inherit both sys.path and PYTHONPATH
    parts = [p for p in os.environ.get("PYTHONPATH", "").split(os.pathsep) if p]
    for e in extra:
        if e not in parts:
            parts.append(e)
        if e not in sys.path:
            sys.path.append(e)
    os.environ["PYTHONPATH"] = os.pathsep.join(parts)
    import amep  # noqa: E402  (the reference, unmodified)
    return amep
