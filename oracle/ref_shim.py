"""TEST INFRASTRUCTURE — import the UNMODIFIED reference from /root/reference in THIS container.

`import fast_forward` from `/root/reference/src` fails three ways here:
  * `importlib.metadata.version("fast-forward-indexes")` (src/fast_forward/__init__.py:11) —
    the package is not pip-installed;
  * `import h5py` (src/fast_forward/index/disk.py:5) — no h5py / libhdf5 in the image;
  * `import nanopq` (src/fast_forward/quantizer/nanopq.py:3) — not installed.
`load_reference()` patches exactly those three things (version string "0.8.0", an empty
`h5py` stub so only `OnDiskIndex` is unusable, `oracle/nanopq_port.py` as `nanopq`) and
returns the genuine `fast_forward` module.  It is used ONLY by `oracle/gen_golden.py` to
produce the committed fixtures in `tests/golden/`, and by the optional cross-check tests
that skip when `/root/reference` is absent (it does not exist on the GPU box).

The reference package has the same import name as our drop-in package; never call this in
a process that has already imported ours.
"""

import importlib
import importlib.metadata
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "fast_forward"))


def load_reference():
    if not reference_available():
        raise RuntimeError("/root/reference is not mounted here")
    if "fast_forward" in sys.modules:
        mod = sys.modules["fast_forward"]
        if not getattr(mod, "__file__", "").startswith(REFERENCE_SRC):
            raise RuntimeError("another `fast_forward` package is already imported")
        return mod

    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import nanopq_port

    sys.modules.setdefault("nanopq", nanopq_port)
    if "h5py" not in sys.modules:
        try:
            import h5py  # noqa: F401
        except ImportError:
            stub = types.ModuleType("h5py")
            stub.File = None  # OnDiskIndex unusable; InMemoryIndex/Ranking unaffected
            sys.modules["h5py"] = stub

    real_version = importlib.metadata.version

    def _version(name):
        if name == "fast-forward-indexes":
            return "0.8.0"  # pyproject.toml:9
        return real_version(name)

    importlib.metadata.version = _version
    sys.path.insert(0, REFERENCE_SRC)
    try:
        return importlib.import_module("fast_forward")
    finally:
        importlib.metadata.version = real_version
