"""Run the REFERENCE's own, unmodified test-suite against this drop-in package.

    python tools/run_reference_suite.py [--native-writer] /path/to/fast-forward-indexes/tests [pytest args]

The tests are copied to a temporary directory outside this repository (nothing of the reference
is kept here), a conftest puts `fast-forward-indexes_b200/` in front of `sys.path` — so that
`import fast_forward` resolves to this package — and installs tests/fake_h5py.py as `h5py` when
the real one is missing; with --native-writer no stand-in is installed, so that `OnDiskIndex`
writes its files with the package's own HDF5 writer (what happens on a box without h5py).
tests/test_encoder.py is left out: it downloads checkpoints.
Needs a CUDA device (the package has no CPU scoring path).
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFTEST = f'''import sys
sys.path.insert(0, {os.path.join(ROOT, "fast-forward-indexes_b200")!r})
sys.path.insert(0, {os.path.join(ROOT, "tests")!r})
'''
H5PY_STUB = '''try:
    import h5py  # noqa: F401
except ImportError:
    import fake_h5py
    sys.modules["h5py"] = fake_h5py
'''


def main() -> int:
    native = "--native-writer" in sys.argv
    if native:
        sys.argv.remove("--native-writer")
    if len(sys.argv) < 2 or not os.path.isdir(sys.argv[1]):
        print(__doc__)
        return 2
    with tempfile.TemporaryDirectory() as tmp:
        shutil.copytree(sys.argv[1], os.path.join(tmp, "tests"),
                        ignore=shutil.ignore_patterns("test_encoder.py", "_constants.py", "__pycache__"))
        with open(os.path.join(tmp, "conftest.py"), "w") as fh:
            fh.write(CONFTEST + ("" if native else H5PY_STUB))
        return subprocess.call([sys.executable, "-m", "pytest", "tests", "-q", "-p", "no:cacheprovider", *sys.argv[2:]],
                               cwd=tmp)


if __name__ == "__main__":
    sys.exit(main())
