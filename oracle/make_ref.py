"""make_ref.py -- TEST INFRASTRUCTURE: put the UNMODIFIED reference package where the GPU box can run it.

    python oracle/make_ref.py            # /root/reference/detex -> oracle/_ref/detex

`/root/reference` only exists in the build container.  `bench.py --impl reference` (and the
`cpu_baseline` leg) must time the reference's own `_SSDetex._MPXDS` (detect.py:559-578) /
`construct._CCX2` (construct.py:425-466) on the GPU box's host cores, so the package has to
travel: `oracle/_ref/` is git-ignored (no reference source ever enters the history) but NOT
gpurun-ignored.

The reference is pure Python: there is nothing to compile.  The sanctioned offline install
    python -m pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>
is tried first; it fails while generating metadata (setup.py:21-22 passes the whole line
"__version__ = 1.0.9" as the version -> packaging.version.InvalidVersion), so the recipe then does
what that install would have done for a pure-Python distribution: it copies the `detex/` package
directory, byte for byte.  `oracle/ref_shim.py` imports it from there (Python-3 / pandas-3 aliases
only, no change to the arithmetic).
"""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("DETEX_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def _same_tree(a, b):
    if not os.path.isdir(b):
        return False
    c = filecmp.dircmp(a, b, ignore=["__pycache__"])
    if c.left_only or c.right_only or c.diff_files or c.funny_files:
        return False
    return all(_same_tree(os.path.join(a, d), os.path.join(b, d)) for d in c.common_dirs)


def make(verbose=False):
    """Returns the path of oracle/_ref if it holds the reference package, else None."""
    pkg = os.path.join(SRC, "detex")
    if not os.path.isdir(pkg):
        return DST if os.path.isdir(os.path.join(DST, "detex")) else None
    if _same_tree(pkg, os.path.join(DST, "detex")):
        return DST
    os.makedirs(DST, exist_ok=True)
    how = "pip"
    with tempfile.TemporaryDirectory() as td:
        cp = os.path.join(td, "reference")          # /root/reference is read-only: install from a copy
        shutil.copytree(SRC, cp, ignore=shutil.ignore_patterns(".git", "__pycache__"))
        r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation",
                            "--no-deps", "--target", DST, cp], capture_output=True, text=True)
        if r.returncode != 0 or not os.path.isdir(os.path.join(DST, "detex")):
            why = [l.strip() for l in r.stderr.splitlines() if "Invalid" in l or "Error" in l or "error:" in l]
            how = "copy (pip install failed: %s)" % (why[0] if why else "see oracle/make_ref.py")[:160]
            shutil.rmtree(os.path.join(DST, "detex"), ignore_errors=True)
            shutil.copytree(pkg, os.path.join(DST, "detex"), ignore=shutil.ignore_patterns("__pycache__"))
    assert _same_tree(pkg, os.path.join(DST, "detex")), "oracle/_ref/detex differs from the reference"
    with open(os.path.join(DST, "HOW"), "w") as f:
        f.write("unmodified copy of %s/detex, made by oracle/make_ref.py via %s\n" % (SRC, how))
    if verbose:
        print("oracle/_ref ready (%s)" % how)
    return DST


if __name__ == "__main__":
    p = make(verbose=True)
    print(p if p else "reference tree not present and oracle/_ref empty")
