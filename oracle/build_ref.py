"""Compile the reference's own hot-path modules into oracle/_ref/ (TEST / BASELINE INFRASTRUCTURE ONLY).

    python oracle/build_ref.py

The reference (punkduckable/PDE-READ, read-only at /root/reference in the build container) is pure Python, so its
"build" is byte-compilation: this recipe compiles the modules of the path where they lie -- Network.py,
PDE_Residual.py, Loss_Functions.py, Extraction.py, Points.py under /root/reference/Code -- and writes ONLY the
compiled .pyc files into oracle/_ref/Code/ (git-ignored, not gpurun-ignored: like a built .so they travel to the
GPU box, where /root/reference does not exist).  No reference source is copied into the repository.

``bench.py --impl reference`` imports these modules (sourceless import) to time the UNMODIFIED reference on the
box's host cores (cpu_baseline.kind = "reference"); without them it falls back to the oracle port.  Nothing under
pde_read_b200/ may import them.
"""
import os
import py_compile
import sys

REF_CODE = "/root/reference/Code"
MODULES = ("Network", "PDE_Residual", "Loss_Functions", "Extraction", "Points")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "Code")


def build(verbose: bool = False) -> bool:
    """Returns True when oracle/_ref/Code holds all modules (freshly built, or built earlier)."""
    if not os.path.isdir(REF_CODE):
        return available()
    os.makedirs(OUT, exist_ok=True)
    for m in MODULES:
        src, dst = os.path.join(REF_CODE, m + ".py"), os.path.join(OUT, m + ".pyc")
        py_compile.compile(src, cfile=dst, dfile="<reference>/Code/%s.py" % m, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        if verbose:
            print("compiled", src, "->", dst)
    # the same compiled modules once more as one archive (zipimport reads .pyc members): a snapshot tool that
    # drops *.pyc files then still carries the build
    import zipfile
    with zipfile.ZipFile(os.path.join(OUT, "reference_code.zip"), "w") as zf:
        for m in MODULES:
            zf.write(os.path.join(OUT, m + ".pyc"), m + ".pyc")
    with open(os.path.join(OUT, "PYTHON_VERSION"), "w") as fh:
        fh.write("%d.%d\n" % sys.version_info[:2])
    return True


def available() -> bool:
    try:
        with open(os.path.join(OUT, "PYTHON_VERSION")) as fh:
            ok = fh.read().strip() == "%d.%d" % sys.version_info[:2]
    except OSError:
        return False
    return ok and (all(os.path.exists(os.path.join(OUT, m + ".pyc")) for m in MODULES)
                   or os.path.exists(os.path.join(OUT, "reference_code.zip")))


def import_reference():
    """(Network, PDE_Residual, Loss_Functions, Extraction, Points) of the unmodified reference, or None."""
    if not available():
        return None
    loose = all(os.path.exists(os.path.join(OUT, m + ".pyc")) for m in MODULES)
    entry = OUT if loose else os.path.join(OUT, "reference_code.zip")
    if entry not in sys.path:
        sys.path.insert(0, entry)
    import importlib
    return tuple(importlib.import_module(m) for m in MODULES)


if __name__ == "__main__":
    print("oracle/_ref ready:", build(verbose=True))
