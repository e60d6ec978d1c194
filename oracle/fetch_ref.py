"""TEST / BENCH INFRASTRUCTURE ONLY: make the UNMODIFIED Python reference available to ``bench.py --impl reference``
and to ``bench.py``'s ``cpu_baseline`` on the GPU box, where ``/root/reference`` does not exist.

    python oracle/fetch_ref.py            # run in the authoring container (called by __graft_entry__.build())

Installs the reference package with pip (offline, no dependencies) into ``baseline/_ref/`` -- the directory the
bench contract names for the reference install; it is git-ignored (nothing of the reference enters the history) but
NOT gpurun-ignored, so it travels with the snapshot like the built ``.so`` files.  ``/root/reference`` is read-only,
so the build runs on a throw-away copy under /tmp.  If pip cannot build it, the package directory is copied as is.
The reference's two non-arithmetic dependencies (``pyfile_utils``, ``matplotlib``) stay absent; ``oracle/ref_stub.py``
stands in for them at import time (``NDS_REFERENCE_ROOT`` points it at ``baseline/_ref``).
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("NDS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def fetch(verbose=True) -> bool:
    if not os.path.isdir(os.path.join(SRC, "ndsimulator")):
        if verbose:
            print(f"fetch_ref: {SRC} not present (GPU box?): keeping {DST} as it is")
        return os.path.isdir(os.path.join(DST, "ndsimulator"))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="nds_ref_build_")
    work = os.path.join(tmp, "reference")
    shutil.copytree(SRC, work, ignore=shutil.ignore_patterns(".git", "reference", "docs"))
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
           "--find-links", "/opt/wheelhouse", "--target", DST, work]
    rc = subprocess.call(cmd)
    how = "pip install --no-index --no-build-isolation --no-deps --target baseline/_ref"
    if rc != 0 or not os.path.isdir(os.path.join(DST, "ndsimulator")):
        shutil.copytree(os.path.join(work, "ndsimulator"), os.path.join(DST, "ndsimulator"), dirs_exist_ok=True)
        how = "plain copy of the package directory (pip build failed)"
    shutil.rmtree(tmp, ignore_errors=True)
    with open(os.path.join(DST, "HOW.txt"), "w") as f:
        f.write(f"installed from {SRC} by oracle/fetch_ref.py: {how}\n")
    if verbose:
        print(f"fetch_ref: {how}")
    return True


if __name__ == "__main__":
    sys.exit(0 if fetch() else 1)
