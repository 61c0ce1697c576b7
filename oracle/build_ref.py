"""Recipe for ``oracle/_ref/``: a run-time copy of the reference's own Python package, so that bench.py's
``--impl reference`` arm and ``cpu_baseline`` time the REFERENCE'S classes (``kind: "reference"``) on the GPU box,
where ``/root/reference`` does not exist.

    python oracle/build_ref.py          (also called by __graft_entry__.build() when /root/reference is present)

What it does: copies the ``*.py`` files of ``/root/reference/src/csromer`` -- unmodified, byte for byte -- into
``oracle/_ref/csromer/`` and records their sha256 in ``oracle/_ref/MANIFEST.json``.  ``oracle/_ref/`` is listed in
``.gitignore`` (reference sources never enter this repository's history) but not in ``.gpurunignore``, so the copy
travels to the GPU box like the built ``.so``.  ``oracle/ref_shim.py`` imports the package from ``/root/reference/src``
when that exists and from ``oracle/_ref`` otherwise, through the same stub modules (astropy / pywt / pynufft /
prox_tv are absent) and the same one-line ``NDFT1DFixed`` subclass.  TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing
under ``csromer_b200/`` imports it.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("CSROMER_REFERENCE_SRC", "/root/reference/src")
DST = os.path.join(HERE, "_ref")


def build(verbose=False):
    """Returns True when oracle/_ref holds a copy of the reference package afterwards."""
    pkg = os.path.join(SRC, "csromer")
    if not os.path.isdir(pkg):
        return os.path.isdir(os.path.join(DST, "csromer"))
    manifest = {}
    tmp = DST + ".tmp.%d" % os.getpid()
    shutil.rmtree(tmp, ignore_errors=True)
    for root, _dirs, files in os.walk(pkg):
        for name in sorted(files):
            if not name.endswith(".py"):
                continue
            src = os.path.join(root, name)
            rel = os.path.relpath(src, SRC)
            dst = os.path.join(tmp, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(tmp, "MANIFEST.json"), "w") as f:
        json.dump({"source": pkg, "files": manifest}, f, indent=1, sort_keys=True)
    shutil.rmtree(DST, ignore_errors=True)
    os.replace(tmp, DST)
    if verbose:
        print("oracle/_ref: %d reference files copied from %s" % (len(manifest), pkg), file=sys.stderr)
    return True


if __name__ == "__main__":
    ok = build(verbose=True)
    sys.exit(0 if ok else 1)
