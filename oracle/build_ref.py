"""Recipe for oracle/_ref/: a verbatim, git-ignored copy of the reference's Python package, so that the reference's
own Numba path can be timed on the GPU box (where /root/reference does not exist).

    python oracle/build_ref.py            # needs /root/reference; no-op (exit 0) when it is absent

TEST / BENCH INFRASTRUCTURE ONLY.  Nothing is copied into the git history: oracle/_ref/ is listed in .gitignore
(not in .gpurunignore, so it travels with the gpurun snapshot like the built .so files).  Only the modules the two
hot paths import are copied -- pypairs/{__init__,_version,pairs,utils,settings,log}.py, tools/, datasets/ (the
default marker JSON), and the two empty stub packages `pypairs/__init__` imports -- unmodified; the two imports
missing from this image (`anndata`, `colorama`) are served by the inert stand-ins in oracle/shims/.
oracle/ref_numba.py is the loader; bench.py --impl reference times it (cpu_baseline.kind = "reference").
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PP_REFERENCE_SRC", "/root/reference/pypairs")
DST = os.path.join(HERE, "_ref", "pypairs")
FILES = ["__init__.py", "_version.py", "pairs.py", "utils.py", "settings.py", "log.py",
         "tools/__init__.py", "tools/sandbag.py", "tools/cyclone.py",
         "datasets/__init__.py", "datasets/leng15_default_markers.json",
         "plotting/__init__.py", "preprocessing/__init__.py"]


def build(force=False):
    """Copy the reference package when its source tree is present; returns the path of oracle/_ref or None."""
    if not os.path.isdir(SRC):
        return os.path.dirname(DST) if os.path.isdir(DST) else None
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            continue
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            os.chmod(dst, 0o644)
    return os.path.dirname(DST)


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print("oracle/_ref: {}".format(out if out else "reference source tree not found, nothing copied"))
