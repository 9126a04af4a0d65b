"""Recipe that makes the UNMODIFIED reference importable on the GPU box — TEST / BENCH INFRASTRUCTURE ONLY.

    python oracle/vendor_reference.py            # build container only: /root/reference is mounted here, not on the GPU box

The reference is 26 loose .py files without setup.py / pyproject.toml, so `pip install --target` has nothing to install; this
recipe is the equivalent: it mirrors the two importable packages of the hot path (`src/`, `diffusion/`) byte for byte from where
they lie under /root/reference into `oracle/_ref/`, which is git-ignored (no reference source enters the history) but NOT
gpurun-ignored (it travels to the GPU box like the built .so).  Nothing under mapdit_b200/ imports it.  Users:
`tools/ref_on_b200.py` / `bench.py`'s informational `ref_on_b200` leg, which time the reference's own PyTorch path
(train.py:46,86-96,222-223; sample.py:25,52-61) on the B200 next to the hand-written kernels.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
PACKAGES = ("src", "diffusion")


def vendor(ref_root="/root/reference", verbose=True):
    if not os.path.isdir(ref_root):
        return False
    n = 0
    for pkg in PACKAGES:
        for root, _, files in os.walk(os.path.join(ref_root, pkg)):
            rel = os.path.relpath(root, ref_root)
            for f in files:
                if not f.endswith(".py"):
                    continue
                os.makedirs(os.path.join(DST, rel), exist_ok=True)
                src, dst = os.path.join(root, f), os.path.join(DST, rel, f)
                if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
                    shutil.copyfile(src, dst)
                n += 1
    if verbose:
        print(f"[vendor_reference] {n} files of {ref_root}/{{{','.join(PACKAGES)}}} mirrored to {DST}")
    return True


def available():
    return os.path.isfile(os.path.join(DST, "src", "models.py")) and os.path.isfile(os.path.join(DST, "diffusion", "__init__.py"))


if __name__ == "__main__":
    sys.exit(0 if vendor(os.environ.get("MAPDIT_REFERENCE", "/root/reference")) else 1)
