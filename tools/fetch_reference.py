"""Copy the UNMODIFIED reference scripts (pure Python, no build step: there is no setup.py to pip-install) from
/root/reference into git-ignored baseline/_ref/ so that they travel to the GPU box with the gpurun snapshot.
tests/test_gpu_dropin.py imports them from there; nothing in the product imports or copies reference code.
Run automatically by __graft_entry__.build() when /root/reference exists."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def fetch(verbose=True):
    if not os.path.isdir(SRC):
        return False
    os.makedirs(DST, exist_ok=True)
    for f in sorted(os.listdir(SRC)):
        if f.endswith(".py"):
            shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
    os.makedirs(os.path.join(DST, "configs"), exist_ok=True)
    for f in sorted(os.listdir(os.path.join(SRC, "configs"))):
        shutil.copy2(os.path.join(SRC, "configs", f), os.path.join(DST, "configs", f))
    if verbose:
        print("[fetch_reference] copied", sorted(os.listdir(DST)))
    return True


if __name__ == "__main__":
    sys.exit(0 if fetch() else 1)
