"""Places the UNMODIFIED reference (its .py files) under baseline/_ref/ — git-ignored, but shipped to the GPU box with the
working tree — so that bench.py's reference arm and the drop-in tests can run the reference's own code where /root/reference does not
exist.  `pip install --target baseline/_ref /root/reference` is what the contract names, but the reference is a directory of
scripts without setup.py / pyproject.toml ("Directory '/root/reference' is not installable"), so the files are copied verbatim.
Nothing under baseline/_ref is tracked or edited.  python tools/install_reference.py"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def install(verbose=True):
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    os.makedirs(DST, exist_ok=True)
    n = 0
    for dirpath, dirnames, filenames in os.walk(SRC):
        dirnames[:] = [d for d in dirnames if not d.startswith(".") and d not in ("figures", "data", "__pycache__")]
        rel = os.path.relpath(dirpath, SRC)
        for f in filenames:
            if not f.endswith((".py", ".txt", ".md")):
                continue
            os.makedirs(os.path.join(DST, rel), exist_ok=True)
            shutil.copy2(os.path.join(dirpath, f), os.path.join(DST, rel, f))
            n += 1
    if verbose:
        print("baseline/_ref: %d reference files copied verbatim from %s" % (n, SRC))
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
