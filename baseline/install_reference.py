"""Fill baseline/_ref/ (git-ignored, shipped to the GPU box) with the two reference source files of the hot path.

    python baseline/install_reference.py [/root/reference]

`pip install --target baseline/_ref /root/reference` is not possible: the reference is a flat directory of scripts
without setup.py / pyproject.toml.  Outcome is recorded in DESIGN.md."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ("encoding.py", "models.py")


def install(src_root: str = "/root/reference") -> bool:
    if not all(os.path.isfile(os.path.join(src_root, f)) for f in FILES):
        return False
    dst = os.path.join(HERE, "_ref")
    os.makedirs(dst, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(src_root, f), os.path.join(dst, f))
    return True


if __name__ == "__main__":
    ok = install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("baseline/_ref installed" if ok else "reference sources not found; baseline/_ref left as it is")
