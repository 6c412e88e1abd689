"""Place the UNMODIFIED reference modules of the hot path under baseline/_ref/ (git-ignored, travels to the GPU box).

    python baseline/fetch_reference.py            # needs /root/reference (build container only)

`bench.py --impl reference` imports models_twomodalinputs / models_singlemodalinput / utils from there and times the
reference's own CPU path.  The reference is a tree of plain Python scripts without packaging
metadata, so "installing" it is a byte-for-byte copy of the three packages the hot path imports
(train_files/trainchaos_proposed_30cases1labeled.py:20-23); nothing under baseline/_ref/ is tracked by git.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("AIDE_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
PACKAGES = ("models_twomodalinputs", "models_singlemodalinput", "utils")


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"{SRC} not present: nothing to fetch (the GPU box uses the copy made in the build container)")
        return 0
    os.makedirs(DST, exist_ok=True)
    digest = hashlib.sha256()
    for pkg in PACKAGES:
        dst = os.path.join(DST, pkg)
        shutil.rmtree(dst, ignore_errors=True)
        shutil.copytree(os.path.join(SRC, pkg), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        for root, _, files in sorted(os.walk(dst)):
            for f in sorted(files):
                with open(os.path.join(root, f), "rb") as fh:
                    digest.update(fh.read())
    for extra in ("LICENSE", "README.md"):
        if os.path.exists(os.path.join(SRC, extra)):
            shutil.copy(os.path.join(SRC, extra), os.path.join(DST, extra))
    with open(os.path.join(DST, "SOURCE.txt"), "w") as fh:
        fh.write(f"lich0031/AIDE (unmodified copy of {', '.join(PACKAGES)} from {SRC})\nsha256 {digest.hexdigest()}\n")
    print(f"reference packages copied to {DST} (sha256 {digest.hexdigest()[:16]})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
