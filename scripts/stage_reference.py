#!/usr/bin/env python
"""Stage the UNMODIFIED reference scripts for a drop-in run on the GPU box.

/root/reference does not exist on the GPU box and reference sources must never enter the repository's history, so the
byte-identical files are copied into baseline/_ref/ (git-ignored, but shipped by gpurun with the working tree):

    baseline/_ref/PointNetGPD/main_1v.py, main_1v_mc.py, main_fullv.py, main_fullv_mc.py, main_test.py
    baseline/_ref/data/pointnetgpd_3class.model          (the shipped checkpoint main_test.py loads)

and their SHA-256 digests are checked against scripts/reference_sha256.txt (committed), so the logs under profiles/
provably come from the unmodified files.  The reference's own model/ package is deliberately NOT copied: the launcher
resolves `model.*` to pointnetgpd_b200."""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
FILES = ["PointNetGPD/main_1v.py", "PointNetGPD/main_1v_mc.py", "PointNetGPD/main_fullv.py", "PointNetGPD/main_fullv_mc.py",
         "PointNetGPD/main_test.py", "data/pointnetgpd_3class.model"]
SHA = os.path.join(ROOT, "scripts", "reference_sha256.txt")


def digest(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    dst_root = os.path.join(ROOT, "baseline", "_ref")
    if "--verify" in sys.argv:            # on the GPU box: check the staged copies against the committed digests
        want = dict(line.split()[::-1] for line in open(SHA).read().splitlines() if line.strip())
        for rel in FILES:
            got = digest(os.path.join(dst_root, rel))
            assert got == want[rel], "staged %s differs from the reference (sha256 %s != %s)" % (rel, got, want[rel])
            print("ok  %s  %s" % (got[:16], rel))
        return 0
    lines = []
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        lines.append("%s  %s" % (digest(src), rel))
    with open(SHA, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))
    return 0


if __name__ == "__main__":
    sys.exit(main())
