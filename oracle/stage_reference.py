"""Stages the reference's Python sources for the O3 oracle run (SURVEY §8c "O3") — TEST INFRASTRUCTURE ONLY.

The GPU box has no /root/reference.  To execute the reference's OWN torch code (occupancy / occlusion targets,
PassOccVox, OccVFE, the spconv_backbone.py forwards) on the B200 against this library, its `.py` files are copied,
unmodified, from where they lie under /root/reference into `oracle/_ref/reference_py/` — a directory that is
git-ignored (never enters history) but not gpurun-ignored (travels to the GPU box with the snapshot, like
oracle/liboracle.so).  Nothing under btcdet_b200/ or spconv/ reads it; tests/golden/ref_loader.py is the only reader.

    python oracle/stage_reference.py          # in the build container (needs /root/reference)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("BTC_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref", "reference_py")
KEEP_EXT = (".py", ".yaml")
TOP = ("btcdet", "tools/cfgs")


def stage(verbose=False):
    """Copy `btcdet/**/*.py` and `tools/cfgs/**/*.yaml`; returns the number of files staged (0 without a checkout)."""
    if not os.path.isdir(os.path.join(SRC, "btcdet")):
        return 0
    n = 0
    for top in TOP:
        for dirpath, _, files in os.walk(os.path.join(SRC, top)):
            rel = os.path.relpath(dirpath, SRC)
            for f in files:
                if not f.endswith(KEEP_EXT):
                    continue
                os.makedirs(os.path.join(DST, rel), exist_ok=True)
                s, d = os.path.join(dirpath, f), os.path.join(DST, rel, f)
                if not os.path.exists(d) or os.path.getmtime(d) < os.path.getmtime(s):
                    shutil.copyfile(s, d)
                n += 1
    if verbose:
        print("staged %d reference files under %s" % (n, DST))
    return n


if __name__ == "__main__":
    sys.exit(0 if stage(verbose=True) else 1)
