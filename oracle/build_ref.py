"""Stages the UNMODIFIED reference files the reference arm needs under oracle/_ref/ (TEST INFRASTRUCTURE; build artefact).

`/root/reference` exists only in the builder container; the GPU box gets a snapshot of the repo.  oracle/_ref/ is git-ignored
(never committed: no reference source enters the history) but NOT gpurun-ignored, so -- like the built .so -- it travels with
the snapshot, and `bench.py --impl reference` / the `gpu_eager_reference` leg can run the reference's own modules there
(kind "reference").  Without it they fall back to the oracle port (kind "port").  Called by __graft_entry__.build().
The files are copied byte for byte, relative paths preserved, so oracle/ref_import.py works on either root."""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = "/root/reference"
WANT = [
    "binary_seg/lib",                       # pranet.py, PraNet_Res2Net.py, Res2Net_v1b.py, pvtv2.py, nn/ (imported by pranet.py:7)
    "binary_seg/MyTrain_med.py",            # structure_loss (:19-38), the training loop the arm restates (:59-86)
    "binary_seg/MyTest_med.py",
    "binary_seg/utils/utils.py",            # clip_gradient (:7-17)
    "multiclass_seg/EMCAD/lib",
    "multiclass_seg/EMCAD/trainer.py",
    "multiclass_seg/EMCAD/utils/utils.py",
    "multiclass_seg/MERIT/lib/decoders.py",
    "multiclass_seg/MIST/lib/MIST.py",
]


def stage(verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(SRC, "binary_seg", "lib")):
        return os.path.isdir(os.path.join(DEST, "binary_seg", "lib"))
    n = 0
    for rel in WANT:
        s, d = os.path.join(SRC, rel), os.path.join(DEST, rel)
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.pth"))
            n += sum(len(f) for _, _, f in os.walk(d))
        elif os.path.isfile(s):
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copy2(s, d)
            n += 1
    with open(os.path.join(DEST, "README"), "w") as f:
        f.write("Byte-for-byte copies of reference files staged by oracle/build_ref.py for the reference arm of bench.py.\n"
                "Build artefact: git-ignored, never committed.\n")
    if verbose:
        print(f"[build] staged {n} reference files under oracle/_ref/")
    return True


if __name__ == "__main__":
    stage()
