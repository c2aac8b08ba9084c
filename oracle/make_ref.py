#!/usr/bin/env python
"""Materialise the UNMODIFIED reference into ``oracle/_ref/`` (TEST INFRASTRUCTURE).

    python oracle/make_ref.py            # build container only: needs /root/reference

``oracle/_ref/`` is git-ignored (reference sources never enter this repository's history) but NOT
gpurun-ignored, so the copy travels to the GPU box with the snapshot.  There it serves as
  * the CPU baseline of ``bench.py --impl reference`` (``cpu_baseline.kind == "reference"``): the reference's own
    ``networks.ANPDistractor`` + ``trainer.losses.LossFunc`` + ``torch.optim.Adam`` on the host cores,
  * the on-box "reference modules on B200 via cuDNN/cuBLAS" bar of the bench line,
  * the unmodified ``trainer/model_trainer.py`` that tests/test_gpu_dropin.py drives on the drop-in modules.
Only Python files of the packages the hot path touches are copied, byte for byte; a MANIFEST with their sha256
is written next to them so a test can tell that nothing was edited.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("B200NP_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
PACKAGES = ("networks", "trainer", "evaluator", "utils", "configs", "dataset")
SCRIPTS = ("train.py", "evaluation.py", "refinement.py")


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def main():
    if not os.path.isdir(os.path.join(SRC, "networks")):
        print(f"make_ref: {SRC} not present -- nothing to do (the GPU box uses the copy shipped with the snapshot)")
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    manifest = []
    for pkg in PACKAGES:
        for dp, _, files in os.walk(os.path.join(SRC, pkg)):
            for f in sorted(files):
                if not f.endswith(".py"):
                    continue
                s = os.path.join(dp, f)
                rel = os.path.relpath(s, SRC)
                d = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(s, d)
                manifest.append((rel, sha256(d)))
    for f in SCRIPTS:
        s = os.path.join(SRC, f)
        if os.path.isfile(s):
            shutil.copyfile(s, os.path.join(DST, f))
            manifest.append((f, sha256(s)))
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as f:
        for rel, h in sorted(manifest):
            f.write(f"{h}  {rel}\n")
    print(f"make_ref: {len(manifest)} files -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
