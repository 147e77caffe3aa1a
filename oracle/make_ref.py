#!/usr/bin/env python
"""ORACLE (test infrastructure): snapshot recipe for the reference arm.

The reference is pure Python, so there is nothing to compile: this script copies the UNMODIFIED files of its
hot path (acr/*.py, mano/manolayer.py, configs/demo.yml) from /root/reference into ``oracle/_ref/``, which is
git-ignored (no reference source ever enters the history) but NOT gpurun-ignored, so the snapshot travels to the
GPU box with the repo exactly like a built ``.so``.  ``bench.py --impl reference`` and its ``cpu_baseline`` leg
then time the reference's OWN code on the host cores (``cpu_baseline.kind = "reference"``) through
oracle/ref_worker.py; without a snapshot they fall back to the oracle port (kind "port").

    python oracle/make_ref.py            # also run by __graft_entry__.build() when /root/reference exists
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ACR_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["acr/config.py", "acr/main.py", "acr/mano_wrapper.py", "acr/model.py", "acr/result_parser.py", "acr/utils.py",
         "acr/visualization.py", "mano/manolayer.py", "configs/demo.yml"]


def make(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print(f"make_ref: {REF} not present (GPU box): keeping the snapshot that travelled with the repo")
        return os.path.isdir(os.path.join(DST, "acr"))
    h = hashlib.sha1()
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(src, "rb") as f:
            h.update(f.read())
    with open(os.path.join(DST, "SNAPSHOT"), "w") as f:
        f.write(f"unmodified copy of {len(FILES)} files of {REF}; sha1 of their concatenation {h.hexdigest()}\n")
    if verbose:
        print(f"make_ref: {len(FILES)} reference files -> {DST} ({h.hexdigest()[:12]})")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
