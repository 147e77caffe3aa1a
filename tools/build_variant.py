#!/usr/bin/env python
"""Build an A/B variant of the library next to the product one:
    python tools/build_variant.py single      -> lib/libacr_b200_single.so  (conv_tc.cu with -DACR_DUAL_ISSUER=0)
Select it at run time with ACR_B200_LIB=<path>.  Variants are measurement tools, never loaded by default."""
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "arbitrary-hands-3d-reconstruction_b200"))
from acr_b200 import build as B  # noqa: E402

VARIANTS = {"single": {"conv_tc.cu": ["-DACR_DUAL_ISSUER=0"]}}
name = sys.argv[1]
B.build()
objs = []
for f in sorted(x for x in os.listdir(B.CSRC) if x.endswith((".cu", ".cpp"))):
    obj = os.path.join(B.OBJDIR, os.path.splitext(f)[0] + ".o")
    extra = VARIANTS[name].get(f)
    if extra:
        obj = os.path.join(B.OBJDIR, os.path.splitext(f)[0] + f"_{name}.o")
        subprocess.run([B.NVCC] + B.FLAGS + extra + ["-c", os.path.join(B.CSRC, f), "-o", obj], check=True)
    objs.append(obj)
out = os.path.join(B.LIBDIR, f"libacr_b200_{name}.so")
subprocess.run([B.NVCC, "-shared", "-o", out] + objs + ["-cudart", "static", "-Xcompiler", "-fPIC"], check=True)
print(out)
