#!/usr/bin/env python
"""Build an A/B variant of libtbk_b200.so with extra -D flags:  python profiles/build_variant.py NAME -DFOO=1 ...
writes profiles/ab/libtbk_NAME.so (git-ignored, travels with gpurun); run with PYTHTB_B200_LIB=profiles/ab/libtbk_NAME.so."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pythtb_b200 import build as B
name, flags = sys.argv[1], sys.argv[2:]
os.makedirs(os.path.join(ROOT, "profiles", "ab"), exist_ok=True)
out = os.path.join(ROOT, "profiles", "ab", "libtbk_%s.so" % name)
print(B.build(force=True, extra_flags=flags, out=out))
