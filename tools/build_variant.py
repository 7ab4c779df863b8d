#!/usr/bin/env python3
"""Build a variant of libgato_b200.so with extra nvcc flags into gato_b200/lib/variants/ (A/B runs on the GPU box through GATO_B200_LIB).
usage: tools/build_variant.py <name> [-DFLAG=..] ..."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out = ROOT / "gato_b200" / "lib" / "variants" / f"libgato_b200_{name}.so"
out.parent.mkdir(parents=True, exist_ok=True)
g.build_library(out, extra_flags=flags)
print(out)
