#!/usr/bin/env python
"""Developer tool: build a VARIANT of the library next to the product build, for A/B runs inside one gpurun call.

    python tools/build_variant.py nostore -DPDSC_EXP_SOMETHING=1       ->  tools/bin/lib_nostore.so
    POINTDSC_B200_LIB=$PWD/tools/bin/lib_nostore.so python tools/stage_profile.py 3dmatch 1000 256
(timing experiments guarded by a macro are added to the sources for the duration of the experiment only; see profiles/README.md)

The product build (__graft_entry__.build) is untouched; tools/bin/ is git-ignored but travels to the GPU box."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "tools", "bin", f"lib_{name}.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + G.NVCC_FLAGS + extra + ["-o", out] + [os.path.join(G.CSRC, s) for s in G.SOURCES]
subprocess.run(cmd, check=True, cwd=G.CSRC)
print(out)
