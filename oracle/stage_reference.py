"""Stages the UNMODIFIED reference hot-path sources for the GPU box (test infrastructure only).

/root/reference does not exist on the gpurun box and the reference's path is CUDA-only, so the only
place the real reference can execute is the B200.  This script copies the few files of the path
(byte for byte, no edits) into ``baseline/_ref/`` - git-ignored, not gpurun-ignored - together
with the already JIT-built torch extensions when present.  It is used for exactly two things:
generating the golden fixtures (tests/golden/make_golden_ref_gpu.py) and timing the reference's
own GPU path beside ours (bench_reference_gpu.py).  Nothing in the product imports it.
"""
import os
import shutil
import sys

SRC = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")

FILES = [
    "graphs/__init__.py",
    "graphs/stylegan_v2_real/__init__.py",
    "graphs/stylegan_v2_real/networks.py",
    "graphs/stylegan_v2_real/transform_base.py",
    "graphs/stylegan_v2_real/constants.py",
    "graphs/stylegan_v2_real/stylegan2.py",
    "graphs/stylegan_v2_real/graph_util.py",
    "graphs/stylegan_v2_real/op/__init__.py",
    "graphs/stylegan_v2_real/op/fused_act.py",
    "graphs/stylegan_v2_real/op/fused_bias_act.cpp",
    "graphs/stylegan_v2_real/op/fused_bias_act_kernel.cu",
    "graphs/stylegan_v2_real/op/upfirdn2d.py",
    "graphs/stylegan_v2_real/op/upfirdn2d.cpp",
    "graphs/stylegan_v2_real/op/upfirdn2d_kernel.cu",
]


def stage() -> bool:
    if not os.path.isdir(SRC):
        print("stage_reference: /root/reference not present, nothing staged")
        return False
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copy2(os.path.join(SRC, rel), dst)
    print(f"stage_reference: staged {len(FILES)} files into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
