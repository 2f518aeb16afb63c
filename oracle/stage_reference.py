"""Stage the reference's own Python sources for this path under `oracle/_ref/` (TEST INFRASTRUCTURE).

    python -m oracle.stage_reference            # needs /root/reference (the build container)

`oracle/_ref/` is git-ignored (reference sources never enter the history) but it is NOT
gpurun-ignored, so it travels to the GPU box, where `/root/reference` does not exist. There
`oracle/reference_shim.py` imports `lightgcn.py` / `utils_v2.py` / `train_lightgcn.py` UNMODIFIED
from this copy, so `bench.py --impl reference` times the reference's own training-step code on the
box's host cores (`cpu_baseline.kind = "reference"`) and the drop-in test drives the reference's
own `TrainLightGCN` loop. Only the third-party operator `torch_geometric.nn.conv.LGConv`, which is
not installable here, is the op-for-op restatement of `oracle/lgconv.py`.

`__graft_entry__.build()` runs this whenever `/root/reference` is present."""
from __future__ import annotations

import os
import shutil
import sys

REFERENCE_ROOT = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["src/lightgcn.py", "src/utils_v2.py", "src/train_lightgcn.py", "src/inference_lightgcn.py",
         "torchserve/lightgcn_handler.py"]


def stage(verbose: bool = False) -> bool:
    """Copy the path's source files; returns False when the reference is not on this machine."""
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, FILES[0])):
        return False
    for rel in FILES:
        src = os.path.join(REFERENCE_ROOT, rel)
        dst = os.path.join(DEST, rel)
        if not os.path.isfile(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src) \
                or os.path.getsize(dst) != os.path.getsize(src):
            shutil.copyfile(src, dst)
            if verbose:
                print(f"staged {rel}")
    return True


if __name__ == "__main__":
    ok = stage(verbose=True)
    print("oracle/_ref ready" if ok else "/root/reference is not present: nothing staged")
    sys.exit(0)
