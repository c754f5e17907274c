"""TEST / BENCH INFRASTRUCTURE (not product code): stage the UNMODIFIED reference files of the hot path under the
git-ignored ``baseline/_ref/`` so that they travel to the GPU box with the repo snapshot (SURVEY.md 8c "GPU-box caveat").

Nothing is edited: the files are byte-for-byte copies, re-made whenever ``/root/reference`` is present (``build()`` calls
this). ``bench.py --impl reference`` / ``--impl reference-gpu`` and the reference legs of the default bench import them
through ``oracle/ref_shims.py`` (``VLSAT_REFERENCE_ROOT=baseline/_ref``); when the directory is absent those legs fall back
to the oracle port and say so (``kind: "port"``). No product module imports anything from here.
"""
from __future__ import annotations

import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")

FILES = [
    "src/model/model_utils/network_PointNet.py", "src/model/model_utils/network_MMG.py", "src/model/model_utils/network_GNN.py",
    "src/model/model_utils/network_util.py", "src/model/model_utils/networks_base.py", "src/model/model_utils/model_base.py",
    "src/model/transformer/attention.py", "src/model/SGFN_MMG/model.py", "src/model/SGFN_MMG/baseline_sgfn.py",
    "src/utils/op_utils.py", "src/utils/config.py", "src/utils/eva_utils_acc.py",
    "clip_adapter/model.py", "clip_adapter/checkpoint/origin_mean.pth", "config/mmgnet.json",
]


def staged() -> bool:
    return os.path.isfile(os.path.join(DST, "src", "model", "SGFN_MMG", "model.py"))


def stage() -> bool:
    """Copy the files listed above from /root/reference. Returns True when the staged tree is complete."""
    if not os.path.isdir(SRC):
        return staged()
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.isfile(dst) or os.path.getsize(dst) != os.path.getsize(src) or os.path.getmtime(dst) < os.path.getmtime(src):
            shutil.copy2(src, dst)
    # package markers exactly where the reference has them (the others are namespace packages there too)
    for d in ("src", "src/model", "src/model/model_utils", "src/model/transformer", "src/model/SGFN_MMG", "clip_adapter"):
        init = os.path.join(SRC, d, "__init__.py")
        if os.path.isfile(init):
            shutil.copy2(init, os.path.join(DST, d, "__init__.py"))
    return staged()


if __name__ == "__main__":
    print("staged" if stage() else "reference not available", DST)
