"""TEST INFRASTRUCTURE: object-preparation fixtures from the reference's own functions -> tests/golden/object_prep.pt.

Run in the build container only (needs /root/reference):  python oracle/make_golden_prep.py
Per case (tests/golden/cases.py PREP_CASES) and object: ``op_utils.gen_descriptor`` (src/utils/op_utils.py:47-64) and
``zero_mean`` (utils/util_data.py:53-59, the same body as dataset_3dssg.py:189-195) of the UNMODIFIED reference on the
gathered points, in float64 like the loader (numpy clouds are float64), then the trainer's permute (model.py:71)."""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import ref_shims  # noqa: E402
import cases  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "object_prep.pt")


def reference_prepare(cloud, choice):
    from src.utils import op_utils
    from utils.util_data import zero_mean
    n, p = choice.shape
    obj_points = torch.zeros([n, p, cloud.shape[1]])                         # float32, as dataset_3dssg.py:274
    descriptor = torch.zeros([n, 11])
    c64 = cloud.double().numpy()
    for i in range(n):
        obj_pointset = c64[choice[i].numpy(), :]
        descriptor[i] = op_utils.gen_descriptor(torch.from_numpy(obj_pointset)[:, :3])
        obj_pointset = torch.from_numpy(obj_pointset.astype("float32"))
        obj_pointset[:, :3] = zero_mean(obj_pointset[:, :3])
        obj_points[i] = obj_pointset
    return obj_points.permute(0, 2, 1).contiguous(), descriptor


def main():
    ref_shims.install()
    out = {}
    for name in cases.PREP_CASES:
        cloud, choice = cases.prep_inputs(name)
        pts, desc = reference_prepare(cloud, choice)
        out[name] = dict(obj_points=pts, descriptor=desc)
        print(name, tuple(pts.shape), tuple(desc.shape))
    torch.save(out, OUT)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
