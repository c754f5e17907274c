"""N2: per-object preparation of the input pipeline (gather sampled points, gen_descriptor, zero_mean,
channels-first). CPU: the oracle against fixtures made by the reference's own functions (oracle/make_golden_prep.py);
GPU (-m gpu): the kernel, through the C ABI, against the fixtures and the oracle."""
import pytest
import torch

import cases
from oracle import vlsat_oracle as O


def _close(got, want, what, rtol=1e-5, atol_scale=5e-6):      # fp32 centring of metre-scale coordinates: a few ulp of the mean
    got, want = got.double().cpu(), want.double().cpu()
    assert got.shape == want.shape, what
    both_nan = torch.isnan(got) & torch.isnan(want)
    tol = rtol * want.abs() + atol_scale * want[~torch.isnan(want)].abs().max().clamp_min(1e-30)
    bad = ~both_nan & ~((got - want).abs() <= tol)
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.numel()} off, worst {(got - want).abs()[bad].max().item():.3g}"


@pytest.mark.parametrize("name", list(cases.PREP_CASES))
def test_oracle_object_prep_matches_reference_functions(name, golden):
    gold = golden("object_prep")[name]
    cloud, choice = cases.prep_inputs(name)
    pts, desc = O.prepare_objects(cloud, choice)
    _close(pts, gold["obj_points"], "obj_points")
    _close(desc, gold["descriptor"], "descriptor")
    assert torch.equal(pts[:, 3:], gold["obj_points"][:, 3:])             # non-xyz channels are a pure gather
    if name == "prep_one_point_pool":
        assert desc[0, 3:].abs().max().item() == 0.0                      # one repeated point: std, extent, volume, length 0


def test_object_prep_rejects_cpu_tensors():
    from vlsat_b200 import data_prep
    cloud, choice = cases.prep_inputs("prep_xyz")
    with pytest.raises(TypeError):
        data_prep.prepare_objects(cloud, choice)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(cases.PREP_CASES))
def test_kernel_object_prep_matches_reference_fixture(name, golden):
    from vlsat_b200 import data_prep
    gold = golden("object_prep")[name]
    cloud, choice = cases.prep_inputs(name)
    pts, desc = data_prep.prepare_objects(cloud.cuda(), choice.cuda())
    _close(pts, gold["obj_points"], "obj_points")
    _close(desc, gold["descriptor"], "descriptor")
    assert torch.equal(pts[:, 3:].cpu(), gold["obj_points"][:, 3:])       # bit-exact gather bookkeeping
    again = data_prep.prepare_objects(cloud.cuda(), choice.cuda())
    assert torch.equal(again[0], pts) and torch.equal(again[1], desc)      # fixed-order reductions: reproducible


@pytest.mark.gpu
def test_kernel_object_prep_config2_shape_and_strided_cloud():
    from vlsat_b200 import data_prep
    g = torch.Generator().manual_seed(5)
    wide = torch.randn(200_000, 12, generator=g)
    wide[:, :3] = wide[:, :3] * 0.5 + torch.tensor([40.0, -25.0, 3.0])      # far from the origin: the two-pass variance matters
    cloud = wide[:, :9]                                                    # row stride 12
    choice = torch.randint(0, 200_000, (640, 256), generator=g)
    choice[0, :5] = torch.tensor([-3, 200_000, 10**12, 0, 199_999])        # out of range: clamped, never out of bounds
    pts, desc = data_prep.prepare_objects(cloud.cuda(), choice.cuda())
    want_pts, want_desc = O.prepare_objects(cloud.double(), choice.clamp(0, 199_999))
    _close(pts, want_pts, "obj_points", rtol=1e-5, atol_scale=1e-5)       # ulp(40 m) = 3.8e-6: the fp32 mean is good to ~1e-5
    _close(desc, want_desc, "descriptor", rtol=1e-4, atol_scale=1e-6)
    p1, d1 = data_prep.prepare_objects(cloud.cuda(), choice[:3, :1].cuda())   # P = 1: torch.std of one sample is NaN
    assert torch.isnan(d1[:, 3:6]).all() and torch.equal(d1[:, 6:].cpu(), torch.zeros(3, 5)) and p1[:, :3].abs().max().item() == 0.0
    e0, e1 = data_prep.prepare_objects(cloud.cuda(), choice[:0].cuda())
    assert e0.shape == (0, 9, 256) and e1.shape == (0, 11)


@pytest.mark.gpu
def test_prepare_scene_hook_matches_the_reference_loop():
    """The loader hook end to end: host sampling with the reference's RNG stream + the device kernel against the reference's
    per-object numpy / torch loop (dataset_3dssg.py:279-293, op_utils.py:47-64) restated by the oracle."""
    import numpy as np
    from oracle import vlsat_oracle as O
    from vlsat_b200 import data_prep
    rs = np.random.RandomState(5)
    m = 20000
    points = (rs.randn(m, 3) * 0.4 + rs.randn(1, 3) * 3).astype(np.float32)
    instances = rs.randint(1, 12, size=m)
    nodes = [2, 9, 4, 11, 6]
    np.random.seed(21)
    idx = data_prep.sample_object_indices(instances, nodes, 128)
    np.random.seed(21)
    obj_points, desc = data_prep.prepare_scene(points, instances, nodes, 128)
    want_pts, want_desc = O.prepare_objects(torch.from_numpy(points), torch.from_numpy(idx))
    assert torch.allclose(obj_points.cpu(), want_pts, rtol=1e-4, atol=1e-5)
    assert torch.allclose(desc.cpu(), want_desc, rtol=1e-4, atol=1e-5)
