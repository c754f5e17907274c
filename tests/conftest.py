import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

# Parity tolerance: the reference's own notion of "equal" (src/utils/op_utils.py:281) and the
# north_star bar (1e-3 relative, fp32).
RTOL, ATOL = 1e-3, 1e-5


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "gpu_next: needs a CUDA device AND has not passed on hardware yet (kernels written after the "
                                       "round's GPU budget was spent); run with -m gpu_next, promote to gpu once green")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords or "gpu_next" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"), map_location="cpu")
        return cache[name]
    return load


# Un-normalised feature tensors (values of either sign, crossing zero) are compared with an absolute floor
# that scales with the tensor: |err| <= 1e-3 |ref| + 1e-4 max|ref|. The tcgen05 engine accumulates in fp32
# with truncation, so its noise floor sits ~5x above the FFMA engine's (still ~2e-5 of the tensor scale).
FEATURE_ATOL_SCALE = 1e-4


def assert_close(actual, expected, what="", rtol=RTOL, atol=ATOL, atol_scale=None):
    actual = actual.detach().float().cpu()
    expected = expected.detach().float().cpu()
    if atol_scale is not None and expected.numel():
        atol = max(atol, atol_scale * expected.abs().max().item())
    assert actual.shape == expected.shape, f"{what}: shape {tuple(actual.shape)} vs {tuple(expected.shape)}"
    assert torch.isfinite(actual).all(), f"{what}: non-finite values"
    err = (actual - expected).abs()
    tol = atol + rtol * expected.abs()
    bad = err > tol
    if bad.any():
        i = torch.argmax(err - tol)
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} outside rtol={rtol} atol={atol}; worst: "
                             f"got {actual.flatten()[i].item():.6g} want {expected.flatten()[i].item():.6g} "
                             f"(max abs err {err.max().item():.3g})")


# Gradients: rtol 1e-3 plus an absolute floor of GRAD_ATOL_SCALE x max|reference gradient| of the same tensor (gradient
# entries are sums of many signed terms, so small entries carry the rounding noise of the large ones).
GRAD_ATOL_SCALE = 2e-4


def grad_floor(summaries: dict) -> float:
    """Absolute floor shared by all gradients of one case: 1e-6 x the largest gradient entry of the case. Some
    gradients are identically zero in exact arithmetic (e.g. the key bias of a softmax attention), so their fp32
    values are pure rounding noise of that magnitude."""
    mx = 0.0
    for w in summaries.values():
        mx = max(mx, float(w["full"].abs().max()) if "full" in w else w["absmax"])
    return 1e-6 * mx


def assert_grad_summary_close(got: torch.Tensor, want: dict, what: str, rtol=RTOL, atol_scale=GRAD_ATOL_SCALE, floor=0.0):
    """Compare a gradient with a ``cases.grad_summary`` fixture (whole tensor, or head + sum + norm)."""
    g = got.detach().double().cpu().reshape(-1)
    assert torch.isfinite(g).all(), f"{what}: non-finite gradient"
    if "full" in want:
        ref = want["full"].double()
        assert g.numel() == ref.numel(), f"{what}: {g.numel()} vs {ref.numel()} elements"
        assert_close(g, ref, what, rtol=rtol, atol=floor, atol_scale=atol_scale)
        return
    head = want["head"].double()
    floor = max(floor, atol_scale * want["absmax"])
    err = (g[:head.numel()] - head).abs()
    tol = floor + rtol * head.abs()
    assert (err <= tol).all(), f"{what}: head differs (max err {err.max().item():.3g}, absmax {want['absmax']:.3g})"
    assert abs(float(g.norm()) - want["norm"]) <= 2 * rtol * want["norm"] + floor * g.numel() ** 0.5, f"{what}: norm {float(g.norm()):.6g} vs {want['norm']:.6g}"
    # the plain sum cancels heavily; bound it by the norm-scaled floor
    n = g.numel()
    assert abs(float(g.sum()) - want["sum"]) <= rtol * abs(want["sum"]) + floor * n ** 0.5 * 4, \
        f"{what}: sum {float(g.sum()):.6g} vs {want['sum']:.6g}"
