import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "hyperbolic-vae_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "reference: needs /root/reference (authoring container only)")


def pytest_collection_modifyitems(config, items):
    import torch

    has_gpu = torch.cuda.is_available()
    has_ref = os.path.isfile("/root/reference/hyperbolic_vae/layers.py")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not has_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present"))


@pytest.fixture(scope="session")
def golden_ops():
    import torch

    return torch.load(os.path.join(ROOT, "tests", "golden", "ops_golden.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_models():
    import torch

    return torch.load(os.path.join(ROOT, "tests", "golden", "models_golden.pt"), weights_only=False)


def pytest_sessionfinish(session, exitstatus):
    """Write the strict-1e-5 audit of every parity comparison made in this session (tests/util_parity.py)."""
    try:
        import json

        import util_parity as U

        if not U.AUDIT:
            return
        tot = dict(comparisons=len(U.AUDIT), elements=sum(r["n"] for r in U.AUDIT),
                   strict_fail=sum(r["strict_fail"] for r in U.AUDIT), well_elements=sum(r["n_well"] for r in U.AUDIT),
                   strict_fail_well=sum(r["strict_fail_well"] for r in U.AUDIT),
                   worst_well=max(r["worst_well"] for r in U.AUDIT), worst_all=max(r["worst_all"] for r in U.AUDIT))
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        worst = sorted(U.AUDIT, key=lambda r: -r["worst_all"])[:40]
        with open(os.path.join(out, "parity_audit_%d.json" % os.getpid()), "w") as f:
            json.dump(dict(total=tot, worst=worst, strict_rtol=U.STRICT_RTOL,
                           all=[[r["what"], r["n"], r["strict_fail"], r["n_well"], r["strict_fail_well"], round(r["worst_well"], 2),
                                 round(r["worst_all"], 2)] for r in U.AUDIT]), f)
        print("\n[parity audit] %(comparisons)d comparisons, %(elements)d elements: %(strict_fail)d fail plain 1e-5*scale "
              "(%(strict_fail_well)d of %(well_elements)d well-conditioned); worst ratio well %(worst_well).3g / all %(worst_all).3g" % tot)
    except Exception as ex:  # the audit never breaks a run
        print("[parity audit] not written: %s" % ex)


@pytest.fixture(scope="session")
def golden_riemannian():
    import torch

    return torch.load(os.path.join(ROOT, "tests", "golden", "riemannian_golden.pt"), weights_only=False)
