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


@pytest.fixture(scope="session")
def golden_riemannian():
    import torch

    return torch.load(os.path.join(ROOT, "tests", "golden", "riemannian_golden.pt"), weights_only=False)


def pytest_sessionfinish(session, exitstatus):
    """Write the strict-1e-5 audit of every parity comparison made in this session (tests/util_parity.py)."""
    try:
        import json

        import util_parity as U

        if not U.AUDIT:
            return
        keys = ("n", "fail32", "fail64", "unexplained", "n_well", "unexplained_well")
        tot = {k: sum(r[k] for r in U.AUDIT) for k in keys}
        tot.update(comparisons=len(U.AUDIT), worst_well=max(r["worst_well"] for r in U.AUDIT), worst_all=max(r["worst_all"] for r in U.AUDIT))
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_audit_%d.json" % os.getpid()), "w") as f:
            json.dump(dict(total=tot, strict_rtol=U.STRICT_RTOL, columns=["what"] + list(keys) + ["worst_well", "worst_all"],
                           all=[[r["what"]] + [r[k] for k in keys] + [round(r["worst_well"], 2), round(r["worst_all"], 2)] for r in U.AUDIT]), f)
        print("\n[parity audit] %(comparisons)d comparisons, %(n)d elements; plain 1e-5*scale: %(fail32)d fail vs the fp32 reference, "
              "%(fail64)d vs float64, %(unexplained)d vs both; well-conditioned: %(unexplained_well)d of %(n_well)d fail both "
              "(worst ratio %(worst_well).3g)" % tot)
    except Exception as ex:  # the audit never breaks a run
        print("[parity audit] not written: %s" % ex)
