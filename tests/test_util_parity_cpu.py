"""CPU checks of the parity helper itself: the conditioned criterion and the strict plain-1e-5 audit."""
import pytest
import torch

import util_parity as U


def test_strict_audit_counts_and_asserts():
    torch.manual_seed(0)
    o64 = torch.randn(50, 8, dtype=torch.float64)
    o32 = o64.float()
    n0 = len(U.AUDIT)
    U.assert_parity(o32.clone(), o32, o64, what="exact")                      # identical to the fp32 reference
    assert U.AUDIT[-1]["fail32"] == 0 and U.AUDIT[-1]["n_well"] == 400
    # per-row rtol tensor: rows 0..24 well-conditioned (1e-5), the rest granted 1e-3
    rt = torch.full((50, 1), 1e-5, dtype=torch.float64)
    rt[25:] = 1e-3
    off = o32.clone()
    off[30] *= 1 + 2e-4                                                        # inside the conditioned bound, outside strict
    U.assert_parity(off, o32, o64, rtol=rt, what="ill rows")
    r = U.AUDIT[-1]
    assert r["unexplained"] > 0 and r["unexplained_well"] == 0 and r["n_well"] == 200
    off2 = o32.clone()
    off2[3] *= 1 + 5e-5
    old = U.STRICT_ASSERT
    try:
        U.STRICT_ASSERT = True
        with pytest.raises(AssertionError):
            U.assert_parity(off2, o32, o64, rtol=rt, atol=1e-6, slack_mult=1e4, what="well row off")  # only strict catches it
        U.STRICT_ASSERT = False
        U.assert_parity(off2, o32, o64, rtol=rt, atol=1e-6, slack_mult=1e4, what="well row off (report only)")
        assert U.AUDIT[-1]["unexplained_well"] > 0
    finally:
        U.STRICT_ASSERT = old
    # norm-relative (batch-summed gradients) and scalar outputs go through the audit too
    U.assert_parity(o32[:, 0], o32[:, 0], o64[:, 0], what="scalars")
    U.assert_parity(o32, o32, o64, norm_relative=True, what="norm")
    assert len(U.AUDIT) >= n0 + 5
    del U.AUDIT[n0:]
