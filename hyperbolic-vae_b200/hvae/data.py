"""On-device input pipeline pieces: the reference normalises the RNA-seq matrix on the host with pandas
/ scipy (hyperbolic_vae/datasets/jerby_arnon.py:97-106); here the raw batch is copied to HBM once and normalised there."""
from __future__ import annotations

import torch
from torch import Tensor

from . import _cabi as C
from .ops import _workspace


def normalize_rnaseq(x: Tensor, method: str, out: Tensor = None) -> Tensor:
    """Mirror of hyperbolic_vae.datasets.jerby_arnon.normalize_rnaseq on a CUDA (cells, genes) fp32 matrix:
    "sum_to_one" | "sum_to_million" (per cell) | "z_score" (per gene over the cells, scipy.stats.zscore)."""
    if x.dim() != 2:
        raise ValueError("normalize_rnaseq expects a (cells, genes) matrix")
    C.require_cuda(x)
    x = x.contiguous()
    out = torch.empty_like(x) if out is None else out
    R, G = x.shape
    if method in ("sum_to_one", "sum_to_million"):
        C.call("hvae_rows_sum_normalize_f32", C.ptr(x), C.ptr(out), R, G, 1.0 if method == "sum_to_one" else 1e6, C.stream())
    elif method == "z_score":
        ws = _workspace(C.lib().hvae_cols_zscore_workspace_bytes(G), x.device)
        C.call("hvae_cols_zscore_f32", C.ptr(x), C.ptr(out), None, None, R, G, C.ptr(ws), ws.numel(), C.stream())
    else:
        raise ValueError(f"rnaseq_normalize_method {method} not recognized")
    return out
