"""ORACLE: geoopt.utils.size2shape (used at /root/reference/hyperbolic_vae/layers.py:177)."""
import torch


def size2shape(*size):
    if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
        size = size[0]
    return tuple(int(s) for s in size)
