"""ORACLE: minimal ManifoldTensor / ManifoldParameter (geoopt/tensor.py surface used at
/root/reference/hyperbolic_vae/layers.py:53,184 and manifolds.py:5)."""
import torch


class ManifoldTensor(torch.Tensor):
    def __new__(cls, *args, manifold=None, requires_grad=False, **kwargs):
        if len(args) == 1 and isinstance(args[0], torch.Tensor):
            data = args[0].data
        else:
            data = torch.Tensor(*args, **kwargs)
        if kwargs.get("device") is not None:
            data = data.to(kwargs["device"])
        instance = torch.Tensor._make_subclass(cls, data, requires_grad)
        instance.manifold = manifold
        return instance

    def proj_(self):
        return self.copy_(self.manifold.projx(self))


class ManifoldParameter(ManifoldTensor, torch.nn.Parameter):
    def __new__(cls, data=None, manifold=None, requires_grad=True):
        if data is None:
            data = torch.empty(0)
        instance = ManifoldTensor.__new__(cls, data, manifold=manifold, requires_grad=requires_grad)
        return instance

    def __repr__(self):
        return "ManifoldParameter on {}:\n".format(self.manifold) + torch.Tensor.__repr__(self)
