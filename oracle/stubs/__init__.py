"""ORACLE: import stubs for the glue packages the reference imports at module top but which are
absent here (pytorch_lightning 2.2.1, plotly, imageio, termcolor).  They carry NO arithmetic —
only enough surface for `import hyperbolic_vae.models.*` to succeed."""
import sys
import types

import torch


def _mod(name):
    m = types.ModuleType(name)
    m.__path__ = []  # behave like a package so `import a.b` works
    return m


class _AttrDict(dict):
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


class LightningModule(torch.nn.Module):
    """nn.Module plus the handful of Lightning methods the reference's models call."""

    def __init__(self, *a, **k):
        super().__init__()
        self._hparams = _AttrDict()
        self.logged = {}

    @property
    def hparams(self):
        return self._hparams

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def save_hyperparameters(self, *names, **kw):
        import inspect

        frame = inspect.currentframe().f_back
        loc = frame.f_locals
        if names:
            for n in names:
                if n in loc:
                    self._hparams[n] = loc[n]
        else:
            for n, v in loc.items():
                if n not in ("self", "__class__"):
                    self._hparams[n] = v

    def log(self, name, value, *a, **k):
        self.logged[name] = value

    def log_dict(self, d, *a, **k):
        self.logged.update(d)


class Callback:
    pass


class LightningDataModule:
    def __init__(self, *a, **k):
        pass


class Trainer:
    def __init__(self, *a, **k):
        self.callback_metrics = {}

    def fit(self, *a, **k):
        raise RuntimeError("pytorch_lightning stub: Trainer.fit is out of scope (SURVEY.md §2 #12)")


def seed_everything(seed, workers=False):
    torch.manual_seed(seed)
    return seed


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, n):
        return _Dummy()


def _lazy(n):
    if n.startswith("__"):
        raise AttributeError(n)
    return _Dummy


def install():
    """Put the stubs into sys.modules (idempotent; never overrides a real install)."""
    def have(name):
        try:
            __import__(name)
            return True
        except Exception:
            return False

    if not have("pytorch_lightning"):
        pl = _mod("pytorch_lightning")
        pl.LightningModule, pl.Callback, pl.Trainer = LightningModule, Callback, Trainer
        pl.LightningDataModule, pl.seed_everything = LightningDataModule, seed_everything
        cb = _mod("pytorch_lightning.callbacks")
        for n in ("EarlyStopping", "LearningRateMonitor", "ModelCheckpoint", "Callback"):
            setattr(cb, n, _Dummy)
        lg = _mod("pytorch_lightning.loggers")
        lg.TensorBoardLogger = _Dummy
        pl.callbacks, pl.loggers = cb, lg
        sys.modules.update({"pytorch_lightning": pl, "pytorch_lightning.callbacks": cb, "pytorch_lightning.loggers": lg})
    if not have("plotly"):
        p = _mod("plotly")
        for sub in ("express", "graph_objects", "io", "subplots"):
            m = _mod("plotly." + sub)
            m.__getattr__ = _lazy  # type: ignore
            setattr(p, sub, m)
            sys.modules["plotly." + sub] = m
        sys.modules["plotly"] = p
    if not have("imageio"):
        i = _mod("imageio")
        v3 = _mod("imageio.v3")
        v3.__getattr__ = _lazy  # type: ignore
        i.v3 = v3
        sys.modules.update({"imageio": i, "imageio.v3": v3})
    if not have("termcolor"):
        t = _mod("termcolor")
        t.colored = lambda s, *a, **k: s
        sys.modules["termcolor"] = t
