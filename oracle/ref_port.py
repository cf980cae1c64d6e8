"""ORACLE (test infrastructure; never imported by the product): CPU restatement of the
reference-OWNED hot path, built on geoopt_min / pvae_min.  It exists because /root/reference does
not travel to the GPU box: tests, smoke() and bench.py's cpu_baseline / --impl reference legs use
this file there.  It is pinned against the reference's real files (executed verbatim through
oracle/reference_loader.py in the authoring container) by tests/golden/*.pt.

What is restated, and where it lives in the reference:
  logdetexp                      hyperbolic_vae/manifolds.py:25-35
  normdist2plane                 hyperbolic_vae/manifolds.py:41-65
  RiemannianLayer/Geodesic/Mobius hyperbolic_vae/layers.py:35-147 (GeodesicLayer with pvae's unsqueeze,
                                 App. A.2 — the reference's expand at layers.py:98-102 only accepts B==1)
  ExpMap0                        hyperbolic_vae/layers.py:124-130
  Distance2PoincareHyperplanes   hyperbolic_vae/layers.py:150-228
  WrappedNormal                  hyperbolic_vae/distributions/wrapped_normal.py:14-89
  ModelA  VAEHyperbolicGyroplaneDecoder   models/vae_hyperbolic_gyroplane_decoder.py:36-152
  ModelB  ImageVAEHyperbolic + VAEHyperbolicExperiment.loss   models/vae_hyperbolic.py:38-233
  ModelC  VAEHyperbolicRNASeq    models/vae_hyperbolic_rnaseq.py:22-118
  ModelOneB  vae_one_b.VAE (hyperbolic, learned scale)   models/vae_one_b.py:17-250
  PvaeMnist  the pvae MNIST graph scripts/_9_pvae_replicate.py transcribes (:5-29,:124-158) with
             the objective of training/old_pvae_train.py:53-58

Noise is always INJECTED (eps / alpha / r arguments) so the CUDA path can be fed identical noise.
Every model exposes `loss(x, **noise) -> dict` returning the same keys as the reference.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .geoopt_min import ManifoldParameter, PoincareBall
from .geoopt_min.layers.stereographic import Distance2StereographicHyperplanes
from .geoopt_min.manifolds.stereographic import math as gmath
from .pvae_min.manifolds import PoincareBall as PvaeBall
from .pvae_min.manifolds import normdist2plane as _normdist2plane
from .pvae_min.ops.manifold_layers import GeodesicLayer, MobiusLayer, RiemannianLayer  # noqa: F401
from .pvae_min.distributions.hyperbolic_radius import HyperbolicRadius, impl_rsample
from .pvae_min.distributions.hyperspherical_uniform import HypersphericalUniform

MIN_NORM = 1e-15
GyroplaneLayer = GeodesicLayer


# ---- manifolds.py ---------------------------------------------------------------------------
def logdetexp(manifold, x, y, keepdim=False):
    d = manifold.dist(x, y, keepdim=keepdim)
    n = x.shape[-1]
    sc = manifold.c.sqrt()
    return (n - 1) * (torch.sinh(sc * d).log() - sc.log() - d.log())


def normdist2plane(manifold, x, a, p, keepdim=False, signed=False, dim=-1, norm=False):
    return _normdist2plane(manifold, x, a, p, keepdim=keepdim, signed=signed, dim=dim, norm=norm)


# ---- layers.py ------------------------------------------------------------------------------
class ExpMap0(nn.Module):
    def __init__(self, manifold):
        super().__init__()
        self.manifold = manifold

    def forward(self, input):
        return self.manifold.expmap0(input)


class Distance2PoincareHyperplanes(nn.Module):
    def __init__(self, plane_shape, num_planes, bias=True, signed=True, squared=False, *, ball, std=1.0):
        super().__init__()
        self.signed, self.squared, self.ball = signed, squared, ball
        self.num_planes, self.std = num_planes, std
        self.points = ManifoldParameter(torch.empty(num_planes, plane_shape), manifold=ball)
        if bias:
            self.bias = nn.Parameter(torch.empty(num_planes))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def forward(self, input):
        x = input.unsqueeze(-1)
        pts = self.points.permute(1, 0)
        d = self.ball.dist2plane(x=x, p=pts, a=pts, signed=self.signed, dim=-2)
        if self.squared and self.signed:
            d = d**2 * d.sign()
        elif self.squared:
            d = d**2
        if self.bias is not None:
            d = d + self.bias
        return d

    @torch.no_grad()
    def reset_parameters(self):
        direction = torch.randn_like(self.points)
        direction /= direction.norm(dim=-1, keepdim=True)
        distance = torch.empty_like(self.points[..., 0]).normal_(std=self.std)
        self.points.set_(self.ball.expmap0(direction * distance.unsqueeze(-1)))
        if self.bias is not None:
            nn.init.uniform_(self.bias, -1.0, 1.0)


# ---- distributions/wrapped_normal.py -------------------------------------------------------------
class WrappedNormal:
    """rsample/log_prob of the reference's WrappedNormal with the noise passed in."""

    def __init__(self, loc, scale, manifold, softplus=False):
        self.loc, self._scale = torch.broadcast_tensors(loc, scale)
        self.softplus = softplus
        self.manifold = manifold
        manifold.assert_check_point_on_manifold(self.loc)
        self.batch_shape = self.loc.shape[:-1]
        self.event_shape = self.loc.shape[-1:]

    @property
    def scale(self):
        return F.softplus(self._scale) if self.softplus else self._scale

    def rsample(self, sample_shape=torch.Size(), eps=None):
        shape = torch.Size(sample_shape) + self.loc.shape
        if eps is None:
            eps = torch.randn(shape, dtype=self.loc.dtype)
        assert eps.shape == shape, (eps.shape, shape)
        m = self.manifold
        zero = m.origin(self.event_shape, dtype=self.loc.dtype)
        v = self.scale * eps
        v = v / m.lambda_x(zero, keepdim=True)
        u = m.transp(zero, self.loc, v)
        return m.expmap(self.loc, u)

    def log_prob(self, x):
        m = self.manifold
        n = int(self.event_shape[0])
        zero = m.origin(self.event_shape, dtype=self.loc.dtype)
        loc = self.loc.unsqueeze(0).expand(x.shape[0], *self.batch_shape, n)
        if x.dim() < loc.dim():
            x = x.unsqueeze(1)
        v = m.logmap(loc, x)
        v = m.transp(loc, zero, v)
        u = v * m.lambda_x(zero, keepdim=True)
        sc = self.scale
        norm_pdf = (-(u**2) / (2 * sc**2) - sc.log() - math.log(math.sqrt(2 * math.pi))).sum(-1, keepdim=True)
        return norm_pdf - logdetexp(m, loc, x, keepdim=True)


# ---- RiemannianNormal with injected (alpha, r) -----------------------------------------------
class RiemannianNormal:
    """hyperbolic_vae/distributions/old_pvae_riemannian_normal.py:12-52 over pvae (App. A.2)."""

    def __init__(self, loc, scale, manifold: PvaeBall):
        assert not (torch.isnan(loc).any() or torch.isnan(scale).any())
        self.manifold, self.loc = manifold, loc
        manifold.assert_check_point_on_manifold(loc)
        self.scale = scale.clamp(min=0.1, max=7.0)
        self.dim = loc.shape[-1]
        self.radius = HyperbolicRadius(self.dim, manifold.c, self.scale)
        self.direction = HypersphericalUniform(self.dim - 1)

    def rsample(self, sample_shape=torch.Size(), alpha=None, r=None):
        if alpha is None:
            alpha = self.direction.sample(torch.Size([*sample_shape, *self.loc.shape[:-1]]))
        if r is None:
            r = self.radius.sample(sample_shape)
        r = impl_rsample.apply(r, self.scale, self.manifold.c, self.dim)
        return self.manifold.expmap_polar(self.loc, alpha, r)

    def log_prob(self, value):
        loc = self.loc.expand(value.shape)
        d2 = self.manifold.dist(loc, value, keepdim=True).pow(2)
        return -d2 / 2 / self.scale.pow(2) - self.direction._log_normalizer() - self.radius.log_normalizer


# ---- helpers -------------------------------------------------------------------------------
def relaxed_bernoulli_log_prob(value, temperature, probs=None, logits=None):
    """torch.distributions.RelaxedBernoulli(T, probs|logits).log_prob(value) written out:
    LogitRelaxedBernoulli density at logit(value) plus the sigmoid-transform log|det|."""
    if logits is None:
        eps = torch.finfo(probs.dtype).eps
        ps = probs.clamp(min=eps, max=1 - eps)
        logits = torch.log(ps) - torch.log1p(-ps)
    t = torch.as_tensor(temperature, dtype=value.dtype)
    finfo = torch.finfo(value.dtype)
    v = value.clamp(min=finfo.tiny, max=1.0 - finfo.eps)
    y = v.log() - (-v).log1p()
    diff = logits - y * t
    base = t.log() + diff - 2 * diff.exp().log1p()
    ladj = -F.softplus(-y) - F.softplus(y)
    return base - ladj


class _Base(nn.Module):
    def _prior(self, z, scale_value=1.0):
        m = self.manifold
        origin = m.origin(z.shape[-1], dtype=z.dtype)
        return WrappedNormal(origin, torch.ones_like(origin) * scale_value, m)


# ---- Model A: models/vae_hyperbolic_gyroplane_decoder.py -------------------------------------------
class ModelA(_Base):
    def __init__(self, data_shape=torch.Size([1, 32, 32]), latent_dim=2, manifold_curvature=1.0, beta=1.0, prior_scale=1.0):
        super().__init__()
        data_shape = torch.Size(data_shape)
        self.beta, self.latent_dim, self.prior_scale = beta, latent_dim, prior_scale
        self.manifold = PoincareBall(c=manifold_curvature)
        n = data_shape.numel()
        self.encoder = nn.Sequential(nn.Flatten(), nn.Linear(n, 64), nn.GELU(), nn.Linear(64, 16), nn.GELU())
        self.mu = nn.Sequential(nn.Linear(16, latent_dim), ExpMap0(self.manifold))
        self.scale = nn.Sequential(nn.Linear(16, latent_dim), nn.Softplus())
        self.decoder = nn.Sequential(
            Distance2StereographicHyperplanes(latent_dim, 16, ball=self.manifold),
            nn.GELU(), nn.Linear(16, 64), nn.GELU(), nn.Linear(64, n), nn.Sigmoid(),
            nn.Unflatten(dim=-1, unflattened_size=data_shape),
        )

    def forward(self, x, eps=None):
        h = self.encoder(x)
        mu, scale = self.mu(h), self.scale(h)
        z = WrappedNormal(mu, scale, self.manifold).rsample(torch.Size([1]), eps=eps).squeeze(0)
        return mu, scale, z, self.decoder(z)

    def loss(self, x, eps=None):
        mu, scale, z, x_hat = self.forward(x, eps)
        recon = -relaxed_bernoulli_log_prob(x.flatten(1), 1.0, probs=x_hat.flatten(1)).sum(-1)
        z1 = z.unsqueeze(0)
        logq = WrappedNormal(mu, scale, self.manifold).log_prob(z1)
        logp = self._prior(z, self.prior_scale).log_prob(z1)
        kl = (logq - logp).sum(-1).squeeze(0)
        return dict(loss_total=(recon + self.beta * kl).mean(), recon_loss=recon.mean(), kl_loss=kl.mean())


# ---- Model B: models/vae_hyperbolic.py ---------------------------------------------------------
class ImageVAEHyperbolic(nn.Module):
    def __init__(self, latent_dim, act_fn, image_shape, encoder_last_layer_module, decoder_first_layer_module,
                 manifold_curvature, loss_recon):
        super().__init__()
        self.latent_dim = latent_dim
        self.encoder_last_layer_module = encoder_last_layer_module
        self.decoder_first_layer_module = decoder_first_layer_module
        self.loss_recon = loss_recon
        ch, w, h = image_shape
        self.manifold = PoincareBall(c=manifold_curvature)
        self.encoder = nn.Sequential(
            nn.Conv2d(ch, 16, 3, 2, 1), act_fn(), nn.Conv2d(16, 32, 3, 2, 1), act_fn(),
            nn.Conv2d(32, 32, 3, 2, 1), act_fn(), nn.Flatten(),
        )
        feat = 32 * (w // 8) * (h // 8)
        if encoder_last_layer_module == "linear":
            self.mu = nn.Linear(feat, latent_dim)
        elif encoder_last_layer_module == "mobius":
            self.mu = MobiusLayer(feat, latent_dim, self.manifold)
        else:
            raise ValueError(f"encoder_last_layer_module {encoder_last_layer_module} not supported")
        self.log_var = nn.Linear(feat, latent_dim)
        if decoder_first_layer_module == "linear":
            first = nn.Linear(latent_dim, feat)
        elif decoder_first_layer_module == "geodesic":
            first = GeodesicLayer(latent_dim, feat, self.manifold)
        elif decoder_first_layer_module == "mobius":
            first = MobiusLayer(latent_dim, feat, self.manifold)
        elif decoder_first_layer_module == "geoopt_gyroplane":
            first = Distance2StereographicHyperplanes(latent_dim, feat, ball=self.manifold)
        else:
            raise ValueError(f"decoder_first_layer {decoder_first_layer_module} not supported")
        layers = [
            first, act_fn(), nn.Unflatten(-1, (32, w // 8, h // 8)),
            nn.ConvTranspose2d(32, 32, 3, 2, 1, output_padding=1), act_fn(), nn.Conv2d(32, 32, 3, 1, 1), act_fn(),
            nn.ConvTranspose2d(32, 16, 3, 2, 1, output_padding=1), act_fn(), nn.Conv2d(16, 16, 3, 1, 1), act_fn(),
            nn.ConvTranspose2d(16, ch, 3, 2, 1, output_padding=1),
        ]
        if loss_recon == "mse":
            layers.append(nn.Sigmoid())
        self.decoder = nn.Sequential(*layers)

    def on_manifold(self, mu):
        return self.manifold.expmap0(mu) if self.encoder_last_layer_module == "linear" else mu

    def forward(self, x, eps=None):
        e = self.encoder(x)
        mu = self.mu(e)
        log_var = torch.zeros_like(mu) if self.loss_recon == "bernoulli" else self.log_var(e)
        scale = torch.exp(0.5 * log_var)
        z = WrappedNormal(self.on_manifold(mu), scale, self.manifold).rsample(torch.Size([1]), eps=eps).squeeze(0)
        return mu, log_var, z, self.decoder(z)


class ModelB(_Base):
    def __init__(self, image_shape=(1, 32, 32), latent_dim=2, manifold_curvature=1.0, encoder_last_layer_module="linear",
                 decoder_first_layer_module="linear", beta=1.0, loss_recon="mse"):
        super().__init__()
        self.model = ImageVAEHyperbolic(latent_dim, nn.GELU, image_shape, encoder_last_layer_module,
                                        decoder_first_layer_module, manifold_curvature, loss_recon)
        self.beta, self.loss_recon = beta, loss_recon

    @property
    def manifold(self):
        return self.model.manifold

    def forward(self, x, eps=None):
        return self.model(x, eps)

    def loss(self, x, eps=None):
        mu, log_var, z, x_hat = self.model(x, eps)
        q = WrappedNormal(self.model.on_manifold(mu), torch.exp(0.5 * log_var), self.manifold)
        z1 = z.unsqueeze(0)
        kl = (q.log_prob(z1) - self._prior(z).log_prob(z1)).sum()
        if self.loss_recon == "mse":
            recon = F.mse_loss(x_hat, x, reduction="sum")
        elif self.loss_recon == "bernoulli":
            recon = -relaxed_bernoulli_log_prob(x.flatten(1), 0.1, logits=x_hat.flatten(1)).mean()
        else:
            raise ValueError(f"loss_recon {self.loss_recon} not supported")
        return dict(loss_total=recon + self.beta * kl, loss_recon=recon, loss_kl=kl)


# ---- Model C: models/vae_hyperbolic_rnaseq.py ---------------------------------------------------
class ModelC(_Base):
    def __init__(self, input_data_shape, latent_dim, manifold_curvature, hidden_layer_dim, beta):
        super().__init__()
        n = torch.Size(input_data_shape).numel()
        self.beta, self.latent_dim, self.prior_scale = beta, latent_dim, 1.0
        self.manifold = PoincareBall(c=manifold_curvature)
        self.encoder = nn.Sequential(nn.Linear(n, hidden_layer_dim), nn.GELU())
        self.mu = nn.Sequential(nn.Linear(hidden_layer_dim, latent_dim), ExpMap0(self.manifold))
        self.scale = nn.Sequential(nn.Linear(hidden_layer_dim, latent_dim), nn.Softplus())
        self.decoder = nn.Sequential(
            Distance2StereographicHyperplanes(latent_dim, hidden_layer_dim, ball=self.manifold),
            nn.GELU(), nn.Linear(hidden_layer_dim, n), nn.Sigmoid(),
        )

    def forward(self, x, eps=None):
        h = self.encoder(x)
        mu, scale = self.mu(h), self.scale(h)
        z = WrappedNormal(mu, scale, self.manifold).rsample(torch.Size([1]), eps=eps).squeeze(0)
        return mu, scale, z, self.decoder(z)

    def loss(self, x, eps=None):
        mu, scale, z, x_hat = self.forward(x, eps)
        recon = (x_hat.flatten(1) - x.flatten(1)).pow(2).sum(-1)
        z1 = z.unsqueeze(0)
        kl = (WrappedNormal(mu, scale, self.manifold).log_prob(z1) - self._prior(z, self.prior_scale).log_prob(z1))
        kl = kl.sum(-1).squeeze(0)
        return dict(loss_total=(recon + self.beta * kl).mean(), recon_loss=recon.mean(), kl_loss=kl.mean())


# ---- vae_one_b.VAE (hyperbolic latent, learned scale) ---------------------------------------------
class ModelOneB(_Base):
    def __init__(self, input_size, hidden_layer_dim, latent_dim, latent_curvature, prior_scale, beta,
                 kl_loss_method="logmap0_analytic", last_activation="none", loss_recon_method="MSE"):
        super().__init__()
        input_size = torch.Size(input_size)
        n = input_size.numel()
        self.latent_manifold = PoincareBall(latent_curvature)
        self.prior_scale, self.beta, self.kl_loss_method = prior_scale, beta, kl_loss_method
        self.last_activation, self.loss_recon_method = last_activation, loss_recon_method
        self.encoder = nn.Sequential(*([] if len(input_size) == 1 else [nn.Flatten()]), nn.Linear(n, hidden_layer_dim), nn.GELU())
        self.mu = nn.Sequential(nn.Linear(hidden_layer_dim, latent_dim), ExpMap0(self.latent_manifold))
        self.scale = nn.Sequential(nn.Linear(hidden_layer_dim, latent_dim), nn.Softplus())
        tail = [] if len(input_size) == 1 else [nn.Unflatten(1, input_size)]
        if last_activation == "sigmoid":
            tail.append(nn.Sigmoid())
        elif last_activation == "softplus":
            tail.append(nn.Softplus())
        self.decoder = nn.Sequential(
            Distance2PoincareHyperplanes(latent_dim, hidden_layer_dim, ball=self.latent_manifold),
            nn.GELU(), nn.Linear(hidden_layer_dim, n), *tail,
        )

    @property
    def manifold(self):
        return self.latent_manifold

    def forward(self, x, eps=None):
        h = self.encoder(x)
        mu, scale = self.mu(h), self.scale(h)
        z = WrappedNormal(mu, scale, self.manifold).rsample(eps=eps)
        return mu, scale, z, self.decoder(z)

    def loss_kl(self, mu, scale, z):
        m = self.manifold
        if self.kl_loss_method == "logmap0_analytic":
            mu0 = m.logmap0(mu)
            ps = torch.ones_like(scale) * self.prior_scale
            var_ratio = (scale / ps).pow(2)
            t1 = (mu0 / ps).pow(2)
            return (0.5 * (var_ratio + t1 - 1 - var_ratio.log())).mean()
        if self.kl_loss_method == "log_prob":
            q = WrappedNormal(mu, scale, m)
            p = WrappedNormal(m.origin(mu.shape[-1]), torch.ones_like(scale) * self.prior_scale, m)
            lq, lp = q.log_prob(z), p.log_prob(z)
            return (lq.exp() * (lq - lp)).mean()
        if self.kl_loss_method == "logmap0_log_prob":
            mu0, z0 = m.logmap0(mu), m.logmap0(z)
            ps = torch.ones_like(scale) * self.prior_scale
            lp = torch.distributions.Normal(torch.zeros_like(mu0), ps).log_prob(z0).sum(-1)
            lq = torch.distributions.Normal(mu0, scale).log_prob(z0).sum(-1)
            return (lq.exp() * (lq - lp)).mean()
        raise ValueError(f"Unrecognized kl_loss_method: {self.kl_loss_method}")

    def loss(self, x, eps=None):
        mu, scale, z, out = self.forward(x, eps)
        if self.loss_recon_method == "MSE":
            recon = F.mse_loss(out, x, reduction="mean")
        elif self.loss_recon_method == "binary_cross_entropy_with_logits":
            recon = F.binary_cross_entropy_with_logits(out, x, reduction="mean")
        else:
            raise ValueError(f"Unrecognized loss_recon_method: {self.loss_recon_method}")
        kl = self.loss_kl(mu, scale, z)
        return dict(loss_reconstruction=recon, loss_kl=kl, loss_total=recon + self.beta * kl)


# ---- pvae MNIST graph (config 2) ------------------------------------------------------------------
class PvaeMnist(nn.Module):
    """EncMob(784->h->Mobius(h,D)+expmap0, sigma: Linear(h,1)->softplus+eta) / DecGeo(Geodesic(D,h)->
    ReLU->Linear(h,784)); RiemannianNormal prior & posterior; Bernoulli (BCE-with-logits) likelihood;
    objective = -E[log p(x|z)] + beta * (log q(z|x) - log p(z)), K=1 (App. A.2 vae_objective)."""

    def __init__(self, latent_dim=10, hidden_dim=600, c=1.0, prior_std=1.0, beta=1.0, data_size=(1, 28, 28)):
        super().__init__()
        self.data_size = torch.Size(data_size)
        n = self.data_size.numel()
        self.manifold = PvaeBall(latent_dim, c)
        self.beta, self.prior_std, self.latent_dim = beta, prior_std, latent_dim
        self.enc = nn.Sequential(nn.Linear(n, hidden_dim), nn.ReLU())
        self.fc21 = MobiusLayer(hidden_dim, latent_dim, self.manifold)
        self.fc22 = nn.Linear(hidden_dim, 1)
        self.dec0 = GeodesicLayer(latent_dim, hidden_dim, self.manifold)
        self.fc31 = nn.Linear(hidden_dim, n)
        self._pz_mu = nn.Parameter(torch.zeros(1, latent_dim), requires_grad=False)
        self._pz_logvar = nn.Parameter(torch.zeros(1, 1), requires_grad=False)

    def encode(self, x):
        e = self.enc(x.view(x.shape[0], -1))
        mu = self.manifold.expmap0(self.fc21(e))
        return mu, F.softplus(self.fc22(e)) + 1e-5

    def decode(self, z):
        return self.fc31(F.relu(self.dec0(z)))

    def loss(self, x, alpha=None, r=None):
        B = x.shape[0]
        mu, sigma = self.encode(x)
        q = RiemannianNormal(mu, sigma, self.manifold)
        zs = q.rsample(torch.Size([1]), alpha=alpha, r=r)  # (1, B, D)
        logits = self.decode(zs)  # (1, B, 784)
        lpx_z = -F.binary_cross_entropy_with_logits(logits, x.view(1, B, -1).expand_as(logits), reduction="none").sum(-1)
        pz_scale = F.softplus(self._pz_logvar) / math.log(2) * self.prior_std
        p = RiemannianNormal(self._pz_mu, pz_scale, self.manifold)
        kld = q.log_prob(zs).sum(-1) - p.log_prob(zs).sum(-1)
        obj = -lpx_z.mean(0).sum() + self.beta * kld.mean(0).sum()
        return dict(loss_total=obj, recon_loss=-lpx_z.mean(0).sum(), kl_loss=kld.mean(0).sum())
