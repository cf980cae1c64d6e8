"""Training-step harness: the reference's model graphs as plain nn.Modules over the hvae kernels.

These are the CALLERS of the hot path ("re-expressed as plain nn.Module harness only";
Lightning itself is out of scope).  Module/parameter names follow the reference so its state_dicts load:
  ModelA    models/vae_hyperbolic_gyroplane_decoder.py:36-152   (VAEHyperbolicGyroplaneDecoder)
  ModelB    models/vae_hyperbolic.py:38-233                     (ImageVAEHyperbolic + VAEHyperbolicExperiment.loss)
  ModelC    models/vae_hyperbolic_rnaseq.py:22-118              (VAEHyperbolicRNASeq)
  ModelOneB models/vae_one_b.py:17-250                          (VAE, hyperbolic latent, learned scale)
  PvaeMnist scripts/_9_pvae_replicate.py:5-29,124-158 + training/old_pvae_train.py:53-58 (config 2)

The Euclidean trunk (Linear / Conv / GELU) is torch's library path (cuBLAS/cuDNN) — not this repo's
product.  Everything hyperbolic is one of our kernels: with `fused=True` the
posterior sample + both log-densities + the KL subtraction are ONE kernel forward and ONE backward
(K4+K5 "latent head"); with `fused=False` the step goes through the drop-in WrappedNormal exactly as
the reference's code does (rsample, log_prob x2), each call one kernel.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .distributions.wrapped_normal import WrappedNormal
from .layers import Distance2PoincareHyperplanes, Distance2StereographicHyperplanes, ExpMap0, GeodesicLayer, Linear, MobiusLayer
from .manifolds import PoincareBall


def relaxed_bernoulli_nll(value, temperature, probs=None, logits=None):
    """-RelaxedBernoulli(T, probs|logits).log_prob(value), elementwise (torch.distributions semantics)."""
    if logits is None:
        eps = torch.finfo(probs.dtype).eps
        ps = probs.clamp(min=eps, max=1 - eps)
        logits = torch.log(ps) - torch.log1p(-ps)
    t = float(temperature)  # a Python scalar: no host->device copy (a tensor built here would break CUDA-graph capture)
    finfo = torch.finfo(value.dtype)
    v = value.clamp(min=finfo.tiny, max=1.0 - finfo.eps)
    y = v.log() - (-v).log1p()
    diff = logits - y * t
    base = math.log(t) + diff - 2 * diff.exp().log1p()
    ladj = -F.softplus(-y) - F.softplus(y)
    return -(base - ladj)


class _LatentMixin:
    fused = True

    def _sample_and_kl(self, mu, scale, eps, prior_scale):
        """-> z (B,D), kl (B,) = log q(z|x) - log p(z).  eps: (B,D) standard normal noise or None."""
        m = self.manifold
        if eps is None:
            eps = torch.randn_like(mu)
        eps = eps.reshape(mu.shape)
        if self.fused:
            return ops.latent_head(mu, scale, eps, prior_scale, m.c_value)
        q = WrappedNormal(mu, scale, m)
        z1 = q.rsample(torch.Size([1]), eps=eps.unsqueeze(0))
        p = WrappedNormal.origin_prior(mu.shape[-1], prior_scale, m, device=mu.device)
        kl = (q.log_prob(z1) - p.log_prob(z1)).sum(-1).squeeze(0)
        return z1.squeeze(0), kl


class ModelA(nn.Module, _LatentMixin):
    def __init__(self, data_shape=torch.Size([1, 32, 32]), latent_dim=2, manifold_curvature=1.0, beta=1.0, prior_scale=1.0,
                 fused=True):
        super().__init__()
        data_shape = torch.Size(data_shape)
        self.beta, self.latent_dim, self.prior_scale, self.fused = beta, latent_dim, prior_scale, fused
        self.manifold = PoincareBall(c=manifold_curvature)
        n = data_shape.numel()
        self.encoder = nn.Sequential(nn.Flatten(), Linear(n, 64), nn.GELU(), Linear(64, 16), nn.GELU())
        self.mu = nn.Sequential(Linear(16, latent_dim), ExpMap0(self.manifold))
        self.scale = nn.Sequential(Linear(16, latent_dim), nn.Softplus())
        self.decoder = nn.Sequential(
            Distance2StereographicHyperplanes(latent_dim, 16, ball=self.manifold),
            nn.GELU(), Linear(16, 64), nn.GELU(), Linear(64, n), nn.Sigmoid(),
            nn.Unflatten(dim=-1, unflattened_size=data_shape),
        )

    def loss(self, x, eps=None):
        h = self.encoder(x)
        mu, scale = self.mu(h), self.scale(h)
        z, kl = self._sample_and_kl(mu, scale, eps, self.prior_scale)
        if self.fused and z.is_cuda:
            # the decoder's final Sigmoid (+ Unflatten) folded into the RelaxedBernoulli head: one row kernel per direction
            recon = ops.recon_rows(self.decoder[:-2](z), x.flatten(1), ops.RECON_RB_SIGMOID, 1.0)
        else:
            x_hat = self.decoder(z)
            recon = relaxed_bernoulli_nll(x.flatten(1), 1.0, probs=x_hat.flatten(1)).sum(-1)
        return dict(loss_total=(recon + self.beta * kl).mean(), recon_loss=recon.mean(), kl_loss=kl.mean())


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
            self.mu = Linear(feat, latent_dim)
        elif encoder_last_layer_module == "mobius":
            self.mu = MobiusLayer(feat, latent_dim, self.manifold)
        else:
            raise ValueError(f"encoder_last_layer_module {encoder_last_layer_module} not supported")
        self.log_var = Linear(feat, latent_dim)
        if decoder_first_layer_module == "linear":
            first = Linear(latent_dim, feat)
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

    def encode(self, x):
        e = self.encoder(x)
        mu = self.mu(e)
        log_var = torch.zeros_like(mu) if self.loss_recon == "bernoulli" else self.log_var(e)
        mu_m = self.manifold.expmap0(mu) if self.encoder_last_layer_module == "linear" else mu
        return mu_m, torch.exp(0.5 * log_var)


class ModelB(nn.Module, _LatentMixin):
    def __init__(self, image_shape=(1, 32, 32), latent_dim=2, manifold_curvature=1.0, encoder_last_layer_module="linear",
                 decoder_first_layer_module="linear", beta=1.0, loss_recon="mse", fused=True):
        super().__init__()
        self.model = ImageVAEHyperbolic(latent_dim, nn.GELU, image_shape, encoder_last_layer_module,
                                        decoder_first_layer_module, manifold_curvature, loss_recon)
        self.beta, self.loss_recon, self.fused = beta, loss_recon, fused

    @property
    def manifold(self):
        return self.model.manifold

    def loss(self, x, eps=None):
        mu_m, scale = self.model.encode(x)
        z, kl_rows = self._sample_and_kl(mu_m, scale, eps, 1.0)
        kl = kl_rows.sum()
        head = self.fused and z.is_cuda
        if self.loss_recon == "mse" and head:
            pre = self.model.decoder[:-1](z)   # the final Sigmoid is folded into the MSE head
            recon = ops.recon_rows(pre.flatten(1), x.flatten(1), ops.RECON_SIGMOID_MSE).sum()
        elif self.loss_recon == "bernoulli" and head:
            logits = self.model.decoder(z).flatten(1)
            recon = ops.recon_rows(logits, x.flatten(1), ops.RECON_RB_LOGITS, 0.1).sum() / logits.numel()
        elif self.loss_recon == "mse":
            recon = F.mse_loss(self.model.decoder(z), x, reduction="sum")
        elif self.loss_recon == "bernoulli":
            recon = relaxed_bernoulli_nll(x.flatten(1), 0.1, logits=self.model.decoder(z).flatten(1)).mean()
        else:
            raise ValueError(f"loss_recon {self.loss_recon} not supported")
        return dict(loss_total=recon + self.beta * kl, loss_recon=recon, loss_kl=kl)


class ModelC(nn.Module, _LatentMixin):
    def __init__(self, input_data_shape, latent_dim, manifold_curvature, hidden_layer_dim, beta, fused=True):
        super().__init__()
        n = torch.Size(input_data_shape).numel()
        self.beta, self.latent_dim, self.prior_scale, self.fused = beta, latent_dim, 1.0, fused
        self.manifold = PoincareBall(c=manifold_curvature)
        self.encoder = nn.Sequential(Linear(n, hidden_layer_dim), nn.GELU())
        self.mu = nn.Sequential(Linear(hidden_layer_dim, latent_dim), ExpMap0(self.manifold))
        self.scale = nn.Sequential(Linear(hidden_layer_dim, latent_dim), nn.Softplus())
        self.decoder = nn.Sequential(
            Distance2StereographicHyperplanes(latent_dim, hidden_layer_dim, ball=self.manifold),
            nn.GELU(), Linear(hidden_layer_dim, n), nn.Sigmoid(),
        )

    def loss(self, x, eps=None):
        h = self.encoder(x)
        mu, scale = self.mu(h), self.scale(h)
        z, kl = self._sample_and_kl(mu, scale, eps, self.prior_scale)
        if self.fused and z.is_cuda:
            recon = ops.recon_rows(self.decoder[:-1](z).flatten(1), x.flatten(1), ops.RECON_SIGMOID_MSE)  # Sigmoid folded in
        else:
            x_hat = self.decoder(z)
            recon = (x_hat.flatten(1) - x.flatten(1)).pow(2).sum(-1)
        return dict(loss_total=(recon + self.beta * kl).mean(), recon_loss=recon.mean(), kl_loss=kl.mean())


class ModelOneB(nn.Module, _LatentMixin):
    def __init__(self, input_size, hidden_layer_dim, latent_dim, latent_curvature, prior_scale, beta,
                 kl_loss_method="logmap0_analytic", last_activation="none", loss_recon_method="MSE"):
        super().__init__()
        input_size = torch.Size(input_size)
        n = input_size.numel()
        self.latent_manifold = PoincareBall(latent_curvature)
        self.prior_scale, self.beta, self.kl_loss_method = prior_scale, beta, kl_loss_method
        self.last_activation, self.loss_recon_method = last_activation, loss_recon_method
        self.encoder = nn.Sequential(*([] if len(input_size) == 1 else [nn.Flatten()]), Linear(n, hidden_layer_dim), nn.GELU())
        self.mu = nn.Sequential(Linear(hidden_layer_dim, latent_dim), ExpMap0(self.latent_manifold))
        self.scale = nn.Sequential(Linear(hidden_layer_dim, latent_dim), nn.Softplus())
        tail = [] if len(input_size) == 1 else [nn.Unflatten(1, input_size)]
        if last_activation == "sigmoid":
            tail.append(nn.Sigmoid())
        elif last_activation == "softplus":
            tail.append(nn.Softplus())
        self.decoder = nn.Sequential(
            Distance2PoincareHyperplanes(latent_dim, hidden_layer_dim, ball=self.latent_manifold),
            nn.GELU(), Linear(hidden_layer_dim, n), *tail,
        )

    @property
    def manifold(self):
        return self.latent_manifold

    def loss_kl(self, mu, scale, z):
        m = self.latent_manifold
        if self.kl_loss_method == "logmap0_analytic":
            mu0 = m.logmap0(mu)
            var_ratio = (scale / self.prior_scale).pow(2)
            t1 = (mu0 / self.prior_scale).pow(2)
            return (0.5 * (var_ratio + t1 - 1 - var_ratio.log())).mean()
        if self.kl_loss_method == "log_prob":
            # NB (vae_one_b.py:193-213): z is (B,D) here, so WrappedNormal.log_prob broadcasts to ALL PAIRS
            # (z_i under q_j) -> (B,B,1); the prior is built with a (B,D) scale so it broadcasts the same way.
            q = WrappedNormal(mu, scale, m)
            p = WrappedNormal(m.origin(mu.shape[-1], device=mu.device), torch.ones_like(scale) * self.prior_scale, m)
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
        h = self.encoder(x)
        mu, scale = self.mu(h), self.scale(h)
        z = WrappedNormal(mu, scale, self.latent_manifold).rsample(eps=eps)
        out = self.decoder(z)
        if self.loss_recon_method == "MSE":
            recon = F.mse_loss(out, x, reduction="mean")
        elif self.loss_recon_method == "binary_cross_entropy_with_logits":
            recon = F.binary_cross_entropy_with_logits(out, x, reduction="mean")
        else:
            raise ValueError(f"Unrecognized loss_recon_method: {self.loss_recon_method}")
        kl = self.loss_kl(mu, scale, z)
        return dict(loss_reconstruction=recon, loss_kl=kl, loss_total=recon + self.beta * kl)


class PvaeMnist(nn.Module):
    """Config 2 — the pvae MNIST graph the reference transcribes in scripts/_9_pvae_replicate.py:5-29,124-158
    with the objective of training/old_pvae_train.py:53-58 (App. A.2):
      enc: Linear(784,h) ReLU -> mu = expmap0(MobiusLayer(h,D)(e)), sigma = softplus(Linear(h,1)) + 1e-5
      posterior / prior: RiemannianNormal (HyperbolicRadius rejection sampler), K = 1
      dec: GeodesicLayer(D,h) ReLU -> Linear(h,784) logits; Bernoulli likelihood (BCE with logits)
      loss = -E log p(x|z) + beta (log q(z|x) - log p(z)), summed over the batch."""

    def __init__(self, latent_dim=10, hidden_dim=600, c=1.0, prior_std=1.0, beta=1.0, data_size=(1, 28, 28), fused=True):
        super().__init__()
        from .distributions.riemannian_normal import RiemannianNormal  # noqa: F401

        self.fused = fused

        self.data_size = torch.Size(data_size)
        n = self.data_size.numel()
        self.manifold = PoincareBall(c)
        self.beta, self.prior_std, self.latent_dim = beta, prior_std, latent_dim
        self.enc = nn.Sequential(Linear(n, hidden_dim), nn.ReLU())
        self.fc21 = MobiusLayer(hidden_dim, latent_dim, self.manifold)
        self.fc22 = Linear(hidden_dim, 1)
        self.dec0 = GeodesicLayer(latent_dim, hidden_dim, self.manifold)
        self.fc31 = Linear(hidden_dim, n)
        self._pz_mu = nn.Parameter(torch.zeros(1, latent_dim), requires_grad=False)
        self._pz_logvar = nn.Parameter(torch.zeros(1, 1), requires_grad=False)

    def _prior(self):
        """RiemannianNormal(0, softplus(logvar)/ln2 * prior_std).  The reference rebuilds it (softplus + the float64
        normaliser series) every step; its parameters are frozen (requires_grad=False), so the built object is cached
        and rebuilt only when they change (tensor version counters) or start requiring grad."""
        from .distributions.riemannian_normal import RiemannianNormal

        lv, mu = self._pz_logvar, self._pz_mu
        frozen = not (lv.requires_grad or mu.requires_grad)
        key = (lv._version, mu._version, lv.data_ptr(), mu.data_ptr(), self.prior_std)
        cache = getattr(self, "_prior_cache", None)
        if frozen and cache is not None and cache[0] == key:
            return cache[1]
        with torch.set_grad_enabled(not frozen):
            p = RiemannianNormal(mu, F.softplus(lv) / math.log(2) * self.prior_std, self.manifold)
        capturing = lv.is_cuda and torch.cuda.is_current_stream_capturing()
        if frozen and not capturing:
            self._prior_cache = (key, p)
        return p

    def encode(self, x, clamp_sigma=False):
        """clamp_sigma: return sigma already clamped to RiemannianNormal's [0.1, 7] (one fused kernel with the softplus)."""
        if self.fused and x.is_cuda:
            e = self.enc[0](x.view(x.shape[0], -1), relu=True)   # Linear + ReLU: one GEMM, the mask inside the gradient split
        else:
            e = self.enc(x.view(x.shape[0], -1))
        mu = self.manifold.expmap0(self.fc21(e))
        h = self.fc22(e)
        if clamp_sigma and h.is_cuda:
            return mu, ops.sigma_head(h, 1e-5, 0.1, 7.0)
        return mu, F.softplus(h) + 1e-5

    def decode(self, z):
        if self.fused and z.is_cuda:
            return self.fc31(self.dec0(z, relu=True))   # the ReLU runs inside the gyroplane kernels
        return self.fc31(F.relu(self.dec0(z)))

    def loss(self, x, alpha=None, r=None):
        from .distributions.riemannian_normal import RiemannianNormal

        B = x.shape[0]
        mu, sigma = self.encode(x, clamp_sigma=self.fused)
        q = RiemannianNormal(mu, sigma, self.manifold, scale_is_clamped=self.fused and sigma.is_cuda)
        p = self._prior()
        head = self.fused and mu.is_cuda and not (p.scale.requires_grad or p.loc.requires_grad)
        if head:
            zs, kld = q.rsample_kl(p, alpha=alpha, r=r)   # sample + both log-densities: one kernel per direction
        else:
            zs = q.rsample(torch.Size([1]), alpha=alpha, r=r)  # (1,B,D)
        logits = self.decode(zs)
        if self.fused:
            nll = ops.bernoulli_nll_rows(logits, x.view(B, -1))  # one row kernel per direction
            if not head:
                kld = q.kl_mc(zs, p)  # one kernel: both log-densities and their difference
            total, recon, kl = ops.pvae_loss(nll, kld, self.beta)   # the three scalars in one launch (double accumulation)
            return dict(loss_total=total, recon_loss=recon, kl_loss=kl)
        lpx_z = -F.binary_cross_entropy_with_logits(logits, x.view(1, B, -1).expand_as(logits), reduction="none").sum(-1)
        kld = q.log_prob(zs).sum(-1) - p.log_prob(zs).sum(-1)
        recon = -lpx_z.mean(0).sum()
        kl = kld.mean(0).sum()
        return dict(loss_total=recon + self.beta * kl, recon_loss=recon, kl_loss=kl)
