"""Gaussian-mixture regression heads (probabilistic centre / size / heading).

Parameter layout of /root/reference/models/p2rnet/modules/mdn.py:17-161 (CategoryEmbeddingMDN with a
'cbr' backbone 128->128, MixtureDensityHead with pi = 1x1 conv 128->G, mu (G,d), log_sigma (G,d)).
predict()  = training forward: sum_g sigmoid(pi)_g * (mu_g + exp(log_sigma)_g * eps), eps ~ N(0,1) drawn with
             the same torch call, shape (rows, G, n_samples, d) and order as mdn.py:44, so a fixed torch
             RNG state reproduces the reference's draw;
generate() = eval: the mixture mean sum_g sigmoid(pi)_g * mu_g (mdn.py:85-99), deterministic.
The 1x1 convs run on channel-last rows with the B200 GEMM; the mixing itself is a (rows x G) . (G x d)
product on <= 4096 rows.
"""
import torch
import torch.nn as nn
from torch.distributions.bernoulli import Bernoulli

from .sub_modules import SingleConv


class Struct:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def update(self, **kw):
        self.__dict__.update(kw)


class MixtureDensityHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.hparams = config
        self.pi = SingleConv(config.input_dim, config.num_gaussian, order="c")
        self.log_sigma = nn.Parameter(torch.zeros(config.num_gaussian, config.out_dim))
        self.mu = nn.Parameter(config.mu_bias_init)

    def forward_rows(self, x):
        return torch.sigmoid(self.pi.forward_rows(x).float())          # (rows, G)

    def sample(self, num_samples, n_rows):
        sigma = torch.exp(self.log_sigma)[None, :, None, :].expand(n_rows, -1, num_samples, -1)
        mu = self.mu[None, :, None, :].expand(n_rows, -1, num_samples, -1)
        eps = mu.data.new(mu.size()).normal_()                          # same RNG consumption as mdn.py:44
        return eps * sigma + mu

    def generate_samples(self, pi_rows, n_samples=None, sample_pi=False):
        n_samples = self.hparams.n_samples if n_samples is None else n_samples
        samples = self.sample(n_samples, pi_rows.size(0))               # (rows, G, n, d)
        if sample_pi:
            w = Bernoulli(pi_rows).sample((n_samples,)).permute(1, 2, 0)
            w = w.unsqueeze(-1)
        else:
            w = pi_rows[:, :, None, None]
        return torch.sum(samples * w, dim=1)                            # (rows, n, d)

    def point_prediction(self, pi_rows, n_samples=None, sample_pi=False):
        s = self.generate_samples(pi_rows, n_samples, sample_pi)
        if self.hparams.central_tendency == "mean":
            return torch.mean(s, dim=1)
        if self.hparams.central_tendency == "median":
            return torch.median(s, dim=1).values
        raise NotImplementedError

    def get_mean(self, pi_rows):
        return torch.sum(self.mu[None] * pi_rows[:, :, None], dim=1)    # (rows, d)


class CategoryEmbeddingMDN(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.hparams = config
        assert not config.batch_norm_continuous_input
        self.backbone = SingleConv(config.continuous_dim, config.hidden_dim, order="cbr")
        config.mdn_config.update(input_dim=config.hidden_dim)
        self.mdn = MixtureDensityHead(config.mdn_config)

    def pi_rows(self, x_rows):
        return self.mdn.forward_rows(self.backbone.forward_rows(x_rows))

    def predict_rows(self, x_rows):
        return self.mdn.point_prediction(self.pi_rows(x_rows))

    def generate_rows(self, x_rows, multi_modes=False, n_samples=10):
        pi = self.pi_rows(x_rows)
        if multi_modes:
            return self.mdn.point_prediction(pi, n_samples=n_samples, sample_pi=True), pi
        return self.mdn.get_mean(pi), pi
