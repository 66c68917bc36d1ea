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
import os

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.distributions.bernoulli import Bernoulli

from .. import _lib, ops
from .sub_modules import SingleConv


# sigmoid + sampling + mixing of a head as one forward and one backward launch (csrc/gmm_ops.cu).  Default since round 2
# (parity-green on a B200 with the flag on); P2R_FUSED_GMM=0 selects the chain of torch kernels it is tested against.
# Read at call time (a test / bench child flips it per process).
def fused_gmm_enabled():
    return os.environ.get("P2R_FUSED_GMM", "1") != "0"


class _FusedGMMPredict(Function):
    """logits [rows,G] (f32 / bf16), mu [G,d] (f32 / f64), log_sigma f32 [G,d], eps [rows,G,1,d] (mu's type)
    -> sum_g sigmoid(logit) * (mu + exp(log_sigma) * eps), [rows,d] in mu's type (mdn.py:36-84, n_samples = 1)."""

    @staticmethod
    def forward(ctx, logits, mu, log_sigma, eps):
        logits = logits if logits.is_contiguous() else logits.contiguous()
        if logits.dtype not in (torch.float32, torch.bfloat16):
            logits = logits.float()
        mu_c = mu.detach().contiguous()
        ls_c = log_sigma.detach().float().contiguous()
        eps = eps.contiguous()
        assert mu_c.dtype in (torch.float32, torch.float64) and eps.dtype == mu_c.dtype
        rows, g = logits.shape
        d = mu_c.shape[1]
        out = torch.empty(rows, d, dtype=mu_c.dtype, device=logits.device)
        with torch.cuda.device(logits.device):
            _lib.call("p2r_gmm_mix", logits.data_ptr(), int(logits.dtype == torch.bfloat16), mu_c.data_ptr(),
                      int(mu_c.dtype == torch.float64), ls_c.data_ptr(), eps.data_ptr(), rows, g, d, out.data_ptr(),
                      torch.cuda.current_stream().cuda_stream)
        ctx.save_for_backward(logits, mu_c, ls_c, eps)
        ctx.ls_dtype = log_sigma.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        logits, mu, ls, eps = ctx.saved_tensors
        rows, g = logits.shape
        d = mu.shape[1]
        dout = dout.to(mu.dtype).contiguous()
        dlogits = torch.empty_like(logits)
        dmu = torch.empty_like(mu)
        dls = torch.empty_like(ls)
        n_ws = int(_lib.query("p2r_gmm_mix_workspace", rows, g, d))
        ws = ops.zeros_ws(n_ws, torch.float64, logits.device)
        with torch.cuda.device(logits.device):
            _lib.call("p2r_gmm_mix_grad", logits.data_ptr(), int(logits.dtype == torch.bfloat16), mu.data_ptr(),
                      int(mu.dtype == torch.float64), ls.data_ptr(), eps.data_ptr(), dout.data_ptr(), rows, g, d,
                      dlogits.data_ptr(), dmu.data_ptr(), dls.data_ptr(), ws.data_ptr(), n_ws,
                      torch.cuda.current_stream().cuda_stream)
        return dlogits, dmu, dls.to(ctx.ls_dtype), None


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

    def fused_point_prediction(self, x):
        """point_prediction(forward_rows(x)) for n_samples = 1 / 'mean' as one kernel; eps drawn exactly like sample()."""
        logits = self.pi.forward_rows(x)
        g, d = self.mu.shape
        eps = self.mu.data.new(logits.size(0), g, 1, d).normal_()        # same RNG consumption as mdn.py:44
        return _FusedGMMPredict.apply(logits, self.mu, self.log_sigma, eps)

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
        hp = self.mdn.hparams
        if fused_gmm_enabled() and hp.n_samples == 1 and hp.central_tendency == "mean" and x_rows.is_cuda:
            return self.mdn.fused_point_prediction(self.backbone.forward_rows(x_rows))
        return self.mdn.point_prediction(self.pi_rows(x_rows))

    def generate_rows(self, x_rows, multi_modes=False, n_samples=10):
        pi = self.pi_rows(x_rows)
        if multi_modes:
            return self.mdn.point_prediction(pi, n_samples=n_samples, sample_pi=True), pi
        return self.mdn.get_mean(pi), pi
