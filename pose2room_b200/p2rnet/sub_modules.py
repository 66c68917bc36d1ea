"""SingleConv: conv(1x1) [+ BatchNorm] [+ ReLU] with the reference's parameter names
(/root/reference/models/p2rnet/modules/sub_modules.py:88-113; submodules 'conv', 'batchnorm', 'ReLU' in the
order given by `order`, conv bias only when no norm layer is present), executed on channel-last rows by the
B200 kernels.  Only the orders the hot path uses are built: 'cbr' and 'c' with kernel_size 1."""
import torch
import torch.nn as nn

from .. import ops


class SingleConv(nn.Sequential):
    def __init__(self, in_channels, out_channels, kernel_size=1, order="cbr", num_groups=8, padding=0, ndim=1):
        super().__init__()
        assert kernel_size == 1 and padding == 0 and order in ("cbr", "c"), \
            "the P2RNet hot path only uses 1x1 'cbr' / 'c' blocks"
        conv_cls = {1: nn.Conv1d, 2: nn.Conv2d}[ndim]
        bn_cls = {1: nn.BatchNorm1d, 2: nn.BatchNorm2d}[ndim]
        self.order = order
        self.add_module("conv", conv_cls(in_channels, out_channels, 1, padding=0, bias="b" not in order))
        if "b" in order:
            self.add_module("batchnorm", bn_cls(out_channels))
        if "r" in order:
            self.add_module("ReLU", nn.ReLU(inplace=True))

    def forward_rows(self, x, act=None):
        """x [M, Cin] channel-last rows -> [M, Cout].  act: activation dtype of the layer's output when it differs from
        x's (throughput mode feeds float32 point coordinates into the first layer of an MLP and gets bf16 out)."""
        w = self.conv.weight.reshape(self.conv.out_channels, self.conv.in_channels)
        if act is not None and act != x.dtype:
            if act == torch.bfloat16 and self.order == "cbr" and self.conv.bias is None and \
                    ops.embed_l1_ok(x, self.conv.out_channels, self.conv.in_channels):
                return ops.embed_l1(x, w, self.batchnorm)       # conv + BN + ReLU, the pre-activation never stored
            if act == torch.bfloat16 and ops.smallk_mixed_ok(x, self.conv.out_channels, self.conv.in_channels):
                y, sums = ops.linear_coords(x, w, self.conv.bias, want_stats=True)
                if "b" in self.order:
                    return ops.batchnorm_act(y, self.batchnorm, relu="r" in self.order, sums=sums)
                return y
            x = x.to(act)
        if "b" in self.order and self.conv.out_channels == 64 and self.batchnorm.training:
            # 64-channel layers (pos_embed / sk_feat over all T*J points): the BatchNorm statistics come out of the GEMM
            # epilogue (an empty tensor when the kernel in use has no fused statistics: batchnorm_act then runs its pass)
            y, sums = ops.linear(x, w, self.conv.bias, want_stats=True)
            return ops.batchnorm_act(y, self.batchnorm, relu="r" in self.order, sums=sums)
        y = ops.linear(x, w, self.conv.bias)
        if "b" in self.order:
            y = ops.batchnorm_act(y, self.batchnorm, relu="r" in self.order)
        return y


def run_rows(seq, x, act=None):
    """act: activation dtype of the stack (see SingleConv.forward_rows); None = x's own."""
    for i, m in enumerate(seq):
        x = m.forward_rows(x, act) if (i == 0 and act is not None) else m.forward_rows(x)
    return x
