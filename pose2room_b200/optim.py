"""AdamW with the update rule of `torch.optim.AdamW` (what the reference's optimiser factory builds,
models/optimizers.py:90: amsgrad = False, maximize = False), applied to all parameters by a few launches of
`p2r_adamw_step` (csrc/optim_ops.cu) and capturable in a CUDA graph: the step count lives on the device.

torch's own capturable fused AdamW evaluates the bias corrections per element (0.16 ms of a 7.1 ms train step for the
1.44 M parameters of P2RNet); this one evaluates them once per thread.  State layout and `state_dict()` keys are torch's
(`step`, `exp_avg`, `exp_avg_sq` per parameter), so checkpoints written by either optimiser load into the other
(`CheckpointIO`, net_utils/utils.py:57-78).  CUDA dense parameters only -- there is no CPU path; parameters that are not
float32 or float64 are refused."""
import ctypes

import numpy as np
import torch

from . import _lib


class AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("pose2room_b200.optim.AdamW: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def _init_state(self, p, group_step):
        st = self.state[p]
        if "exp_avg" not in st:
            st["step"] = group_step
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not p.is_cuda or p.grad.is_sparse:
                    raise RuntimeError("pose2room_b200.optim.AdamW: CUDA dense parameters only (there is no CPU path)")
            # float32 and float64 contiguous parameters go through the kernel (P2RNet has ONE float64 parameter,
            # detection.gmm_heading.mdn.mu); anything else is refused
            for p in ps:
                if p.dtype not in (torch.float32, torch.float64) or p.grad.dtype != p.dtype or not p.is_contiguous():
                    raise RuntimeError("pose2room_b200.optim.AdamW: float32 / float64 contiguous parameters with gradients "
                                       "of their own dtype only")
            # one device step counter per group, shared by its parameters' state entries (torch keeps one per parameter;
            # after load_state_dict they are separate tensors again: re-share the first one)
            gstep = None
            for p in ps:
                st = self.state.get(p)
                if st and "step" in st:
                    gstep = st["step"]
                    break
            if gstep is None or not torch.is_tensor(gstep) or not gstep.is_cuda:
                v = float(gstep) if gstep is not None else 0.0
                gstep = torch.full((), v, dtype=torch.float32, device=ps[0].device)
            if gstep.dtype != torch.float32:
                gstep = gstep.float()
            b1, b2 = group["betas"]
            grads, ms, vs = [], [], []
            for p in ps:
                st = self._init_state(p, gstep)
                st["step"] = gstep
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                grads.append(g)
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
            n = len(ps)
            ptrs = lambda ts: np.array([t.data_ptr() for t in ts], dtype=np.uint64)
            a_p, a_g, a_m, a_v = ptrs(ps), ptrs(grads), ptrs(ms), ptrs(vs)
            numel = np.array([p.numel() if p.dtype == torch.float32 else -p.numel() for p in ps], dtype=np.int64)
            with torch.cuda.device(ps[0].device):
                _lib.call("p2r_adamw_step", n, a_p.ctypes.data_as(ctypes.c_void_p), a_g.ctypes.data_as(ctypes.c_void_p),
                          a_m.ctypes.data_as(ctypes.c_void_p), a_v.ctypes.data_as(ctypes.c_void_p),
                          numel.ctypes.data_as(ctypes.c_void_p), gstep.data_ptr(), float(group["lr"]), float(b1), float(b2),
                          float(group["eps"]), float(group["weight_decay"]), torch.cuda.current_stream().cuda_stream)
        return loss
