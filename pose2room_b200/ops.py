"""Autograd operators of the dense (channel-last) part of the P2RNet hot path, on libp2r_b200.so.

Every forward AND backward here is a hand-written kernel call through the C ABI; PyTorch only owns
the memory, the stream and the autograd graph.  Layout convention: activations are [rows, channels]
(rows = points / frames), i.e. the reference's (B, C, N) Conv1d / (B, C, T, V) Conv2d tensors transposed
to channel-last, which turns every 1x1 conv into a row-major GEMM and every BatchNorm into a column
statistic ("data layout").

Precision: activations may be float32 (parity mode: fp32-exact SIMT GEMM, fixed summation order) or
bfloat16 (throughput mode: tensor-core GEMM from gemm_sm100.cu, fp32 accumulate).  Parameters stay fp32.
"""
import os

import torch
from torch.autograd import Function

from . import _lib

_DT = {torch.float32: 0, torch.bfloat16: 1}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _splits_for(out_rows, out_cols, reduce_dim):
    tiles = ((out_rows + 63) // 64) * ((out_cols + 63) // 64)
    want = max(1, (148 * 4) // max(tiles, 1))
    return int(max(1, min(want, (reduce_dim + 255) // 256)))


# gemm backend hook: gemm_sm100 installs a tensor-core implementation for bf16 operands here
_TC_GEMM = {"fn": None}

# optional per-GEMM device timing (bench.py's roofline leg): when PROFILE["on"], every GEMM launched through
# linear() is bracketed by CUDA events on the launching stream and logged as (tag, M, N, K, ev0, ev1).
PROFILE = {"on": False, "log": []}


# ---- optional overlap of weight-gradient work with the rest of the backward pass -------------------------------
# The input-gradient chain (dX GEMM -> BatchNorm backward -> ...) is the critical path of the backward pass and is
# mostly HBM-bound; the weight-gradient GEMMs (dW = dz^T.x) are tensor-bound and nobody needs their result before
# the optimiser.  Inside `with overlap_weight_grads():` every dW / db is launched on a side stream and handed to
# autograd only when the context exits (one join), so the two kinds of work share the GPU.  The context must span
# forward AND backward: inside it the layers hand autograd a DETACHED weight (so the main backward never walks -- and
# frees -- the graph behind the weight, e.g. the einsum that builds W_eff) and remember the real one as the target of
# the deferred gradient.  Off by default: a plain forward + `loss.backward()` (what the reference's trainer does,
# models/training.py:25-43) behaves exactly as before.
DEFER = {"on": False, "stream": None, "items": [], "branch_streams": [], "keep": [], "ws": {}, "ws_off": {},
         "streams_by_device": {}, "device": None, "bn_counted": set(), "lazy_casts": [], "in_deferred_fn": False}
_WS_BYTES = 8 << 20
_DIRECT_LEAF_GRADS = os.environ.get("P2R_DIRECT_LEAF_GRADS", "1") != "0"


# ---- bf16 shadow copies of the weights (throughput mode) ---------------------------------------------------------
# Every tensor-core layer needs its fp32 weight in bf16.  Converting layer by layer costs ~45 tiny cast kernels per step;
# with shadows registered, ONE multi-tensor copy refreshes all of them when the step context is entered (after the
# previous optimiser step, inside the captured graph) and the layers look their weight up by address.
_SHADOW = {"params": [], "copies": [], "by_ptr": {}, "bn": [], "parents": {}}


def register_weight_shadows(module):
    """Create bf16 shadows for every float32 parameter of `module` with two or more dimensions (call once, after the
    module is on its device; harmless for modules that never run in bf16 mode)."""
    params = [p for p in module.parameters() if p.dtype == torch.float32 and p.dim() >= 2 and p.is_cuda]
    # a weight whose row count is not a multiple of 8 (TMA row pitch of the transposed roles: the 259-row vote head, the
    # 100 mixture weights) gets its shadow inside a zero-padded parent, so the GEMM wrappers need no per-step padding copy
    parents = [torch.zeros((-(-p.shape[0] // 8) * 8,) + tuple(p.shape[1:]), dtype=torch.bfloat16, device=p.device)
               for p in params]
    copies = [q[:p.shape[0]] for p, q in zip(params, parents)]
    _SHADOW["params"], _SHADOW["copies"] = params, copies
    _SHADOW["by_ptr"] = {p.data_ptr(): c for p, c in zip(params, copies)}
    _SHADOW["parents"] = {c.data_ptr(): q for c, q in zip(copies, parents) if q.shape[0] != c.shape[0]}
    # the BatchNorm layers' `num_batches_tracked += 1` (29 one-element kernels per step) become one multi-tensor add when
    # the step context is entered; batchnorm_act skips its own increment for these modules while the context is open
    _SHADOW["bn"] = [m for m in module.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)
                     and m.track_running_stats and m.num_batches_tracked is not None and m.num_batches_tracked.is_cuda]
    refresh_weight_shadows()


def clear_weight_shadows():
    _SHADOW["params"], _SHADOW["copies"], _SHADOW["by_ptr"], _SHADOW["bn"], _SHADOW["parents"] = [], [], {}, [], {}
    DEFER["bn_counted"] = set()


def refresh_weight_shadows():
    if _SHADOW["params"]:
        with torch.no_grad():
            torch._foreach_copy_(_SHADOW["copies"], _SHADOW["params"])


def bf16_weight(w):
    """`w` (a parameter or a reshaped view of one) in bf16: its registered shadow if there is one, else a conversion."""
    if w.dtype == torch.bfloat16:
        return w
    c = _SHADOW["by_ptr"].get(w.data_ptr()) if w.is_contiguous() else None
    if c is not None and c.numel() == w.numel():
        return c.view(w.shape)
    return w.to(torch.bfloat16)


def padded_shadow(w, rows):
    """The zero-padded [rows, k] parent of the registered bf16 shadow `w` ([n, k], n <= rows), or None."""
    if w.dtype != torch.bfloat16 or w.dim() != 2 or not w.is_contiguous():
        return None
    q = _SHADOW["parents"].get(w.data_ptr())
    if q is None or q.shape[0] != rows or q[0].numel() != w.shape[1] or w.shape[0] > rows:
        return None
    return q.view(rows, w.shape[1])


def zeros_ws(shape, dtype, device):
    """Zero-filled scratch for column sums (BatchNorm statistics, bias gradients).  Inside the multi-stream step
    context the ~100 small buffers of a step are carved from one workspace that is cleared by a single memset when the
    context is entered, instead of one fill kernel each; the result must not outlive the step.  Outside: torch.zeros."""
    if not DEFER["on"]:
        return torch.zeros(shape, dtype=dtype, device=device)
    n = 1
    for d in (shape if isinstance(shape, (tuple, list)) else (shape,)):
        n *= int(d)
    nbytes = (n * torch.empty(0, dtype=dtype).element_size() + 255) // 256 * 256
    key = str(device)
    off = DEFER["ws_off"].get(key, 0)
    if key not in DEFER["ws"] or off + nbytes > _WS_BYTES:
        return torch.zeros(shape, dtype=dtype, device=device)
    DEFER["ws_off"][key] = off + nbytes
    return DEFER["ws"][key][off:off + nbytes].view(dtype)[:n].view(shape)


class overlap_weight_grads:
    """Ordering contract (also INTEGRATION.md): parameters receive their `.grad` only when the context EXITS (or at an
    explicit `flush()`), so `optimizer.step()`, gradient clipping and any gradient all-reduce belong AFTER the `with`
    block; `parallel.allreduce_gradients` raises if it is called while deferred gradients are pending.  The model must
    live on the CUDA device that is current when the context is entered (the side / branch streams and the workspace
    are per device)."""

    def __enter__(self):
        dev = torch.cuda.current_device() if torch.cuda.is_available() else None
        per_dev = DEFER["streams_by_device"].setdefault(dev, {"stream": None, "branch_streams": []})
        if per_dev["stream"] is None:
            per_dev["stream"] = torch.cuda.Stream()
        DEFER["stream"], DEFER["branch_streams"], DEFER["device"] = per_dev["stream"], per_dev["branch_streams"], dev
        DEFER["items"] = []
        DEFER["on"] = True
        refresh_weight_shadows()               # (no-op unless register_weight_shadows was called)
        counted = [m for m in _SHADOW["bn"] if m.training]
        DEFER["bn_counted"] = set(id(m) for m in counted)
        if counted:
            torch._foreach_add_([m.num_batches_tracked for m in counted], 1)
        if torch.cuda.is_available():          # one memset for every small zero-initialised buffer of the step
            key = str(torch.device("cuda", torch.cuda.current_device()))
            if key not in DEFER["ws"]:
                DEFER["ws"][key] = torch.empty(_WS_BYTES, dtype=torch.uint8, device=key)
            DEFER["ws"][key].zero_()
            DEFER["ws_off"][key] = 0
        return self

    @staticmethod
    def flush():
        """Join the side stream and hand every deferred gradient to autograd now (parameters get their `.grad`)."""
        items, DEFER["items"] = DEFER["items"], []
        torch.cuda.current_stream().wait_stream(DEFER["stream"])      # join: every deferred dW / db is complete
        lazy, DEFER["lazy_casts"] = DEFER["lazy_casts"], []
        if lazy:
            with torch.no_grad():
                torch._foreach_copy_([d for d, _ in lazy], [s for _, s in lazy])     # float64 sums -> float32 gradients
        tensors, grads = [], []
        for t, g, _ in items:
            # A parameter that is itself the target takes its gradient directly: through autograd, AccumulateGrad would
            # COPY it (the Python list here holds a second reference, so the engine may not steal the tensor) -- one small
            # copy kernel per parameter and step.  Tensor hooks are honoured by taking the autograd route; hooks on the
            # accumulator node (DistributedDataParallel) are not visible from here: P2R_DIRECT_LEAF_GRADS=0 for those.
            if _DIRECT_LEAF_GRADS and t.is_leaf and t._backward_hooks is None and \
                    getattr(t, "_post_accumulate_grad_hooks", None) is None and g.shape == t.shape and \
                    g.dtype == t.dtype and g.device == t.device and g.is_contiguous():
                if t.grad is None:
                    t.grad = g
                else:
                    with torch.no_grad():
                        t.grad.add_(g)
            else:
                tensors.append(t)
                grads.append(g)
        if tensors:
            torch.autograd.backward(tensors, grads)                    # views / einsum backward -> leaf .grad

    def __exit__(self, exc_type, exc, tb):
        DEFER["on"] = False
        DEFER["keep"] = []
        DEFER["bn_counted"] = set()
        _COLSUM.clear()
        _GCN_PREBUILT.clear()
        for hook in _STEP_END_HOOKS:
            hook()
        if exc_type is None:
            self.flush()
        else:
            DEFER["items"], DEFER["lazy_casts"] = [], []
            torch.cuda.current_stream().wait_stream(DEFER["stream"])
        return False


def deferred_gradients_pending():
    return bool(DEFER["items"])


def parallel_branches(fns):
    """Run independent chains of small kernels (the four box heads: ~90 launches each, forward and again in the
    backward) side by side.  Inside `overlap_weight_grads()` every chain but the first gets its own stream, forked from
    and joined back to the current one (autograd replays each chain's backward on the stream of its forward); outside
    it the chains simply run one after the other.  Outputs are kept alive until the context exits, so memory handed
    out on a branch stream is never recycled while the main stream still reads it."""
    if not DEFER["on"] or PROFILE["on"] or len(fns) < 2:
        return [fn() for fn in fns]
    main = torch.cuda.current_stream()
    while len(DEFER["branch_streams"]) < len(fns):
        DEFER["branch_streams"].append(torch.cuda.Stream())
    outs = [None] * len(fns)
    for i in range(len(fns)):
        DEFER["branch_streams"][i].wait_stream(main)          # fork point: before anything of this group is enqueued
    # EVERY chain gets a branch stream, the first one too: in the backward autograd runs the chains in reverse issue order
    # and accumulates their gradients of the shared input on the main stream as they arrive -- a chain that lives on the
    # main stream is queued behind those accumulations, i.e. behind all the other chains (117 us in the step's timeline)
    for i, fn in enumerate(fns):                              # host issue order = list order (RNG consumption order)
        with torch.cuda.stream(DEFER["branch_streams"][i]):
            outs[i] = fn()
    for i in range(len(fns)):
        main.wait_stream(DEFER["branch_streams"][i])
    DEFER["keep"].append(outs)
    return outs


class fork_branch:
    """Run fn() on a forked stream NOW and hand its result over at `.join()`: for a kernel whose result is needed much
    later than its inputs are ready (the arc-length seed sampling: inputs = the raw joints, consumer = the gather at the
    end of the backbone; 40 us of a dependent chain otherwise).  Outside the multi-stream step it simply runs in line."""

    def __init__(self, fn):
        self.stream = None
        if not DEFER["on"] or PROFILE["on"]:
            self.out = fn()
            return
        per_dev = DEFER["streams_by_device"][DEFER["device"]]
        if per_dev.get("fork_stream") is None:
            per_dev["fork_stream"] = torch.cuda.Stream()
        self.stream = per_dev["fork_stream"]
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.out = fn()
        DEFER["keep"].append(self.out)      # allocated on the forked stream: alive until the step context exits

    def join(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        return self.out


def _defer(fn, targets, keep, inline=False):
    """Run fn() -> list of gradients on the side stream; `targets` are the autograd-connected tensors they belong to;
    `keep` are the operands the side-stream kernels read (kept alive until the join so the caching allocator cannot
    hand their memory to main-stream tensors in the meantime).  inline=True: run on the current stream (kernels that
    own whole SMs gain nothing from a second stream) but still hand the gradients over at the join."""
    DEFER["in_deferred_fn"] = True
    try:
        grads = _run_deferred(fn, inline)
    finally:
        DEFER["in_deferred_fn"] = False
    for t, g in zip(targets, grads):
        if t is not None and g is not None:
            DEFER["items"].append((t, g.to(t.dtype) if g.dtype != t.dtype else g, keep))


def _run_deferred(fn, inline):
    if inline:
        grads = fn()
    else:
        if torch.cuda.current_device() != DEFER["device"]:
            raise RuntimeError("pose2room_b200.ops.overlap_weight_grads was entered on cuda:%s but a layer runs on cuda:%s; "
                               "enter the context with the model's device current" % (DEFER["device"], torch.cuda.current_device()))
        side = DEFER["stream"]
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            grads = fn()
    return grads


class _Timed:
    def __init__(self, tag, m, n, k):
        self.rec = (tag, m, n, k)

    def __enter__(self):
        if PROFILE["on"]:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if PROFILE["on"]:
            self.e1.record()
            PROFILE["log"].append(self.rec + (self.e0, self.e1))
        return False


def sgemm(a, b, trans_a=False, trans_b=True, bias=None, relu=False, out_dtype=None, splits=1, out=None):
    """C = op(a) @ op(b) (+bias)(ReLU) with the SIMT fp32-accumulate kernel (see include/p2r_b200.h)."""
    assert a.is_cuda and b.is_cuda and a.dim() == 2 and b.dim() == 2
    a = a if a.is_contiguous() else a.contiguous()
    b = b if b.is_contiguous() else b.contiguous()
    m, k = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    n = b.shape[0] if trans_b else b.shape[1]
    kb = b.shape[1] if trans_b else b.shape[0]
    assert k == kb, (a.shape, b.shape, trans_a, trans_b)
    out_dtype = out_dtype or a.dtype
    if splits > 1:
        assert out_dtype == torch.float32 and bias is None and not relu
        c = torch.zeros(m, n, dtype=torch.float32, device=a.device) if out is None else out
    else:
        c = torch.empty(m, n, dtype=out_dtype, device=a.device) if out is None else out
    if bias is not None:
        bias = bias.float().contiguous()
    with torch.cuda.device(a.device):
        _lib.call("p2r_sgemm", m, n, k, a.data_ptr(), a.stride(0), int(trans_a), _DT[a.dtype], b.data_ptr(),
                  b.stride(0), int(trans_b), _DT[b.dtype], c.data_ptr(), c.stride(0), _DT[c.dtype], _ptr(bias),
                  int(relu), 0, int(splits), _stream())
    return c


def _to_float_at_join(s64):
    """float32 copy of a float64 sum that nobody reads before the step context joins its side stream: inside a deferred
    weight-gradient function the ~16 conversions of a step become ONE multi-tensor copy at the join (flush)."""
    if DEFER["on"] and DEFER.get("in_deferred_fn"):
        out = torch.empty(s64.shape, dtype=torch.float32, device=s64.device)
        DEFER["lazy_casts"].append((out, s64))
        return out
    return s64.float()


def _col_sum(dy, y=None, relu=False, at_join=False):
    """sum over rows of dz = relu ? dy*(y>0) : dy -> float32 [C] (bias gradients).  at_join=True: the caller only hands the
    result to the deferred-gradient list (see _to_float_at_join); it must not read it."""
    cast = _to_float_at_join if at_join else (lambda t: t.float())
    m, c = dy.shape
    vec = 8 if dy.dtype == torch.bfloat16 else 4
    if not relu and c > 256 and c % vec == 0:     # wide matrices (the 1600-column graph-conv output)
        s1 = zeros_ws(c, torch.float64, dy.device)
        with torch.cuda.device(dy.device):
            _lib.call("p2r_col_sum_wide", dy.data_ptr(), _DT[dy.dtype], m, c, s1.data_ptr(), _stream())
        return cast(s1)
    if (c > 256 and c % 256 or c <= 256 and 256 % c) and (c > 8192 or dy.dtype not in _DT or not dy.is_contiguous()):
        return (dy.float() if not relu else dy.float() * (y > 0)).sum(0)
    s1 = zeros_ws(c, torch.float64, dy.device)
    with torch.cuda.device(dy.device):
        _lib.call("p2r_col_bwd_stats", dy.data_ptr(), None, _ptr(y), _DT[dy.dtype], m, c, None, None, int(relu),
                  s1.data_ptr(), None, None, None, _stream())
    return cast(s1)


def _smallk_ok(x, n, k, relu):
    v = 8 if x.dtype == torch.bfloat16 else 4
    return k <= 4 and not relu and n % v == 0 and (256 * v) % n == 0 and 256 % (n // v) == 0


class _Linear(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu, targets=None, sparsity=None, want_stats=False):
        ctx.targets = targets      # (weight, bias) with their autograd history when the gradient is deferred
        ctx.sparsity = sparsity    # 64x64 block pattern of `weight` (gemm_sm100.BlockSparsity) or None
        x = x if x.is_contiguous() else x.contiguous()
        tc = _TC_GEMM["fn"]
        sums = None
        w_lp = None
        with _Timed("fwd", x.shape[0], weight.shape[0], x.shape[1]):
            if _smallk_ok(x, weight.shape[0], x.shape[1], relu):
                y = torch.empty(x.shape[0], weight.shape[0], dtype=x.dtype, device=x.device)
                wf = weight.float().contiguous()
                bf = bias.float().contiguous() if bias is not None else None
                with torch.cuda.device(x.device):
                    _lib.call("p2r_smallk_linear", x.data_ptr(), wf.data_ptr(), _ptr(bf), _DT[x.dtype], x.shape[0],
                              weight.shape[0], x.shape[1], y.data_ptr(), _stream())
            elif tc is not None and x.dtype == torch.bfloat16 and tc.supports(x.shape[0], weight.shape[0], x.shape[1]):
                w_lp = bf16_weight(weight)   # once per step (or the registered shadow): reused by dx
                if sparsity is not None or want_stats:
                    y, sums = tc.linear_fwd_ex(x, w_lp, bias, relu, sparsity, want_stats)
                else:
                    y = tc.linear_fwd(x, w_lp, bias, relu)
            else:
                y = sgemm(x, weight, False, True, bias, relu, out_dtype=x.dtype)
        ctx.save_for_backward(x, weight, y if relu else None)
        ctx.w_lp = w_lp
        ctx.relu = relu
        ctx.has_bias = bias is not None
        if want_stats:
            if sums is None:     # kernels without the fused epilogue statistics: the caller runs its own pass
                sums = torch.empty(0, dtype=torch.float64, device=x.device)
            ctx.mark_non_differentiable(sums)
            ctx.set_materialize_grads(False)      # no zero-filled "gradient" tensor for the statistics output
            return y, sums
        return y

    @staticmethod
    def backward(ctx, dy, _dsums=None):
        if dy is None:
            return (None,) * 7
        x, weight, y = ctx.saved_tensors
        sp = ctx.sparsity
        dy = dy if dy.is_contiguous() else dy.contiguous()
        if ctx.relu:
            dz = torch.empty_like(dy)
            with torch.cuda.device(dy.device):
                _lib.call("p2r_relu_bwd", dy.data_ptr(), y.data_ptr(), _DT[dy.dtype], dy.numel(), dz.data_ptr(), _stream())
        else:
            dz = dy
        m, n = dz.shape
        k = x.shape[1]
        dx = dw = db = None
        tc = _TC_GEMM["fn"]
        use_tc = tc is not None and x.dtype == torch.bfloat16 and tc.supports(m, n, k)
        if ctx.needs_input_grad[0]:
            with _Timed("dx", m, n, k):
                w_dx = ctx.w_lp if (use_tc and ctx.w_lp is not None) else weight
                dx = tc.linear_dx(dz, w_dx, sp) if use_tc else sgemm(dz, weight, False, False, out_dtype=x.dtype)
        if ctx.targets is not None:
            tw, tb = ctx.targets
            need_w = tw is not None and tw.requires_grad
            need_b = tb is not None and tb.requires_grad

            def weight_grads():
                gw = gb = None
                if need_w:
                    if _smallk_ok(x, n, k, ctx.relu):
                        gw = torch.zeros(n, k, dtype=torch.float32, device=x.device)
                        with torch.cuda.device(x.device):
                            _lib.call("p2r_smallk_dw", dz.data_ptr(), x.data_ptr(), _DT[x.dtype], m, n, k,
                                      gw.data_ptr(), _stream())
                    elif use_tc:
                        gw = tc.linear_dw(dz, x, sp)
                    else:
                        gw = sgemm(dz, x, True, False, out_dtype=torch.float32, splits=_splits_for(n, k, m))
                if need_b:
                    gb = _col_sum(dz, at_join=True)
                return [gw, gb]
            _defer(weight_grads, [tw if need_w else None, tb if need_b else None], (dz, x))
            return dx, None, None, None, None, None, None
        if ctx.needs_input_grad[1]:
            with _Timed("dw", m, n, k):
                if _smallk_ok(x, n, k, ctx.relu):
                    dw = torch.zeros(n, k, dtype=torch.float32, device=x.device)
                    with torch.cuda.device(x.device):
                        _lib.call("p2r_smallk_dw", dz.data_ptr(), x.data_ptr(), _DT[x.dtype], m, n, k, dw.data_ptr(),
                                  _stream())
                elif use_tc:
                    dw = tc.linear_dw(dz, x, sp)
                else:
                    dw = sgemm(dz, x, True, False, out_dtype=torch.float32, splits=_splits_for(n, k, m))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _col_sum(dz)
        return dx, dw, db, None, None, None, None


class _SmallKMixed(Function):
    """First layer of a point MLP in throughput mode: x [M, K <= 4] float32 COORDINATES (never rounded to bf16), bf16
    output; dW from the float32 x.  The input carries no gradient (it is the pose data)."""

    @staticmethod
    def forward(ctx, x, weight, bias, targets, want_stats):
        x = x if x.is_contiguous() else x.contiguous()
        m, k = x.shape
        n = weight.shape[0]
        y = torch.empty(m, n, dtype=torch.bfloat16, device=x.device)
        wf = weight.float().contiguous()
        bf = bias.float().contiguous() if bias is not None else None
        with torch.cuda.device(x.device):
            _lib.call("p2r_smallk_linear_mixed", x.data_ptr(), wf.data_ptr(), _ptr(bf), m, n, k, y.data_ptr(), _stream())
        ctx.save_for_backward(x)
        ctx.targets, ctx.dims, ctx.has_bias = targets, (m, n, k), bias is not None
        if want_stats:
            sums = torch.empty(0, dtype=torch.float64, device=x.device)      # no fused statistics: the caller runs its pass
            ctx.mark_non_differentiable(sums)
            ctx.set_materialize_grads(False)
            return y, sums
        return y

    @staticmethod
    def backward(ctx, dy, _dsums=None):
        if dy is None:
            return (None,) * 5
        if ctx.needs_input_grad[0]:
            raise RuntimeError("pose2room_b200: the float32-coordinate first layer has no input gradient")
        (x,) = ctx.saved_tensors
        m, n, k = ctx.dims
        dz = dy if dy.is_contiguous() else dy.contiguous()
        dz = dz if dz.dtype == torch.bfloat16 else dz.to(torch.bfloat16)

        def weight_grads(need_w=True, need_b=ctx.has_bias):
            gw = gb = None
            if need_w:
                gw = torch.zeros(n, k, dtype=torch.float32, device=x.device)
                with torch.cuda.device(x.device):
                    _lib.call("p2r_smallk_dw_mixed", dz.data_ptr(), x.data_ptr(), m, n, k, gw.data_ptr(), _stream())
            if need_b:
                gb = _col_sum(dz, at_join=True)
            return [gw, gb]
        if ctx.targets is not None:
            tw, tb = ctx.targets
            need_w = tw is not None and tw.requires_grad
            need_b = tb is not None and tb.requires_grad
            _defer(lambda: weight_grads(need_w, need_b), [tw if need_w else None, tb if need_b else None], (dz, x))
            return None, None, None, None, None
        gw, gb = weight_grads(ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2])
        return None, gw, gb, None, None


class _EmbedL1(Function):
    """Conv (K <= 4 -> N) + BatchNorm + ReLU on float32 coordinates -> bf16, the pre-activation never stored
    (csrc/embed_ops.cu): statistics from the moments of x, z recomputed in the backward, dW accumulated in the pass that
    computes dz.  The input carries no gradient (it is the pose data)."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, running_mean, running_var, training, momentum, eps):
        x = x if x.is_contiguous() else x.contiguous()
        m, k = x.shape
        n = weight.shape[0]
        dev = x.device
        wf = weight.float().contiguous()
        stats = torch.empty(4, n, dtype=torch.float32, device=dev)       # mean, rstd, scale, shift
        y = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
        with torch.cuda.device(dev):
            if training:
                mom = zeros_ws(k + k * k, torch.float64, dev)
                _lib.call("p2r_coord_moments", x.data_ptr(), m, k, mom.data_ptr(), _stream())
                _lib.call("p2r_embed_l1_finalize", mom.data_ptr(), m, k, wf.data_ptr(), n, _ptr(gamma), _ptr(beta), float(eps),
                          float(momentum), _ptr(running_mean), _ptr(running_var), stats[0].data_ptr(), stats[1].data_ptr(),
                          stats[2].data_ptr(), stats[3].data_ptr(), _stream())
            else:
                rstd = torch.rsqrt(running_var + eps)
                g_ = gamma if gamma is not None else torch.ones_like(rstd)
                stats[0] = running_mean
                stats[1] = rstd
                stats[2] = g_ * rstd
                stats[3] = (beta if beta is not None else 0.0) - running_mean * g_ * rstd
            _lib.call("p2r_embed_l1_fwd", x.data_ptr(), wf.data_ptr(), stats[2].data_ptr(), stats[3].data_ptr(), m, n, k,
                      y.data_ptr(), _stream())
        ctx.save_for_backward(x, wf, stats)
        ctx.training, ctx.w_shape = training, weight.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        if ctx.needs_input_grad[0]:
            raise RuntimeError("pose2room_b200: the float32-coordinate first layer has no input gradient")
        x, wf, stats = ctx.saved_tensors
        m, k = x.shape
        n = wf.shape[0]
        dev = x.device
        dy = dy if dy.is_contiguous() else dy.contiguous()
        dy = dy if dy.dtype == torch.bfloat16 else dy.to(torch.bfloat16)
        sums = zeros_ws((2, n), torch.float64, dev)
        dw = torch.zeros(n, k, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("p2r_embed_l1_bwd_stats", dy.data_ptr(), x.data_ptr(), wf.data_ptr(), stats[0].data_ptr(),
                      stats[1].data_ptr(), stats[2].data_ptr(), stats[3].data_ptr(), m, n, k, sums[0].data_ptr(),
                      sums[1].data_ptr(), _stream())
            _lib.call("p2r_embed_l1_bwd_dw", dy.data_ptr(), x.data_ptr(), wf.data_ptr(), stats[0].data_ptr(),
                      stats[1].data_ptr(), stats[2].data_ptr(), stats[3].data_ptr(),
                      sums[0].data_ptr() if ctx.training else None, sums[1].data_ptr() if ctx.training else None, m, n, k,
                      dw.data_ptr(), _stream())
        sums_f = sums.float()
        return (None, dw.reshape(ctx.w_shape) if ctx.needs_input_grad[1] else None,
                sums_f[1] if ctx.needs_input_grad[2] else None, sums_f[0] if ctx.needs_input_grad[3] else None,
                None, None, None, None, None)


def embed_l1_ok(x, n, k):
    return (x.is_cuda and x.dtype == torch.float32 and k <= 4 and n % 8 == 0 and 2048 % n == 0 and 256 % (n // 8) == 0
            and os.environ.get("P2R_FUSED_EMBED_L1", "1") != "0")


def embed_l1(x, weight, bn):
    """relu(bn(x @ weight^T)) for float32 coordinates x [M, K <= 4] -> bf16 [M, N] (see _EmbedL1); `bn` an nn.BatchNorm
    module (its mode, parameters, running statistics, momentum and eps are honoured like in batchnorm_act)."""
    training = bn.training or not bn.track_running_stats
    if training and bn.track_running_stats and bn.num_batches_tracked is not None and \
            id(bn) not in DEFER.get("bn_counted", ()):
        bn.num_batches_tracked += 1
    if bn.momentum is None:
        momentum = 1.0 / float(bn.num_batches_tracked) if (training and bn.track_running_stats) else 0.0
    else:
        momentum = bn.momentum
    return _EmbedL1.apply(x, weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, training, momentum, bn.eps)


def smallk_mixed_ok(x, n, k):
    return x.is_cuda and x.dtype == torch.float32 and k <= 4 and n % 8 == 0 and 2048 % n == 0 and 256 % (n // 8) == 0


def linear_coords(x, weight, bias=None, want_stats=False):
    """y bf16 [M, N] = x f32 [M, K <= 4] @ weight^T (+ bias): see _SmallKMixed.  Never deferred: the input carries no
    gradient, so with a detached weight the node would drop out of the autograd graph (same rule as linear())."""
    return _SmallKMixed.apply(x, weight, bias, None, want_stats)


def linear(x, weight, bias=None, relu=False, sparsity=None, want_stats=False):
    """y[M,N] = x[M,K] @ weight[N,K]^T (+bias)(ReLU): a 1x1 Conv1d/Conv2d in channel-last form.
    sparsity: optional 64x64 block pattern of `weight` (zero blocks are skipped by the tensor-core kernel).
    want_stats: also return the per-channel (column % 64) sum / sum of squares of y from the GEMM epilogue, shaped
    [copies, 2, 64] float64 (an EMPTY tensor when the kernel in use cannot produce them) -> returns (y, sums)."""
    # (a layer whose input carries no gradient would drop out of the graph with a detached weight: keep it on the plain path)
    if DEFER["on"] and not PROFILE["on"] and x.shape[0] >= 4096 and torch.is_grad_enabled() and x.requires_grad:
        return _Linear.apply(x, weight.detach(), bias.detach() if bias is not None else None, relu, (weight, bias),
                             sparsity, want_stats)
    return _Linear.apply(x, weight, bias, relu, None, sparsity, want_stats)


# column sums of a gradient tensor produced as a by-product by the BatchNorm backward that wrote it (keyed by data_ptr,
# consumed by the layer before it in the same backward pass: the bias gradient of the graph convolution)
_COLSUM = {}
# OFF by default: the kernel is correct in isolation and in an eager step (tests), but bench.py's
# back-to-back warm-up steps did not finish with it enabled (two runs, cause not established) -- see DESIGN.md section 3.
_FUSED_COLSUM = os.environ.get("P2R_FUSED_COLSUM", "0") != "0"
_FUSED_COLSUM1 = os.environ.get("P2R_FUSED_COLSUM1", "1") != "0"


_GCN_PREBUILT = {}
_STEP_END_HOOKS = []      # callables run when the multi-stream step context exits (caches of per-step prebuilt operands)


def _gcn_build(conv_w, conv_b, a_eff):
    """(cw, cb, ae, W_eff, W_eff^T, b_eff) of one graph convolution: W_eff[(w,co),(v,ci)] = sum_k W_k[co,ci] A_k[v,w]
    (ref: ConvTemporalGraphical.forward, stgcn_layers.py:58-67, as one matrix), both orientations in bf16, one launch."""
    k, v = a_eff.shape[0], a_eff.shape[1]
    co, ci = conv_w.shape[0] // k, conv_w.shape[1]
    dev = conv_w.device
    cw = conv_w.reshape(k * co, ci).float().contiguous()
    cb = conv_b.float().contiguous() if conv_b is not None else None
    ae = a_eff.float().contiguous()
    w_eff = torch.empty(v * co, v * ci, dtype=torch.bfloat16, device=dev)
    w_eff_t = torch.empty(v * ci, v * co, dtype=torch.bfloat16, device=dev)
    b_eff = torch.empty(v * co, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("p2r_gcn_build_weight", cw.data_ptr(), _ptr(cb), ae.data_ptr(), k, v, co, ci, w_eff.data_ptr(),
                  w_eff_t.data_ptr(), b_eff.data_ptr(), _stream())
    return cw, cb, ae, w_eff, w_eff_t, b_eff


def graph_conv_prebuild(specs):
    """Build the effective weights of several graph convolutions NOW, on the current stream (a forked branch: they only
    depend on parameters, so the six builds of the backbone -- 20 us each, serial in front of their GEMMs otherwise -- run
    beside the embedding layers).  specs: [(conv_weight, conv_bias, a_eff)]; the next graph_conv() call with the same
    conv_weight / a_eff tensors takes the prebuilt matrices.  Entries not consumed are dropped when the step ends."""
    out = []
    _GCN_PREBUILT.clear()          # (whatever an earlier forward pass did not consume)
    with torch.no_grad():
        for conv_w, conv_b, a_eff in specs:
            built = _gcn_build(conv_w.detach(), conv_b.detach() if conv_b is not None else None, a_eff.detach())
            _GCN_PREBUILT[(conv_w.data_ptr(), a_eff.data_ptr())] = built
            out.append(built)
    if DEFER["on"]:
        DEFER["keep"].append(out)  # allocated on the forked stream, read on the main one: alive until the step ends
    return out


def graph_conv_prebuild_ok(rows):
    """Will graph_conv() take the tensor-core path for `rows` bf16 frames of 64-channel joints (graph_conv_available)?"""
    return _TC_GEMM["fn"] is not None and rows >= 128


class _GraphConv(Function):
    """The graph convolution of st_gcn_block as ONE tensor-core GEMM (bf16 mode): builds W_eff / W_eff^T / b_eff from
    the conv parameters and A = adjacency * importance with one kernel, runs the block-sparse GEMM (statistics of the
    following BatchNorm from its epilogue), and in the backward folds dW_eff back onto conv weight, conv bias and A with
    one kernel.  ref: ConvTemporalGraphical.forward, stgcn_layers.py:58-67."""

    @staticmethod
    def forward(ctx, x, conv_w, conv_b, a_eff, targets, sparsity, want_stats, with_alias=False):
        tc = _TC_GEMM["fn"]
        k, v = a_eff.shape[0], a_eff.shape[1]
        co, ci = conv_w.shape[0] // k, conv_w.shape[1]
        x = x if x.is_contiguous() else x.contiguous()
        dev = x.device
        built = _GCN_PREBUILT.pop((conv_w.data_ptr(), a_eff.data_ptr()), None)     # graph_conv_prebuild ran ahead
        cw, cb, ae, w_eff, w_eff_t, b_eff = built if built is not None else _gcn_build(conv_w, conv_b, a_eff)
        with _Timed("fwd", x.shape[0], v * co, v * ci):
            y, sums = tc.linear_fwd_ex(x, w_eff, b_eff, False, sparsity, want_stats)
        ctx.save_for_backward(x, cw, cb, ae, w_eff_t)
        ctx.targets, ctx.sparsity, ctx.dims = targets, sparsity, (k, v, co, ci)
        ctx.w_shape = conv_w.shape
        if sums is None:
            sums = torch.empty(0, dtype=torch.float64, device=dev)
        ctx.mark_non_differentiable(sums)
        ctx.set_materialize_grads(False)
        if with_alias:
            # third output: x itself (autograd hands it out as an alias) for the block's residual branch.  Its gradient
            # then arrives HERE, next to dy, and the input-gradient GEMM adds its product onto it in place (reduce-add
            # stores) -- instead of autograd summing the two 105 MB gradients of x with a separate elementwise kernel.
            return y, sums, x
        return y, sums

    @staticmethod
    def backward(ctx, dy, _dsums=None, d_alias=None):
        if dy is None:
            return (d_alias,) + (None,) * 7
        x, cw, cb, ae, w_eff_t = ctx.saved_tensors
        k, v, co, ci = ctx.dims
        sp = ctx.sparsity
        tc = _TC_GEMM["fn"]
        dy = dy if dy.is_contiguous() else dy.contiguous()
        fused_cs = _COLSUM.pop(dy.data_ptr(), None)              # [V, Co] sums of dy from the BatchNorm backward, if any
        if fused_cs is not None and fused_cs.numel() != v * co:
            fused_cs = None
        dx = None
        if ctx.needs_input_grad[0]:
            acc = None
            if d_alias is not None:
                if d_alias.dtype == torch.bfloat16 and d_alias.is_contiguous() and d_alias.shape == x.shape:
                    acc = d_alias          # fresh tensor written by the BatchNorm backward of the block: ours to add onto
            with _Timed("dx", dy.shape[0], v * co, v * ci):
                dx = tc.linear_dx_pretransposed(dy, w_eff_t, sp, accumulate_into=acc)
            if d_alias is not None and acc is None:
                dx = dx + d_alias.to(dx.dtype)
        elif d_alias is not None:
            dx = d_alias

        def weight_grads():
            with _Timed("gcn_dw", dy.shape[0], v * co, v * ci):
                # fp32 [V*Co, V*Ci]; the structurally-zero tiles are neither computed nor zero-filled: the two reducers below
                # only read blocks (w, v) with A[k, v, w] != 0 for some k (csrc/graph_conv.cu)
                dw_eff = tc.linear_dw(dy, x, sp, zero_skipped=False)
            db_eff = None
            if cb is not None:
                db_eff = fused_cs.reshape(-1).float() if fused_cs is not None else _col_sum(dy)
            d_w = torch.empty(k * co, ci, dtype=torch.float32, device=dy.device)
            d_b = torch.empty(k * co, dtype=torch.float32, device=dy.device) if cb is not None else None
            d_a = torch.empty(k, v, v, dtype=torch.float32, device=dy.device)
            with torch.cuda.device(dy.device):
                _lib.call("p2r_gcn_reduce_weight_grad", dw_eff.data_ptr(), _ptr(db_eff), cw.data_ptr(), _ptr(cb),
                          ae.data_ptr(), k, v, co, ci, d_w.data_ptr(), _ptr(d_b), d_a.data_ptr(), _stream())
            return [d_w.reshape(ctx.w_shape), d_b, d_a]

        if ctx.targets is not None:
            _defer(weight_grads, [t if (t is not None and t.requires_grad) else None for t in ctx.targets], (dy, x),
                   inline=getattr(tc, "DW_INLINE", False))
            return dx, None, None, None, None, None, None, None
        gw, gb, ga = weight_grads()
        return dx, gw, gb, ga, None, None, None, None


_FUSED_RESADD = os.environ.get("P2R_FUSED_RESADD", "1") != "0"


def graph_conv(x, conv_weight, conv_bias, a_eff, sparsity=None, want_stats=True, residual_alias=False):
    """x [M, V*Ci] frames (channel-last joints) -> (y [M, V*Co], sums): the st_gcn graph convolution with conv weight
    (K*Co, Ci, 1, 1), conv bias (K*Co,), A_eff (K, V, V).  Tensor-core (bf16) path only -- callers check
    `graph_conv_available(x)` and use the generic linear() otherwise.
    residual_alias: also return x (as an autograd alias) for the block's residual branch -> (y, sums, x_alias); the
    gradient of the alias is then folded into the input-gradient GEMM (see _GraphConv.forward)."""
    alias = bool(residual_alias and _FUSED_RESADD)
    if DEFER["on"] and not PROFILE["on"] and torch.is_grad_enabled() and x.requires_grad:
        out = _GraphConv.apply(x, conv_weight.detach(), conv_bias.detach() if conv_bias is not None else None,
                               a_eff.detach(), (conv_weight, conv_bias, a_eff), sparsity, want_stats, alias)
    else:
        out = _GraphConv.apply(x, conv_weight, conv_bias, a_eff, None, sparsity, want_stats, alias)
    if residual_alias and not alias:
        return out[0], out[1], x
    return out


def graph_conv_available(x, co, ci):
    tc = _TC_GEMM["fn"]
    return tc is not None and x.dtype == torch.bfloat16 and co == 64 and ci == 64 and x.shape[0] >= 128


class _BatchNormAct(Function):
    """Training- or eval-mode BatchNorm over the rows of x[M,C] (+residual)(+ReLU).
    ref: nn.BatchNorm1d/2d inside SingleConv 'cbr' (sub_modules.py:64-70) and st_gcn_block.tcn
    (stgcn_layers.py:402-414) followed by `+ res` and ReLU (:436-438)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, residual, training, momentum, eps, relu, sums=None,
                colsum_period=0):
        ctx.colsum_period = colsum_period
        x = x if x.is_contiguous() else x.contiguous()
        m, c = x.shape
        dev = x.device
        dt = _DT[x.dtype]
        stats = torch.empty(4, c, dtype=torch.float32, device=dev)  # mean, rstd, scale, shift
        with torch.cuda.device(dev):
            if training:
                if sums is None or sums.numel() == 0:     # no statistics from the producing GEMM's epilogue
                    sums = zeros_ws((1, 2, c), torch.float64, dev)
                    _lib.call("p2r_col_stats", x.data_ptr(), dt, m, c, sums[0, 0].data_ptr(), sums[0, 1].data_ptr(), _stream())
                assert sums.dim() == 3 and sums.shape[1] == 2 and sums.shape[2] == c and sums.is_contiguous()
                _lib.call("p2r_bn_finalize", c, m, sums[0, 0].data_ptr(), sums[0, 1].data_ptr(), sums.shape[0], 2 * c,
                          _ptr(gamma), _ptr(beta),
                          float(eps), float(momentum), _ptr(running_mean), _ptr(running_var), stats[0].data_ptr(),
                          stats[1].data_ptr(), stats[2].data_ptr(), stats[3].data_ptr(), _stream())
            else:
                rstd = torch.rsqrt(running_var + eps)
                g_ = gamma if gamma is not None else torch.ones_like(rstd)        # affine=False
                stats[0] = running_mean
                stats[1] = rstd
                stats[2] = g_ * rstd
                stats[3] = (beta if beta is not None else 0.0) - running_mean * g_ * rstd
            y = torch.empty_like(x)
            if residual is not None:
                residual = residual if residual.is_contiguous() else residual.contiguous()
            # ReLU mask for the backward: recomputed from x (mode 2) unless a residual was added; then either the output
            # itself (mode 1) or -- with the streaming kernels -- a 1-bit-per-element mask written by this pass (mode 3):
            # the two backward passes then read M*8 bytes instead of the whole M*128-byte output
            mask = None
            if relu and residual is not None and torch.is_grad_enabled() and c % 8 == 0 and \
                    _lib.query("p2r_stream_bn_supported", dt, m, c) == 1:
                mask = torch.empty(m, c // 8, dtype=torch.uint8, device=dev)
            _lib.call("p2r_affine_act", x.data_ptr(), dt, m, c, stats[2].data_ptr(), stats[3].data_ptr(),
                      _ptr(residual), int(relu), y.data_ptr(), _ptr(mask), _stream())
        ctx.relu_mode = 0 if not relu else ((3 if mask is not None else 1) if residual is not None else 2)
        ctx.save_for_backward(x, mask if ctx.relu_mode == 3 else (y if ctx.relu_mode == 1 else None), stats)
        ctx.training, ctx.has_res = training, residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, stats = ctx.saved_tensors
        dy = dy if dy.is_contiguous() else dy.contiguous()
        m, c = x.shape
        dev = x.device
        dt = _DT[x.dtype]
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if ctx.has_res else None
        sums = zeros_ws((2, c), torch.float64, dev)
        period = ctx.colsum_period
        cs = None
        # period 1 (plain column sums: the bias gradient of the conv right in front of this BatchNorm) costs the pass
        # nothing -- per-thread register sums -- and is on by default; longer periods go through shared-memory atomics
        if 0 < period <= 32 and DEFER["on"] and (_FUSED_COLSUM or (period == 1 and _FUSED_COLSUM1)) and \
                _lib.query("p2r_stream_bn_supported", dt, m, c) == 1:
            cs = zeros_ws((period, c), torch.float64, dev)      # per-(row % period, channel) sums of dx, by-product
        with torch.cuda.device(dev):
            _lib.call("p2r_col_bwd_stats", dy.data_ptr(), x.data_ptr(), _ptr(y), dt, m, c, stats[0].data_ptr(),
                      stats[1].data_ptr(), ctx.relu_mode, sums[0].data_ptr(), sums[1].data_ptr(), stats[2].data_ptr(),
                      stats[3].data_ptr(), _stream())
            want_f = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
            sums_f = torch.empty(2, c, dtype=torch.float32, device=dev) if want_f else None     # written by the same launch
            _lib.call("p2r_bn_bwd_apply_ex", dy.data_ptr(), x.data_ptr(), _ptr(y), dt, m, c, stats[0].data_ptr(),
                      stats[1].data_ptr(), stats[2].data_ptr(), sums[0].data_ptr() if ctx.training else None,
                      sums[1].data_ptr() if ctx.training else None, ctx.relu_mode, dx.data_ptr(), _ptr(dres),
                      stats[3].data_ptr(), _ptr(cs), int(period if cs is not None else 0),
                      sums.data_ptr() if want_f else None, _ptr(sums_f), _stream())
        if cs is not None:
            _COLSUM[dx.data_ptr()] = cs
        dgamma = sums_f[1] if ctx.needs_input_grad[1] else None
        dbeta = sums_f[0] if ctx.needs_input_grad[2] else None
        return dx, dgamma, dbeta, None, None, dres, None, None, None, None, None, None


def batchnorm_act(x, bn, relu=False, residual=None, sums=None, colsum_period=0):
    """Apply nn.BatchNorm{1,2}d module `bn` (its parameters / running stats / momentum / eps / mode) to the
    channel-last matrix x[M,C], optionally adding `residual` and a ReLU -- one fused elementwise pass.
    sums: per-channel [copies, 2, C] float64 sum / sum of squares of x already produced by the GEMM that wrote x.
    colsum_period: the backward also leaves sum_rows dx per (row % period, channel) for the layer that produced x (the
    graph convolution's bias gradient: rows cycle through the joints) -- streaming kernels, multi-stream step only."""
    c = x.shape[1]
    if not (256 % c == 0 or c % 256 == 0):
        raise RuntimeError("pose2room_b200.ops.batchnorm_act: %d channels -- the column-statistics kernels take widths that "
                           "divide 256 or are multiples of 256 (every BatchNorm of the P2RNet path does)" % c)
    training = bn.training or not bn.track_running_stats
    if training and bn.track_running_stats and bn.num_batches_tracked is not None and \
            id(bn) not in DEFER.get("bn_counted", ()):
        bn.num_batches_tracked += 1
    if bn.momentum is None:      # nn.BatchNorm: cumulative moving average (a host read; no module of the path uses it)
        momentum = 1.0 / float(bn.num_batches_tracked) if (training and bn.track_running_stats) else 0.0
    else:
        momentum = bn.momentum
    return _BatchNormAct.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, residual, training,
                               momentum, bn.eps, relu, sums, colsum_period)


class _TemporalUnfold(Function):
    @staticmethod
    def forward(ctx, x, kt):
        b, t, v, c = x.shape
        x = x if x.is_contiguous() else x.contiguous()
        col = torch.empty(b * t * v, kt * c, dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("p2r_temporal_unfold", x.data_ptr(), _DT[x.dtype], b, t, v, c, kt, col.data_ptr(), _stream())
        ctx.shape = (b, t, v, c, kt)
        return col

    @staticmethod
    def backward(ctx, dcol):
        b, t, v, c, kt = ctx.shape
        dcol = dcol if dcol.is_contiguous() else dcol.contiguous()
        dx = torch.empty(b, t, v, c, dtype=dcol.dtype, device=dcol.device)
        with torch.cuda.device(dcol.device):
            _lib.call("p2r_temporal_fold", dcol.data_ptr(), _DT[dcol.dtype], b, t, v, c, kt, dx.data_ptr(), _stream())
        return dx, None


def temporal_conv(x, weight, bias, want_stats=False):
    """(KT x 1) temporal convolution, zero padding (KT-1)/2, stride 1 (stgcn_layers.py:405-411).
    x [B,T,V,Ci] channel-last, weight (Co,Ci,KT,1) as stored by nn.Conv2d -> [B*T*V, Co].
    want_stats: -> (y, sums) like linear()."""
    co, ci, kt, _ = weight.shape
    tc = _TC_GEMM["fn"]
    if tc is not None and x.dtype == torch.bfloat16 and tc.supports_tconv(x.shape, co):
        return tc.temporal_conv(x, weight, bias, want_stats)
    col = _TemporalUnfold.apply(x, kt)
    w2 = weight[:, :, :, 0].permute(0, 2, 1).reshape(co, kt * ci)  # column = dt*Ci + ci
    return linear(col, w2, bias, want_stats=want_stats)


class _GroupRows(Function):
    @staticmethod
    def forward(ctx, feats, idx):
        b, n, c = feats.shape
        _, p, s = idx.shape
        feats = feats if feats.is_contiguous() else feats.contiguous()
        out = torch.empty(b, p, s, c, dtype=feats.dtype, device=feats.device)
        with torch.cuda.device(feats.device):
            _lib.call("p2r_group_rows", feats.data_ptr(), _DT[feats.dtype], idx.data_ptr(), b, n, c, p, s,
                      out.data_ptr(), _stream())
        ctx.save_for_backward(idx)
        ctx.shape = (b, n, c, p, s)
        ctx.dtype = feats.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        b, n, c, p, s = ctx.shape
        g = g if g.is_contiguous() else g.contiguous()
        d = torch.zeros(b, n, c, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            _lib.call("p2r_group_rows_grad", g.data_ptr(), _DT[g.dtype], idx.data_ptr(), b, n, c, p, s, d.data_ptr(),
                      _stream())
        return d.to(ctx.dtype), None


def group_rows(feats, idx):
    """feats [B,N,C], idx [B,P,S] int32 -> [B,P,S,C]: channel-last twin of grouping_operation."""
    return _GroupRows.apply(feats, idx.contiguous())


class _SelectRows(Function):
    """out[b][p] = feats[b][idx[b][p]] (the seed-frame pick before conv_joint, stgcn.py:136-139).  The adjoint is written
    destination-major by one launch (csrc/dense_ops.cu select_rows_grad_kernel): no zero-fill of [B,N,C], no atomics."""

    @staticmethod
    def forward(ctx, feats, idx):
        b, n, c = feats.shape
        p = idx.shape[1]
        feats = feats if feats.is_contiguous() else feats.contiguous()
        idx32 = idx.to(torch.int32).contiguous()
        out = torch.empty(b, p, c, dtype=feats.dtype, device=feats.device)
        with torch.cuda.device(feats.device):
            _lib.call("p2r_group_rows", feats.data_ptr(), _DT[feats.dtype], idx32.data_ptr(), b, n, c, p, 1,
                      out.data_ptr(), _stream())
        ctx.save_for_backward(idx32)
        ctx.shape = (b, n, c, p)
        return out

    @staticmethod
    def backward(ctx, g):
        (idx32,) = ctx.saved_tensors
        b, n, c, p = ctx.shape
        g = g if g.is_contiguous() else g.contiguous()
        d = torch.empty(b, n, c, dtype=g.dtype, device=g.device)
        with torch.cuda.device(g.device):
            _lib.call("p2r_select_rows_grad", g.data_ptr(), _DT[g.dtype], idx32.data_ptr(), b, n, c, p, d.data_ptr(),
                      _stream())
        return d, None


def select_rows(feats, idx):
    """feats [B,N,C] (float32 / bfloat16), idx [B,P] integer -> [B,P,C]; torch.gather along dim 1 with whole rows."""
    return _SelectRows.apply(feats, idx)


class _MaxPoolRows(Function):
    @staticmethod
    def forward(ctx, x):
        r, s, c = x.shape
        x = x if x.is_contiguous() else x.contiguous()
        out = torch.empty(r, c, dtype=x.dtype, device=x.device)
        arg = torch.empty(r, c, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("p2r_maxpool_rows", x.data_ptr(), _DT[x.dtype], r, s, c, out.data_ptr(), arg.data_ptr(), _stream())
        ctx.save_for_backward(arg)
        ctx.shape = (r, s, c)
        return out

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        r, s, c = ctx.shape
        g = g if g.is_contiguous() else g.contiguous()
        dx = torch.empty(r, s, c, dtype=g.dtype, device=g.device)
        with torch.cuda.device(g.device):
            _lib.call("p2r_maxpool_rows_grad", g.data_ptr(), _DT[g.dtype], arg.data_ptr(), r, s, c, dx.data_ptr(), _stream())
        return dx


def maxpool_rows(x):
    """x [R,S,C] -> max over S -> [R,C] (F.max_pool2d over nsample, pointnet2_modules.py:243-247)."""
    return _MaxPoolRows.apply(x)


class _SAFused(Function):
    """group -> Conv2d 1x1 + ReLU -> Conv2d 1x1 + ReLU -> max over nsample of the set-abstraction layer as one kernel
    (csrc/sa_fused.cu; ref: pointnet2_modules.py:220-256).  Training keeps the first activation (written by the same
    kernel) and the arg-max; the backward pass re-gathers the grouped rows instead of having kept them."""

    @staticmethod
    def forward(ctx, feats, idx, w1, b1, w2, b2, targets):
        b, n, c = feats.shape
        _, p, s = idx.shape
        dev = feats.device
        feats = feats if feats.is_contiguous() else feats.contiguous()
        w1b, w2b = bf16_weight(w1).contiguous(), bf16_weight(w2).contiguous()
        b1f = b1.float().contiguous() if b1 is not None else None
        b2f = b2.float().contiguous() if b2 is not None else None
        need_grad = targets is not None or any(ctx.needs_input_grad)   # (grad mode is always off inside forward)
        out = torch.empty(b * p, c, dtype=torch.bfloat16, device=dev)
        arg = torch.empty(b * p, c, dtype=torch.uint8, device=dev) if need_grad else None
        h1 = torch.empty(b * p * s, c, dtype=torch.bfloat16, device=dev) if need_grad else None
        with torch.cuda.device(dev), _Timed("sa_fused", b * p * s, c, c):
            _lib.call("p2r_sa_fused", feats.data_ptr(), idx.data_ptr(), w1b.data_ptr(), _ptr(b1f), w2b.data_ptr(), _ptr(b2f),
                      b, n, p, s, c, out.data_ptr(), 1, _ptr(arg), _ptr(h1), _stream())
        ctx.save_for_backward(feats, idx, w1b, w2b, h1, arg, out)
        ctx.targets = targets
        ctx.has_bias = (b1 is not None, b2 is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        feats, idx, w1b, w2b, h1, arg, out = ctx.saved_tensors
        b, n, c = feats.shape
        _, p, s = idx.shape
        dev = feats.device
        tc = _TC_GEMM["fn"]
        g = g if g.is_contiguous() else g.contiguous()
        g = g if g.dtype == torch.bfloat16 else g.to(torch.bfloat16)
        rows = b * p * s
        with torch.cuda.device(dev):
            gz = torch.empty_like(g)                    # ReLU of layer 2 at the pooled position: out > 0
            _lib.call("p2r_relu_bwd", g.data_ptr(), out.data_ptr(), _DT[g.dtype], g.numel(), gz.data_ptr(), _stream())
            dz2 = torch.empty(rows, c, dtype=torch.bfloat16, device=dev)
            _lib.call("p2r_maxpool_rows_grad", gz.data_ptr(), _DT[gz.dtype], arg.data_ptr(), b * p, s, c, dz2.data_ptr(),
                      _stream())
            dh1 = tc.linear_dx(dz2, w2b)
            dz1 = torch.empty_like(dh1)
            _lib.call("p2r_relu_bwd", dh1.data_ptr(), h1.data_ptr(), _DT[dh1.dtype], dh1.numel(), dz1.data_ptr(), _stream())
            dfeats = None
            if ctx.needs_input_grad[0]:
                dxg = tc.linear_dx(dz1, w1b)
                d = torch.zeros(b, n, c, dtype=torch.float32, device=dev)
                _lib.call("p2r_group_rows_grad", dxg.data_ptr(), _DT[dxg.dtype], idx.data_ptr(), b, n, c, p, s, d.data_ptr(),
                          _stream())
                dfeats = d.to(feats.dtype)
        has_b1, has_b2 = ctx.has_bias

        def weight_grads():
            xg = torch.empty(rows, c, dtype=feats.dtype, device=dev)       # re-gather (33 MB) instead of having kept it
            with torch.cuda.device(dev):
                _lib.call("p2r_group_rows", feats.data_ptr(), _DT[feats.dtype], idx.data_ptr(), b, n, c, p, s, xg.data_ptr(),
                          _stream())
            return [tc.linear_dw(dz1, xg), _col_sum(dz1, at_join=True) if has_b1 else None,
                    tc.linear_dw(dz2, h1), _col_sum(dz2, at_join=True) if has_b2 else None]

        if ctx.targets is not None:
            # (idx too: the re-gather runs on the side stream after this node's saved tensors have been released -- without
            # it a captured step re-used the index buffer and the gather read garbage row numbers: illegal address)
            _defer(weight_grads, [t if (t is not None and t.requires_grad) else None for t in ctx.targets],
                   (dz1, dz2, h1, feats, idx))
            return dfeats, None, None, None, None, None, None
        gw1, gb1, gw2, gb2 = weight_grads()
        return dfeats, None, gw1, gb1, gw2, gb2, None


def sa_fused_available(feats, idx, convs):
    """The one-kernel set-abstraction path: tensor-core (bf16) mode, two 256 -> 256 layers, nsample a power of two <= 128."""
    tc = _TC_GEMM["fn"]
    s = idx.shape[2]
    return (tc is not None and feats.is_cuda and feats.dtype == torch.bfloat16 and len(convs) == 2 and feats.shape[2] == 256
            and all(cv.in_channels == 256 and cv.out_channels == 256 for cv in convs) and 0 < s <= 128 and (s & (s - 1)) == 0
            and os.environ.get("P2R_FUSED_SA", "1") != "0")


def sa_fused(feats, idx, conv1, conv2):
    """feats [B,N,256] bf16 rows, idx [B,P,S] int32, two nn.Conv2d(256, 256, 1) -> [B*P, 256] bf16 (see _SAFused)."""
    w1, b1 = conv1.weight.reshape(conv1.out_channels, conv1.in_channels), conv1.bias
    w2, b2 = conv2.weight.reshape(conv2.out_channels, conv2.in_channels), conv2.bias
    idx = idx.contiguous()
    if DEFER["on"] and not PROFILE["on"] and torch.is_grad_enabled() and feats.requires_grad:
        det = lambda t: t.detach() if t is not None else None
        return _SAFused.apply(feats, idx, det(w1), det(b1), det(w2), det(b2), (w1, b1, w2, b2))
    return _SAFused.apply(feats, idx, w1, b1, w2, b2, None)


class _EmbedSum(Function):
    @staticmethod
    def forward(ctx, sk, pos):
        f, j, c = sk.shape
        k = pos.shape[1]
        sk = sk if sk.is_contiguous() else sk.contiguous()
        pos = pos if pos.is_contiguous() else pos.contiguous()
        x = torch.empty_like(sk)
        with torch.cuda.device(sk.device):
            _lib.call("p2r_embed_sum", sk.data_ptr(), pos.data_ptr(), _DT[sk.dtype], f, j, k, c, x.data_ptr(), _stream())
        ctx.shape = (f, j, k, c)
        return x

    @staticmethod
    def backward(ctx, dx):
        f, j, k, c = ctx.shape
        dx = dx if dx.is_contiguous() else dx.contiguous()
        dpos = torch.empty(f, k, c, dtype=dx.dtype, device=dx.device)
        with torch.cuda.device(dx.device):
            _lib.call("p2r_embed_sum_grad", dx.data_ptr(), _DT[dx.dtype], f, j, k, c, dpos.data_ptr(), _stream())
        return dx, dpos


def embed_sum(sk, pos):
    """sk [F,J,C] joint features, pos [F,K,C] window embeddings -> sk + mean_k(pos) broadcast over joints."""
    return _EmbedSum.apply(sk, pos)
