"""Data-parallel plumbing of the hot path: the reference's DDP axis (net_utils/utils.py:251, one gradient all-reduce
per step) as two explicit calls that are safe inside a captured CUDA graph, plus the cross-rank gather the reference's
sharded evaluation lacks."""
import torch
import torch.distributed as dist


def broadcast_parameters(module, src=0):
    """Identical replicas at start-up (what DistributedDataParallel does in its constructor)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


def allreduce_gradients(params):
    """Average the gradients of `params` over all ranks with ONE flat all-reduce (the whole model is 1.44 M parameters
    = 5.8 MB: a single bucket; message-latency bound on NVLink).  Parameters without a gradient are skipped
    consistently on every rank (same model, same graph)."""
    from . import ops
    if ops.deferred_gradients_pending():
        raise RuntimeError("allreduce_gradients inside ops.overlap_weight_grads(): the weight gradients of this step are "
                           "still deferred (parameters get their .grad when the context exits); call it after the block")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    # one bucket per dtype (the heading GMM's mu is float64)
    by_dtype = {}
    for g in grads:
        by_dtype.setdefault(g.dtype, []).append(g)
    world = dist.get_world_size()
    for gs in by_dtype.values():
        flat = torch._utils._flatten_dense_tensors(gs)
        if dist.get_backend() == "gloo":
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.div_(world)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        # one multi-tensor copy back into the ~150 gradient tensors (a copy kernel each costs more than the all-reduce)
        torch._foreach_copy_(gs, list(torch._utils._unflatten_dense_tensors(flat, gs)))


def gather_ap_state(ap_calculator):
    """all_gather_object the (pred_map_cls, gt_map_cls) of every rank into `ap_calculator` on all ranks, so that
    compute_metrics() sees the whole evaluation set (the reference computes AP per shard)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ap_calculator
    mine = (ap_calculator.pred_map_cls, ap_calculator.gt_map_cls)
    states = [None] * dist.get_world_size()
    dist.all_gather_object(states, mine)
    ap_calculator.reset()
    for st in states:
        ap_calculator.merge(st)
    return ap_calculator
