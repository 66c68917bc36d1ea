"""Graph / loss / eval geometry of the P2RNet hot path on the B200 kernels.

Host-side mirror of (citations into /root/reference):
  knn, get_graph_offset          net_utils/vn_dgcnn_util.py:4-10,70-95
  nn_distance, huber_loss        net_utils/nn_distance.py:15-61
  nms_3d_faster[_samecls]        net_utils/nms.py:41-119
  box3d_iou                      net_utils/box_util.py:90-118
  parse_predictions              net_utils/ap_helper.py:133-255
Same names, argument meaning and return types as the reference functions.
"""
import numpy as np
import torch
from torch.autograd import Function

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("pose2room_b200: CUDA tensor required (there is no CPU path)")


# ----------------------------------------------------------------------------------------- graph
def knn(x, k):
    """x (B,C,N) float32 -> idx (B,N,k) int64 of the k nearest points (self included)."""
    _need_cuda(x)
    x = x.contiguous().float()
    b, c, n = x.shape
    idx = torch.empty(b, n, k, dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("p2r_knn_graph", x.data_ptr(), b, c, n, int(k), idx.data_ptr(), _stream())
    return idx


class _GraphOffset(Function):
    @staticmethod
    def forward(ctx, x, idx):
        b, d3, n = x.shape
        k = idx.shape[2]
        out = torch.empty(b, n, k, d3 // 3, 3, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("p2r_graph_offset", x.data_ptr(), idx.data_ptr(), b, d3, n, k, out.data_ptr(), _stream())
        ctx.save_for_backward(idx)
        ctx.shape = (b, d3, n, k)
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        b, d3, n, k = ctx.shape
        g = g.reshape(b, n, k, d3)
        gx = torch.zeros(b, n, d3, dtype=g.dtype, device=g.device)
        gx.scatter_add_(1, idx.reshape(b, n * k, 1).expand(b, n * k, d3), g.reshape(b, n * k, d3))
        gx -= g.sum(dim=2)
        return gx.transpose(1, 2).contiguous(), None


def get_graph_offset(x, k=20, idx=None, x_coord=None):
    """x (B,3d,N) -> (B,N,k,d,3) = x[idx] - x  (vn_dgcnn_util.py:70-95)."""
    _need_cuda(x)
    b = x.size(0)
    n = x.size(2)
    x = x.reshape(b, -1, n).contiguous().float()
    if idx is None:
        idx = knn(x if x_coord is None else x_coord, k=k)
    return _GraphOffset.apply(x, idx.contiguous().long())


def get_graph_feature(x, k=20, idx=None, x_coord=None):
    """vn_dgcnn_util.py:13-40: x (B,d,3,N) -> (B,2d,3,N,k) = cat(x[idx]-x, x) (edge features of the VN-DGCNN stage)."""
    b, n = x.size(0), x.size(3)
    flat = x.reshape(b, -1, n)
    off = get_graph_offset(flat, k=k, idx=idx, x_coord=x_coord)            # (B,N,k,d,3)
    d = flat.size(1) // 3
    centre = flat.transpose(2, 1).reshape(b, n, 1, d, 3).expand(b, n, off.size(2), d, 3)
    return torch.cat((off, centre), dim=3).permute(0, 3, 4, 1, 2).contiguous()


def get_graph_feature_cross(x, k=20, idx=None):
    """vn_dgcnn_util.py:97-121: as get_graph_feature plus the cross product (x[idx] x x) as a third block."""
    b, n = x.size(0), x.size(3)
    flat = x.reshape(b, -1, n)
    off = get_graph_offset(flat, k=k, idx=idx)
    d = flat.size(1) // 3
    centre = flat.transpose(2, 1).reshape(b, n, 1, d, 3).expand(b, n, off.size(2), d, 3)
    cross = torch.cross(off + centre, centre, dim=-1)
    return torch.cat((off, centre, cross), dim=3).permute(0, 3, 4, 1, 2).contiguous()


# ----------------------------------------------------------------------------------------- losses
def huber_loss(error, delta=1.0):
    """nn_distance.py:15-32 (elementwise torch; tiny tensors)."""
    abs_error = torch.abs(error)
    quadratic = torch.clamp(abs_error, max=delta)
    linear = abs_error - quadratic
    return 0.5 * quadratic ** 2 + delta * linear


class _NNDistance(Function):
    @staticmethod
    def forward(ctx, pc1, pc2, mode, delta):
        b, n, c = pc1.shape
        m = pc2.shape[1]
        dev = pc1.device
        dist1 = torch.empty(b, n, dtype=torch.float32, device=dev)
        idx1 = torch.empty(b, n, dtype=torch.int64, device=dev)
        dist2 = torch.empty(b, m, dtype=torch.float32, device=dev)
        idx2 = torch.empty(b, m, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.call("p2r_nn_distance", pc1.data_ptr(), pc2.data_ptr(), b, n, m, c, mode, float(delta),
                      dist1.data_ptr(), idx1.data_ptr(), dist2.data_ptr(), idx2.data_ptr(), _stream())
        ctx.save_for_backward(pc1, pc2, idx1, idx2)
        ctx.mode, ctx.delta = mode, float(delta)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, idx1, dist2, idx2

    @staticmethod
    def backward(ctx, g1, gi1, g2, gi2):
        pc1, pc2, idx1, idx2 = ctx.saved_tensors
        b, n, c = pc1.shape
        m = pc2.shape[1]
        gp1 = torch.zeros_like(pc1)
        gp2 = torch.zeros_like(pc2)
        g1 = g1.contiguous().float() if g1 is not None else None
        g2 = g2.contiguous().float() if g2 is not None else None
        with torch.cuda.device(pc1.device):
            _lib.call("p2r_nn_distance_grad", pc1.data_ptr(), pc2.data_ptr(), idx1.data_ptr(), idx2.data_ptr(),
                      g1.data_ptr() if g1 is not None else None, g2.data_ptr() if g2 is not None else None,
                      b, n, m, c, ctx.mode, ctx.delta, gp1.data_ptr(), gp2.data_ptr(), _stream())
        return gp1, gp2, None, None


def nn_distance(pc1, pc2, l1smooth=False, delta=1.0, l1=False):
    """pc1 (B,N,C), pc2 (B,M,C) -> dist1 (B,N) f32, idx1 (B,N) i64, dist2 (B,M) f32, idx2 (B,M) i64."""
    _need_cuda(pc1, pc2)
    mode = 2 if l1smooth else (1 if l1 else 0)
    return _NNDistance.apply(pc1.contiguous().float(), pc2.contiguous().float(), mode, delta)


# ----------------------------------------------------------------------------------------- eval
def decode_boxes(center, log_size, heading_sincos, hip, contact=1.0):
    """center/log_size f32 (B,K,3), heading (B,K,2) (sin,cos), hip f32 (B,T,3) (may be a strided view of
    input_joints[:, :, origin_joint]) -> corners f64 (B,K,8,3), aabb f64 (B,K,6), nonempty u8 (B,K)."""
    _need_cuda(center, log_size, heading_sincos, hip)
    center = center.detach().contiguous().float()
    log_size = log_size.detach().contiguous().float()
    heading = heading_sincos.detach().contiguous().double()
    b, k, _ = center.shape
    hip = hip.detach()
    t = hip.shape[1]
    if not (hip.dtype == torch.float32 and hip.stride(2) == 1 and hip.stride(0) == t * hip.stride(1)):
        hip = hip.float().contiguous()
    dev = center.device
    corners = torch.empty(b, k, 8, 3, dtype=torch.float64, device=dev)
    aabb = torch.empty(b, k, 6, dtype=torch.float64, device=dev)
    nonempty = torch.empty(b, k, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.call("p2r_decode_boxes", center.data_ptr(), log_size.data_ptr(), heading.data_ptr(), hip.data_ptr(),
                  int(hip.stride(1)), b, k, t, float(contact), corners.data_ptr(), aabb.data_ptr(),
                  nonempty.data_ptr(), _stream())
    return corners, aabb, nonempty


def nms3d_batched(aabb, score, valid=None, cls=None, thr=0.10, old_type=False):
    """Batched greedy NMS on the device.  aabb f64 (B,K,6), score (B,K) -> keep u8 (B,K), order i32 (B,K)."""
    _need_cuda(aabb, score)
    aabb = aabb.contiguous().double()
    score = score.contiguous().double()
    b, k, _ = aabb.shape
    dev = aabb.device
    keep = torch.empty(b, k, dtype=torch.uint8, device=dev)
    order = torch.empty(b, k, dtype=torch.int32, device=dev)
    v = valid.contiguous().to(torch.uint8) if valid is not None else None
    c = cls.contiguous().to(torch.int32) if cls is not None else None
    with torch.cuda.device(dev):
        _lib.call("p2r_nms3d", aabb.data_ptr(), score.data_ptr(), v.data_ptr() if v is not None else None,
                  c.data_ptr() if c is not None else None, b, k, float(thr), int(bool(old_type)),
                  keep.data_ptr(), order.data_ptr(), _stream())
    return keep, order


def _nms_numpy_front(boxes, thr, old_type, with_cls):
    boxes = np.asarray(boxes, dtype=np.float64)
    if boxes.shape[0] == 0:
        return []
    t = torch.from_numpy(boxes).cuda()
    cls = t[None, :, 7].to(torch.int32) if with_cls else None
    keep, order = nms3d_batched(t[None, :, 0:6], t[None, :, 6], None, cls, thr, old_type)
    order = order[0].cpu().numpy()
    return [int(i) for i in order if i >= 0]


def nms_2d_faster(boxes, overlap_threshold, old_type=False):
    """nms.py:7-39: boxes (K,5) [x1,y1,x2,y2,score].  Runs on the 3-D kernel with unit z extent: area*(1-0) and
    (w*h)*1 are the same fp64 values as the reference's 2-D products, so the selection is identical."""
    boxes = np.asarray(boxes, dtype=np.float64)
    if boxes.shape[0] == 0:
        return []
    k = boxes.shape[0]
    b3 = np.zeros((k, 7))
    b3[:, 0:2] = boxes[:, 0:2]
    b3[:, 3:5] = boxes[:, 2:4]
    b3[:, 5] = 1.0
    b3[:, 6] = boxes[:, 4]
    return _nms_numpy_front(b3, overlap_threshold, old_type, False)


def nms_3d_faster(boxes, overlap_threshold, old_type=False):
    """nms.py:41-77: boxes (K,7) [x1,y1,z1,x2,y2,z2,score] numpy -> list of kept rows, descending score."""
    return _nms_numpy_front(boxes, overlap_threshold, old_type, False)


def nms_3d_faster_samecls(boxes, overlap_threshold, old_type=False):
    """nms.py:79-119: boxes (K,8) with the class in column 7; only same-class boxes suppress each other."""
    return _nms_numpy_front(boxes, overlap_threshold, old_type, True)


def box3d_iou_matrix(corners1, corners2):
    """corners1 (P,8,3), corners2 (G,8,3) (tensor or ndarray) -> (iou3d, iou2d) f64 (P,G) CUDA tensors."""
    c1 = torch.as_tensor(corners1, dtype=torch.float64).cuda().contiguous()
    c2 = torch.as_tensor(corners2, dtype=torch.float64).cuda().contiguous()
    p, g = c1.shape[0], c2.shape[0]
    i3 = torch.empty(p, g, dtype=torch.float64, device=c1.device)
    i2 = torch.empty(p, g, dtype=torch.float64, device=c1.device)
    with torch.cuda.device(c1.device):
        _lib.call("p2r_box3d_iou", c1.data_ptr(), c2.data_ptr(), p, g, i3.data_ptr(), i2.data_ptr(), _stream())
    return i3, i2


def box3d_iou(corners1, corners2):
    """box_util.py:90-118 single pair -> (iou3d, iou2d) python floats."""
    i3, i2 = box3d_iou_matrix(np.asarray(corners1)[None], np.asarray(corners2)[None])
    return float(i3[0, 0]), float(i2[0, 0])
