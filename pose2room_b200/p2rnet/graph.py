"""Skeleton graph and its spatial-configuration adjacency stack A (K, V, V).

Same result as the reference's `Graph(layout, strategy='spatial', max_hop=5)`
(/root/reference/models/p2rnet/modules/stgcn_layers.py:69-208): hop distances up to max_hop over the
undirected skeleton, column-normalised reachability, then per hop one "root" matrix (hop 0) or a
"root + closer-to-centre" matrix and a "further-from-centre" matrix (hop >= 1): K = 1 + 2*max_hop = 11.
The edge lists are data of the two skeleton layouts the hot path uses (53-joint VirtualHome rig,
25-joint NTU-RGB+D rig; stgcn_layers.py:119-129,151-161).
"""
import numpy as np

_NTU_1BASE = [(1, 2), (2, 21), (3, 21), (4, 3), (5, 21), (6, 5), (7, 6), (8, 7), (9, 21), (10, 9), (11, 10),
              (12, 11), (13, 1), (14, 13), (15, 14), (16, 15), (17, 1), (18, 17), (19, 18), (20, 19), (22, 23),
              (23, 8), (24, 25), (25, 12)]
_VROOM = [(0, 1), (1, 3), (3, 5), (5, 19), (0, 2), (2, 4), (4, 6), (6, 20), (0, 7), (7, 8), (8, 9), (9, 10),
          (10, 21), (10, 22), (8, 11), (11, 13), (13, 15), (15, 17), (8, 12), (12, 14), (14, 16), (16, 18),
          (17, 23), (23, 24), (24, 25), (17, 26), (26, 27), (27, 28), (17, 29), (29, 30), (30, 31), (17, 32),
          (32, 33), (33, 34), (17, 35), (35, 36), (36, 37), (18, 38), (38, 39), (39, 40), (18, 41), (41, 42),
          (42, 43), (18, 44), (44, 45), (45, 46), (18, 47), (47, 48), (48, 49), (18, 50), (50, 51), (51, 52)]

LAYOUTS = {
    "virtualroom": dict(num_node=53, edges=_VROOM, center=0),
    "ntu-rgb+d": dict(num_node=25, edges=[(i - 1, j - 1) for i, j in _NTU_1BASE], center=20),
}


def layout_for_joints(joint_num):
    return {53: "virtualroom", 25: "ntu-rgb+d"}[joint_num]


def hop_distance(num_node, edges, max_hop):
    """Shortest-path length between joints, inf beyond max_hop (stgcn_layers.py:208-221)."""
    adj = np.eye(num_node, dtype=bool)
    for i, j in edges:
        adj[i, j] = adj[j, i] = True
    hop = np.full((num_node, num_node), np.inf)
    reach = np.eye(num_node, dtype=bool)
    hop[reach] = 0
    for d in range(1, max_hop + 1):
        nxt = (reach.astype(np.int64) @ adj.astype(np.int64)) > 0
        hop[nxt & np.isinf(hop)] = d
        reach = nxt
    return hop


def spatial_adjacency(layout="virtualroom", max_hop=5):
    spec = LAYOUTS[layout]
    n, center = spec["num_node"], spec["center"]
    hop = hop_distance(n, spec["edges"], max_hop)
    reach = (hop <= max_hop).astype(np.float64)
    deg = reach.sum(0)
    norm = reach / np.where(deg > 0, deg, 1.0)[None, :]   # A . D^-1  (column normalisation)
    mats = []
    dc = hop[:, center]
    for h in range(max_hop + 1):
        sel = hop == h                                      # sel[j, i]: joint j is h hops from joint i
        same = sel & (dc[:, None] == dc[None, :])
        closer = sel & (dc[:, None] > dc[None, :])
        further = sel & (dc[:, None] < dc[None, :])
        root = np.where(same, norm, 0.0)
        if h == 0:
            mats.append(root)
        else:
            mats.append(root + np.where(closer, norm, 0.0))
            mats.append(np.where(further, norm, 0.0))
    return np.stack(mats)
