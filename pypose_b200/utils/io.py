"""Tensor layouts of the two dataset formats the reference's examples feed to the LM routes:

* g2o pose graphs  (examples/module/pgo/pgo_dataset.py:8-51):  `VERTEX_SE3:QUAT id x y z qx qy qz qw`,
  `EDGE_SE3:QUAT i j x y z qx qy qz qw  <21 upper-triangular entries of the 6x6 information matrix>`
  -> nodes (N,7) SE3, edges (E,2) int64, poses (E,7) SE3, infos (E,6,6)       = `optimizer.step((edges, poses), weight=infos)`
* BAL bundle adjustment (examples/module/ba/bal_dataset.py:96-137): header `cameras points observations`, then
  `cam point x y` rows, 9 numbers per camera (rotation vector, translation, f, k1, k2), 3 per point
  -> cameras (C,7) SE3 [t, q], intrinsics (C,3), points (P,3), pixels (M,2), cidx (M,), pidx (M,)

Parsing is vectorised host code (numpy); nothing here touches the GPU — the tensors go straight into
`pp.module.PoseGraph` / `pp.module.BundleAdjustment`.
"""
import numpy as np
import torch

from ..lietensor import SE3


def _info_matrix(upper):
    """(E,21) upper-triangular rows -> (E,6,6) symmetric (pgo_dataset.py:22-29)."""
    iu = np.triu_indices(6)
    A = np.zeros((upper.shape[0], 6, 6), dtype=upper.dtype)
    A[:, iu[0], iu[1]] = upper
    A[:, iu[1], iu[0]] = upper
    return A


def read_g2o(path, dtype=None, device="cpu"):
    dtype = torch.get_default_dtype() if dtype is None else dtype
    ids, nodes, edges, poses, infos = [], [], [], [], []
    with open(path) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "VERTEX_SE3:QUAT":
                ids.append(int(tok[1]))
                nodes.append(tok[2:9])
            elif tok[0] == "EDGE_SE3:QUAT":
                edges.append(tok[1:3])
                poses.append(tok[3:10])
                infos.append(tok[10:31])
    nodes = torch.from_numpy(np.asarray(nodes, dtype=np.float64).reshape(-1, 7))
    poses = torch.from_numpy(np.asarray(poses, dtype=np.float64).reshape(-1, 7))
    infos = torch.from_numpy(_info_matrix(np.asarray(infos, dtype=np.float64).reshape(-1, 21)))
    edges = torch.from_numpy(np.asarray(edges, dtype=np.int64).reshape(-1, 2))
    assert len(ids) == nodes.shape[0] and edges.shape[0] == poses.shape[0] == infos.shape[0]
    return {"ids": torch.tensor(ids, dtype=torch.int64), "nodes": SE3(nodes.to(dtype).to(device)), "edges": edges.to(device),
            "poses": SE3(poses.to(dtype).to(device)), "infos": infos.to(dtype).to(device)}


def _rotvec_to_quat(rv):
    th = np.linalg.norm(rv, axis=-1, keepdims=True)
    with np.errstate(invalid="ignore", divide="ignore"):
        scale = np.where(th > 1e-12, np.sin(0.5 * th) / th, 0.5 - th ** 2 / 48.0)
    return np.concatenate([rv * scale, np.cos(0.5 * th)], -1)


def read_bal(path, dtype=torch.float64, device="cpu"):
    with open(path) as f:
        C, P, M = map(int, f.readline().split())
        rest = np.array(f.read().split(), dtype=np.float64)
    obs = rest[:4 * M].reshape(M, 4)
    cams = rest[4 * M:4 * M + 9 * C].reshape(C, 9)
    pts = rest[4 * M + 9 * C:4 * M + 9 * C + 3 * P].reshape(P, 3)
    pose = np.concatenate([cams[:, 3:6], _rotvec_to_quat(cams[:, :3])], -1)
    t = lambda a, dt=dtype: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(device)
    return {"cameras": SE3(t(pose)), "intrinsics": t(cams[:, 6:]), "points": t(pts), "pixels": t(obs[:, 2:4]),
            "cidx": t(obs[:, 0], torch.int64), "pidx": t(obs[:, 1], torch.int64)}
