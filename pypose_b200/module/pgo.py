"""Pose-graph model (reference: examples/module/pgo/pgo.py:15-25).

    forward(edges, poses) = Log(poses^-1 @ nodes[edges[:,0]]^-1 @ nodes[edges[:,1]])        (E, 6)

Written with ordinary LieTensor ops (runs through the generic dense route and through the reference);
`pp.optim.LM(..., solver=PCG(), sparse=True)` recognises the type and takes the block-sparse route
(optim/structured.py: per-edge J^T J blocks + matrix-free block-Jacobi PCG)."""
from torch import nn

from ..lietensor.lietensor import Parameter


class PoseGraph(nn.Module):
    def __init__(self, nodes):
        super().__init__()
        self.nodes = Parameter(nodes, sjac=True)

    def forward(self, edges, poses):
        node1 = self.nodes[edges[..., 0]]
        node2 = self.nodes[edges[..., 1]]
        return (poses.Inv() @ node1.Inv() @ node2).Log().tensor()
