"""Small conversion helpers on the op path (reference: pypose/lietensor/convert.py:830-862).

The matrix / Euler converters (mat2SE3, euler2SO3, ...) are I/O-boundary code and not part of the
hot path."""
from torch.nn.functional import normalize

from .lietensor import LieTensor, RxSO3_type, SE3_type, SO3_type, Sim3_type


def quat2unit(input, eps=1e-12):
    """Normalise the quaternion part of a group LieTensor in place and return it (convert.py:830-862)."""
    if isinstance(input, LieTensor) and input.ltype in (SO3_type, RxSO3_type, SE3_type, Sim3_type):
        data = input.tensor()
        sl = slice(0, 4) if input.ltype in (SO3_type, RxSO3_type) else slice(3, 7)
        data[..., sl] = normalize(data[..., sl], p=2, dim=-1, eps=eps)
        out = LieTensor(data, ltype=input.ltype)
        if (out.rotation().tensor().norm(p=2, dim=-1) < eps).any():
            raise ValueError("Detected zero quaternions, which cannot be normalized.")
        return out
    import warnings
    warnings.warn("Input is not Lie group, doing thing and returning input..")
    return input


# ---------------------------------------------------------------------------------------------------------------
# matrix / Euler converters (reference: pypose/lietensor/convert.py:8-258, 607-663).  I/O-boundary helpers written
# with torch ops: same argument meaning, checks and case selection as the reference, so the sign of
# the returned quaternion is the same.
# ---------------------------------------------------------------------------------------------------------------
import torch
import warnings


def _as_matrix(mat):
    if not torch.is_tensor(mat):
        mat = torch.tensor(mat)
    if mat.dim() < 2:
        raise ValueError("Input size must be at least 2 dimensions. Got {}".format(mat.shape))
    if tuple(mat.shape[-2:]) not in ((3, 3), (3, 4), (4, 4)):
        raise ValueError("Input size must be a * x 3 x 3 or * x 3 x 4 or * x 4 x 4 tensor. Got {}".format(mat.shape))
    return mat


def mat2SO3(mat, check=True, rtol=1e-5, atol=1e-5):
    """Rotation matrix (*, 3, 3) -> SO3 (Shepperd's four cases, selected as convert.py:60-90 does)."""
    from .utils import SO3
    R = _as_matrix(mat)[..., :3, :3]
    if check:
        with torch.no_grad():
            eye = torch.eye(3, dtype=R.dtype, device=R.device).expand_as(R)
            if not torch.allclose(R @ R.mT, eye, rtol=rtol, atol=atol):
                raise ValueError("Input rotation matrices are not all orthogonal matrix")
            if not torch.allclose(torch.det(R), torch.ones(R.shape[:-2], dtype=R.dtype, device=R.device), rtol=rtol, atol=atol):
                raise ValueError("Input rotation matrices' determinant are not all equal to 1")
    m = lambda i, j: R[..., i, j]
    t = [1 + m(0, 0) - m(1, 1) - m(2, 2), 1 - m(0, 0) + m(1, 1) - m(2, 2), 1 - m(0, 0) - m(1, 1) + m(2, 2),
         1 + m(0, 0) + m(1, 1) + m(2, 2)]
    cand = [torch.stack([t[0], m(0, 1) + m(1, 0), m(0, 2) + m(2, 0), m(2, 1) - m(1, 2)], -1),
            torch.stack([m(0, 1) + m(1, 0), t[1], m(1, 2) + m(2, 1), m(0, 2) - m(2, 0)], -1),
            torch.stack([m(0, 2) + m(2, 0), m(1, 2) + m(2, 1), t[2], m(1, 0) - m(0, 1)], -1),
            torch.stack([m(2, 1) - m(1, 2), m(0, 2) - m(2, 0), m(1, 0) - m(0, 1), t[3]], -1)]
    small_zz, x_gt_y, x_lt_ny = m(2, 2) < atol, m(0, 0) > m(1, 1), m(0, 0) < -m(1, 1)
    case = torch.where(small_zz, torch.where(x_gt_y, 0, 1), torch.where(x_lt_ny, 2, 3))
    q = torch.zeros_like(cand[0])
    tt = torch.zeros_like(t[0])
    for k in range(4):
        sel = case == k
        q = torch.where(sel.unsqueeze(-1), cand[k], q)
        tt = torch.where(sel, t[k], tt)
    return SO3(q / (2 * tt.sqrt()).unsqueeze(-1))


def _translation_of(mat):
    if mat.shape[-1] == 3:
        return torch.zeros(mat.shape[:-2] + (3,), dtype=mat.dtype, device=mat.device, requires_grad=mat.requires_grad)
    return mat[..., :3, 3]


def _check_last_row(mat, check, rtol, atol):
    if tuple(mat.shape[-2:]) == (4, 4) and check:
        e = torch.tensor([0, 0, 0, 1], dtype=mat.dtype, device=mat.device).expand_as(mat[..., 3, :])
        if not torch.allclose(mat[..., 3, :], e, rtol=rtol, atol=atol):
            warnings.warn("input of shape 4x4 last rows are not all equal [0, 0, 0, 1]")


def mat2SE3(mat, check=True, rtol=1e-5, atol=1e-5):
    from .utils import SE3
    mat = _as_matrix(mat)
    _check_last_row(mat, check, rtol, atol)
    q = mat2SO3(mat[..., :3, :3], check=check, rtol=rtol, atol=atol).tensor()
    return SE3(torch.cat([_translation_of(mat), q], dim=-1))


def _scale_of(mat, rtol, atol):
    rot = mat[..., :3, :3]
    s = torch.pow(torch.det(rot), 1 / 3).unsqueeze(-1)
    if torch.allclose(s, torch.zeros_like(s), rtol=rtol, atol=atol):
        raise ValueError("Rotation matrix not full rank.")
    return rot, s


def mat2Sim3(mat, check=True, rtol=1e-5, atol=1e-5):
    from .utils import Sim3
    mat = _as_matrix(mat)
    _check_last_row(mat, check, rtol, atol)
    rot, s = _scale_of(mat, rtol, atol)
    q = mat2SO3(rot / s.unsqueeze(-1), check=check, rtol=rtol, atol=atol).tensor()
    return Sim3(torch.cat([_translation_of(mat), q, s], dim=-1))


def mat2RxSO3(mat, check=True, rtol=1e-5, atol=1e-5):
    from .utils import RxSO3
    mat = _as_matrix(mat)
    rot, s = _scale_of(mat, rtol, atol)
    q = mat2SO3(rot / s.unsqueeze(-1), check=check, rtol=rtol, atol=atol).tensor()
    return RxSO3(torch.cat([q, s], dim=-1))


def from_matrix(mat, ltype, check=True, rtol=1e-5, atol=1e-5):
    mat = _as_matrix(mat)
    table = {SO3_type: mat2SO3, SE3_type: mat2SE3, Sim3_type: mat2Sim3, RxSO3_type: mat2RxSO3}
    if ltype not in table:
        raise ValueError("Input ltype must be one of SO3_type, SE3_type, Sim3_type or RxSO3_type. Got {}".format(ltype))
    return table[ltype](mat, check=check, rtol=rtol, atol=atol)


def euler2SO3(euler):
    """(roll, pitch, yaw) -> SO3 (convert.py:607-663)."""
    from .utils import SO3
    if not torch.is_tensor(euler):
        euler = torch.tensor(euler)
    assert euler.shape[-1] == 3
    half = 0.5 * euler
    (sr, sp, sy), (cr, cp, cy) = half.sin().unbind(-1), half.cos().unbind(-1)
    return SO3(torch.stack([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                            cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy], dim=-1))
