"""Device-decided LM trials (csrc/lmstep.cu): one C call per trial, one host read per step.

The host loop keeps the shape of optimizer.py:659-680; what moved to the device is everything inside one trial —
linearise, solve, retract, trial loss, `strategy.update`, the accept test and the parameter update.  The host owns
`param_groups` (users and schedulers edit the damping between steps): the current values travel as arguments of the
call and the updated ones come back in the 16-double state, which is the step's only device->host read.
"""
import ctypes

import torch

from .. import _C
from . import _fused
from .strategy import Adaptive, Constant, TrustRegion

ST_STATUS, ST_LOSS, ST_LAST, ST_DAMPING, ST_RADIUS, ST_DOWN, ST_REJECT, ST_CUR, ST_TRIAL, ST_PRED, ST_FAILED = range(11)
_KIND = {Constant: 0, Adaptive: 1, TrustRegion: 2}


def strategy_kind(strategy):
    """0 / 1 / 2 for exactly the three reference strategies (a subclass may override update -> host route), else None."""
    return _KIND.get(type(strategy))


def fill_ctl(ctl, strategy, pg, last, cached, reject_count, reject_limit):
    """The 14 host doubles of b200_lm_*_step's `ctl` (layout: include/b200pose.h)."""
    kind = _KIND[type(strategy)]
    ctl[0], ctl[1], ctl[2] = last, 1.0 if cached else 0.0, pg['damping']
    ctl[4], ctl[5], ctl[6] = float(reject_count), float(reject_limit), float(kind)
    if kind == 0:
        ctl[3] = 0.0
    else:
        ctl[3], ctl[7], ctl[8], ctl[9] = pg['down'], pg['high'], pg['low'], pg['up']
        ctl[12], ctl[13] = strategy.min, strategy.max
        if kind == 2:
            ctl[10], ctl[11] = strategy.down, pg['factor']


def apply_state(strategy, pg, st):
    """param_groups <- what strategy.update would have written (strategy.py:41-46, 134-151, 248-274)."""
    kind = _KIND[type(strategy)]
    if kind == 1:
        pg['damping'] = st[ST_DAMPING]
    elif kind == 2:
        pg['radius'], pg['down'], pg['damping'] = st[ST_RADIUS], st[ST_DOWN], st[ST_DAMPING]


class DeviceStep:
    """Per-problem buffers of the device-decided route: pinned host copy of the state, host `ctl`, bound entry point."""

    def __init__(self, device, dtype):
        self.device, self.dtype = device, dtype
        self.host = torch.zeros(16, dtype=torch.float64).pin_memory()
        self.host_ptr = self.host.data_ptr()
        self.ctl = (ctypes.c_double * 14)()
        self.ctl_ptr = ctypes.addressof(self.ctl)
        self.W = _fused._workspaces(device)
        self.sfx = _C.suffix(dtype)
        self.state = None
        self.comm = None
        self.seq = 0

    def next_seq(self):
        self.seq += 1
        return self.seq

    def attach_comm(self, comm, part_off=0, pt_off=0):
        self.comm, self.part_off, self.pt_off = comm, part_off, pt_off
        self.epoch0 = self.epoch1 = 0

    def new_state(self):
        """A fresh 16-double device state per step: `optimizer.loss` / `.last` are views of it, so losses a caller keeps
        from earlier steps do not change under them."""
        self.state = torch.empty(16, dtype=torch.float64, device=self.device)
        return self.state

    def read(self):
        return self.host.tolist()


def reproj_trial(ds, prob, scale, dmin, dmax, retry):
    poses = prob._poses()
    H, g, Pt = prob._buf
    _C.enqueue("b200_lm_reproj_step_" + ds.sfx, poses, poses.data_ptr(), prob.pts.data_ptr(), prob.pix.data_ptr(),
               prob.seg.data_ptr(), H.data_ptr(), g.data_ptr(), Pt.data_ptr(), ds.W[0].data_ptr(), ds.W[1].data_ptr(),
               ds.state.data_ptr(), ds.host_ptr, ds.next_seq(), ds.ctl_ptr, int(prob.robust[0]), float(prob.robust[1]), float(scale),
               float(dmin), float(dmax), 1 if retry else 0, prob.pts.shape[0], poses.shape[0])
    return ds.read()


def poseinv_trial(ds, prob, scale, dmin, dmax, retry):
    P, X = prob._rows()
    if prob._trial is None or prob._trial.shape != P.shape:
        prob._trial = torch.empty_like(P)
    _C.enqueue("b200_lm_poseinv_step_" + ds.sfx, P, P.data_ptr(), X.data_ptr(), prob._trial.data_ptr(), ds.W[0].data_ptr(),
               ds.state.data_ptr(), ds.host_ptr, ds.next_seq(), ds.ctl_ptr, int(prob.robust[0]), float(prob.robust[1]), float(scale),
               float(dmin), float(dmax), P.shape[0])
    return ds.read()


def _align(n, a=256):
    return (n + a - 1) // a * a


def reproj_payload(ncam, world, itemsize):
    """(part_off, pt_off, bytes) of the exchange payload of b200_lm_reproj_step_peer."""
    q = (ncam + world - 1) // world
    part = _align(world * q * 27 * itemsize)
    return 0, part, part + _align(ncam * 7 * itemsize)


def reproj_trial_peer(ds, prob, scale, dmin, dmax, retry):
    poses = prob._poses()
    H, g, _ = prob._buf
    c = ds.comm
    if not retry:
        ds.epoch0 += 1
    ds.epoch1 += 1
    _C.enqueue("b200_lm_reproj_step_peer_" + ds.sfx, poses, poses.data_ptr(), prob.pts.data_ptr(), prob.pix.data_ptr(),
               prob.seg.data_ptr(), H.data_ptr(), g.data_ptr(), c.bases_ptr, c.rank, c.world, ds.part_off, ds.pt_off,
               ds.epoch0, ds.epoch1, ds.W[0].data_ptr(), ds.W[1].data_ptr(), ds.W[2].data_ptr(), ds.state.data_ptr(),
               ds.host_ptr, ds.next_seq(), ds.ctl_ptr, int(prob.robust[0]), float(prob.robust[1]), float(scale), float(dmin), float(dmax),
               1 if retry else 0, poses.shape[0])
    return ds.read()


def poseinv_trial_peer(ds, prob, scale, dmin, dmax, retry):
    P, X = prob._rows()
    if prob._trial is None or prob._trial.shape != P.shape:
        prob._trial = torch.empty_like(P)
    c = ds.comm
    ds.epoch1 += 1
    _C.enqueue("b200_lm_poseinv_step_peer_" + ds.sfx, P, P.data_ptr(), X.data_ptr(), prob._trial.data_ptr(), c.bases_ptr,
               c.rank, c.world, ds.epoch1, ds.W[0].data_ptr(), ds.state.data_ptr(), ds.host_ptr, ds.next_seq(), ds.ctl_ptr,
               int(prob.robust[0]), float(prob.robust[1]), float(scale), float(dmin), float(dmax), P.shape[0])
    return ds.read()
