"""OptimizablePose with the interface of src/se3pose.py:8-92: a 6-vector parameter [t(3), w(3)],
R(w) = I + A(|w|) [w]x + B(|w|) [w]x^2 with the reference's 11-term Taylor polynomials.

These torch methods are the host-side API (pose bookkeeping in the tracker/mapper); inside the fused
optimisation loop the same maps and their Jacobian run as CUDA kernels (csrc/pose.cu).
"""
from copy import deepcopy
from math import factorial, pi

import torch
import torch.nn as nn

_NTH = 10


def _series(x, first):
    """sum_{i<=10} (-1)^i x^(2i) / (2i+first)!  -- first=1: sin(x)/x, first=2: (1-cos x)/x^2, first=3: (x-sin x)/x^3."""
    out = torch.zeros_like(x)
    for i in range(_NTH + 1):
        out = out + (-1) ** i * x ** (2 * i) / float(factorial(2 * i + first))
    return out


class OptimizablePose(nn.Module):
    def __init__(self, init_pose):
        super().__init__()
        assert isinstance(init_pose, torch.FloatTensor)
        self.register_parameter("data", nn.Parameter(init_pose))

    def copy_from(self, pose):
        self.data = deepcopy(pose.data)

    @classmethod
    def taylor_A(cls, x, nth=_NTH):
        return _series(x, 1)

    @classmethod
    def taylor_B(cls, x, nth=_NTH):
        return _series(x, 2)

    @classmethod
    def taylor_C(cls, x, nth=_NTH):
        return _series(x, 3)

    @classmethod
    def skew_symmetric(cls, w):
        a, b, c = w.unbind(dim=-1)
        z = torch.zeros_like(a)
        return torch.stack([torch.stack([z, -c, b], -1), torch.stack([c, z, -a], -1), torch.stack([-b, a, z], -1)], -2)

    def rotation(self):
        w = self.data[3:]
        K = self.skew_symmetric(w)
        th = w.norm(dim=-1)[..., None, None]
        eye = torch.eye(3, device=w.device, dtype=torch.float32)
        return eye + self.taylor_A(th) * K + self.taylor_B(th) * K @ K

    def translation(self):
        return self.data[:3]

    def matrix(self):
        Rt = torch.eye(4)
        Rt[:3, :3] = self.rotation()
        Rt[:3, 3] = self.translation()
        return Rt

    @classmethod
    def log(cls, R, eps=1e-7):
        tr = R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]
        th = ((tr - 1) / 2).clamp(-1 + eps, 1 - eps).acos_()[..., None, None] % pi
        lnR = 1 / (2 * cls.taylor_A(th) + 1e-8) * (R - R.transpose(-2, -1))
        return torch.stack([lnR[..., 2, 1], lnR[..., 0, 2], lnR[..., 1, 0]], dim=-1)

    @classmethod
    def from_matrix(cls, Rt, eps=1e-8):
        return OptimizablePose(torch.cat([Rt[:3, 3], cls.log(Rt[:3, :3])], dim=-1))
