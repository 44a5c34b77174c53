"""Host-side mirror of the reference's geometry interface (SMNet/projector/{core,projector,point_cloud}.py)
on top of the eod_backproject_quantize kernel.

Scalars stay on the host exactly as in the reference: the 4x4 pose comes from fp32 torch.cos/sin of
(heading, elevation + pi) (core.py:6-34) and the intrinsics from Python-double math rounded once to fp32
(core.py:68-77).  The kernel receives them as inputs and never recomputes trig or intrinsics, which is
what makes the cell indices bit-exact with torch-CPU execution of the reference.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch

from . import ops
from ._lib import ORDER_XZ, ORDER_ZX


def transform3d(xyzhe: torch.Tensor, axis_swap: bool = False) -> torch.Tensor:
    """(N,5) x,y,z,heading,elevation -> (N,4,4) camera-to-world, fp32 on the host (core.py:6-34).
    axis_swap: the online robot variant (robot_demo.py:40-90), T @ R with R exchanging the camera's x and z axes."""
    xyzhe = xyzhe.detach().to("cpu", torch.float32)
    cx, sx = torch.cos(xyzhe[:, 4]), torch.sin(xyzhe[:, 4])
    cy, sy = torch.cos(xyzhe[:, 3]), torch.sin(xyzhe[:, 3])
    T = torch.zeros(xyzhe.shape[0], 4, 4, dtype=torch.float32)
    T[:, 0, 0] = cy
    T[:, 0, 1] = sx * sy
    T[:, 0, 2] = cx * sy
    T[:, 0, 3] = xyzhe[:, 0]
    T[:, 1, 1] = cx
    T[:, 1, 2] = -sx
    T[:, 1, 3] = xyzhe[:, 1]
    T[:, 2, 0] = -sy
    T[:, 2, 1] = cy * sx
    T[:, 2, 2] = cy * cx
    T[:, 2, 3] = xyzhe[:, 2]
    T[:, 3, 3] = 1
    if axis_swap:
        # torch.matmul(T, R) with the 0/1 permutation R of robot_demo.py:68-88: columns 0 and 2 change places.  Every
        # product with a 0/1 entry and every sum with the resulting +0 is exact, so the column swap is bit-identical
        # (tests/golden/robot.npz holds the matmul's own output).
        T = T[:, :, [2, 1, 0, 3]].contiguous()
    return T


def compute_intrinsics(width: int, height: int, vfov: float) -> Tuple[float, float, float, float]:
    """(fx, fy, cx, cy) rounded to fp32 like torch.Tensor([[f_x, 0, cx], ...]) does (core.py:68-77)."""
    hfov = width / height * vfov
    vals = torch.tensor([width / (2.0 * math.tan(hfov / 2.0)), height / (2.0 * math.tan(vfov / 2.0)), width / 2.0,
                         height / 2.0], dtype=torch.float64).to(torch.float32)
    return tuple(float(v) for v in vals)


class Projector:
    """Drop-in for SMNet/projector/projector.py:Projector (and PointCloud via ``point_cloud``).

    forward(depth (B,1,H,W), T (B,4,4), return_heights=False) ->
        (projection_indices_2D int64 (B,H,W,2), outliers bool (B,H,W)[, heights f32 (B,H,W)])
    Extra: ``flat_indices`` returns the clipped flat int32 cell index of build_memory_data.py:135-143 /
    robot_demo.py:526-533 in the same launch.
    """

    def __init__(self, vfov: float, batch_size: int, feature_map_height: int, feature_map_width: int, output_height: int,
                 output_width: int, gridcellsize: float, world_shift_origin, z_clip_threshold: float,
                 device: torch.device = torch.device("cuda"), intrinsics: Optional[Sequence[float]] = None, depth_div: float = 1000.0):
        self.depth_div = float(depth_div)                        # divisor of uint16 depth inputs (robot_demo.py:515: millimetres)
        self.vfov, self.batch_size = vfov, batch_size
        self.fmh, self.fmw = feature_map_height, feature_map_width
        self.output_height, self.output_width = output_height, output_width
        self.gridcellsize, self.z_clip_threshold = gridcellsize, z_clip_threshold
        self.device = torch.device(device)
        self.world_shift_origin = torch.as_tensor(world_shift_origin, dtype=torch.float32).reshape(3).cpu()
        # intrinsics: (fx, fy, cx, cy) of a calibrated camera instead of the vfov model - the robot demo hard-codes its K
        # (robot_demo.py:123-125); rounded to fp32 as torch.tensor(K) holds them
        self.intrinsics = compute_intrinsics(feature_map_width, feature_map_height, vfov) if intrinsics is None else \
            tuple(float(v) for v in torch.tensor([float(x) for x in intrinsics], dtype=torch.float32))

    def _run(self, depth: torch.Tensor, T: torch.Tensor, map_world_shift=None, order: int = ORDER_ZX, **want):
        assert depth.shape[2] == self.fmh and depth.shape[3] == self.fmw
        # raw uint16 sensor depth stays uint16 (divided by depth_div inside the kernel: robot_demo.py:515-517)
        d = depth[:, 0].to(self.device).contiguous() if depth.dtype == torch.uint16 else depth[:, 0].to(self.device, torch.float32).contiguous()
        B = d.shape[0]
        pose = T.detach().to("cpu", torch.float32)[:, :3, :].reshape(B, 12)
        s1 = torch.zeros(3) if map_world_shift is None else torch.as_tensor(map_world_shift, dtype=torch.float32).reshape(3).cpu()
        shifts = torch.cat([self.world_shift_origin, s1]).unsqueeze(0).repeat(B, 1)
        return ops.backproject_quantize(d, pose.to(self.device), shifts.to(self.device), self.intrinsics, self.gridcellsize,
                                        self.output_width, self.output_height, order, self.z_clip_threshold, depth_div=self.depth_div, **want)

    def forward(self, depth: torch.Tensor, T: torch.Tensor, obs_per_map: int = 1, return_heights: bool = False):
        r = self._run(depth, T, want_idx=False, want_q2=True, want_outlier=True, want_height=return_heights)
        idx2d, outliers = r["q2"].to(torch.int64), r["outlier"].view(torch.bool)
        return (idx2d, outliers, r["height"]) if return_heights else (idx2d, outliers)

    __call__ = forward

    def point_cloud(self, depth: torch.Tensor, T: torch.Tensor):
        """PointCloud.forward (point_cloud.py:56-85): (world xyz (B,H,W,3) f32, no_depth_mask bool)."""
        r = self._run(depth, T, want_idx=False, want_world=True)
        return r["world"], depth[:, 0].to(self.device) == 0

    def flat_indices(self, depth: torch.Tensor, T: torch.Tensor, map_world_shift=None, order: str = "zx") -> torch.Tensor:
        """(B,H,W,1) int32 clipped flat cell indices, the ``proj_indices`` the model consumes."""
        r = self._run(depth, T, map_world_shift, ORDER_ZX if order == "zx" else ORDER_XZ)
        return r["idx"].unsqueeze(-1)
