"""Synthetic MP3D-shaped RGB-D episodes (host side, numpy) - the inputs of the bench and the parity tests.

There is no simulator and no dataset in this environment, so episodes are synthesised to
the reference's shapes and conventions:
  * camera: 480x640, vfov 67.5 deg (SMNet/build_data.py:75-80), sensor 1.25 m above the floor, depth in
    metres clipped to [0, 10] (habitat normalised depth x10, build_data.py:205-207), 0 == no depth;
  * pose: (x, y, z, heading, elevation + pi) exactly as the builders assemble it (build_data.py:190-196);
  * trajectory: random walk of 'forward 0.1 m' / 'turn +-9 deg' steps (SMNet/utils/habitat_utils.py:31-32);
  * scene: piecewise-planar room (floor, ceiling, four walls, a few boxes) ray-cast per pixel, so that
    neighbouring pixels fall into neighbouring map cells as in a real scan.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Tuple

import numpy as np

VFOV_DEG = 67.5
SENSOR_HEIGHT = 1.25
MAX_DEPTH = 10.0


@dataclass
class Room:
    x0: float
    x1: float
    z0: float
    z1: float
    ceil: float
    boxes: np.ndarray            # (nb, 6): xmin, xmax, ymin, ymax, zmin, zmax


@dataclass
class Episode:
    depth: np.ndarray            # (T, H, W) f32 metres
    xyzhe: np.ndarray            # (T, 5) f32: x, y, z, heading, elevation + pi
    map_world_shift: np.ndarray  # (3,) f32
    cell: float
    map_w: int
    map_h: int
    room: Room
    seed: int
    meta: dict = field(default_factory=dict)


def make_room(rng: np.random.Generator, size_x: float = 12.0, size_z: float = 9.0, n_boxes: int = 5) -> Room:
    boxes = []
    for _ in range(n_boxes):
        w, d, h = rng.uniform(0.4, 1.6), rng.uniform(0.4, 1.6), rng.uniform(0.4, 1.8)
        cx, cz = rng.uniform(1.0, size_x - 1.0), rng.uniform(1.0, size_z - 1.0)
        boxes.append([cx - w / 2, cx + w / 2, 0.0, h, cz - d / 2, cz + d / 2])
    return Room(0.0, size_x, 0.0, size_z, 2.6, np.asarray(boxes, np.float64).reshape(-1, 6))


def _ray_dirs(H: int, W: int, vfov: float, heading: float) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """World-space ray directions with unit camera-z component, so that the ray parameter IS the depth
    (camera looks along (-sin h, 0, -cos h); image v grows downwards)."""
    hfov = W / H * vfov
    fx = W / (2.0 * math.tan(hfov / 2.0))
    fy = H / (2.0 * math.tan(vfov / 2.0))
    xs = (np.arange(W) + 0.5 - W / 2.0) / fx
    ys = (np.arange(H) + 0.5 - H / 2.0) / fy
    X, Y = np.meshgrid(xs, ys)
    ch, sh = math.cos(heading), math.sin(heading)
    return ch * X - sh, -Y, -sh * X - ch


def render_depth(room: Room, pos: np.ndarray, heading: float, H: int = 480, W: int = 640,
                 vfov: float = math.radians(VFOV_DEG)) -> np.ndarray:
    dx, dy, dz = _ray_dirs(H, W, vfov, heading)
    px, py, pz = (float(v) for v in pos)
    big = np.float64(1e30)
    t = np.full((H, W), big)

    def plane(num, den):
        with np.errstate(divide="ignore", invalid="ignore"):
            tt = num / den
        return np.where((den != 0) & (tt > 1e-6), tt, big)

    t = np.minimum(t, plane(0.0 - py, dy))
    t = np.minimum(t, plane(room.ceil - py, dy))
    t = np.minimum(t, plane(room.x0 - px, dx))
    t = np.minimum(t, plane(room.x1 - px, dx))
    t = np.minimum(t, plane(room.z0 - pz, dz))
    t = np.minimum(t, plane(room.z1 - pz, dz))
    for b in room.boxes:                       # slab test per axis-aligned box
        with np.errstate(divide="ignore", invalid="ignore"):
            tx0, tx1 = (b[0] - px) / dx, (b[1] - px) / dx
            ty0, ty1 = (b[2] - py) / dy, (b[3] - py) / dy
            tz0, tz1 = (b[4] - pz) / dz, (b[5] - pz) / dz
        tn = np.maximum(np.maximum(np.minimum(tx0, tx1), np.minimum(ty0, ty1)), np.minimum(tz0, tz1))
        tf = np.minimum(np.minimum(np.maximum(tx0, tx1), np.maximum(ty0, ty1)), np.maximum(tz0, tz1))
        hit = (tf >= tn) & (tn > 1e-6) & np.isfinite(tn)
        t = np.where(hit, np.minimum(t, tn), t)
    return np.minimum(t, MAX_DEPTH).astype(np.float32)


def random_walk(rng: np.random.Generator, room: Room, n_frames: int, y: float = SENSOR_HEIGHT) -> np.ndarray:
    """(n_frames, 5) f32 xyzhe with elevation + pi (build_data.py:190-196)."""
    m = 1.0
    x, z = rng.uniform(room.x0 + m, room.x1 - m), rng.uniform(room.z0 + m, room.z1 - m)
    h = rng.uniform(0.0, 2.0 * math.pi)
    out = np.zeros((n_frames, 5), np.float64)
    for t in range(n_frames):
        out[t] = (x, y, z, h, math.pi)
        if rng.uniform() < 0.6:
            nx, nz = x - 0.1 * math.sin(h), z - 0.1 * math.cos(h)
            if room.x0 + m < nx < room.x1 - m and room.z0 + m < nz < room.z1 - m:
                x, z = nx, nz
            else:
                h = (h + math.radians(9.0)) % (2.0 * math.pi)
        else:
            h = (h + math.radians(9.0) * (1.0 if rng.uniform() < 0.5 else -1.0)) % (2.0 * math.pi)
    return out.astype(np.float32)


def make_episode(seed: int, n_frames: int = 20, H: int = 480, W: int = 640, map_w: int = 500, map_h: int = 500,
                 cell: float = 0.02 * 10, room_size: Tuple[float, float] = (12.0, 9.0), zero_frac: float = 0.02,
                 vfov: float = math.radians(VFOV_DEG)) -> Episode:
    """Episode ``seed`` (the bench uses 1234 + episode id).  The room is centred in the map, so with the
    default 0.2 m cells every point is in-map; with 0.02 m cells and a room larger than the map the
    clip-to-border (build_memory_data.py:141-142) and out-of-map mask (core.py:258-261) paths are hit."""
    rng = np.random.default_rng(seed)
    room = make_room(rng, *room_size)
    poses = random_walk(rng, room, n_frames)
    depth = np.empty((n_frames, H, W), np.float32)
    for t in range(n_frames):
        d = render_depth(room, poses[t, :3].astype(np.float64), float(poses[t, 3]), H, W, vfov)
        if zero_frac > 0:
            holes = rng.uniform(size=(H // 8 + 1, W // 8 + 1)) < zero_frac      # 8x8 no-depth patches
            d = np.where(np.kron(holes, np.ones((8, 8), bool))[:H, :W], np.float32(0), d)
        depth[t] = d
    return Episode(depth, poses, map_shift(room, map_w, map_h, cell), float(cell), map_w, map_h, room, seed)


def map_shift(room: Room, map_w: int, map_h: int, cell: float) -> np.ndarray:
    """map_world_shift (3,) f32 that centres the room in a map_w x map_h grid of ``cell`` metres (the same depth maps can be
    replayed into grids of another size or resolution: only this shift changes)."""
    span_x, span_z = map_w * cell, map_h * cell
    return np.array([(room.x0 + room.x1) / 2 - span_x / 2, 0.0, (room.z0 + room.z1) / 2 - span_z / 2], np.float32)


class DeviceEpisodes:
    """Synthetic episodes rendered ON THE DEVICE, one frame-step at a time (torch, fp32): for workloads whose depth maps would
    not fit or would take minutes to ray-cast on the host (BASELINE configs[4]: 512 episodes x 100 frames).  Same scene model
    as ``make_episode`` (room + boxes + random walk + 8x8 no-depth patches); values differ from the numpy renderer in the
    last bits (fp32 vs fp64 ray parameters), which is irrelevant for synthetic inputs.  Input generation, not the hot path."""

    def __init__(self, n_episodes: int, n_frames: int, device, H: int = 480, W: int = 640, map_w: int = 500, map_h: int = 500,
                 cell: float = 0.2, seed0: int = 1234, zero_frac: float = 0.02, vfov: float = math.radians(VFOV_DEG), n_boxes: int = 5):
        import torch
        self.torch, self.device = torch, torch.device(device)
        self.n, self.T, self.H, self.W, self.zero_frac = n_episodes, n_frames, H, W, zero_frac
        rooms, poses = [], []
        for e in range(n_episodes):
            rng = np.random.default_rng(seed0 + e)
            room = make_room(rng, n_boxes=n_boxes)
            rooms.append(room)
            poses.append(random_walk(rng, room, n_frames))
        self.xyzhe = np.stack(poses)                                                       # (n, T, 5) f32
        self.shift = np.stack([map_shift(r, map_w, map_h, cell) for r in rooms])           # (n, 3) f32
        self.walls = torch.tensor([[r.x0, r.x1, r.z0, r.z1, r.ceil] for r in rooms], dtype=torch.float32, device=self.device)
        self.boxes = torch.tensor(np.stack([r.boxes for r in rooms]), dtype=torch.float32, device=self.device)      # (n, nb, 6)
        hfov = W / H * vfov
        fx, fy = W / (2.0 * math.tan(hfov / 2.0)), H / (2.0 * math.tan(vfov / 2.0))
        xs = (torch.arange(W, device=self.device, dtype=torch.float32) + 0.5 - W / 2.0) / fx
        ys = (torch.arange(H, device=self.device, dtype=torch.float32) + 0.5 - H / 2.0) / fy
        self.X, self.Y = xs[None, None, :], ys[None, :, None]
        self.gen = torch.Generator(device=self.device).manual_seed(seed0)

    def render(self, episodes, frames, out=None):
        """depth (R,H,W) f32 for slot s = frame frames[s] of episode episodes[s]."""
        torch = self.torch
        ep = torch.as_tensor(np.asarray(episodes, np.int64), device=self.device)
        pose = torch.from_numpy(self.xyzhe[np.asarray(episodes), np.asarray(frames)]).to(self.device)            # (R,5)
        px, py, pz, h = (pose[:, k, None, None] for k in range(4))
        ch, sh = torch.cos(h), torch.sin(h)
        dx, dy, dz = ch * self.X - sh, (-self.Y).expand(len(episodes), -1, self.W), -sh * self.X - ch
        big = torch.tensor(1e30, device=self.device)
        wl = self.walls[ep]

        def plane(num, den):
            tt = num / den
            return torch.where((den != 0) & (tt > 1e-6), tt, big)

        t = plane(0.0 - py, dy)
        t = torch.minimum(t, plane(wl[:, 4, None, None] - py, dy))
        t = torch.minimum(t, plane(wl[:, 0, None, None] - px, dx))
        t = torch.minimum(t, plane(wl[:, 1, None, None] - px, dx))
        t = torch.minimum(t, plane(wl[:, 2, None, None] - pz, dz))
        t = torch.minimum(t, plane(wl[:, 3, None, None] - pz, dz))
        bx = self.boxes[ep]
        for b in range(bx.shape[1]):
            c = [bx[:, b, k, None, None] for k in range(6)]
            tx0, tx1 = (c[0] - px) / dx, (c[1] - px) / dx
            ty0, ty1 = (c[2] - py) / dy, (c[3] - py) / dy
            tz0, tz1 = (c[4] - pz) / dz, (c[5] - pz) / dz
            tn = torch.maximum(torch.maximum(torch.minimum(tx0, tx1), torch.minimum(ty0, ty1)), torch.minimum(tz0, tz1))
            tf = torch.minimum(torch.minimum(torch.maximum(tx0, tx1), torch.maximum(ty0, ty1)), torch.maximum(tz0, tz1))
            hit = (tf >= tn) & (tn > 1e-6) & torch.isfinite(tn)
            t = torch.where(hit, torch.minimum(t, tn), t)
        d = torch.clamp(t, max=MAX_DEPTH)
        if self.zero_frac > 0:
            holes = torch.rand((len(episodes), self.H // 8 + 1, self.W // 8 + 1), device=self.device, generator=self.gen) < self.zero_frac
            holes = holes.repeat_interleave(8, 1).repeat_interleave(8, 2)[:, : self.H, : self.W]
            d = torch.where(holes, torch.zeros((), device=self.device), d)
        if out is not None:
            out.copy_(d)
            return out
        return d.contiguous()


def make_detections(rng: np.random.Generator, H: int = 480, W: int = 640, C: int = 512, k_range=(4, 16)):
    """Kept detections of one frame as the write consumes them (custom_rcnn.py:848,880): box_features
    (K,C) = 50 * normalize(N(0,1)), masks (K,H,W) bool (rectangles / ellipses 40-200 px)."""
    K = int(rng.integers(k_range[0], k_range[1] + 1))
    f = rng.standard_normal((K, C)).astype(np.float32)
    f = (50.0 * f / np.linalg.norm(f, axis=1, keepdims=True)).astype(np.float32)
    masks = np.zeros((K, H, W), bool)
    vv, uu = np.mgrid[0:H, 0:W]
    for k in range(K):
        hw, hh = rng.integers(20, 101), rng.integers(20, 101)
        cu, cv = rng.integers(0, W), rng.integers(0, H)
        if rng.uniform() < 0.5:
            masks[k] = (np.abs(uu - cu) <= hw) & (np.abs(vv - cv) <= hh)
        else:
            masks[k] = ((uu - cu) / hw) ** 2 + ((vv - cv) / hh) ** 2 <= 1.0
    return f, masks


def make_mask_head_detections(rng: np.random.Generator, H: int = 480, W: int = 640, C: int = 512, k_range=(4, 16), S: int = 28,
                              edge_cases: bool = False):
    """Kept detections as the detector hands them over BEFORE pasting (custom_rcnn.py:876-880): box_features (K,C) =
    50 * normalize(N(0,1)), mask_probs (K,S,S) f32 = sigmoid of a smooth blob logit (what mask_rcnn_inference leaves in
    pred_masks), boxes (K,4) f32 XYXY in pixels with fractional corners.  edge_cases adds boxes that stick out of the
    image, sub-pixel boxes and a saturated mask."""
    K = int(rng.integers(k_range[0], k_range[1] + 1))
    f = rng.standard_normal((K, C)).astype(np.float32)
    f = (50.0 * f / np.linalg.norm(f, axis=1, keepdims=True)).astype(np.float32)
    yy, xx = np.mgrid[0:S, 0:S].astype(np.float32)
    probs = np.empty((K, S, S), np.float32)
    boxes = np.empty((K, 4), np.float32)
    for k in range(K):
        cx, cy = rng.uniform(0.3 * S, 0.7 * S, 2)
        rx, ry = rng.uniform(0.2 * S, 0.55 * S, 2)
        logit = 6.0 * (1.0 - ((xx - cx) / rx) ** 2 - ((yy - cy) / ry) ** 2) + rng.standard_normal((S, S)) * 0.5
        probs[k] = (1.0 / (1.0 + np.exp(-logit))).astype(np.float32)
        bw, bh = rng.uniform(0.06 * W, 0.35 * W), rng.uniform(0.08 * H, 0.45 * H)
        x0, y0 = rng.uniform(-0.1 * W, 0.9 * W) if edge_cases else rng.uniform(0, W - bw), \
            rng.uniform(-0.1 * H, 0.9 * H) if edge_cases else rng.uniform(0, H - bh)
        boxes[k] = (x0, y0, x0 + bw, y0 + bh)
    if edge_cases and K >= 3:
        boxes[0] = (W * 0.5 + 0.25, H * 0.5 + 0.25, W * 0.5 + 0.75, H * 0.5 + 0.6)      # thinner than a pixel
        probs[1] = 1.0                                                                  # saturated: value == 0.5 on the box edge
        boxes[1] = (np.floor(W * 0.25), np.floor(H * 0.25), np.floor(W * 0.25) + 28.0, np.floor(H * 0.25) + 56.0)
        boxes[2] = (-30.5, -12.25, W + 17.0, H + 3.5)                                   # larger than the image
    return f, probs, boxes
