"""Synthetic COCO-shape inputs for the box-fitting path.

``depth[B,H,W]`` float32: tilted plane ``z0 + ax*u + ay*v`` (z0 ~ U[2,6] m, at
most 1 m of tilt across the image), one Gaussian bump per instance (sigma =
0.25 x mask radius, 0.2-1.0 m towards the camera), +-5 mm uniform noise, and 3 %
of the pixels that lie outside every mask set to the ``10000.0`` sentinel the
reference's depth stage writes for invalid pixels
(``src/batch_scripts/depth.py:82`` of the reference).

``K[B,3,3]`` float64: ``fx = fy = 0.9 W``, ``cx = W/2``, ``cy = H/2`` (the shape
MoGe produces, ``external/MoGe/infer_moge.py:29-30`` of the reference).

``masks[B,I,H,W]`` bool: rotated ellipses, centre at least 10 px from the border,
area U[2 %, 12 %] of the image (overlaps allowed).

``ground[B,I,3]`` float64: ``normalise((0,-1,0) + N(0, 0.1^2))``.

Everything is produced with torch on the requested device from one seeded
generator, in chunks of images so the temporaries stay small.
"""

from __future__ import annotations

import math

import torch

SENTINEL = 10000.0


def make_inputs(B, H, W, I, seed=1234, device="cpu", chunk=16, area=(0.02, 0.12), with_ground=True,
                sentinel_frac=0.03):
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(int(seed))

    def U(shape, lo, hi, dtype=torch.float32):
        return torch.rand(shape, generator=gen, device=dev, dtype=dtype) * (hi - lo) + lo

    depth = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    masks = torch.empty((B, I, H, W), dtype=torch.bool, device=dev)
    K = torch.zeros((B, 3, 3), dtype=torch.float64, device=dev)
    K[:, 0, 0] = 0.9 * W
    K[:, 1, 1] = 0.9 * W
    K[:, 0, 2] = W / 2
    K[:, 1, 2] = H / 2
    K[:, 2, 2] = 1.0

    u = torch.arange(W, device=dev, dtype=torch.float32).view(1, 1, 1, W)
    v = torch.arange(H, device=dev, dtype=torch.float32).view(1, 1, H, 1)
    for b0 in range(0, B, chunk):
        n = min(chunk, B - b0)
        z0 = U((n, 1, 1), 2.0, 6.0)
        ax = U((n, 1, 1), -1.0, 1.0) / W
        ay = U((n, 1, 1), -1.0, 1.0) / H
        frac = U((n, I, 1, 1), area[0], area[1])
        aspect = U((n, I, 1, 1), 1.0, 3.0)
        theta = U((n, I, 1, 1), 0.0, math.pi)
        cx = U((n, I, 1, 1), 10.0, W - 10.0)
        cy = U((n, I, 1, 1), 10.0, H - 10.0)
        amp = U((n, I, 1, 1), 0.2, 1.0)
        semi_a = torch.sqrt(frac * (H * W) * aspect / math.pi)
        semi_b = semi_a / aspect
        du = u - cx
        dv = v - cy
        ct, st = torch.cos(theta), torch.sin(theta)
        xr = (du * ct + dv * st) / semi_a
        yr = (dv * ct - du * st) / semi_b
        m = (xr * xr + yr * yr) <= 1.0
        sigma = 0.25 * torch.sqrt(semi_a * semi_b)
        bump = (amp * torch.exp(-(du * du + dv * dv) / (2.0 * sigma * sigma))).sum(dim=1)
        d = z0 + ax * u[0] + ay * v[0] - bump
        d = d + U((n, H, W), -0.005, 0.005)
        if sentinel_frac > 0:
            outside = ~m.any(dim=1)
            hit = (torch.rand((n, H, W), generator=gen, device=dev) < sentinel_frac) & outside
            d = torch.where(hit, torch.full_like(d, SENTINEL), d)
        depth[b0:b0 + n] = d
        masks[b0:b0 + n] = m

    ground = None
    if with_ground:
        g = torch.randn((B, I, 3), generator=gen, device=dev, dtype=torch.float64) * 0.1
        g[..., 1] -= 1.0
        ground = g / g.norm(dim=-1, keepdim=True)
    return depth, K, masks, ground


# The benchmark configurations of BASELINE.json (index = position in "configs").
CONFIGS = {
    1: dict(B=1, H=480, W=640, I=4, method="pca", yaw_steps=0, gpus=1),
    2: dict(B=256, H=480, W=640, I=8, method="sweep", yaw_steps=36, gpus=1),
    3: dict(B=2048, H=480, W=640, I=10, method="pca", yaw_steps=0, gpus=8),
    4: dict(B=128, H=1536, W=1536, I=20, method="pca", yaw_steps=0, gpus=1),
    5: dict(B=1024, H=480, W=640, I=32, method="sweep", yaw_steps=360, gpus=8),
}
