"""Deterministic synthetic TSDF pairs: TSDF = clip(signed distance / 10 voxels, -1, 1),
far field exactly +-1.0 as produced by the reference's generators (tsdf/generation.py:238,282)."""
import numpy as np

NARROW_BAND_HALF_WIDTH = 10.0


def sphere_plane_pair_3d(size=128, shift=(2.5, -1.5, 1.0), radius_scale=1.04, plane_shift=-2.5, xp=np, device=None,
                         planes=None):
    """C2 geometry: canonical = sphere(centre size/2, r = 0.3 size) U half-space below the plane axis0 = 0.75 size;
    live = the same scene with the sphere moved by `shift` (scaled by size/128), grown by `radius_scale` and the
    plane moved by `plane_shift`. Returns (canonical, live) float32 [size]^3. `xp` is numpy or torch.
    `planes=(lo, hi)` generates only the axis-0 planes [lo, hi) (slab-sharded volumes: identical values)."""
    s = size / 128.0
    lo, hi = (0, size) if planes is None else planes
    if xp is np:
        axis = np.arange(size, dtype=np.float32)
        i, j, k = np.meshgrid(axis[lo:hi], axis, axis, indexing="ij")
        sqrt, minimum, clip = np.sqrt, np.minimum, np.clip
    else:
        axis = xp.arange(size, dtype=xp.float32, device=device)
        i, j, k = xp.meshgrid(axis[lo:hi], axis, axis, indexing="ij")
        sqrt, minimum, clip = xp.sqrt, xp.minimum, xp.clamp

    def scene(cx, cy, cz, r, plane):
        sphere = sqrt((i - cx) ** 2 + (j - cy) ** 2 + (k - cz) ** 2) - r
        half_space = plane - i
        sd = minimum(sphere, half_space) / NARROW_BAND_HALF_WIDTH
        return clip(sd, -1.0, 1.0)

    c = 64.0 * s
    canonical = scene(c, c, c, 38.4 * s, 96.0 * s)
    live = scene(c + shift[0] * s, c + shift[1] * s, c + shift[2] * s, 38.4 * s * radius_scale,
                 (96.0 + plane_shift) * s)
    if xp is np:
        return canonical.astype(np.float32), live.astype(np.float32)
    return canonical.float().contiguous(), live.float().contiguous()


def circle_line_pair_2d(size=128, shift=(2.0, -1.2), radius_scale=1.04, line_shift=-1.3):
    """2D analogue (numpy): circle U half-plane."""
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    s = size / 128.0

    def scene(cx, cy, r, line):
        circle = np.sqrt((xx - cx) ** 2 + (yy - cy) ** 2) - r
        half_plane = line - yy
        return np.clip(np.minimum(circle, half_plane) / NARROW_BAND_HALF_WIDTH, -1.0, 1.0).astype(np.float32)

    c = 64.0 * s
    canonical = scene(c, 58.0 * s, 28.0 * s, 102.0 * s)
    live = scene(c + shift[0] * s, 58.0 * s + shift[1] * s, 28.0 * s * radius_scale, (102.0 + line_shift) * s)
    return canonical, live


def multipair_batch_3d(pair_count, size=128, seed=1234):
    """C4: `pair_count` independent pairs, per-pair shift ~U[-3,3]^3 and radius scale ~U[0.95,1.05]."""
    rng = np.random.default_rng(seed)
    shifts = rng.uniform(-3.0, 3.0, size=(pair_count, 3))
    scales = rng.uniform(0.95, 1.05, size=pair_count)
    canonicals, lives = [], []
    for p in range(pair_count):
        canonical, live = sphere_plane_pair_3d(size, tuple(shifts[p]), float(scales[p]))
        canonicals.append(canonical)
        lives.append(live)
    return np.stack(canonicals), np.stack(lives)


# Sobolev kernel generator (reference nonrigid_opt/slavcheva/sobolev_filter.py:208-252). The reference solves
# (I - strength * L) S = e in float32, where its 7-point-stencil matrix L (sobolev_filter.py:136-159) links every
# voxel to the flat indices +-1, +-size, +-size^2 whenever they fall inside [0, size^3) -- i.e. neighbours wrap
# around row ends -- and takes the leading mode-1 factor of S (sktensor HOOI == leading left singular vector of
# the mode unfolding). SURVEY.md F15: this reproduces the reference's hard-coded 7-tap and 3-tap constants
# (math_utils/convolution.py:20-26, cpp/tests/test_slavcheva_optimizer.cpp:297-298) bit for bit.
def sobolev_kernel_1d(size=7, strength=0.1, precision=np.float32, mode=0):
    """mode=0 reproduces the reference's 7-tap default bit for bit; mode=1 its 3-tap test kernel."""
    n3 = size ** 3
    laplacian = np.zeros((n3, n3), dtype=precision)
    for voxel in range(n3):
        laplacian[voxel, voxel] = -6.0
        for offset in (-1, 1, size, -size, -size * size, size * size):
            neighbour = voxel + offset
            if 0 <= neighbour < n3:
                laplacian[voxel, neighbour] = 1.0
    one_hot = np.zeros((n3, 1), dtype=precision)
    one_hot[n3 // 2] = 1.0
    solution = np.linalg.solve(np.identity(n3, dtype=precision) - precision(strength) * laplacian, one_hot)
    kernel3d = solution.reshape(size, size, size)
    unfolding = np.moveaxis(kernel3d, mode, 0).reshape(size, size * size)
    u, _, _ = np.linalg.svd(unfolding, full_matrices=False)
    return np.abs(u[:, 0]).astype(np.float32)
