"""Deterministic synthetic inputs for the hot path.

Pure numpy, seeded; used by tests/ and bench.py on both the oracle and the CUDA
path so they see byte-identical inputs.  Nothing here is part of the product's
compute path and nothing here imports the oracle.

Scene: the viewer's default camera (renderer/camera.hpp:105-108, application/
scene_viewer_application.cpp:347,362: fovy = pi/4, near = 1/16, infinite reverse-Z,
eye (0,0,8) -> look (0,0,0)), a bumpy ground plane, a far wall and a few spheres so
view-Z spans ~2..190 m with ~10 % sky (depth == 0).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

FLT_MAX = float(np.finfo(np.float32).max)
GBUFFER_SEED = 0x6B0F
LIGHT_SEED = 0x1167

CLUSTER_RES = (128, 64, 4096)  # application/scene_viewer_application.cpp:407


def perspective_inf(fovy: float, aspect: float, near: float) -> np.ndarray:
    """math/muglm/muglm.cpp:319-345 with far == InfiniteFarPlane. Column-major 4x4 (m[c, r])."""
    t = np.tan(np.float32(fovy) / np.float32(2.0), dtype=np.float32)
    m = np.zeros((4, 4), np.float32)
    m[0, 0] = np.float32(1.0) / (np.float32(aspect) * t)
    m[1, 1] = -(np.float32(1.0) / t)
    m[0, 1] = m[2, 1] = m[3, 1] = np.float32(-0.0)  # the reference's "result[i].y *= -1" on zeros
    m[3, 2] = np.float32(near)
    m[2, 3] = np.float32(-1.0)
    return m


def look_at_view(eye, at, up=(0.0, 1.0, 0.0)) -> np.ndarray:
    """Right-handed view matrix (column-major m[c, r]); identity rotation for the default pose."""
    eye = np.asarray(eye, np.float64)
    f = np.asarray(at, np.float64) - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, np.asarray(up, np.float64))
    r /= np.linalg.norm(r)
    u = np.cross(r, f)
    rot = np.stack([r, u, -f])  # rows
    m = np.eye(4)
    m[:3, :3] = rot
    m[:3, 3] = -rot @ eye
    return np.ascontiguousarray(m.T.astype(np.float32)) + np.float32(0.0)  # column-major out[c, r]; -0 -> +0


@dataclass
class Scene:
    width: int
    height: int
    projection: np.ndarray  # (4,4) f32 column-major [c, r]
    view: np.ndarray
    albedo: np.ndarray      # (H,W) u32 R8G8B8A8_SRGB
    normal: np.ndarray      # (H,W) u32 A2B10G10R10_UNORM
    pbr: np.ndarray         # (H,W) u16 R8G8_UNORM
    depth: np.ndarray       # (H,W) f32 reverse-Z, 0 = sky
    emissive: np.ndarray    # (H,W) u32 B10G11R11_UFLOAT
    world_pos: np.ndarray = field(repr=False, default=None)  # (H,W,3) f64, helper for light placement
    world_nrm: np.ndarray = field(repr=False, default=None)
    dir_color: tuple = (6.0, 5.5, 4.5)  # scene_viewer_application.cpp:380
    dir_direction: tuple = field(default_factory=lambda: tuple(
        (np.array([0.3, 0.8, 0.5]) / np.linalg.norm([0.3, 0.8, 0.5])).astype(np.float32).tolist()))


def _pack_ufloat(v: np.ndarray, mbits: int) -> np.ndarray:
    """float32 -> unsigned small float (5-bit exponent), truncating; inputs here are finite >= 0."""
    x = v.astype(np.float32).view(np.uint32).astype(np.int64)
    e = (x >> 23) - 127
    m = (x & 0x7FFFFF) | 0x800000
    normal = ((e + 15) << mbits) | ((m >> (23 - mbits)) & ((1 << mbits) - 1))
    shift = np.clip((23 - mbits) + (-14 - e), 0, 31)
    denorm = np.where(shift > 24, 0, m >> shift)
    out = np.where(e >= -14, normal, denorm)
    out = np.where(e > 15, (30 << mbits) | ((1 << mbits) - 1), out)
    out = np.where(v <= 0, 0, out)
    return out.astype(np.uint32)


def pack_r11g11b10(rgb: np.ndarray) -> np.ndarray:
    return (_pack_ufloat(rgb[..., 0], 6) | (_pack_ufloat(rgb[..., 1], 6) << np.uint32(11))
            | (_pack_ufloat(rgb[..., 2], 5) << np.uint32(22))).astype(np.uint32)


def make_scene(width: int, height: int, seed: int = GBUFFER_SEED) -> Scene:
    rng = np.random.default_rng(seed)
    aspect = width / height
    proj = perspective_inf(math.pi / 4.0, aspect, 1.0 / 16.0)
    eye = np.array([0.0, 0.0, 8.0])
    view = look_at_view(eye, (0.0, 0.0, 0.0))

    # primary rays through pixel centres (y down, projection carries the Y flip)
    xs = (np.arange(width) + 0.5) / width * 2.0 - 1.0
    ys = (np.arange(height) + 0.5) / height * 2.0 - 1.0
    tan_half = math.tan(math.pi / 8.0)
    dx = xs[None, :] * tan_half * aspect
    dy = -ys[:, None] * tan_half  # ndc y = -1 is the top row, which looks up
    dirs = np.stack(np.broadcast_arrays(dx, dy, -np.ones_like(dx * dy)), -1).astype(np.float64)
    # dirs are in view space with view_depth == 1 per unit t (|z| component is 1)

    t_hit = np.full((height, width), np.inf)
    nrm = np.zeros((height, width, 3))
    # ground plane y = -2 (world == view rotation here)
    with np.errstate(divide="ignore", invalid="ignore"):
        t_g = (-2.0 - eye[1]) / dirs[..., 1]
    ok = (dirs[..., 1] < 0) & (t_g > 0)
    t_hit = np.where(ok, t_g, t_hit)
    nrm[ok] = (0.0, 1.0, 0.0)
    # far wall z = -180 (view depth 188), up to y = 62
    t_w = (-180.0 - eye[2]) / dirs[..., 2]
    y_w = eye[1] + t_w * dirs[..., 1]
    ok = (t_w < t_hit) & (y_w < 62.0)
    t_hit = np.where(ok, t_w, t_hit)
    nrm[ok] = (0.0, 0.0, 1.0)
    # spheres resting on the ground
    n_sph = 24
    sph_r = rng.uniform(1.0, 6.0, n_sph)
    sph_z = -rng.uniform(4.0, 150.0, n_sph)
    sph_x = rng.uniform(-0.6, 0.6, n_sph) * (eye[2] - sph_z) * tan_half * aspect
    a_all = np.einsum("...k,...k", dirs, dirs)
    for cx, cz, r in zip(sph_x, sph_z, sph_r):
        c = np.array([cx, -2.0 + r, cz])
        oc = eye - c
        # conservative screen-space box of the sphere so only its pixels are intersected
        dz = eye[2] - cz
        if dz > r + 0.5:
            k = 1.2 * r / (dz - r)
            px0 = ((-oc[0] / dz - k) / (tan_half * aspect) * 0.5 + 0.5) * width
            px1 = ((-oc[0] / dz + k) / (tan_half * aspect) * 0.5 + 0.5) * width
            py0 = ((oc[1] / dz - k) / tan_half * 0.5 + 0.5) * height
            py1 = ((oc[1] / dz + k) / tan_half * 0.5 + 0.5) * height
            x0, x1 = int(max(px0 - 2, 0)), int(min(px1 + 3, width))
            y0, y1 = int(max(py0 - 2, 0)), int(min(py1 + 3, height))
        else:
            x0, x1, y0, y1 = 0, width, 0, height
        if x1 <= x0 or y1 <= y0:
            continue
        sub = (slice(y0, y1), slice(x0, x1))
        d_sub = dirs[sub]
        b = d_sub @ oc
        a = a_all[sub]
        cc = oc @ oc - r * r
        disc = b * b - a * cc
        with np.errstate(invalid="ignore"):
            t_s = (-b - np.sqrt(disc)) / a
        ok = (disc > 0) & (t_s > 0.2) & (t_s < t_hit[sub])
        t_hit[sub] = np.where(ok, t_s, t_hit[sub])
        p = eye + t_s[..., None] * d_sub
        n = (p - c) / r
        nrm[sub] = np.where(ok[..., None], n, nrm[sub])

    sky = ~np.isfinite(t_hit)
    t_safe = np.where(sky, 1.0, t_hit)
    pos = eye + t_safe[..., None] * dirs
    # analytic bump on the normals (ground/wall only get a gentle ripple)
    ripple = 0.25 * np.stack([np.sin(pos[..., 0] * 0.9 + pos[..., 2] * 0.3),
                              np.zeros_like(t_safe),
                              np.cos(pos[..., 2] * 0.7 - pos[..., 0] * 0.2)], -1)
    nrm = nrm + ripple * (np.abs(nrm[..., 1:2]) > 0.5)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=-1, keepdims=True), 1e-9)

    # reverse-Z infinite: depth = near / view_depth ; view_depth == t for these rays
    depth = np.where(sky, 0.0, (1.0 / 16.0) / t_safe).astype(np.float32)

    n10 = np.clip(np.rint((nrm * 0.5 + 0.5) * 1023.0), 0, 1023).astype(np.uint32)
    normal = (n10[..., 0] | (n10[..., 1] << np.uint32(10)) | (n10[..., 2] << np.uint32(20)) | np.uint32(3 << 30)).astype(np.uint32)

    alb = rng.integers(13, 243, size=(height, width, 3), dtype=np.uint32)  # ~[0.05, 0.95] in sRGB8
    albedo = (alb[..., 0] | (alb[..., 1] << np.uint32(8)) | (alb[..., 2] << np.uint32(16)) | np.uint32(0xFF000000)).astype(np.uint32)
    metallic = np.where(rng.random((height, width)) < 0.2, 255, 0).astype(np.uint16)
    rough = rng.integers(26, 256, size=(height, width), dtype=np.uint16)
    pbr = (metallic | (rough << np.uint16(8))).astype(np.uint16)

    emis = np.zeros((height, width, 3), np.float32)
    hot = rng.random((height, width)) < 0.001
    emis[hot] = rng.uniform(0.0, 50.0, size=(int(hot.sum()), 3)).astype(np.float32)
    emissive = pack_r11g11b10(emis)

    return Scene(width, height, proj, view, albedo, normal, pbr, depth, emissive,
                 world_pos=np.where(sky[..., None], np.nan, pos), world_nrm=nrm)


@dataclass
class Lights:
    """Raw light descriptions (what the application owns); host prep turns them into
    PositionalFragmentInfo records, sorted front-to-back."""
    color: np.ndarray      # (N,3) f32
    position: np.ndarray   # (N,3) f32
    is_point: np.ndarray   # (N,) bool
    rot: np.ndarray        # (N,3,3) f32 column-major [c, r] node rotation (spots)
    inner_cone: np.ndarray  # (N,) f32 (cos of half angle)
    outer_cone: np.ndarray


def make_lights(n: int, spot_fraction: float = 0.0, seed: int = LIGHT_SEED,
                aspect: float = 16.0 / 9.0) -> Lights:
    """N lights uniform in the visible frustum volume z in [2,150] m, snapped to hug the ground
    (+0.5..3 m), colour uniform [0.5, 20]^3 (=> falloff radius ~2.2..14 m). Distinct sort keys."""
    rng = np.random.default_rng(seed + n * 7 + int(spot_fraction * 1000))
    tan_half = math.tan(math.pi / 8.0)
    # uniform in frustum volume: p(z) ~ z^2
    z = (rng.random(n) * (150.0 ** 3 - 2.0 ** 3) + 2.0 ** 3) ** (1.0 / 3.0)
    z = np.sort(z) + np.arange(n) * 1e-4  # distinct, already front-to-back
    x = rng.uniform(-1.0, 1.0, n) * z * tan_half * aspect
    y = -2.0 + rng.uniform(0.5, 3.0, n)
    pos = np.stack([x, y, 8.0 - z], -1).astype(np.float32)
    color = rng.uniform(0.5, 20.0, size=(n, 3)).astype(np.float32)
    is_point = rng.random(n) >= spot_fraction
    # spot orientation: mostly downwards with some tilt
    d = np.stack([rng.uniform(-0.5, 0.5, n), -np.ones(n), rng.uniform(-0.5, 0.5, n)], -1)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    rot = np.zeros((n, 3, 3), np.float32)
    for i in range(n):
        f = d[i]
        up = np.array([0.0, 0.0, 1.0]) if abs(f[1]) > 0.9 else np.array([0.0, 1.0, 0.0])
        r = np.cross(f, up); r /= np.linalg.norm(r)
        u = np.cross(r, f)
        rot[i, 0] = r; rot[i, 1] = u; rot[i, 2] = -f  # columns: right, up, -forward
    outer = rng.uniform(0.6, 0.9, n).astype(np.float32)
    inner = np.minimum(outer + rng.uniform(0.02, 0.08, n), 0.999).astype(np.float32)
    return Lights(color, pos, is_point, rot, inner, outer)
