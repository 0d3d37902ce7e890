"""Seeded synthetic LiDAR frames of the shapes BASELINE.json names.

A spinning sensor at the origin: ``n_beams`` elevations x ``n_az`` azimuth steps are ray-cast against
a ground plane at z = -1.8 m and ``n_boxes`` car-sized oriented boxes; rays that hit nothing return a
far wall so every ray yields a point (Waymo-like: 64 x 2812 = 180k returns; nuScenes-like:
32 x 1090 = 35k).  Frame f of a stream uses seed 1000 + f.  numpy only — input generation is not
part of any timed region.
"""
import numpy as np

GROUND_Z = -1.8


def make_boxes(rng, n_boxes, r_min=5.0, r_max=75.0):
    """(n,7) [x,y,z,dx,dy,dz,heading] car boxes on the ground, centres >= 6 m apart."""
    boxes = []
    tries = 0
    while len(boxes) < n_boxes and tries < 100000:
        tries += 1
        r = rng.uniform(r_min, r_max)
        a = rng.uniform(-np.pi, np.pi)
        size = np.array([4.2, 2.0, 1.6]) * (1.0 + 0.1 * np.clip(rng.standard_normal(3), -2, 2))
        c = np.array([r * np.cos(a), r * np.sin(a), GROUND_Z + size[2] / 2])
        if all(np.hypot(c[0] - b[0], c[1] - b[1]) > 6.0 for b in boxes):
            boxes.append(np.concatenate([c, size, [rng.uniform(-np.pi, np.pi)]]))
    return np.asarray(boxes, dtype=np.float32).reshape(-1, 7)


def make_frame(seed, n_beams=64, n_az=2812, n_boxes=50, r_max=75.0, box_r_max=50.0, el_lo=-17.6, el_hi=2.4, noise=0.01):
    """-> points (n_beams*n_az, 3) float32, boxes (n_boxes, 7) float32"""
    rng = np.random.default_rng(seed)
    boxes = make_boxes(rng, n_boxes, r_max=box_r_max)   # cars occlude each other: ~40 of 50 keep >= 30 returns
    el = np.deg2rad(np.linspace(el_lo, el_hi, n_beams))
    az = np.linspace(-np.pi, np.pi, n_az, endpoint=False) + rng.uniform(0, 2 * np.pi / n_az)
    el, az = np.meshgrid(el, az, indexing="ij")
    d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=-1).reshape(-1, 3)
    t = np.full(len(d), r_max * 1.05)
    down = d[:, 2] < -1e-6
    t[down] = np.minimum(t[down], GROUND_Z / d[down, 2])
    for b in boxes.astype(np.float64):   # slab test in the box frame (origin ray)
        c, s = np.cos(b[6]), np.sin(b[6])
        R = np.array([[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]])   # world -> box
        o = -(R @ b[:3])
        dl = d @ R.T
        half = b[3:6] / 2
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = (-half - o) / dl
            t2 = (half - o) / dl
        tn = np.nanmax(np.minimum(t1, t2), axis=1)
        tf = np.nanmin(np.maximum(t1, t2), axis=1)
        hit = (tn <= tf) & (tn > 0) & (tn < t)
        t[hit] = tn[hit]
    t = t + noise * rng.standard_normal(len(t))
    pts = (d * t[:, None]).astype(np.float32)
    return pts, boxes


def make_stream(n_frames, first_seed=1000, **kw):
    """-> points (F, P, 3), boxes (F, T, 7)"""
    frames = [make_frame(first_seed + f, **kw) for f in range(n_frames)]
    return np.stack([f[0] for f in frames]), np.stack([f[1] for f in frames])


def rotate_stream(points, boxes, angle):
    """The same frames seen by a sensor mounted with a yaw offset: points (F,P,3) and boxes (F,T,7) rotated about z by
    ``angle`` (rigid, so every point keeps its box).  A cheap way to get many distinct frames with identical statistics."""
    c, s = np.cos(angle), np.sin(angle)
    p = points.astype(np.float64)
    out_p = np.stack([p[..., 0] * c - p[..., 1] * s, p[..., 0] * s + p[..., 1] * c, p[..., 2]], axis=-1).astype(np.float32)
    b = boxes.astype(np.float64).copy()
    b[..., 0], b[..., 1] = boxes[..., 0] * c - boxes[..., 1] * s, boxes[..., 0] * s + boxes[..., 1] * c
    b[..., 6] = (boxes[..., 6] + angle + np.pi) % (2 * np.pi) - np.pi
    return out_p, b.astype(np.float32)


def make_object_clouds(seed, n_obj, n_in=1024, n_dense=0):
    """Per-object clouds for the MLP / FPS / kNN stages: ``n_in`` points on the two sensor-facing
    faces of a car box (+1 cm noise) and optionally an ``n_dense``-point full surface (C4)."""
    rng = np.random.default_rng(seed)
    boxes = make_boxes(rng, n_obj, r_min=6.0, r_max=60.0) if n_obj <= 100 else np.concatenate(
        [make_boxes(rng, 100, r_min=6.0, r_max=70.0) for _ in range((n_obj + 99) // 100)])[:n_obj]
    part = np.empty((n_obj, n_in, 3), np.float32)
    dense = np.empty((n_obj, n_dense, 3), np.float32) if n_dense else None
    for i, b in enumerate(boxes.astype(np.float64)):
        c, s = np.cos(b[6]), np.sin(b[6])
        R = np.array([[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]])
        view = R @ (-b[:3])           # sensor direction in the box frame
        u = rng.uniform(-0.5, 0.5, (n_in, 3)) * b[3:6]
        face = rng.integers(0, 2, n_in)
        u[face == 0, 0] = np.sign(view[0]) * b[3] / 2
        u[face == 1, 1] = np.sign(view[1]) * b[4] / 2
        part[i] = (u @ R + b[:3] + 0.01 * rng.standard_normal((n_in, 3))).astype(np.float32)
        if n_dense:
            v = rng.uniform(-0.5, 0.5, (n_dense, 3)) * b[3:6]
            ax = rng.integers(0, 3, n_dense)
            sgn = rng.choice([-1.0, 1.0], n_dense)
            v[np.arange(n_dense), ax] = sgn * b[3:6][ax] / 2
            dense[i] = (v @ R + b[:3]).astype(np.float32)
    return part, dense, boxes
