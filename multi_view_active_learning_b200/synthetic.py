"""Seeded synthetic pools: ring-camera rigs, random skeletons, heatmap bump centres.

Host-side helpers shared by tests, ``bench.py`` and ``oracle/make_golden.py``.  Everything is
numpy + ``default_rng(seed)``; dense heatmaps for large pools are rendered on the device by
``mval_synth_heatmaps`` (csrc/synth.cu) from the bump centres produced here.
"""
import numpy as np

IMAGE_SIZE = 256
HEATMAP_SIZE = 64
STRIDE = 4


def ring_cameras(n_views, radius=3000.0, focal=600.0, rng=None, jitter=0.15):
    """[V, 3, 4] float64 projection matrices K [R|t] of cameras on a ring looking at the origin."""
    rng = np.random.default_rng(0) if rng is None else rng
    K = np.array([[focal, 0.0, IMAGE_SIZE / 2], [0.0, focal, IMAGE_SIZE / 2], [0.0, 0.0, 1.0]])
    P = np.zeros((n_views, 3, 4))
    for v in range(n_views):
        ang = 2 * np.pi * (v + jitter * rng.uniform(-1, 1)) / n_views
        elev = 0.35 * rng.uniform(-1, 1) + (0.25 if v % 2 else -0.1)
        r = radius * (1 + 0.1 * rng.uniform(-1, 1))
        c = np.array([r * np.cos(ang) * np.cos(elev), r * np.sin(ang) * np.cos(elev), r * np.sin(elev)])
        z = -c / np.linalg.norm(c)
        up = np.array([0.0, 0.0, 1.0])
        x = np.cross(z, up)
        x /= np.linalg.norm(x)
        y = np.cross(z, x)
        R = np.stack([x, y, z])
        t = -R @ c
        P[v] = K @ np.concatenate([R, t[:, None]], axis=1)
    return P


def project(P, X):
    """P [..., V, 3, 4], X [..., J, 3] -> [..., V, J, 2] pixel coordinates."""
    Xh = np.concatenate([X, np.ones(X.shape[:-1] + (1,))], axis=-1)
    ph = np.einsum("...vrc,...jc->...vjr", P, Xh)
    return ph[..., :2] / ph[..., 2:3]


def make_pool(n_frames, n_views, n_joints, seed=0, box=400.0, radius=3000.0, p_outlier=0.1,
              per_frame_cameras=True, valid_prob=1.0, subpixel=True):
    """Returns dict(P [N,V,3,4] f64, X [N,J,3], centres [N,V,J,2] f32 heatmap-pixel bump centres,
    valid [N,J] bool).  With probability ``p_outlier`` a (frame, view, joint) bump is displaced by at
    least 10 heatmap pixels so RANSAC has something to reject."""
    rng = np.random.default_rng(seed)
    base = ring_cameras(n_views, radius=radius, rng=rng)
    P = np.broadcast_to(base, (n_frames,) + base.shape).copy()
    if per_frame_cameras:
        # per-frame crop jitter: shifts the principal point like the dataset's bbox crop does
        shift = rng.uniform(-6, 6, size=(n_frames, n_views, 2))
        P[:, :, 0, :] += shift[:, :, 0:1] * P[:, :, 2, :]
        P[:, :, 1, :] += shift[:, :, 1:2] * P[:, :, 2, :]
    X = rng.uniform(-box, box, size=(n_frames, n_joints, 3))
    uv = project(P, X) / STRIDE  # heatmap pixels
    out = rng.uniform(size=uv.shape[:-1]) < p_outlier
    disp = rng.uniform(10, 22, size=uv.shape) * rng.choice([-1.0, 1.0], size=uv.shape)
    uv = np.where(out[..., None], uv + disp, uv)
    if not subpixel:
        uv = np.round(uv)
    uv = np.clip(uv, 1.0, HEATMAP_SIZE - 2.0)
    valid = rng.uniform(size=(n_frames, n_joints)) < valid_prob
    valid[:, 0] = True
    return {"P": P, "X": X, "centres": uv.astype(np.float32), "valid": valid}


def render_heatmaps(centres, sigma=1.0, noise=0.05, seed=0, size=HEATMAP_SIZE):
    """centres [..., 2] (x, y) heatmap pixels -> float32 [..., size, size] Gaussian bump + N(0, noise)."""
    rng = np.random.default_rng(seed)
    c = np.asarray(centres, dtype=np.float32)
    ys, xs = np.mgrid[0:size, 0:size].astype(np.float32)
    d2 = (xs - c[..., 0, None, None]) ** 2 + (ys - c[..., 1, None, None]) ** 2
    hm = np.exp(-d2 / np.float32(2 * sigma * sigma)).astype(np.float32)
    if noise > 0:
        hm += rng.normal(0, noise, size=hm.shape).astype(np.float32)
    return hm


def onehot_heatmaps(keypoints_2d, stride=STRIDE, size=HEATMAP_SIZE):
    """int keypoints [..., 2] = (x, y)*stride -> float32 [..., size, size] with a single 1 at (row y, col x)."""
    kp = np.asarray(keypoints_2d) // stride
    hm = np.zeros(kp.shape[:-1] + (size, size), dtype=np.float32)
    idx = np.indices(kp.shape[:-1])
    hm[tuple(idx) + (kp[..., 1], kp[..., 0])] = 1.0
    return hm
