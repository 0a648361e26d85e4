"""Shared helpers of the parity tests: golden fixtures, leaf matching, image metrics."""
from __future__ import annotations

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(kind: str, scene: str):
    path = os.path.join(GOLDEN, f"{kind}_{scene}.npz")
    if not os.path.exists(path):
        return None
    return np.load(path)


def _apply(xf, p, point: bool):
    if xf is None:
        return np.asarray(p, dtype=np.float64)
    r = np.array(list(xf.r), dtype=np.float64).reshape(3, 3)
    out = r @ np.asarray(p, dtype=np.float64)
    if point:
        out = out + np.array(list(xf.t), dtype=np.float64)
    return out


def flat_leaf_keys(desc) -> np.ndarray:
    """World-space geometry key of every world primitive of a flattened scene, in the same
    10-double layout as oracle/ref_driver.cpp's leaf_key()."""
    keys = np.zeros((desc.n_world, 10), dtype=np.float64)
    for i in range(desc.n_world):
        ref = desc.world[i]
        if ref.type == 0:
            s = desc.spheres[ref.index]
            xf = desc.xforms[s.xform] if s.xform >= 0 else None
            keys[i, 0] = 0
            keys[i, 1:4] = _apply(xf, list(s.center0), True)
            keys[i, 4:7] = _apply(xf, list(s.center_vec), False)
            keys[i, 7] = s.radius
        elif ref.type == 1:
            q = desc.quads[ref.index]
            xf = desc.xforms[q.xform] if q.xform >= 0 else None
            keys[i, 0] = 1
            keys[i, 1:4] = _apply(xf, list(q.Q), True)
            keys[i, 4:7] = _apply(xf, list(q.u), False)
            keys[i, 7:10] = _apply(xf, list(q.v), False)
        else:
            t = desc.triangles[ref.index]
            xf = desc.xforms[t.xform] if t.xform >= 0 else None
            keys[i, 0] = 2
            keys[i, 1:4] = _apply(xf, list(t.p0), True)
            keys[i, 4:7] = _apply(xf, list(t.p1), True)
            keys[i, 7:10] = _apply(xf, list(t.p2), True)
    return keys


def match_leaves(ref_keys: np.ndarray, flat_keys: np.ndarray, tol: float = 1e-6):
    """For every reference leaf, the index of the flattened primitive with the same type and
    geometry.  Returns (mapping, max_distance)."""
    from scipy.spatial import cKDTree

    scale = np.array([1e6] + [1.0] * 9)  # the type column must match exactly
    tree = cKDTree(flat_keys * scale)
    dist, idx = tree.query(ref_keys * scale, k=1)
    if dist.max() > tol * max(1.0, np.abs(ref_keys[:, 1:]).max()):
        raise AssertionError(f"reference leaf without a geometric twin in the flattened scene (distance {dist.max():g})")
    return idx.astype(np.int64), float(dist.max())


def equivalent_ids(flat_keys: np.ndarray) -> np.ndarray:
    """Canonical id per primitive: duplicates of the same geometry get the same id."""
    _, inv = np.unique(np.round(flat_keys, 9), axis=0, return_inverse=True)
    return inv.reshape(-1)


def gamma_image(lin: np.ndarray) -> np.ndarray:
    """Camera.txt:77-84: sqrt gamma, clamp to [0, 0.999]"""
    return np.clip(np.sqrt(np.maximum(lin, 0.0)), 0.0, 0.999)


def psnr_after_gamma(a_lin: np.ndarray, b_lin: np.ndarray) -> float:
    a, b = gamma_image(a_lin.astype(np.float64)), gamma_image(b_lin.astype(np.float64))
    mse = float(np.mean((a - b) ** 2))
    return 99.0 if mse == 0 else float(10.0 * np.log10(1.0 / mse))


def compare_primary(gold, aov: dict, flat_keys: np.ndarray, undecidable=None) -> dict:
    """Primary-hit parity of a device AOV against a golden fixture of the same frame size.
    `undecidable` (see undecidable_pixels) are excluded from the identity count and reported."""
    ids_ref = gold["ids"].reshape(-1)
    mapping, _ = match_leaves(gold["leaves"], flat_keys)
    canon = equivalent_ids(flat_keys)
    ref_flat = np.where(ids_ref >= 0, canon[mapping[np.maximum(ids_ref, 0)]], -1)
    dev_ids = aov["prim_id"].reshape(-1)
    dev_flat = np.where(dev_ids >= 0, canon[np.maximum(dev_ids, 0)], -1)
    same = ref_flat == dev_flat
    # Exact ties: two different primitives hit at the same distance (a ray through the shared
    # edge of two Cornell walls, coplanar overlapping triangles).  The reference resolves those
    # by BVH visiting order at the 1e-13 level; they are not identity errors.
    t_all_ref = gold["t"].reshape(-1)
    t_all_dev = aov["t"].reshape(-1).astype(np.float64)
    tie = (~same) & (ref_flat >= 0) & (dev_flat >= 0) & (np.abs(t_all_dev - t_all_ref) <= 2e-6 * np.abs(t_all_ref))
    if undecidable is None:
        undecidable = np.zeros(same.shape, dtype=bool)
    decided = ~undecidable
    both = same & (ref_flat >= 0)
    t_ref = gold["t"].reshape(-1)[both]
    t_dev = aov["t"].reshape(-1)[both].astype(np.float64)
    t_rel = np.abs(t_dev - t_ref) / np.maximum(np.abs(t_ref), 1e-30)
    n_ref = gold["normal"].reshape(-1, 3)[both].astype(np.float64)
    n_dev = aov["normal"].reshape(-1, 3)[both].astype(np.float64)
    n_err = np.abs(n_dev - n_ref).max(axis=1) if both.any() else np.zeros(0)
    uv_ref = gold["uv"].reshape(-1, 2)[both].astype(np.float64)
    uv_dev = aov["uv"].reshape(-1, 2)[both].astype(np.float64)
    uv_err = np.abs(uv_dev - uv_ref).max(axis=1) if both.any() else np.zeros(0)
    return {
        "pixels": int(ids_ref.size), "id_match": float((same | tie)[decided].mean()), "id_match_strict": float(same.mean()),
        "mismatches": int((~(same | tie) & decided).sum()), "ties": int(tie.sum()), "undecidable": int(undecidable.sum()),
        "t_rel_max": float(t_rel.max()) if t_rel.size else 0.0,
        "t_rel_p9999": float(np.quantile(t_rel, 0.9999)) if t_rel.size else 0.0,
        "t_within_1e5": float((t_rel <= 1e-5).mean()) if t_rel.size else 1.0,
        "n_err_max": float(n_err.max()) if n_err.size else 0.0,
        "n_within_1e5": float((n_err <= 1e-5).mean()) if n_err.size else 1.0,
        "uv_err_max": float(uv_err.max()) if uv_err.size else 0.0,
        "mismatch_pixels": np.nonzero(~(same | tie) & decided)[0][:16].tolist(),
    }


def camera_center_rays(cam, width: int, height: int) -> np.ndarray:
    """Pixel-centre rays of Camera.txt:136-168 in double, as n x 9 rows (o, d, time 0, 0.001, inf)."""
    lookfrom, lookat, vup = (np.array(list(v), dtype=np.float64) for v in (cam.lookfrom, cam.lookat, cam.vup))
    theta = cam.vfov * np.pi / 180.0
    vh = 2 * np.tan(theta / 2) * cam.focus_dist
    vw = vh * (width / height)
    w = (lookfrom - lookat) / np.linalg.norm(lookfrom - lookat)
    u = np.cross(vup, w)
    u /= np.linalg.norm(u)
    v = np.cross(w, u)
    du, dv = vw * u / width, -vh * v / height
    p00 = lookfrom - cam.focus_dist * w - vw * u / 2 + vh * v / 2 + 0.5 * (du + dv)
    jj, ii = np.mgrid[0:height, 0:width]
    d = p00[None, None, :] + ii[..., None] * du + jj[..., None] * dv - lookfrom
    rays = np.zeros((height * width, 9))
    rays[:, 0:3] = lookfrom
    rays[:, 3:6] = d.reshape(-1, 3)
    rays[:, 7] = 0.001
    rays[:, 8] = np.inf
    return rays


def undecidable_pixels(scene, width: int, height: int, eps: float = 3e-7) -> np.ndarray:
    """Pixels whose REFERENCE answer is itself decided by rounding: the hit primitive of the
    double-precision oracle changes when the ray direction is nudged by `eps` relative (a few
    FP32 ulps, i.e. less than the error of merely storing the ray in FP32).  These are the rays
    that run exactly along a shared edge or through a crack between two quads — e.g. the
    diagonals of the symmetric Cornell-box frame — and no FP32 renderer can be asked to
    reproduce the reference's coin flip there.  Returned as a boolean mask (flattened)."""
    from oracle import port

    rays = camera_center_rays(scene.desc.camera, width, height)
    base = port.hit(scene, rays)["prim_id"]
    d = rays[:, 3:6]
    norm = np.linalg.norm(d, axis=1, keepdims=True)
    a = np.cross(d, np.array([0.3, 1.0, 0.2]))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = np.cross(d, a)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    mask = np.zeros(len(base), dtype=bool)
    for sa, sb in ((1, 0), (-1, 0), (0, 1), (0, -1)):
        r2 = rays.copy()
        r2[:, 3:6] = d + eps * norm * (sa * a + sb * b)
        mask |= port.hit(scene, r2)["prim_id"] != base
    return mask
