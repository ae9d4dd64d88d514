"""Seeded synthetic inputs shaped like the reference's data.  Used by the
tests and by bench.py; pure numpy, no GPU."""
import numpy as np


def random_dets(rng, n, extent=(200, 200, 64), side=(8, 40), integer=False):
    """n boxes [x1,y1,z1,x2,y2,z2,score] f32 with DISTINCT scores."""
    ext = np.asarray(extent, dtype=np.float64)
    c = rng.uniform(0, 1, (n, 3)) * ext
    s = rng.uniform(side[0], side[1], (n, 3))
    d = np.hstack([c - s / 2, c + s / 2])
    if integer:
        d = np.round(d)
    scores = (rng.permutation(n).astype(np.float64) + 1.0) / max(n, 1)
    return np.hstack([d, scores[:, None]]).astype(np.float32)


def blob_volume(rng, shape, n_blobs, sigma_xy=(4, 9), sigma_z=(2, 5), amp=(80, 200), noise=40, want_prm=True):
    """uint8 volume [S,H,W] of anisotropic Gaussian blobs + uniform noise, with per-blob tight boxes
    and per-blob PRM-like response crops.  Returns (vol, boxes float [n,6], blobs list of dict)."""
    S, H, W = shape
    vol = rng.uniform(0, noise, size=shape).astype(np.float32)
    boxes = np.zeros((n_blobs, 6), dtype=np.float32)
    blobs = []
    for i in range(n_blobs):
        sxy = rng.uniform(*sigma_xy); sz = rng.uniform(*sigma_z); a = rng.uniform(*amp)
        cz = rng.uniform(3 * sz, max(3 * sz + 1, S - 3 * sz))
        cy = rng.uniform(3 * sxy, max(3 * sxy + 1, H - 3 * sxy))
        cx = rng.uniform(3 * sxy, max(3 * sxy + 1, W - 3 * sxy))
        rz, rxy = int(np.ceil(3 * sz)), int(np.ceil(3 * sxy))
        z0, z1 = max(0, int(cz) - rz), min(S, int(cz) + rz + 1)
        y0, y1 = max(0, int(cy) - rxy), min(H, int(cy) + rxy + 1)
        x0, x1 = max(0, int(cx) - rxy), min(W, int(cx) + rxy + 1)
        zz, yy, xx = np.meshgrid(np.arange(z0, z1), np.arange(y0, y1), np.arange(x0, x1), indexing="ij")
        r2 = ((zz - cz) / sz) ** 2 + ((yy - cy) / sxy) ** 2 + ((xx - cx) / sxy) ** 2
        vol[z0:z1, y0:y1, x0:x1] += (a * np.exp(-0.5 * r2)).astype(np.float32)
        # tight box at ~2 sigma
        bz, bxy = 2.0 * sz, 2.0 * sxy
        boxes[i] = [cx - bxy, cy - bxy, cz - bz, cx + bxy, cy + bxy, cz + bz]
        blobs.append(dict(c=(cz, cy, cx), sz=sz, sxy=sxy, amp=a))
    vol = np.clip(vol, 0, 255).astype(np.uint8)
    return vol, boxes, blobs


def prm_crop(blob, box_int, scale=0.8):
    """uint8 PRM-like crop for one (integer, inclusive) box: a Gaussian at 0.8 sigma, peak 255."""
    x1, y1, z1, x2, y2, z2 = [int(v) for v in box_int]
    zz, yy, xx = np.meshgrid(np.arange(z1, z2 + 1), np.arange(y1, y2 + 1), np.arange(x1, x2 + 1), indexing="ij")
    cz, cy, cx = blob["c"]
    sz, sxy = blob["sz"] * scale, blob["sxy"] * scale
    r2 = ((zz - cz) / sz) ** 2 + ((yy - cy) / sxy) ** 2 + ((xx - cx) / sxy) ** 2
    return (255.0 * np.exp(-0.5 * r2)).astype(np.uint8)


def postproc_case(seed, shape=(64, 256, 256), n_blobs=35, n_dup=10, n_false=5, sigma_xy=(4, 9), sigma_z=(2, 5)):
    """A full binarization_soma-style input: volume, detections (tight blob boxes + jittered
    duplicates + false boxes, distinct scores), int boxes, packed PRM crops and offsets."""
    from .binarization import dets_to_boxes, crop_offsets
    rng = np.random.default_rng(seed)
    S, H, W = shape
    vol, bx, blobs = blob_volume(rng, shape, n_blobs, sigma_xy=sigma_xy, sigma_z=sigma_z)
    owners = list(range(n_blobs))
    allb = [bx]
    if n_dup:
        src = rng.integers(0, n_blobs, n_dup)
        allb.append(bx[src] + rng.uniform(-2, 2, (n_dup, 6)).astype(np.float32))
        owners += [int(s) for s in src]
    if n_false:
        fd = random_dets(rng, n_false, extent=(W, H, S), side=(6, 24))[:, :6]
        # detections reach the scripts clipped to the tile (boxes_3d.py:144-225 clip_tiled_boxes_3d); the reference script
        # indexes with the raw int() box and fails on a box that leaves the volume
        fd = np.clip(fd, 0, np.array([W - 1, H - 1, S - 1, W - 1, H - 1, S - 1], dtype=np.float32))
        allb.append(fd)
        owners += [int(v) for v in rng.integers(0, n_blobs, n_false)]
    allb = np.concatenate(allb, axis=0)
    n = allb.shape[0]
    scores = (rng.permutation(n).astype(np.float64) + 1.0) / n
    dets = np.hstack([allb, scores[:, None]]).astype(np.float32)
    boxes = dets_to_boxes(dets, shape)
    off = crop_offsets(boxes)
    prm = np.empty(int(off[-1]), dtype=np.uint8)
    for i in range(n):
        prm[off[i]:off[i + 1]] = prm_crop(blobs[owners[i]], boxes[i]).ravel()
    return dict(volume=vol, dets=dets, boxes=boxes, prm=prm, crop_off=off, blobs=blobs)


def response_map(rng, shape, n_peaks=200, channels=1, noise=1e-3):
    """fp32 [1,A,S,H,W] response map: blob field + small noise (RPN fg-probability-like)."""
    S, H, W = shape
    out = np.empty((1, channels, S, H, W), dtype=np.float32)
    for a in range(channels):
        m = rng.uniform(0, noise, size=shape).astype(np.float32)
        for _ in range(n_peaks):
            sg = rng.uniform(1.0, 3.0)
            c = rng.uniform(0, 1, 3) * np.array([S, H, W])
            r = int(np.ceil(3 * sg))
            z0, z1 = max(0, int(c[0]) - r), min(S, int(c[0]) + r + 1)
            y0, y1 = max(0, int(c[1]) - r), min(H, int(c[1]) + r + 1)
            x0, x1 = max(0, int(c[2]) - r), min(W, int(c[2]) + r + 1)
            zz, yy, xx = np.meshgrid(np.arange(z0, z1), np.arange(y0, y1), np.arange(x0, x1), indexing="ij")
            r2 = ((zz - c[0]) ** 2 + (yy - c[1]) ** 2 + (xx - c[2]) ** 2) / sg ** 2
            m[z0:z1, y0:y1, x0:x1] += (rng.uniform(0.3, 1.0) * np.exp(-0.5 * r2)).astype(np.float32)
        out[0, a] = m
    return out


def roialign_case(seed, feat_shape=(2, 256, 8, 32, 32), n_rois=512, scale=1.0 / 8, side=(10, 50), frac_outside=0.05):
    """features N(0,1) fp32, rois [R,7] in image coords."""
    rng = np.random.default_rng(seed)
    B, C, S, H, W = feat_shape
    feat = rng.standard_normal(feat_shape).astype(np.float32)
    img = np.array([W, H, S]) / scale
    ctr = rng.uniform(0, 1, (n_rois, 3)) * img
    sd = rng.uniform(side[0], side[1], (n_rois, 3))
    n_out = int(round(frac_outside * n_rois))
    if n_out:
        idx = rng.choice(n_rois, n_out, replace=False)
        ctr[idx] += rng.choice([-1, 1], (n_out, 3)) * img * rng.uniform(0.45, 0.6, (n_out, 3))
    rois = np.hstack([rng.integers(0, B, (n_rois, 1)).astype(np.float64), ctr - sd / 2, ctr + sd / 2]).astype(np.float32)
    return feat, rois
