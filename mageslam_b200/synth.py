"""Seeded synthetic inputs for the two hot paths. numpy only -- usable on the GPU box.

Frames: blurred-noise images (about 1e4 FAST candidates at threshold 10-20) and a structured scene (filled
rectangles / discs on a gradient with fine texture) viewed through a smooth homography trajectory.
BA problems: K cameras on an arc looking at P points in a frustum-visible box, d observations per point
(template: reference Dependencies/g2o/g2o/examples/ba/ba_demo.cpp:114-200).
"""
import numpy as np


def _gauss_kernel(sigma):
    r = max(1, int(3 * sigma + 0.5))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def _sep_filter(img, k):
    r = len(k) // 2
    p = np.pad(img, ((0, 0), (r, r)), mode="reflect")
    out = np.zeros_like(img, dtype=np.float64)
    for i, kv in enumerate(k):
        out += kv * p[:, i:i + img.shape[1]]
    p = np.pad(out, ((r, r), (0, 0)), mode="reflect")
    out2 = np.zeros_like(out)
    for i, kv in enumerate(k):
        out2 += kv * p[i:i + img.shape[0], :]
    return out2


def noise_frame(seed, w=640, h=480, sigma=1.5):
    """uniform noise -> Gaussian sigma -> min-max normalise to [0,255]."""
    rng = np.random.default_rng(seed)
    f = _sep_filter(rng.random((h, w)), _gauss_kernel(sigma))
    f = (f - f.min()) / (f.max() - f.min())
    return np.round(f * 255.0).astype(np.uint8)


def structured_scene(seed, w=1280, h=960, nshapes=160):
    """Filled rectangles / discs on a gradient plus low-amplitude texture: strong, repeatable corners."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = 60.0 + 80.0 * xx / w + 40.0 * yy / h
    for _ in range(nshapes):
        val = float(rng.integers(10, 246))
        if rng.random() < 0.6:
            x0 = int(rng.integers(0, w - 20)); y0 = int(rng.integers(0, h - 20))
            ww = int(rng.integers(12, 140)); hh = int(rng.integers(12, 140))
            img[y0:y0 + hh, x0:x0 + ww] = val
        else:
            cx = int(rng.integers(0, w)); cy = int(rng.integers(0, h)); rad = int(rng.integers(6, 60))
            y0, y1 = max(0, cy - rad), min(h, cy + rad + 1)
            x0, x1 = max(0, cx - rad), min(w, cx + rad + 1)
            sub = img[y0:y1, x0:x1]
            m = (yy[y0:y1, x0:x1] - cy) ** 2 + (xx[y0:y1, x0:x1] - cx) ** 2 <= rad * rad
            sub[m] = val
    tex = _sep_filter(rng.random((h, w)), _gauss_kernel(1.0))
    img += 40.0 * (tex - tex.mean()) / (tex.std() + 1e-9) * 0.5
    return np.clip(np.round(img), 0, 255).astype(np.uint8)


def _warp(scene, Hm, w, h):
    """Sample scene at Hm * (x, y, 1) with bilinear interpolation (numpy)."""
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    X = Hm[0, 0] * xs + Hm[0, 1] * ys + Hm[0, 2]
    Y = Hm[1, 0] * xs + Hm[1, 1] * ys + Hm[1, 2]
    Z = Hm[2, 0] * xs + Hm[2, 1] * ys + Hm[2, 2]
    X /= Z; Y /= Z
    sh, sw = scene.shape
    X = np.clip(X, 0, sw - 1.001); Y = np.clip(Y, 0, sh - 1.001)
    x0 = np.floor(X).astype(np.int64); y0 = np.floor(Y).astype(np.int64)
    fx = X - x0; fy = Y - y0
    s = scene.astype(np.float64)
    v = (s[y0, x0] * (1 - fx) * (1 - fy) + s[y0, x0 + 1] * fx * (1 - fy) +
         s[y0 + 1, x0] * (1 - fx) * fy + s[y0 + 1, x0 + 1] * fx * fy)
    return np.clip(np.round(v), 0, 255).astype(np.uint8)


def video_frames(n, w=640, h=480, seed=0):
    """n frames of the structured scene through a smooth (seeded) homography trajectory; uint8 [n, h, w]."""
    rng = np.random.default_rng(seed + 1000)
    scene = structured_scene(seed, 2 * w, 2 * h)
    ph = rng.random(6) * 2 * np.pi
    out = np.empty((n, h, w), np.uint8)
    for t in range(n):
        a = 0.08 * np.sin(0.021 * t + ph[0])
        s = 1.0 + 0.15 * np.sin(0.013 * t + ph[1])
        tx = 0.5 * w + 0.35 * w * np.sin(0.017 * t + ph[2])
        ty = 0.5 * h + 0.35 * h * np.sin(0.011 * t + ph[3])
        p0 = 1e-5 * np.sin(0.009 * t + ph[4]); p1 = 1e-5 * np.sin(0.007 * t + ph[5])
        c, sn = np.cos(a) * s, np.sin(a) * s
        Hm = np.array([[c, -sn, tx], [sn, c, ty], [p0, p1, 1.0]])
        out[t] = _warp(scene, Hm, w, h)
    return out


def shifted_noisy(img, dx=3, dy=2, amp=2, seed=1):
    """Same frame shifted by (dx, dy) px with +-amp grey-level noise (SURVEY 8(d) config 1 self-match)."""
    rng = np.random.default_rng(seed)
    out = np.roll(np.roll(img, dy, axis=0), dx, axis=1).astype(np.int16)
    out += rng.integers(-amp, amp + 1, size=img.shape, dtype=np.int16)
    return np.clip(out, 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------------------------------------------------
def _rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def ba_problem(K=10, P=2000, obs_per_point=4, seed=1, n_fixed=2, f=500.0, cx=320.0, cy=240.0,
               pixel_sigma=0.5, point_sigma=0.02, pose_sigma=0.01, outlier_frac=0.0, loop=False, info_mode="one"):
    """Synthetic BA window. Returns a dict of float32 arrays in the BundlerLib conventions:
    cam_pos[K,3] + cam_rot[K,9] (column-major 3x3) = world->camera (view) transform (reference BundlerLib.cpp:261-276),
    intrinsics[K,4] = (cx, cy, fx, fy), fixed[K], points[P,3], obs_uv[E,2], obs_cam[E], obs_pt[E], obs_info[E]."""
    rng = np.random.default_rng(seed)
    cams_R = np.zeros((K, 3, 3)); cams_t = np.zeros((K, 3))
    centers = np.zeros((K, 3))
    for k in range(K):
        if loop:
            ang = 2 * np.pi * k / K
            C = np.array([12.0 * np.cos(ang), 0.3 * np.sin(3 * ang), 12.0 * np.sin(ang)])
            yaw = -ang + np.pi / 2 + np.pi / 2     # look outward-tangent mix
            Rwc = _rot(0.0, yaw, 0.0)
        else:
            C = np.array([0.25 * (k - (K - 1) / 2.0), 0.02 * np.sin(k), 0.05 * np.cos(0.7 * k)])
            Rwc = _rot(0.01 * np.sin(k), 0.03 * (k - (K - 1) / 2.0) / max(K, 1), 0.01 * np.cos(k))
        centers[k] = C
        R = Rwc.T                      # world -> camera
        cams_R[k] = R
        cams_t[k] = -R @ C
    pts = np.zeros((P, 3))
    obs_cam = np.zeros(P * obs_per_point, np.int32); obs_pt = np.zeros(P * obs_per_point, np.int32)
    obs_uv = np.zeros((P * obs_per_point, 2))
    e = 0
    for i in range(P):
        for _try in range(200):
            if loop:
                k0 = int(rng.integers(0, K))
                depth = rng.uniform(3.0, 8.0)
                local = np.array([rng.uniform(-1.5, 1.5), rng.uniform(-1.0, 1.0), depth])
                X = cams_R[k0].T @ (local - cams_t[k0])
                cand = [(k0 + j) % K for j in range(-(obs_per_point // 2), obs_per_point - obs_per_point // 2)]
            else:
                X = np.array([rng.uniform(-3.0, 3.0), rng.uniform(-2.0, 2.0), rng.uniform(3.0, 8.0)])
                cand = list(rng.permutation(K))
            seen = []
            for k in cand:
                Xc = cams_R[k] @ X + cams_t[k]
                if Xc[2] <= 0.5:
                    continue
                u = f * Xc[0] / Xc[2] + cx; v = f * Xc[1] / Xc[2] + cy
                if 0 <= u < 2 * cx and 0 <= v < 2 * cy:
                    seen.append((int(k), u, v))
                if len(seen) == obs_per_point:
                    break
            if len(seen) == obs_per_point:
                break
        else:
            raise RuntimeError("could not place point %d" % i)
        pts[i] = X
        for k, u, v in seen:
            obs_cam[e] = k; obs_pt[e] = i
            obs_uv[e] = (u + rng.normal(0, pixel_sigma), v + rng.normal(0, pixel_sigma))
            e += 1
    E = e
    if outlier_frac > 0:
        bad = rng.random(E) < outlier_frac
        obs_uv[bad] += rng.uniform(-60, 60, size=(int(bad.sum()), 2))
    pts_init = pts + rng.normal(0, point_sigma, size=pts.shape)
    cam_pos = np.zeros((K, 3), np.float32); cam_rot = np.zeros((K, 9), np.float32)
    fixed = np.zeros(K, np.int32)
    for k in range(K):
        R, t = cams_R[k], cams_t[k]
        if k < n_fixed:
            fixed[k] = 1
        else:
            dR = _rot(*rng.normal(0, pose_sigma * 0.2, 3))
            R = dR @ R
            t = t + rng.normal(0, pose_sigma, 3)
        cam_pos[k] = t
        cam_rot[k] = R.T.reshape(-1)          # column-major flattening of R
    if info_mode == "one":
        info = np.ones(E, np.float32)
    else:                                      # MapPointRefinementConfidence(0..5), reference Map/MappingMath.h:42-49
        rc = rng.integers(0, 6, size=P)
        conf = 1.0 - 1.0 / (1.5 + rc) ** 2
        info = conf[obs_pt[:E]].astype(np.float32)
    return dict(cam_pos=cam_pos, cam_rot=cam_rot,
                intrinsics=np.tile(np.array([cx, cy, f, f], np.float32), (K, 1)), fixed=fixed,
                points=pts_init.astype(np.float32), obs_uv=obs_uv[:E].astype(np.float32),
                obs_cam=obs_cam[:E].copy(), obs_pt=obs_pt[:E].copy(), obs_info=info,
                true_points=pts.astype(np.float32), true_cam_R=cams_R.copy(), true_cam_t=cams_t.copy())


def _quat_xyzw(R):
    """Rotation matrix -> unit quaternion (x, y, z, w), w >= 0."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R))); j = (i + 1) % 3; k = (j + 1) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s; q[3] = (R[k, j] - R[j, k]) / s; q[j] = (R[j, i] + R[i, j]) / s; q[k] = (R[k, i] + R[i, k]) / s
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def ba_add_tethers(prob, seed=0, n_distance=4, n_rotation=4, n_transform=4, noise=1e-3):
    """Adds tether edges between cameras to a ba_problem dict, in the conventions of BundlerLib's three constraint setters
    (reference BundlerLib.cpp:311-350): tether_distance = [(cam1, cam2, distance, weight)] on the view-transform translations,
    tether_rotation = [(cam1, cam2, q_xyzw of T1^-1 T2, weight)], tether_transform = [(cam1, cam2, t, q_xyzw, weight)] with
    C = T2 T1^-1 (the measurement EdgeSE3Expmap's error log(T2^-1 C T1) vanishes at). Pairs include one with a fixed camera
    and one with both cameras fixed (an inactive edge) when the window has fixed cameras."""
    rng = np.random.default_rng(seed)
    R, t, fixed = prob["true_cam_R"], prob["true_cam_t"], prob["fixed"]
    K = len(R)
    def pairs(n, off):
        out = []
        for i in range(n):
            a = (i * 2 + off) % (K - 1)
            out.append((a, a + 1))
        if n and fixed[0] and fixed[1]:
            out[0] = (0, 1)                                # both fixed: never active
        if n > 1 and K > 2:
            out[1] = (K - 1, 1)                            # reversed index order, one end fixed when n_fixed >= 2
        return out
    dist = [(a, b, float(np.linalg.norm(t[b] - t[a]) * (1 + rng.normal(0, noise))), float(rng.uniform(200.0, 2000.0))) for a, b in pairs(n_distance, 0)]
    rot = []
    for a, b in pairs(n_rotation, 1):
        Rrel = R[a].T @ R[b]
        q = _quat_xyzw(_rot(*rng.normal(0, noise, 3)) @ Rrel)
        rot.append((a, b, q.astype(np.float32), float(rng.uniform(500.0, 5000.0))))
    xf = []
    for a, b in pairs(n_transform, 0):
        Rc = R[b] @ R[a].T
        tc = t[b] - Rc @ t[a]
        q = _quat_xyzw(_rot(*rng.normal(0, noise, 3)) @ Rc)
        xf.append((a, b, (tc + rng.normal(0, noise, 3)).astype(np.float32), q.astype(np.float32), float(rng.uniform(1e5, 1e6))))
    out = dict(prob)
    out["tether_distance"], out["tether_rotation"], out["tether_transform"] = dist, rot, xf
    return out


def local_map_scene(n=4000, seed=0, width=640, height=480, scale=1.2, levels=8):
    """A tracking-thread scene for the map-point projection step: a camera pose (view matrix, calibration, position, forward)
    and n local-map points scattered so that every culling branch of IsGoodCandidate fires (behind the camera, outside the
    image border, viewing angle, scale-invariance range, octave out of range). Returns a dict of float32 arrays."""
    rng = np.random.default_rng(seed)
    ang = rng.normal(0, 0.2, 3)
    cx, sx, cy, sy, cz, sz = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]); Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    R = (Rz @ Ry @ Rx)                       # world -> camera
    C = rng.normal(0, 0.5, 3)                 # camera centre (world)
    view = np.concatenate([R, (-R @ C)[:, None]], axis=1).astype(np.float32)
    K = np.array([[520.0, 0, width / 2 + 3.5], [0, 515.0, height / 2 - 2.25], [0, 0, 1]], np.float32)
    forward = (R.T @ np.array([0, 0, 1.0])).astype(np.float32)
    forward /= np.linalg.norm(forward)
    # points in camera space: mostly in front within the frustum, some behind / outside
    z = rng.uniform(-2.0, 12.0, n)
    z[np.abs(z) < 0.05] = 0.5
    x = rng.uniform(-0.85, 0.85, n) * np.abs(z)
    y = rng.uniform(-0.65, 0.65, n) * np.abs(z)
    Pc = np.stack([x, y, z], axis=1)
    Pw = (R.T @ Pc.T).T + C
    dist = np.linalg.norm(Pw - C, axis=1)
    # mean viewing direction: from a previous observer roughly along the current ray, perturbed up to ~90 degrees
    ray = (Pw - C) / np.maximum(dist[:, None], 1e-6)
    d = ray + rng.normal(0, 0.6, (n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    # scale-invariance range: observed at octave o at distance d0 ~ dist * jitter
    o = rng.integers(0, levels, n)
    d0 = dist * rng.uniform(0.4, 2.5, n)
    dmin = d0 * np.power(scale, -(o + 0.5))
    dmax = d0 * np.power(scale, levels - (o + 0.5))
    pts = np.zeros(n, np.dtype([("position", "<f4", 3), ("mean_view_dir", "<f4", 3), ("dmin", "<f4"), ("dmax", "<f4")]))
    pts["position"] = Pw.astype(np.float32); pts["mean_view_dir"] = d.astype(np.float32)
    pts["dmin"] = dmin.astype(np.float32); pts["dmax"] = dmax.astype(np.float32)
    return dict(view=view, K=K, position=C.astype(np.float32), forward=forward, points=pts, width=width, height=height,
                scale=scale, levels=levels)


def natural_frames(n, w=640, h=480, seed=0, noise_sigma=1.5):
    """n frames with the statistics of an indoor camera image rather than of a corner-dense test chart: the structured scene's
    shapes on a smooth gradient (piecewise-constant regions, so corners sit on object outlines), a 1/f ("pink") texture of a few
    grey levels and independent sensor noise of `noise_sigma` grey levels per frame, seen through the same homography trajectory
    as video_frames. About 2 % of the pixels pass FAST at threshold 10 (video_frames: about a quarter) and the tier configuration still finds its 2000 key points."""
    rng = np.random.default_rng(seed + 7000)
    sw, sh = 2 * w, 2 * h
    yy, xx = np.mgrid[0:sh, 0:sw]
    img = 70.0 + 70.0 * xx / sw + 40.0 * yy / sh
    for _ in range(200):
        val = float(rng.integers(15, 241))
        if rng.random() < 0.65:
            x0 = int(rng.integers(0, sw - 30)); y0 = int(rng.integers(0, sh - 30))
            img[y0:y0 + int(rng.integers(25, 260)), x0:x0 + int(rng.integers(25, 260))] = val
        else:
            cx = int(rng.integers(0, sw)); cy = int(rng.integers(0, sh)); rad = int(rng.integers(12, 110))
            y0, y1 = max(0, cy - rad), min(sh, cy + rad + 1); x0, x1 = max(0, cx - rad), min(sw, cx + rad + 1)
            sub = img[y0:y1, x0:x1]
            sub[(yy[y0:y1, x0:x1] - cy) ** 2 + (xx[y0:y1, x0:x1] - cx) ** 2 <= rad * rad] = val
    fy = np.fft.fftfreq(sh)[:, None]; fx = np.fft.rfftfreq(sw)[None, :]
    amp = 1.0 / np.maximum(np.hypot(fx, fy), 1.0 / max(sw, sh))
    tex = np.fft.irfft2(amp * np.exp(2j * np.pi * rng.random(amp.shape)), s=(sh, sw))
    img += 12.0 * (tex - tex.mean()) / (tex.std() + 1e-9)
    img = _sep_filter(img, _gauss_kernel(0.8))                       # lens blur: edges are a couple of pixels wide
    scene = np.clip(np.round(img), 0, 255).astype(np.uint8)
    ph = rng.random(6) * 2 * np.pi
    out = np.empty((n, h, w), np.uint8)
    for t in range(n):
        a = 0.08 * np.sin(0.021 * t + ph[0]); s = 1.0 + 0.15 * np.sin(0.013 * t + ph[1])
        tx = 0.5 * w + 0.35 * w * np.sin(0.017 * t + ph[2]); ty = 0.5 * h + 0.35 * h * np.sin(0.011 * t + ph[3])
        p0 = 1e-5 * np.sin(0.009 * t + ph[4]); p1 = 1e-5 * np.sin(0.007 * t + ph[5])
        c, sn = np.cos(a) * s, np.sin(a) * s
        f = _warp(scene, np.array([[c, -sn, tx], [sn, c, ty], [p0, p1, 1.0]]), w, h).astype(np.float64)
        out[t] = np.clip(np.round(f + noise_sigma * rng.standard_normal((h, w))), 0, 255).astype(np.uint8)
    return out
