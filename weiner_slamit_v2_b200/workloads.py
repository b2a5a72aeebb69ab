"""Seeded synthetic matching workloads: keypoint sets, descriptor
pairs with controlled bit noise, projected map points.  Pure numpy so oracle and GPU see the same bytes."""
import numpy as np

from ._lib import KP_DTYPE

SCALE_FACTORS_8 = np.array([1.0, 1.2, 1.44, 1.728, 2.0736, 2.48832, 2.985984, 3.5831808], np.float32)


def flip_bits(desc, nbits, rng):
    """Flip nbits[i] random distinct bits of each 256-bit descriptor (vectorised)."""
    n = len(desc)
    rank = np.argsort(rng.random((n, 256)), axis=1).argsort(axis=1)       # a random permutation per row
    mask = np.packbits(rank < np.asarray(nbits).reshape(-1, 1), axis=1, bitorder="little")
    return desc ^ mask


def random_keypoints(n, width, height, rng, nlevels=8, level0_only=False):
    k = np.zeros(n, KP_DTYPE)
    if level0_only:
        octv = np.zeros(n, np.int32)
    else:   # geometric split like the extractor's per-level quota
        p = 1.0 / SCALE_FACTORS_8[:nlevels] ** 1
        octv = np.sort(rng.choice(nlevels, size=n, p=p / p.sum())).astype(np.int32)
    sc = SCALE_FACTORS_8[octv]
    k["x"] = (rng.integers(19, np.maximum(20, (width / sc).astype(int) - 19)) * sc).astype(np.float32)
    k["y"] = (rng.integers(19, np.maximum(20, (height / sc).astype(int) - 19)) * sc).astype(np.float32)
    k["octave"] = octv
    k["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    k["size"] = (31 * sc).astype(np.int32)
    k["response"] = rng.integers(7, 120, n)
    k["class_id"] = -1
    return k


def init_pair(index, n=1000, width=640, height=480, brute_force=False):
    """A frame pair for SearchForInitialization: F2 = F1's keypoints moved by a few pixels, rotated by
    a common angle (so the rotation histogram has a dominant bin), descriptors with 0..60 flipped bits,
    shuffled order.  brute_force: all keypoints on octave 0 (every F1 point is a query)."""
    rng = np.random.default_rng(70000 + index)
    k1 = random_keypoints(n, width, height, rng, level0_only=brute_force)
    d1 = rng.integers(0, 256, (n, 32)).astype(np.uint8)
    perm = rng.permutation(n)
    k2 = k1[perm].copy()
    k2["x"] += rng.normal(0, 3, n).astype(np.float32)
    k2["y"] += rng.normal(0, 3, n).astype(np.float32)
    k2["x"] = np.clip(k2["x"], 0, width - 1); k2["y"] = np.clip(k2["y"], 0, height - 1)
    rot = rng.uniform(0, 30)
    k2["angle"] = np.mod(k2["angle"] - rot + rng.normal(0, 4, n), 360).astype(np.float32)
    d2 = flip_bits(d1[perm], rng.integers(0, 61, n), rng)
    # a few exact duplicates so distance ties and match stealing occur
    dup = rng.choice(n, size=n // 20, replace=False)
    d2[dup] = d2[np.roll(dup, 1)]
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    return k1, d1, k2, d2, prev


def projection_frame(index, n_kp=2000, n_mp=10000, width=1280, height=720, nlevels=8):
    """One frame for SearchByProjection: keypoints + map points that project near them (config 5)."""
    rng = np.random.default_rng(90000 + index)
    kp = random_keypoints(n_kp, width, height, rng, nlevels)
    kd = rng.integers(0, 256, (n_kp, 32)).astype(np.uint8)
    src = rng.integers(0, n_kp, n_mp)
    desc = flip_bits(kd[src], rng.integers(0, 41, n_mp), rng)
    x = (kp["x"][src] + rng.normal(0, 2, n_mp)).astype(np.float32)
    y = (kp["y"][src] + rng.normal(0, 2, n_mp)).astype(np.float32)
    level = np.minimum(kp["octave"][src] + (rng.random(n_mp) < 0.5), nlevels - 1).astype(np.int32)
    mp = dict(x=x, y=y, xr=np.zeros(n_mp, np.float32), level=level,
              viewcos=rng.uniform(0.9, 1.0, n_mp).astype(np.float32), desc=desc,
              in_view=(rng.random(n_mp) < 0.97).astype(np.uint8), bad=(rng.random(n_mp) < 0.02).astype(np.uint8),
              obs=(rng.random(n_mp) < 0.95).astype(np.int32) * rng.integers(1, 9, n_mp).astype(np.int32))
    return kp, kd, mp


def motion_frame(index, n_last=1500, n_cur=2000, width=640, height=480, nlevels=8,
                 K=(526.69, 540.36, 313.07, 238.39)):
    """A frame pair for SearchByProjection(CurrentFrame, LastFrame, th, bMono) (motion-model tracking):
    the last frame's keypoints carry map points (world positions); the current frame's keypoints sit near
    their projections under a small camera motion Tcw.  Includes points behind the camera, projections
    outside the image, keypoints without map point and outliers."""
    rng = np.random.default_rng(120000 + index)
    fx, fy, cx, cy = K
    # small motion: rotation by a few degrees about a random axis + translation
    axis = rng.normal(size=3); axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rng.uniform(0.5, 3.0))
    Kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = (np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx).astype(np.float32)
    t = rng.normal(0, 0.05, 3).astype(np.float32)
    last = random_keypoints(n_last, width, height, rng, nlevels)
    u = rng.uniform(-20, width + 20, n_last); v = rng.uniform(-20, height + 20, n_last)
    z = rng.uniform(1.0, 10.0, n_last)
    z[rng.random(n_last) < 0.03] *= -1                                     # behind the camera
    Xc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)
    wpos = ((Xc - t) @ R).astype(np.float32)                               # Xw = R^T (Xc - t)
    mp_desc = rng.integers(0, 256, (n_last, 32)).astype(np.uint8)
    has_mp = (rng.random(n_last) < 0.9).astype(np.uint8)
    outlier = (rng.random(n_last) < 0.05).astype(np.uint8)
    mp_obs = ((rng.random(n_last) < 0.95) * rng.integers(1, 9, n_last)).astype(np.int32)
    # current frame: most keypoints are noisy re-observations of last-frame map points
    cur = random_keypoints(n_cur, width, height, rng, nlevels)
    src = rng.integers(0, n_last, n_cur)
    re = rng.random(n_cur) < 0.8
    cur["x"] = np.where(re, np.clip(u[src] + rng.normal(0, 2.5, n_cur), 0, width - 1), cur["x"]).astype(np.float32)
    cur["y"] = np.where(re, np.clip(v[src] + rng.normal(0, 2.5, n_cur), 0, height - 1), cur["y"]).astype(np.float32)
    cur["octave"] = np.where(re, np.clip(last["octave"][src] + rng.integers(-1, 2, n_cur), 0, nlevels - 1), cur["octave"])
    rot = rng.uniform(0, 40)
    cur["angle"] = np.where(re, np.mod(last["angle"][src] - rot + rng.normal(0, 5, n_cur), 360), cur["angle"]).astype(np.float32)
    cdesc = rng.integers(0, 256, (n_cur, 32)).astype(np.uint8)
    cdesc[re] = flip_bits(mp_desc[src[re]], rng.integers(0, 70, int(re.sum())), rng)
    return dict(has_mp=has_mp, outlier=outlier, wpos=wpos, mp_desc=mp_desc, mp_obs=mp_obs,
                last_octave=last["octave"].astype(np.int32), last_angle=last["angle"].astype(np.float32),
                Rcw=R.reshape(9), tcw=t, K=np.array(K, np.float32), cur=cur, cdesc=cdesc)


def camera_centre(Rcw, tcw):
    """Ow = -Rcw^T tcw in float32, products and sums in source order (the 3x3 cv::Mat expression of
    S/ORBmatcher.cc:1482 as the oracle's harness evaluates it)."""
    R = np.asarray(Rcw, np.float32).reshape(3, 3)
    t = np.asarray(tcw, np.float32)
    out = np.zeros(3, np.float32)
    for r in range(3):
        acc = np.float32(R[0, r] * t[0])
        acc = np.float32(acc + np.float32(R[1, r] * t[1]))
        acc = np.float32(acc + np.float32(R[2, r] * t[2]))
        out[r] = -acc
    return out


def relocalisation_frame(index, n_kf=1500, n_cur=2000, width=640, height=480, nlevels=8, scale=1.2,
                         K=(526.69, 540.36, 313.07, 238.39)):
    """A (current frame, key frame) pair for SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist)
    (relocalisation): key-frame map-point slots with world positions and scale-invariance distances, some empty
    (valid 0), already found (2) or bad (3); the current frame's keypoints sit near the projections at about the
    predicted level.  Includes projections outside the image, points behind the camera, distances outside
    [0.8 min, 1.2 max] and pre-occupied keypoints; stays off the band [0.8 min, min) where the reference
    indexes mvScaleFactors out of range."""
    rng = np.random.default_rng(130000 + index)
    fx, fy, cx, cy = K
    axis = rng.normal(size=3); axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rng.uniform(2.0, 25.0))
    Kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = (np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx).astype(np.float32)
    t = rng.normal(0, 0.5, 3).astype(np.float32)
    u = rng.uniform(-20, width + 20, n_kf); v = rng.uniform(-20, height + 20, n_kf)
    z = rng.uniform(1.0, 10.0, n_kf)
    z[rng.random(n_kf) < 0.03] *= -1
    Xc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)
    wpos = ((Xc - t) @ R.astype(np.float64)).astype(np.float32)
    Ow = camera_centre(R, t)
    dist = np.linalg.norm(wpos.astype(np.float64) - Ow, axis=1)
    level = rng.integers(0, nlevels, n_kf)
    expo = level - 0.5 + rng.uniform(-0.3, 0.3, n_kf)
    kind = rng.random(n_kf)
    expo = np.where(kind < 0.04, -2.0, expo)                      # too far: dist > 1.2 mfMax
    expo = np.where(kind > 0.96, nlevels + 2.0, expo)             # too close: dist < 0.8 mfMin
    mf_max = (dist * scale ** expo).astype(np.float32)
    mf_min = (mf_max / np.float32(scale ** (nlevels - 1))).astype(np.float32)
    valid = rng.choice(np.array([0, 1, 2, 3], np.uint8), n_kf, p=[0.08, 0.80, 0.07, 0.05])
    mp_desc = rng.integers(0, 256, (n_kf, 32)).astype(np.uint8)
    kf_angle = rng.uniform(0, 360, n_kf).astype(np.float32)
    cur = random_keypoints(n_cur, width, height, rng, nlevels)
    src = rng.integers(0, max(n_kf, 1), n_cur)
    re = (rng.random(n_cur) < 0.8) & (n_kf > 0)
    if n_kf == 0:
        u = v = np.zeros(1); level = np.zeros(1, np.int64); kf_angle = np.zeros(1, np.float32); mp_desc = np.zeros((1, 32), np.uint8)
    cur["x"] = np.where(re, np.clip(u[src] + rng.normal(0, 2.5, n_cur), 0, width - 1), cur["x"]).astype(np.float32)
    cur["y"] = np.where(re, np.clip(v[src] + rng.normal(0, 2.5, n_cur), 0, height - 1), cur["y"]).astype(np.float32)
    cur["octave"] = np.where(re, np.clip(level[src] + rng.integers(-1, 2, n_cur), 0, nlevels - 1), cur["octave"])
    rot = rng.uniform(0, 40)
    cur["angle"] = np.where(re, np.mod(kf_angle[src] - rot + rng.normal(0, 5, n_cur), 360), cur["angle"]).astype(np.float32)
    cdesc = rng.integers(0, 256, (n_cur, 32)).astype(np.uint8)
    cdesc[re] = flip_bits(mp_desc[src[re]], rng.integers(0, 130, int(re.sum())), rng)
    kp_mp = np.where(rng.random(n_cur) < 0.1, -2, -1).astype(np.int32)     # keypoints that already hold a map point
    if n_kf == 0:
        kf_angle = np.zeros(0, np.float32); mp_desc = np.zeros((0, 32), np.uint8)
    return dict(valid=valid, wpos=wpos, mp_desc=mp_desc, mf_max=mf_max, mf_min=mf_min, kf_angle=kf_angle,
                Rcw=R.reshape(9), tcw=t, Ow=Ow, K=np.array(K, np.float32), cur=cur, cdesc=cdesc, kp_mp=kp_mp,
                log_scale=np.float32(np.log(np.float32(scale))))


def flatten_feature_vector(node_of_feature):
    """DBoW2::FeatureVector (std::map<NodeId, vector<unsigned>> filled by addFeature in feature order) as arrays:
    node ids ascending, node_start (nn + 1 offsets) and the feature indices node after node."""
    node_of_feature = np.asarray(node_of_feature, np.int64)
    order = np.argsort(node_of_feature, kind="stable")
    nodes, counts = np.unique(node_of_feature, return_counts=True)
    start = np.zeros(len(nodes) + 1, np.int32)
    start[1:] = np.cumsum(counts)
    return nodes.astype(np.uint32), start, order.astype(np.uint32)


def bow_pair(index, n_kf=2000, n_f=2000, n_nodes=100):
    """A (key frame, frame) pair for SearchByBoW: every feature sits in one vocabulary node (about n_nodes distinct
    node ids, sparse values); most frame features re-observe a key-frame feature (some bits flipped) and usually
    fall into the same node.  Key-frame slots without map point (0) and with a bad one (3) are included."""
    rng = np.random.default_rng(140000 + index)
    ids = np.sort(rng.choice(np.arange(10, 10 + 12 * max(n_nodes, 1)), max(n_nodes, 1), replace=False))
    kf_desc = rng.integers(0, 256, (n_kf, 32)).astype(np.uint8)
    kf_node = ids[rng.integers(0, len(ids), n_kf)] if n_kf else np.zeros(0, np.int64)
    kf_valid = rng.choice(np.array([0, 1, 3], np.uint8), n_kf, p=[0.25, 0.70, 0.05])
    kf_angle = rng.uniform(0, 360, n_kf).astype(np.float32)
    f_desc = rng.integers(0, 256, (n_f, 32)).astype(np.uint8)
    f_node = ids[rng.integers(0, len(ids), n_f)] if n_f else np.zeros(0, np.int64)
    f_angle = rng.uniform(0, 360, n_f).astype(np.float32)
    if n_kf and n_f:
        src = rng.integers(0, n_kf, n_f)
        re = rng.random(n_f) < 0.75
        f_desc[re] = flip_bits(kf_desc[src[re]], rng.integers(0, 70, int(re.sum())), rng)
        same = re & (rng.random(n_f) < 0.85)
        f_node = np.where(same, kf_node[src], f_node)
        rot = rng.uniform(0, 40)
        f_angle = np.where(re, np.mod(kf_angle[src] - rot + rng.normal(0, 5, n_f), 360), f_angle).astype(np.float32)
    kn, ks, kfeat = flatten_feature_vector(kf_node)
    fn, fs, ffeat = flatten_feature_vector(f_node)
    # map-point state of the second side, used when it is a key frame too (SearchByBoW(pKF1, pKF2, ...))
    f_valid = rng.choice(np.array([0, 1, 3], np.uint8), n_f, p=[0.25, 0.70, 0.05])
    return dict(kf_valid=kf_valid, kf_desc=kf_desc, kf_angle=kf_angle, kf_node=kn, kf_start=ks, kf_feat=kfeat,
                f_desc=f_desc, f_angle=f_angle, f_node=fn, f_start=fs, f_feat=ffeat, f_valid=f_valid)


def _rodrigues(rng, lo_deg, hi_deg):
    axis = rng.normal(size=3); axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rng.uniform(lo_deg, hi_deg))
    Kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx


def epipole_in_second(Cw1, R2w, t2w, K):
    """ex, ey of S/ORBmatcher.cc:668-675 in float32, products and sums in source order."""
    f = np.float32
    R = np.asarray(R2w, f).reshape(3, 3); C = np.asarray(Cw1, f); t = np.asarray(t2w, f)
    C2 = [f(f(f(f(R[r, 0] * C[0]) + f(R[r, 1] * C[1])) + f(R[r, 2] * C[2])) + t[r]) for r in range(3)]
    invz = f(f(1.0) / C2[2])
    return np.array([f(f(f(f(K[0]) * C2[0]) * invz) + f(K[2])), f(f(f(f(K[1]) * C2[1]) * invz) + f(K[3]))], f)


def triangulation_pair(index, n1=2000, n2=2000, n_nodes=100, width=640, height=480, nlevels=8, scale=1.2,
                       K=(526.69, 540.36, 313.07, 238.39), stereo_fraction=0.0, forward=False):
    """Two key frames looking at the same random 3-D points for SearchForTriangulation: keypoints = projections plus
    noise, F12 from the two poses (LocalMapping::ComputeF12's formula), corresponding features mostly in the same
    vocabulary node with similar descriptors; slots that already hold a map point; optionally stereo keypoints."""
    rng = np.random.default_rng(150000 + index)
    fx, fy, cx, cy = K
    Km = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    R1, R2 = _rodrigues(rng, 0, 10), _rodrigues(rng, 0, 10)
    t1, t2 = rng.normal(0, 0.3, 3), rng.normal(0, 0.3, 3)
    if forward:                                       # motion along the optical axis: the epipole falls inside the image
        R2 = R1 @ _rodrigues(rng, 0, 1.0)
        t2 = t1 + np.array([0.02, 0.01, -1.0])
    npts = max(n1, n2, 1)
    Xw = np.stack([rng.uniform(-4, 4, npts), rng.uniform(-3, 3, npts), rng.uniform(3, 12, npts)], 1)
    def project(R, t):
        Xc = Xw @ R.T + t
        return fx * Xc[:, 0] / Xc[:, 2] + cx, fy * Xc[:, 1] / Xc[:, 2] + cy
    u1, v1 = project(R1, t1); u2, v2 = project(R2, t2)
    R12 = R1 @ R2.T; t12 = -R12 @ t2 + t1
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    F12 = (np.linalg.inv(Km).T @ tx @ R12 @ np.linalg.inv(Km)).astype(np.float32)
    ids = np.sort(rng.choice(np.arange(10, 10 + 12 * max(n_nodes, 1)), max(n_nodes, 1), replace=False))
    pt_node = ids[rng.integers(0, len(ids), npts)]
    pt_desc = rng.integers(0, 256, (npts, 32)).astype(np.uint8)
    pt_oct = rng.integers(0, nlevels, npts)
    pt_ang = rng.uniform(0, 360, npts)
    rot = rng.uniform(0, 40)
    def side(n, u, v, other_rot):
        src = rng.permutation(npts)[:n] if n else np.zeros(0, np.int64)
        re = rng.random(n) < 0.8                      # the rest are unrelated features
        x = np.where(re, u[src] + rng.normal(0, 1.0, n), rng.uniform(0, width, n)).astype(np.float32)
        y = np.where(re, v[src] + rng.normal(0, 1.0, n), rng.uniform(0, height, n)).astype(np.float32)
        node = np.where(re & (rng.random(n) < 0.85), pt_node[src], ids[rng.integers(0, len(ids), n)])
        desc = rng.integers(0, 256, (n, 32)).astype(np.uint8)
        if n:
            desc[re] = flip_bits(pt_desc[src[re]], rng.integers(0, 45, int(re.sum())), rng)
        octv = np.clip(pt_oct[src] + rng.integers(-1, 2, n), 0, nlevels - 1).astype(np.int32)
        ang = np.mod(pt_ang[src] - other_rot + rng.normal(0, 5, n), 360).astype(np.float32)
        has_mp = (rng.random(n) < 0.4).astype(np.uint8)
        ur = np.where(rng.random(n) < stereo_fraction, x - rng.uniform(1, 30, n), -1).astype(np.float32)
        nn, st, ft = flatten_feature_vector(node)
        return dict(x=x, y=y, octave=octv, angle=ang, desc=desc, has_mp=has_mp, u_right=ur, node=nn, start=st, feat=ft)
    a, b = side(n1, u1, v1, 0.0), side(n2, u2, v2, rot)
    Cw1 = (-R1.T @ t1).astype(np.float32)
    R2f, t2f = R2.astype(np.float32), t2.astype(np.float32)
    sf = np.array([scale ** i for i in range(nlevels)], np.float32)
    return dict(k1=a, k2=b, F12=F12.reshape(9), Cw1=Cw1, R2w=R2f.reshape(9), t2w=t2f, K=np.array(K, np.float32),
                epipole=epipole_in_second(Cw1, R2f, t2f, np.array(K, np.float32)), scale_factors=sf,
                level_sigma2=(sf * sf).astype(np.float32))


def observed_descriptors(index, sizes):
    """Per map point, the descriptors of its observations: noisy copies of one random descriptor (0..90 flipped bits,
    so duplicates and ties occur)."""
    rng = np.random.default_rng(160000 + index)
    out = []
    for n in sizes:
        base = rng.integers(0, 256, (1, 32)).astype(np.uint8)
        out.append(flip_bits(np.repeat(base, n, 0), rng.integers(0, 90, n), rng) if n else np.zeros((0, 32), np.uint8))
    return out


def fuse_frame(index, n_mp=3000, n_kp=2000, width=640, height=480, nlevels=8, scale=1.2,
               K=(526.69, 540.36, 313.07, 238.39), bf=40.0, stereo_fraction=0.2):
    """A (key frame, candidate map points) pair for the search inside ORBmatcher::Fuse: map points with world
    position, mean viewing direction, descriptor and scale-invariance distances; key-frame keypoints near the
    projections at about the predicted level (reprojection noise around the chi-square gates), some with a stereo
    coordinate.  valid: 1 usable, 2 already observed by the key frame, 3 bad, 0 NULL.  Includes points behind the
    camera, outside the image, outside the distance range and seen from too steep an angle."""
    rng = np.random.default_rng(170000 + index)
    fx, fy, cx, cy = K
    R = _rodrigues(rng, 2, 25).astype(np.float32)
    t = rng.normal(0, 0.5, 3).astype(np.float32)
    u = rng.uniform(-20, width + 20, n_mp); v = rng.uniform(-20, height + 20, n_mp)
    z = rng.uniform(1.0, 10.0, n_mp)
    z[rng.random(n_mp) < 0.03] *= -1
    Xc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)
    wpos = ((Xc - t) @ R.astype(np.float64)).astype(np.float32)
    Ow = camera_centre(R, t)
    PO = wpos.astype(np.float64) - Ow
    dist = np.linalg.norm(PO, axis=1)
    nrm = PO / np.maximum(dist, 1e-9)[:, None] + rng.normal(0, 0.25, (n_mp, 3))
    steep = rng.random(n_mp) < 0.1
    nrm[steep] = rng.normal(size=(int(steep.sum()), 3))
    nrm = (nrm / np.linalg.norm(nrm, axis=1)[:, None]).astype(np.float32)
    level = rng.integers(0, nlevels, n_mp)
    expo = level - 0.5 + rng.uniform(-0.3, 0.3, n_mp)
    kind = rng.random(n_mp)
    expo = np.where(kind < 0.04, -2.0, expo)
    expo = np.where(kind > 0.96, nlevels + 2.0, expo)
    mf_max = (dist * scale ** expo).astype(np.float32)
    mf_min = (mf_max / np.float32(scale ** (nlevels - 1))).astype(np.float32)
    valid = rng.choice(np.array([0, 1, 2, 3], np.uint8), n_mp, p=[0.05, 0.82, 0.08, 0.05])
    mp_desc = rng.integers(0, 256, (n_mp, 32)).astype(np.uint8)
    kp = random_keypoints(n_kp, width, height, rng, nlevels)
    src = rng.integers(0, max(n_mp, 1), n_kp)
    re = (rng.random(n_kp) < 0.8) & (n_mp > 0)
    if n_mp == 0:
        u = v = z = np.ones(1); level = np.zeros(1, np.int64); mp_desc0 = np.zeros((1, 32), np.uint8)
    else:
        mp_desc0 = mp_desc
    sig = scale ** np.clip(level[src], 0, nlevels - 1)
    kp["x"] = np.where(re, np.clip(u[src] + rng.normal(0, 1.2, n_kp) * sig, 0, width - 1), kp["x"]).astype(np.float32)
    kp["y"] = np.where(re, np.clip(v[src] + rng.normal(0, 1.2, n_kp) * sig, 0, height - 1), kp["y"]).astype(np.float32)
    kp["octave"] = np.where(re, np.clip(level[src] + rng.integers(-1, 2, n_kp), 0, nlevels - 1), kp["octave"])
    kdesc = rng.integers(0, 256, (n_kp, 32)).astype(np.uint8)
    kdesc[re] = flip_bits(mp_desc0[src[re]], rng.integers(0, 80, int(re.sum())), rng)
    stereo = rng.random(n_kp) < stereo_fraction
    ur = np.where(stereo, kp["x"] - bf / np.abs(z[src]) + rng.normal(0, 1.0, n_kp), -1).astype(np.float32)
    ur = np.where(stereo & (ur < 0), 0.5, ur).astype(np.float32)
    sf = np.array([scale ** i for i in range(nlevels)], np.float32)
    return dict(valid=valid, wpos=wpos, normal=nrm, mp_desc=mp_desc, mf_max=mf_max, mf_min=mf_min, Rcw=R.reshape(9), tcw=t, Ow=Ow,
                K=np.array(K, np.float32), bf=np.float32(bf), kp=kp, kdesc=kdesc, u_right=ur, scale_factors=sf,
                inv_level_sigma2=(np.float32(1.0) / (sf * sf)).astype(np.float32), log_scale=np.float32(np.log(np.float32(scale))))


def sim3_pair(index, n1=1500, n2=1500, width=640, height=480, nlevels=8, scale=1.2, K=(526.69, 540.36, 313.07, 238.39)):
    """Two key frames observing a common cloud, each with its OWN map points (slot i of key frame k holds map point i of
    k: world position, descriptor, scale-invariance distances; valid 0 = empty slot, 1 good, 3 bad), and a relative
    similarity (s12 = 1, R12, t12) close to the true one -- the inputs of SearchBySim3.  Returns the two sides in
    fuse_frame() layout (each side's points are searched in the OTHER side's keypoints) plus the transforms."""
    rng = np.random.default_rng(180000 + index)
    fx, fy, cx, cy = K
    R1, R2 = _rodrigues(rng, 0, 8).astype(np.float32), _rodrigues(rng, 0, 8).astype(np.float32)
    t1, t2 = rng.normal(0, 0.2, 3).astype(np.float32), rng.normal(0, 0.2, 3).astype(np.float32)
    npts = max(n1, n2, 1)
    P = np.stack([rng.uniform(-4, 4, npts), rng.uniform(-3, 3, npts), rng.uniform(3, 12, npts)], 1)
    base_desc = rng.integers(0, 256, (npts, 32)).astype(np.uint8)
    sf = np.array([scale ** i for i in range(nlevels)], np.float32)

    def side(n, R, t):
        src = rng.permutation(npts)[:n] if n else np.zeros(0, np.int64)
        X = (P[src] + rng.normal(0, 0.01, (n, 3))).astype(np.float32)
        Xc = X.astype(np.float64) @ R.astype(np.float64).T + t
        z = np.where(np.abs(Xc[:, 2]) < 1e-3, 1e-3, Xc[:, 2]) if n else np.zeros(0)
        u = fx * Xc[:, 0] / z + cx; v = fy * Xc[:, 1] / z + cy
        dist = np.linalg.norm(Xc, axis=1)
        level = rng.integers(0, nlevels - 1, n)       # seen from the other key frame the level may grow by one: stay in range
        mf_max = (dist * scale ** (level - 0.5 + rng.uniform(-0.3, 0.3, n))).astype(np.float32)
        mf_min = (mf_max / np.float32(scale ** (nlevels - 1))).astype(np.float32)
        mp_desc = flip_bits(base_desc[src], rng.integers(0, 40, n), rng) if n else np.zeros((0, 32), np.uint8)
        kp = random_keypoints(n, width, height, rng, nlevels)
        kp["x"] = np.clip(u + rng.normal(0, 1.5, n), 0, width - 1).astype(np.float32)
        kp["y"] = np.clip(v + rng.normal(0, 1.5, n), 0, height - 1).astype(np.float32)
        kp["octave"] = np.clip(level + rng.integers(-1, 1, n), 0, nlevels - 1)
        kdesc = flip_bits(mp_desc, rng.integers(0, 40, n), rng) if n else np.zeros((0, 32), np.uint8)
        valid = rng.choice(np.array([0, 1, 3], np.uint8), n, p=[0.2, 0.75, 0.05])
        return dict(valid=valid, wpos=X, normal=np.zeros((n, 3), np.float32), mp_desc=mp_desc, mf_max=mf_max, mf_min=mf_min,
                    kp=kp, kdesc=kdesc, u_right=np.full(n, -1, np.float32), Rcw=R.reshape(9), tcw=t)
    a, b = side(n1, R1, t1), side(n2, R2, t2)
    R12 = (R1.astype(np.float64) @ R2.astype(np.float64).T @ _rodrigues(rng, 0, 0.3)).astype(np.float32)
    t12 = (t1 - R12.astype(np.float64) @ t2 + rng.normal(0, 0.005, 3)).astype(np.float32)
    sR21 = np.ascontiguousarray(R12.T)                                   # (1/s12) * R12^T with s12 = 1
    t21 = camera_centre(R12, t12)                                        # -sR21 * t12, float in source order
    common = dict(K=np.array(K, np.float32), bf=np.float32(0), scale_factors=sf, inv_level_sigma2=(np.float32(1) / (sf * sf)).astype(np.float32),
                  log_scale=np.float32(np.log(np.float32(scale))), Ow=np.zeros(3, np.float32))
    # leg 1: key frame 1's map points searched in key frame 2's keypoints; leg 2 the other way round
    leg1 = dict(common, **{k: a[k] for k in ("valid", "wpos", "normal", "mp_desc", "mf_max", "mf_min", "Rcw", "tcw")}, kp=b["kp"], kdesc=b["kdesc"], u_right=b["u_right"])
    leg2 = dict(common, **{k: b[k] for k in ("valid", "wpos", "normal", "mp_desc", "mf_max", "mf_min", "Rcw", "tcw")}, kp=a["kp"], kdesc=a["kdesc"], u_right=a["u_right"])
    return dict(k1=a, k2=b, leg1=leg1, leg2=leg2, R12=R12.reshape(9), t12=t12, sR21=sR21.reshape(9), t21=t21, common=common)


def synthetic_vocabulary(index, k=6, L=4, stopped=0.05):
    """A DBoW2 vocabulary tree in the flattened form of include/orb_b200.h (orbb200_vocabulary_create): breadth-first
    node numbering exactly as TemplatedVocabulary::loadFromTextFile assigns it when the nodes are written in this
    order (node id = line number, word id = running count of leaves).  Inner nodes have 2..k children whose
    descriptors are the parent's with some bits flipped; all leaves sit at level L; a few words are stopped (weight 0)."""
    rng = np.random.default_rng(190000 + index)
    parent, level, desc = [0], [0], [np.zeros(32, np.uint8)]
    frontier = [0]
    for lv in range(1, L + 1):
        nxt = []
        for pnode in frontier:
            nch = int(rng.integers(2, k + 1))
            base = desc[pnode] if pnode else rng.integers(0, 256, 32).astype(np.uint8)
            for _ in range(nch):
                d = flip_bits((rng.integers(0, 256, (1, 32)).astype(np.uint8) if pnode == 0 else base[None]), np.array([int(rng.integers(20, 70))]), rng)[0]
                parent.append(pnode); level.append(lv); desc.append(d); nxt.append(len(parent) - 1)
        frontier = nxt
    n = len(parent)
    parent = np.array(parent, np.int32); level = np.array(level, np.int32)
    is_leaf = level == L
    children = [[] for _ in range(n)]
    for i in range(1, n):
        children[parent[i]].append(i)
    child_start = np.zeros(n + 1, np.int32)
    child_start[1:] = np.cumsum([len(c) for c in children])
    child_list = np.array([c for cs in children for c in cs], np.int32)
    word_id = np.full(n, -1, np.int32)
    word_id[is_leaf] = np.arange(int(is_leaf.sum()))
    weight = np.zeros(n, np.float64)
    weight[is_leaf] = np.round(rng.uniform(0.5, 9.0, int(is_leaf.sum())), 6)
    weight[is_leaf & (rng.random(n) < stopped)] = 0.0
    return dict(k=k, L=L, parent=parent, is_leaf=is_leaf, child_start=child_start, children=child_list,
                desc=np.stack(desc).astype(np.uint8), word_id=word_id, weight=weight)


def write_vocabulary_text(path, voc):
    """ORBvoc.txt format (TemplatedVocabulary::saveToTextFile): header "k L scoring weighting" (0 0 = L1_NORM, TF_IDF),
    then one line per node "parent isLeaf d0 .. d31 weight".  No trailing newline: the reference's loader would turn
    one into an extra, empty child of the root."""
    lines = ["%d %d 0 0" % (voc["k"], voc["L"])]
    for i in range(1, len(voc["parent"])):
        lines.append("%d %d %s %.6f" % (voc["parent"][i], int(voc["is_leaf"][i]), " ".join(str(int(b)) for b in voc["desc"][i]), voc["weight"][i]))
    with open(path, "w") as f:
        f.write("\n".join(lines))


def vocabulary_features(index, voc, n):
    """n descriptors near random leaves of the vocabulary (so that the descent is not trivial)."""
    rng = np.random.default_rng(191000 + index)
    leaves = np.nonzero(voc["is_leaf"])[0]
    pick = leaves[rng.integers(0, len(leaves), n)] if n else np.zeros(0, np.int64)
    return flip_bits(voc["desc"][pick], rng.integers(0, 60, n), rng) if n else np.zeros((0, 32), np.uint8)
