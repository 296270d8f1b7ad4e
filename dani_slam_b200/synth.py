"""Seeded synthetic inputs for the ORB front-end.

Pure-integer numpy only, so that the same seed gives the same bytes everywhere (no cv2, no float
filters).  Throughput frames are corner-dense (every pyramid level fills its quadtree quota); parity
frames add flat, low-contrast, tie-heavy and odd-size content to hit the reference's quirks
(empty cells, the minThFAST retry `ORBextractor.cc:843-846`, equal-score NMS / quadtree ties).
"""
from __future__ import annotations

import numpy as np


def _box3(a: np.ndarray) -> np.ndarray:
    """3×3 box filter with edge replication, integer arithmetic (sum // 9)."""
    p = np.pad(a.astype(np.int32), 1, mode="edge")
    s = np.zeros_like(a, dtype=np.int32)
    for dy in range(3):
        for dx in range(3):
            s += p[dy:dy + a.shape[0], dx:dx + a.shape[1]]
    return (s // 9).astype(np.uint8)


def _rects(img: np.ndarray, rng: np.random.Generator, count: int, lo: int = 4, hi: int = 20) -> None:
    h, w = img.shape
    xs = rng.integers(0, w, size=count)
    ys = rng.integers(0, h, size=count)
    ws = rng.integers(lo, hi + 1, size=count)
    hs = rng.integers(lo, hi + 1, size=count)
    gs = rng.integers(0, 256, size=count)
    for x, y, rw, rh, g in zip(xs, ys, ws, hs, gs):
        img[y:y + rh, x:x + rw] = g


def throughput_frame(seed: int, width: int = 640, height: int = 480) -> np.ndarray:
    """Throughput frame exactly as SURVEY.md §8(d) prescribes: u8 uniform noise low-passed by two 3×3 integer box
    passes, plus ~W·H/1500 random axis-aligned grey rectangles (4–20 px).  Every pyramid level fills its
    quadtree quota (≈21.6k FAST candidates → 1002…1006 keypoints per 640×480 frame at nFeatures=1000)."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(height, width), dtype=np.uint8)
    img = _box3(_box3(img))
    _rects(img, rng, max(1, width * height // 1500))
    return np.ascontiguousarray(img)


def parity_frame(seed: int, width: int = 640, height: int = 480) -> np.ndarray:
    """Quadrants: corner-dense | flat ; low-contrast (amplitude 8–19) | regular dots (ties)."""
    rng = np.random.default_rng(seed)
    img = throughput_frame(seed, width, height)
    hh, hw = height // 2, width // 2
    # top-right: constant grey → empty cells
    img[:hh, hw:] = int(rng.integers(40, 216))
    # bottom-left: texture of amplitude 8..19 → empty at FAST 20, corners at FAST 7
    base = int(rng.integers(60, 180))
    amp = int(rng.integers(8, 20))
    tex = rng.integers(0, 2, size=(height - hh, hw), dtype=np.uint8) * amp
    blk = np.kron(rng.integers(0, 2, size=((height - hh + 5) // 6, (hw + 5) // 6), dtype=np.uint8),
                  np.ones((6, 6), dtype=np.uint8))[: height - hh, :hw] * amp
    img[hh:, :hw] = (base + np.where(rng.integers(0, 4, size=tex.shape) == 0, tex, blk)).astype(np.uint8)
    # bottom-right: regular dots and a checkerboard → equal responses everywhere
    quad = np.full((height - hh, width - hw), 100, dtype=np.uint8)
    quad[4::9, 4::9] = 220
    quad[5::9, 4::9] = 220
    cb = (np.add.outer(np.arange(quad.shape[0]) // 8, np.arange(quad.shape[1]) // 8) & 1).astype(np.uint8)
    half = quad.shape[0] // 2
    quad[half:] = np.where(cb[half:] == 1, 200, 60)
    img[hh:, hw:] = quad
    return np.ascontiguousarray(img)


def stereo_right(left: np.ndarray, seed: int, max_disp: int = 40) -> np.ndarray:
    """Right image = left shifted by a per-row-band constant disparity (so matches exist)."""
    rng = np.random.default_rng(seed + 1_000_000)
    h, w = left.shape
    right = np.empty_like(left)
    band = 94
    for y0 in range(0, h, band):
        d = int(rng.integers(1, max_disp + 1))
        rows = left[y0:y0 + band]
        right[y0:y0 + band, : w - d] = rows[:, d:]
        right[y0:y0 + band, w - d:] = rows[:, w - d - 1: w - d]
    return right


def descriptors(n: int, seed: int) -> np.ndarray:
    """n i.i.d. uniform 256-bit descriptors, shape (n, 32) uint8."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(n, 32), dtype=np.uint8)


def knn_case(nq: int, ndb: int, seed: int = 1234, planted_frac: float = 0.01, dup_rows: int = 4):
    """Config-4 style case: random DB, random queries, a fraction of queries planted as DB rows with
    k∈[0,40] flipped bits, plus a few exact duplicate DB rows (tie handling: lower index first)."""
    rng = np.random.default_rng(seed)
    db = rng.integers(0, 256, size=(ndb, 32), dtype=np.uint8)
    q = rng.integers(0, 256, size=(nq, 32), dtype=np.uint8)
    n_plant = max(1, int(nq * planted_frac)) if ndb > 0 else 0
    for i in range(n_plant):
        src = int(rng.integers(0, ndb))
        row = db[src].copy()
        k = int(rng.integers(0, 41))
        bits = rng.choice(256, size=k, replace=False)
        for b in bits:
            row[b >> 3] ^= np.uint8(1 << (b & 7))
        q[i] = row
        if i < dup_rows and ndb > 8:
            # duplicate the source row at a later (and an earlier) index to exercise ties
            db[min(ndb - 1, src + 1 + int(rng.integers(0, 5)))] = db[src]
    return q, db


def vocabulary(k: int = 10, L: int = 3, seed: int = 7, ragged: float = 0.0, stop_frac: float = 0.05, flips: int = 40,
               scoring: int = 0, weighting: int = 0, min_leaf_level: int = 1) -> dict:
    """Synthetic DBoW2-style vocabulary tree in the node-stream form of the reference's text format
    (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1378-1420): node i+1 has `parent[i]` (0 = root), a leaf flag, a
    256-bit descriptor and a weight; ids are assigned in stream order, children of a node in the order they appear.
    Children are their parent's descriptor with `flips` random bits flipped (so that descents are meaningful and
    ties between siblings occur), laid out breadth first like a trained ORBvoc.  `ragged` makes that fraction of
    the inner nodes at levels >= min_leaf_level leaves early; `stop_frac` of the words get weight 0 (stopped)."""
    rng = np.random.default_rng(seed)
    parent, leaf, desc, weight = [], [], [], []
    frontier = [(0, np.zeros(32, np.uint8), 0)]           # (node id, descriptor, level)
    frontier[0] = (0, rng.integers(0, 256, 32, dtype=np.uint8), 0)
    next_id = 1
    while frontier:
        new_frontier = []
        for nid, d, lvl in frontier:
            nch = k if rng.random() > 0.15 else int(rng.integers(2, k + 1))    # not every node is full
            for _ in range(nch):
                cd = d.copy()
                for b in rng.choice(256, size=int(rng.integers(0, flips + 1)), replace=False):
                    cd[b >> 3] ^= np.uint8(1 << (b & 7))
                is_leaf = (lvl + 1 == L) or (lvl + 1 >= min_leaf_level and rng.random() < ragged)
                parent.append(nid); leaf.append(1 if is_leaf else 0); desc.append(cd)
                if is_leaf:
                    weight.append(0.0 if rng.random() < stop_frac else float(rng.uniform(0.05, 9.0)))
                else:
                    weight.append(0.0)
                    new_frontier.append((next_id, cd, lvl + 1))
                next_id += 1
        frontier = new_frontier
    return dict(k=k, L=L, scoring=scoring, weighting=weighting, parent=np.asarray(parent, np.int32), is_leaf=np.asarray(leaf, np.uint8),
                desc=np.stack(desc).astype(np.uint8), weight=np.asarray(weight, np.float64))


def vocabulary_complete(k: int = 10, L: int = 6, seed: int = 7, flips: int = 48, stop_frac: float = 0.0, scoring: int = 0,
                        weighting: int = 0) -> dict:
    """Complete k-ary tree of depth L in breadth-first order (the shape of ORBvoc.txt: k = 10, L = 6, 1.1 M nodes),
    generated level by level with numpy: a child is its parent XOR a random mask of about `flips` bits."""
    rng = np.random.default_rng(seed)
    root = rng.integers(0, 256, (1, 32), dtype=np.uint8)
    parents, descs, leafs = [], [], []
    cur, first_id = root, 0
    for lvl in range(1, L + 1):
        n = cur.shape[0] * k
        mask = np.packbits(rng.random((n, 256)) < flips / 256.0, axis=1)
        child = np.repeat(cur, k, axis=0) ^ mask
        parents.append(np.repeat(np.arange(first_id, first_id + cur.shape[0], dtype=np.int32), k))
        descs.append(child)
        leafs.append(np.full(n, 1 if lvl == L else 0, np.uint8))
        first_id = 1 + sum(len(p_) for p_ in parents[:-1])
        cur = child
    parent = np.concatenate(parents); desc = np.concatenate(descs); leaf = np.concatenate(leafs)
    weight = np.where(leaf == 1, rng.uniform(0.05, 9.0, len(leaf)), 0.0)
    if stop_frac > 0:
        weight = np.where(rng.random(len(leaf)) < stop_frac, 0.0, weight)
    return dict(k=k, L=L, scoring=scoring, weighting=weighting, parent=parent.astype(np.int32), is_leaf=leaf, desc=desc, weight=weight.astype(np.float64))


def write_vocabulary_text(voc: dict, path: str) -> None:
    """The reference's ORBvoc.txt format (TemplatedVocabulary::saveToTextFile, TemplatedVocabulary.h:1428-1451): header
    "k L scoring weighting", then per node "parent isLeaf d0 … d31 weight".  No trailing newline: the reference's
    `while(!f.eof())` loader reads a phantom node from a final empty line."""
    lines = ["%d %d %d %d" % (voc["k"], voc["L"], voc["scoring"], voc["weighting"])]
    for p_, l_, d_, w_ in zip(voc["parent"], voc["is_leaf"], voc["desc"], voc["weight"]):
        lines.append("%d %d %s %s" % (p_, l_, " ".join(str(int(x)) for x in d_), repr(float(w_))))
    with open(path, "w") as f:
        f.write("\n".join(lines))


def vocabulary_queries(voc: dict, n: int, seed: int = 11, flips: int = 60) -> np.ndarray:
    """n descriptors near random vocabulary nodes (so that all branches get traffic) plus some pure noise."""
    rng = np.random.default_rng(seed)
    src = voc["desc"][rng.integers(0, len(voc["desc"]), size=n)].copy()
    for i in range(n):
        if i % 7 == 0:
            src[i] = rng.integers(0, 256, 32, dtype=np.uint8)
            continue
        for b in rng.choice(256, size=int(rng.integers(0, flips + 1)), replace=False):
            src[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return src


# ---- matcher scenes (SearchForInitialization / SearchByProjection / stereo association) ----
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])


def _flip_bits(rows: np.ndarray, rng: np.random.Generator, max_flips: int) -> np.ndarray:
    out = rows.copy()
    for i in range(len(out)):
        for bit in rng.choice(256, size=int(rng.integers(0, max_flips + 1)), replace=False):
            out[i, bit >> 3] ^= np.uint8(1 << (bit & 7))
    return out


def keypoint_records(n: int, seed: int, width: float = 640.0, height: float = 480.0, levels: int = 8, integer_frac: float = 0.3,
                     spill: float = 4.0) -> np.ndarray:
    """n keypoint records (28-byte cv::KeyPoint layout) scattered over the image (a few outside it, like undistorted keypoints;
    a fraction on integer / half-cell coordinates so the 64×48 grid's round() sees exact .5 cases)."""
    rng = np.random.default_rng(seed)
    k = np.zeros(n, KP_DTYPE)
    k["x"] = rng.uniform(-spill, width + spill, n).astype(np.float32)
    k["y"] = rng.uniform(-spill, height + spill, n).astype(np.float32)
    snap = rng.random(n) < integer_frac
    k["x"][snap] = np.round(k["x"][snap]); k["y"][snap] = np.round(k["y"][snap])
    if n > 40:
        k["x"][:16] = (np.arange(16) * (width / 64) + width / 128).astype(np.float32)      # exactly half a grid cell
        k["y"][16:32] = (np.arange(16) * (height / 48) + height / 96).astype(np.float32)
    k["octave"] = rng.choice(levels, n, p=np.array([1.2 ** -i for i in range(levels)]) / sum(1.2 ** -i for i in range(levels)))
    k["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    k["size"] = 31.0
    k["response"] = rng.integers(7, 120, n).astype(np.float32)
    k["class_id"] = -1
    return k


def init_scene(n1: int, n2: int, seed: int, width: float = 640.0, height: float = 480.0, max_shift: float = 30.0, max_flips: int = 70):
    """Two frames for SearchForInitialization: frame 2 holds moved, noisy copies of frame-1 keypoints (plus clutter), so windows
    overlap, train keypoints get locked and stolen, and rotations cluster in a few histogram bins."""
    rng = np.random.default_rng(seed)
    k1 = keypoint_records(n1, seed * 7 + 1, width, height)
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    k2 = keypoint_records(n2, seed * 7 + 2, width, height)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    src = rng.integers(0, n1, n2) if n1 else np.zeros(n2, np.int64)
    copy = rng.random(n2) < 0.8 if n1 else np.zeros(n2, bool)
    k2["x"][copy] = (k1["x"][src[copy]] + rng.uniform(-max_shift, max_shift, copy.sum())).astype(np.float32)
    k2["y"][copy] = (k1["y"][src[copy]] + rng.uniform(-max_shift, max_shift, copy.sum())).astype(np.float32)
    k2["octave"][copy] = np.maximum(0, k1["octave"][src[copy]] + rng.integers(-1, 2, copy.sum()))
    k2["angle"][copy] = ((k1["angle"][src[copy]] - rng.choice([0.0, 2.0, 14.99, 15.0, 31.0, 200.0], copy.sum(), p=[.5, .2, .05, .05, .1, .1])) % 360).astype(np.float32)
    d2[copy] = _flip_bits(d1[src[copy]], rng, max_flips)
    if n2 > 8 and n1 > 8:                                                  # identical descriptors next to each other: ties
        d2[1] = d2[0]; k2["x"][1] = k2["x"][0] + 1; k2["y"][1] = k2["y"][0]
    return k1, d1, k2, d2


def projection_scene(n: int, m: int, seed: int, width: float = 640.0, height: float = 480.0, levels: int = 8, stereo: bool = False):
    """A frame (n keypoints) and m projected map points for SearchByProjection(Frame&, vector<MapPoint*>&): map points are noisy
    copies of frame keypoints at the same or the next level (so best / second-best fall on equal and on different levels), several
    map points compete for the same keypoint, some keypoints already hold a map point with or without observations."""
    rng = np.random.default_rng(seed)
    k = keypoint_records(n, seed * 5 + 3, width, height, levels)
    d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    if n > 20:                                                             # near-duplicate neighbours: ratio test on / across levels
        for i in range(0, min(n - 1, 400), 2):
            k["x"][i + 1] = k["x"][i] + np.float32(rng.uniform(-2, 2)); k["y"][i + 1] = k["y"][i] + np.float32(rng.uniform(-2, 2))
            k["octave"][i + 1] = k["octave"][i] if rng.random() < 0.5 else min(levels - 1, k["octave"][i] + 1)
            d[i + 1] = _flip_bits(d[i:i + 1], rng, 12)[0]
    sf = np.cumprod(np.r_[np.float32(1.0), np.full(levels - 1, np.float32(1.2), np.float32)]).astype(np.float32)
    src = rng.integers(0, n, m) if n else np.zeros(m, np.int64)
    p5 = np.zeros((m, 5), np.float32)
    if n:
        p5[:, 0] = k["x"][src] + rng.uniform(-3, 3, m); p5[:, 1] = k["y"][src] + rng.uniform(-3, 3, m)
    p5[:, 3] = rng.choice([0.9999, 0.99, 0.998, 0.5], m).astype(np.float32)
    p5[:, 4] = rng.uniform(0.5, 80, m)
    lvl = (np.minimum(levels - 1, k["octave"][src] + rng.integers(0, 2, m)) if n else np.zeros(m)).astype(np.int32)
    flags = (1 | (2 * (rng.random(m) < 0.03))).astype(np.uint8)
    flags[rng.random(m) < 0.05] &= 0xFE
    obs = rng.choice([0, 1, 3, 7], m, p=[.1, .3, .3, .3]).astype(np.int32)
    md = _flip_bits(d[src], rng, 90) if n else rng.integers(0, 256, (m, 32), dtype=np.uint8)
    kp_obs = rng.choice([-1, -1, -1, 0, 2], n).astype(np.int32)
    u_right = None
    if stereo:
        u_right = np.where(rng.random(n) < 0.7, k["x"] - rng.uniform(0.5, 40, n), -1.0).astype(np.float32)
        if n:
            p5[:, 2] = u_right[src] + rng.choice([0.0, 1.0, 6.0, 30.0], m)
    return dict(kps=k, desc=d, scale_factors=sf, mp_proj5=p5, mp_level=lvl, mp_flags=flags, mp_obs=obs, mp_desc=md, kp_obs=kp_obs, u_right=u_right)


def bow_scene(n_kf: int, n_f: int, seed: int, n_nodes: int = 40, max_flips: int = 60):
    """A keyframe and a frame for SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&): frame features are noisy copies of keyframe
    features (same vocabulary node most of the time), several keyframe features compete for one frame feature (the ordered walk
    decides), identical descriptors sit next to each other (ties), some keyframe features hold no or a bad map point, a few nodes
    exist on one side only (the lower_bound jumps), rotations cluster in a few histogram bins.
    Returns dict(kf_kps, kf_desc, kf_mp, kf_fv, f_kps, f_desc, f_fv) with fv = (nodes, off, idx) CSR over ascending node ids."""
    rng = np.random.default_rng(seed)
    kk = keypoint_records(n_kf, seed * 11 + 1)
    kd = rng.integers(0, 256, (n_kf, 32), dtype=np.uint8)
    fk = keypoint_records(n_f, seed * 11 + 2)
    fd = rng.integers(0, 256, (n_f, 32), dtype=np.uint8)
    kf_node = rng.integers(0, n_nodes, n_kf) * 3 + 5                      # node ids are sparse
    f_node = rng.integers(0, n_nodes + 4, n_f) * 3 + 5                    # a few nodes only the frame has
    if n_kf and n_f:
        src = rng.integers(0, n_kf, n_f)
        copy = rng.random(n_f) < 0.75
        fd[copy] = _flip_bits(kd[src[copy]], rng, max_flips)
        f_node[copy] = np.where(rng.random(copy.sum()) < 0.9, kf_node[src[copy]], f_node[copy])
        fk["angle"][copy] = ((kk["angle"][src[copy]] - rng.choice([0.0, 3.0, 14.99, 15.0, 45.0, 180.0], copy.sum(), p=[.5, .2, .05, .05, .1, .1])) % 360).astype(np.float32)
        if n_f > 8:                                                        # exact duplicates in one node: first in list order wins
            fd[1] = fd[0]; f_node[1] = f_node[0]
        if n_kf > 8:                                                       # two keyframe features with the same descriptor compete
            kd[1] = kd[0]; kf_node[1] = kf_node[0]
    kf_mp = rng.choice([0, 1, 1, 1, 1, 1, 2], n_kf).astype(np.uint8)

    def csr(node, n):
        order = rng.permutation(n)                                         # stored order inside a node is not sorted by index
        nodes = np.unique(node) if n else np.zeros(0, np.int64)
        off = [0]; idx = []
        for v in nodes:
            members = [int(i) for i in order if node[i] == v]
            idx += members; off.append(len(idx))
        return nodes.astype(np.int32), np.asarray(off, np.int32), np.asarray(idx, np.int32)

    kf_fv, f_fv = csr(kf_node, n_kf), csr(f_node, n_f)
    f_mp = rng.choice([0, 1, 1, 1, 1, 1, 2], n_f).astype(np.uint8)     # used when the second side is a keyframe too (SearchByBoW(KF, KF))
    return dict(kf_kps=kk, kf_desc=kd, kf_mp=kf_mp, kf_fv=kf_fv, f_kps=fk, f_desc=fd, f_fv=f_fv, f_mp=f_mp)
