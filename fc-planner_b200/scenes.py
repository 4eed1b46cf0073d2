"""Inputs without ROS / PCL: a PCD reader, a voxel-centroid down-sampler, and the deterministic
synthetic "plant" scene of BASELINE.json configs[4].

The reference obtains its model cloud, normals and candidate viewpoints from ROSA skeleton
extraction plus `viewpointGeneration` (hcplanner.cpp:535-749), neither of which is on the hot
path; both are wall-clock seeded, so parity is defined on identical inputs
to the oracle and to the CUDA path.  This module only has to be deterministic and representative.
Everything here is numpy on the host; nothing in it is timed.
"""
import struct

import numpy as np

# ------------------------------------------------------------------------------- PCD reading


def _lzf_decompress(src: bytes, out_len: int) -> bytes:
    """LZF (liblzf) decompression, as used by PCD `DATA binary_compressed`."""
    out = bytearray(out_len)
    ip, op, n = 0, 0, len(src)
    while ip < n:
        ctrl = src[ip]
        ip += 1
        if ctrl < 32:  # literal run of ctrl+1 bytes
            ctrl += 1
            out[op:op + ctrl] = src[ip:ip + ctrl]
            ip += ctrl
            op += ctrl
        else:  # back reference
            length = ctrl >> 5
            ref = op - ((ctrl & 0x1F) << 8) - 1
            if length == 7:
                length += src[ip]
                ip += 1
            ref -= src[ip]
            ip += 1
            length += 2
            for _ in range(length):  # may overlap: byte-wise copy
                out[op] = out[ref]
                op += 1
                ref += 1
    return bytes(out[:op])


def read_pcd_xyz(path) -> np.ndarray:
    """Read the x, y, z float32 fields of a .pcd file (ascii, binary or binary_compressed)."""
    with open(path, "rb") as f:
        raw = f.read()
    header, pos = {}, 0
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line or line.startswith("#"):
            continue
        key, _, val = line.partition(" ")
        header[key.upper()] = val.split()
        if key.upper() == "DATA":
            break
    fields = header["FIELDS"]
    sizes = [int(s) for s in header["SIZE"]]
    counts = [int(c) for c in header.get("COUNT", ["1"] * len(fields))]
    npts = int(header["POINTS"][0])
    mode = header["DATA"][0]
    ix = [fields.index(c) for c in ("x", "y", "z")]
    if mode == "ascii":
        arr = np.loadtxt(raw[pos:].decode("ascii").splitlines(), dtype=np.float64, ndmin=2)
        return np.ascontiguousarray(arr[:npts, ix]).astype(np.float32)
    offs = np.cumsum([0] + [s * c for s, c in zip(sizes, counts)])
    stride = int(offs[-1])
    if mode == "binary":
        buf = np.frombuffer(raw, dtype=np.uint8, count=npts * stride, offset=pos).reshape(npts, stride)
        cols = [buf[:, offs[i]:offs[i] + 4].copy().view(np.float32)[:, 0] for i in ix]
        return np.stack(cols, axis=1).astype(np.float32)
    if mode == "binary_compressed":
        comp, uncomp = struct.unpack_from("<II", raw, pos)
        data = _lzf_decompress(raw[pos + 8:pos + 8 + comp], uncomp)
        cols = []
        for i in ix:  # fields are stored one after another (SoA)
            start = int(offs[i]) * npts
            cols.append(np.frombuffer(data, dtype=np.float32, count=npts, offset=start))
        return np.stack(cols, axis=1).astype(np.float32)
    raise ValueError(f"unsupported PCD DATA mode {mode!r}")


def voxel_centroids(xyz: np.ndarray, leaf: float) -> np.ndarray:
    """Voxel-centroid down-sampling (the effect of pcl::VoxelGrid, hcplanner.cpp:143-148):
    one float32 centroid per occupied leaf-sized voxel, in ascending voxel-key order."""
    xyz = np.asarray(xyz, dtype=np.float32)
    mn = xyz.min(axis=0)
    ijk = np.floor((xyz - mn) / np.float32(leaf)).astype(np.int64)
    dims = ijk.max(axis=0) + 1
    key = (ijk[:, 0] * dims[1] + ijk[:, 1]) * dims[2] + ijk[:, 2]
    order = np.argsort(key, kind="stable")
    key_s, pts_s = key[order], xyz[order].astype(np.float64)
    uniq, start = np.unique(key_s, return_index=True)
    sums = np.add.reduceat(pts_s, start, axis=0)
    cnt = np.diff(np.append(start, len(key_s)))[:, None]
    return (sums / cnt).astype(np.float32)


# ------------------------------------------------------------------------------- synthetic plant


def _plant_primitives(rng, scale):
    """3 x 3 lots of 20 m (x scale): alternating boxes and vertical cylinders, plus two pipe racks."""
    boxes, cyls = [], []
    lot = 20.0 * scale
    k = 0
    for ix in range(3):
        for iy in range(3):
            cx = (ix - 1) * lot + rng.uniform(-1.5, 1.5) * scale
            cy = (iy - 1) * lot + rng.uniform(-1.5, 1.5) * scale
            h = rng.uniform(9.0, 28.0) * scale
            if k % 2 == 0:
                sx, sy = rng.uniform(9.0, 13.0, size=2) * scale
                boxes.append((cx, cy, h / 2, sx, sy, h))
            else:
                r = rng.uniform(4.0, 6.0) * scale
                cyls.append((cx, cy, 0.0, h, r))
            k += 1
    # two horizontal pipe racks (thin long boxes) crossing the site at height
    boxes.append((0.0, -0.5 * lot, 7.0 * scale, 2.6 * lot, 1.6 * scale, 1.6 * scale))
    boxes.append((0.5 * lot, 0.0, 12.0 * scale, 1.6 * scale, 2.6 * lot, 1.6 * scale))
    return boxes, cyls


def _inside_any(p, boxes, cyls, margin, skip=None):
    inside = np.zeros(len(p), dtype=bool)
    for i, (cx, cy, cz, sx, sy, sz) in enumerate(boxes):
        if skip == ("b", i):
            continue
        inside |= ((np.abs(p[:, 0] - cx) < sx / 2 - margin) & (np.abs(p[:, 1] - cy) < sy / 2 - margin)
                   & (np.abs(p[:, 2] - cz) < sz / 2 - margin))
    for i, (cx, cy, z0, z1, r) in enumerate(cyls):
        if skip == ("c", i):
            continue
        inside |= (((p[:, 0] - cx) ** 2 + (p[:, 1] - cy) ** 2 < (r - margin) ** 2)
                   & (p[:, 2] > z0 - 1.0) & (p[:, 2] < z1 - margin))
    return inside


def make_plant(n_points=20000, scale=0.5, seed=0):
    """Deterministic synthetic scene (BASELINE.json configs[4] geometry): the union of axis-aligned
    boxes and vertical cylinders standing on z = 0, surfaces sampled on a jittered lattice with
    analytic outward normals.  Returns (points float32 [n,3], normals float64 [n,3]).
    scale = 1.0 gives the 60 x 60 x 30 m plant (about 1e4 m^2 of surface, 1e6 points at 0.1 m).
    """
    rng = np.random.default_rng(seed)
    boxes, cyls = _plant_primitives(rng, scale)
    area = sum(2 * sz * (sx + sy) + sx * sy for (_, _, _, sx, sy, sz) in boxes) + sum(2 * np.pi * r * (z1 - z0) + np.pi * r * r for (_, _, z0, z1, r) in cyls)
    h = float(np.sqrt(area / (1.25 * n_points)))
    pts, nrm = [], []

    def lattice(a0, a1, b0, b1):
        na, nb = max(1, int(round((a1 - a0) / h))), max(1, int(round((b1 - b0) / h)))
        a = a0 + (np.arange(na) + 0.5) * (a1 - a0) / na
        b = b0 + (np.arange(nb) + 0.5) * (b1 - b0) / nb
        A, B = np.meshgrid(a, b, indexing="ij")
        A = A.ravel() + rng.uniform(-0.25, 0.25, A.size) * h
        B = B.ravel() + rng.uniform(-0.25, 0.25, B.size) * h
        return A, B

    for i, (cx, cy, cz, sx, sy, sz) in enumerate(boxes):
        x0, x1, y0, y1, z0, z1 = cx - sx / 2, cx + sx / 2, cy - sy / 2, cy + sy / 2, cz - sz / 2, cz + sz / 2
        faces = [((1, 0, 0), x1), ((-1, 0, 0), x0), ((0, 1, 0), y1), ((0, -1, 0), y0), ((0, 0, 1), z1)]
        if z0 > 1e-9:
            faces.append(((0, 0, -1), z0))
        for n, c in faces:
            if n[0]:
                A, B = lattice(y0, y1, z0, z1)
                p = np.stack([np.full_like(A, c), A, B], 1)
            elif n[1]:
                A, B = lattice(x0, x1, z0, z1)
                p = np.stack([A, np.full_like(A, c), B], 1)
            else:
                A, B = lattice(x0, x1, y0, y1)
                p = np.stack([A, B, np.full_like(A, c)], 1)
            keep = ~_inside_any(p, boxes, cyls, 0.0, skip=("b", i))
            pts.append(p[keep])
            nrm.append(np.tile(np.array(n, dtype=np.float64), (int(keep.sum()), 1)))
    for i, (cx, cy, z0, z1, r) in enumerate(cyls):
        A, B = lattice(0.0, 2 * np.pi * r, z0, z1)  # arc length x height
        th = A / r
        p = np.stack([cx + r * np.cos(th), cy + r * np.sin(th), B], 1)
        n = np.stack([np.cos(th), np.sin(th), np.zeros_like(th)], 1)
        keep = ~_inside_any(p, boxes, cyls, 0.0, skip=("c", i))
        pts.append(p[keep]); nrm.append(n[keep])
        A, B = lattice(cx - r, cx + r, cy - r, cy + r)  # top cap
        m = (A - cx) ** 2 + (B - cy) ** 2 <= r * r
        p = np.stack([A[m], B[m], np.full(int(m.sum()), z1)], 1)
        keep = ~_inside_any(p, boxes, cyls, 0.0, skip=("c", i))
        pts.append(p[keep]); nrm.append(np.tile(np.array([0.0, 0.0, 1.0]), (int(keep.sum()), 1)))
    pts = np.concatenate(pts).astype(np.float32)
    nrm = np.concatenate(nrm)
    # no duplicate coordinates (setNormals keys a std::map by the exact coordinates)
    _, first = np.unique(pts, axis=0, return_index=True)
    first.sort()
    pts, nrm = pts[first], nrm[first]
    if len(pts) > n_points:
        sel = np.sort(rng.choice(len(pts), size=n_points, replace=False))
        pts, nrm = pts[sel], nrm[sel]
    return np.ascontiguousarray(pts), np.ascontiguousarray(nrm), (boxes, cyls)


def plant_viewpoints(points, normals, prims, n_view, dist_vp, seed=0):
    """Candidate viewpoints as the reference generates them (viewpoint_manager.cpp:149-169):
    vp = float(pt + dist_vp * n), looking back along -n; a seeded subset of the points whose
    candidate is not inside a primitive.  Returns float32 [n_view, 6] (x, y, z, nx, ny, nz)."""
    rng = np.random.default_rng(seed + 1)
    boxes, cyls = prims
    order = rng.permutation(len(points))
    out = []
    got = 0
    for start in range(0, len(order), 1 << 16):
        idx = order[start:start + (1 << 16)]
        vp = (points[idx].astype(np.float64) + dist_vp * normals[idx]).astype(np.float32)
        ok = ~_inside_any(vp.astype(np.float64), boxes, cyls, -0.5) & (vp[:, 2] > 0.5)
        rows = np.concatenate([vp[ok], (-normals[idx][ok]).astype(np.float32)], axis=1)
        out.append(rows)
        got += len(rows)
        if got >= n_view:
            break
    out = np.concatenate(out)[:n_view]
    return np.ascontiguousarray(out, dtype=np.float32)


def poses_from_viewpoints(pos_dir):
    """(x, y, z, pitch, yaw) doubles for float (pos, dir) rows: PitchYaw of the normalised
    direction (viewpoint_manager.h:238-248) — for feeding evaluator passes in tests and bench."""
    pos_dir = np.asarray(pos_dir, dtype=np.float32)
    d = pos_dir[:, 3:6].astype(np.float64)
    nrm = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
    d = d / nrm[:, None]
    n2 = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
    pitch = np.arcsin(d[:, 2] / (n2 + 1e-3))
    yaw = np.arctan2(d[:, 1], d[:, 0])
    return np.ascontiguousarray(np.concatenate([pos_dir[:, :3].astype(np.float64), pitch[:, None], yaw[:, None]], axis=1))
