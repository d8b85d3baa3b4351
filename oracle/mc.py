"""Marching-cubes oracle (TEST INFRASTRUCTURE ONLY): a plain numpy / Python restatement used to check csrc/mc.cu.

The reference meshes with skimage.measure.marching_cubes, once per voxel on the voxel's res^3 lattice
(src/utils/mesh_util.py:145-169).  scikit-image is a third-party dependency that is neither vendored in /root/reference nor
installed here and the reference pins no version (requirements.txt: "scikit-image" without one), so the oracle restates the
PUBLISHED algorithm (Lorensen & Cline 1987, as every implementation including skimage's realises it):
  * one vertex on every lattice edge whose end points have SDF values of different sign, at the linearly interpolated zero
    crossing  (this vertex set is implementation independent -- it is what both skimage variants and this kernel produce);
  * inside each cell, the crossings are joined into closed polygons along the cell's faces and triangulated.
The polygon construction here is written independently of the CUDA code (Python sets / dicts, per cell, no table), so the GPU
tests compare two implementations of the same rule set; the rule-independent properties (closed, consistently oriented 2-manifold,
vertices on the crossing edges, area of analytic shapes) are tested separately in tests/test_mc_cpu.py / test_gpu_mesh.py.
Vertex placement follows mesh_util.py:149-161: spacing 1/(res-1), then (v - 0.5) * voxel_size + voxel centre.
"""
import itertools

import numpy as np

FACES = []
for axis in range(3):
    for side in range(2):
        a, b = (axis + 1) % 3, (axis + 2) % 3
        cyc = []
        for u, v in ((0, 0), (1, 0), (1, 1), (0, 1)):
            p = [0, 0, 0]
            p[axis], p[a], p[b] = side, u, v
            cyc.append(tuple(p))
        FACES.append(cyc)


def crossing_vertices(sdf, centre, voxel_size):
    """dict {(axis, i, j, k): xyz float32} for every crossed lattice edge of one voxel (edge from (i,j,k) along +axis)."""
    res = sdf.shape[0]
    inv = np.float32(1.0) / np.float32(res - 1)
    out = {}
    neg = sdf < 0
    for axis in range(3):
        for i, j, k in itertools.product(range(res), repeat=3):
            q = [i, j, k]
            if q[axis] >= res - 1:
                continue
            q2 = list(q); q2[axis] += 1
            if neg[i, j, k] != neg[tuple(q2)]:
                s0, s1 = np.float32(sdf[i, j, k]), np.float32(sdf[tuple(q2)])
                t = s0 / (s0 - s1)
                p = np.array(q, np.float32)
                p[axis] = p[axis] + t
                out[(axis, i, j, k)] = (p * inv - np.float32(0.5)) * np.float32(voxel_size) + np.asarray(centre, np.float32)
    return out


def cell_polygons(neg8):
    """neg8[(x,y,z)] -> bool for the 8 corners of a cell.  Returns oriented triangles as triples of cell-local edge keys
    (axis, x, y, z): the zero-level segments of each face are joined into loops; on a face with two inside corners on a diagonal
    each inside corner is cut off on its own; loops are fan-triangulated from their first edge and oriented towards the outside."""
    def edge_key(pa, pb):
        axis = [i for i in range(3) if pa[i] != pb[i]][0]
        base = pa if pa[axis] == 0 else pb
        return (axis,) + tuple(base)
    links = {}
    def link(e1, e2):
        links.setdefault(e1, []).append(e2)
        links.setdefault(e2, []).append(e1)
    for cyc in FACES:
        crossed = [i for i in range(4) if neg8[cyc[i]] != neg8[cyc[(i + 1) % 4]]]
        E = lambda i: edge_key(cyc[i], cyc[(i + 1) % 4])
        if len(crossed) == 2:
            link(E(crossed[0]), E(crossed[1]))
        elif len(crossed) == 4:
            for i in range(4):
                if neg8[cyc[i]]:
                    link(E((i + 3) % 4), E(i))
    tris, seen = [], set()
    for e0 in sorted(links):
        if e0 in seen:
            continue
        loop, prev, cur = [], None, e0
        while cur not in seen:
            seen.add(cur)
            loop.append(cur)
            nxt = [x for x in links[cur] if x != prev]
            nxt = nxt[0] if nxt else links[cur][0]
            prev, cur = cur, nxt
        mids, direction = [], np.zeros(3)
        for (axis, x, y, z) in loop:
            p0 = np.array([x, y, z], float); p1 = p0.copy(); p1[axis] += 1
            mids.append(0.5 * (p0 + p1))
            direction += (p1 - p0) if neg8[(x, y, z)] else (p0 - p1)
        n = np.zeros(3)
        for i in range(len(loop)):
            a, b = mids[i], mids[(i + 1) % len(loop)]
            n += np.array([(a[1] - b[1]) * (a[2] + b[2]), (a[2] - b[2]) * (a[0] + b[0]), (a[0] - b[0]) * (a[1] + b[1])])
        if n @ direction < 0:
            loop = loop[::-1]
        for i in range(1, len(loop) - 1):
            tris.append((loop[0], loop[i], loop[i + 1]))
    return tris


def marching_cubes_voxel(sdf, centre, voxel_size):
    """One voxel: (verts dict as crossing_vertices, list of triangles as triples of global edge keys)."""
    res = sdf.shape[0]
    verts = crossing_vertices(sdf, centre, voxel_size)
    neg = sdf < 0
    tris = []
    for ci, cj, ck in itertools.product(range(res - 1), repeat=3):
        neg8 = {(x, y, z): bool(neg[ci + x, cj + y, ck + z]) for x, y, z in itertools.product((0, 1), repeat=3)}
        for t in cell_polygons(neg8):
            tris.append(tuple((a, ci + x, cj + y, ck + z) for (a, x, y, z) in t))
    return verts, tris
