"""CPU tests of the derived marching-cubes case table (nl_mc_case_table, csrc/mc.cu): no GPU needed."""
import ctypes as C
import itertools

import numpy as np

from util import ROOT  # noqa: F401


def _table():
    import nerfloam_b200 as nl
    tri = np.zeros((256, 16), np.uint8); nt = np.zeros(256, np.uint8); ec = np.zeros((12, 2), np.uint8)
    assert nl._capi.lib().nl_mc_case_table(tri.ctypes.data_as(C.c_void_p), nt.ctypes.data_as(C.c_void_p), ec.ctypes.data_as(C.c_void_p)) == 0
    return tri, nt, ec


def test_every_case_uses_exactly_its_crossed_edges():
    tri, nt, ec = _table()
    assert nt.max() == 5 and nt[0] == 0 and nt[255] == 0
    for cfg in range(256):
        crossed = {e for e in range(12) if ((cfg >> ec[e, 0]) & 1) != ((cfg >> ec[e, 1]) & 1)}
        used = set(tri[cfg, :3 * nt[cfg]].tolist())
        assert used == crossed, cfg
        assert all(x == 255 for x in tri[cfg, 3 * nt[cfg]:])
        for q in range(nt[cfg]):
            assert len(set(tri[cfg, 3 * q:3 * q + 3].tolist())) == 3


def test_against_the_python_restatement_case_by_case():
    """The C++ table and oracle/mc.py (independent Python code for the same construction) give the same triangles."""
    from oracle import mc as OM
    tri, nt, ec = _table()

    def key(e):   # table edge id -> oracle edge key (axis, base corner)
        c0 = int(ec[e, 0])
        return (e >> 2, c0 & 1, (c0 >> 1) & 1, (c0 >> 2) & 1)

    def canon(tris):   # triangles up to rotation of their three vertices
        return sorted(min(t[i:] + t[:i] for i in range(3)) for t in tris)
    for cfg in range(256):
        neg8 = {(x, y, z): bool((cfg >> (x | (y << 1) | (z << 2))) & 1) for x, y, z in itertools.product((0, 1), repeat=3)}
        want = canon([tuple(t) for t in OM.cell_polygons(neg8)])
        got = [tuple(key(int(e)) for e in tri[cfg, 3 * q:3 * q + 3]) for q in range(nt[cfg])]
        # the fan may start at a different vertex of a polygon: compare the polygons' triangle fans as edge sets + orientation via area
        assert len(got) == len(want), cfg
        assert sorted(sorted(t) for t in got) == sorted(sorted(t) for t in want) or _same_surface(got, want), cfg


def _same_surface(a, b):
    """Two triangulations of the same oriented polygons have the same boundary-edge multiset (interior diagonals cancel)."""
    def boundary(tris):
        cnt = {}
        for t in tris:
            for i in range(3):
                e = (t[i], t[(i + 1) % 3])
                if (e[1], e[0]) in cnt:
                    cnt[(e[1], e[0])] -= 1
                    if cnt[(e[1], e[0])] == 0:
                        del cnt[(e[1], e[0])]
                else:
                    cnt[e] = cnt.get(e, 0) + 1
        return sorted(cnt.items())
    return boundary(a) == boundary(b)


def test_random_fields_give_closed_oriented_manifolds():
    """A random sign field on a lattice whose border is all positive: the union of all cells' triangles must be a closed,
    consistently oriented surface -- every directed edge is matched by exactly one opposite one, also across cell faces where the
    classic hand-made table can leave holes -- and the normals point towards positive values."""
    tri, nt, ec = _table()
    rng = np.random.default_rng(0)
    for trial in range(30):
        n = 6
        s = rng.normal(size=(n, n, n)).astype(np.float32)
        s[0], s[-1], s[:, 0], s[:, -1], s[:, :, 0], s[:, :, -1] = 1, 1, 1, 1, 1, 1
        neg = s < 0
        directed = {}
        vol6 = 0.0
        for ci, cj, ck in itertools.product(range(n - 1), repeat=3):
            cfg = 0
            for c in range(8):
                cfg |= int(neg[ci + (c & 1), cj + ((c >> 1) & 1), ck + ((c >> 2) & 1)]) << c
            for q in range(nt[cfg]):
                vs, ps = [], []
                for e in tri[cfg, 3 * q:3 * q + 3]:
                    c0, c1 = int(ec[e, 0]), int(ec[e, 1])
                    p0 = np.array([ci + (c0 & 1), cj + ((c0 >> 1) & 1), ck + ((c0 >> 2) & 1)], float)
                    p1 = np.array([ci + (c1 & 1), cj + ((c1 >> 1) & 1), ck + ((c1 >> 2) & 1)], float)
                    vs.append((int(e) >> 2,) + tuple(int(x) for x in p0))
                    s0, s1 = float(s[tuple(p0.astype(int))]), float(s[tuple(p1.astype(int))])
                    ps.append(p0 + (p1 - p0) * (s0 / (s0 - s1)))
                for i in range(3):
                    d = (vs[i], vs[(i + 1) % 3])
                    directed[d] = directed.get(d, 0) + 1
                vol6 += np.dot(ps[0], np.cross(ps[1], ps[2]))
        assert all(c == 1 for c in directed.values())
        assert all((b, a) in directed for (a, b) in directed)            # closed + consistently oriented
        # outward normals (towards positive SDF): the enclosed (negative) volume comes out positive
        assert vol6 > 0
