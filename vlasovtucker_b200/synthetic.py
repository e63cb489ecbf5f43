"""Synthetic meshes for the large configurations.

A periodic box of nx*ny*nz hexahedra, each split into 6 Kuhn (Freudenthal) tetrahedra.  The
split is translation invariant, so opposite box faces triangulate identically and periodic
partners are exact.  The tables follow the reference's index contract (src/mesh.cpp:129-181):
``faces[4t+j]`` uses vertex order j=0:{p1,p2,p3}, 1:{p0,p3,p2}, 2:{p0,p1,p3}, 3:{p0,p2,p1}, tets
are negatively oriented (mesh.cpp:152), geometry uses the formulas of src/primitives.cpp:94-139.
Tests check these tables against the oracle's restatement of ``Mesh::Reconstruct`` fed with the
same nodes/tets/boundary triangles.
"""
import itertools

import numpy as np

from .context import MeshTables

_FACE_VERTS = np.array([[1, 2, 3], [0, 3, 2], [0, 1, 3], [0, 2, 1]])   # mesh.cpp:157-160


def kuhn_box(nx, ny, nz, lengths=(1.0, 1.0, 1.0)):
    """Nodes, tets (reference orientation) and boundary triangles of the Kuhn box.

    Returns (nodes[(nx+1)(ny+1)(nz+1),3], tets[6*nx*ny*nz,4], tris[nb,3], tri_entity[nb]) with
    Gmsh OCC box entity numbering 1: x=min, 2: x=max, 3: y=min, 4: y=max, 5: z=min, 6: z=max.
    """
    dims = np.array([nx, ny, nz])
    L = np.asarray(lengths, float)
    gx, gy, gz = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")

    def nid(i, j, k):
        return (i * (ny + 1) + j) * (nz + 1) + k

    nodes = np.stack([gx.ravel() * (L[0] / nx), gy.ravel() * (L[1] / ny), gz.ravel() * (L[2] / nz)], 1)
    # hexes, x fastest (so that consecutive tets are spatial neighbours)
    hk, hj, hi = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    hi, hj, hk = hi.ravel(), hj.ravel(), hk.ravel()
    tets = np.empty((len(hi), 6, 4), np.int64)
    eye = np.eye(3, dtype=np.int64)
    for q, perm in enumerate(itertools.permutations(range(3))):
        c = np.stack([hi, hj, hk], 1)
        v = [c]
        for ax in perm:
            v.append(v[-1] + eye[ax])
        ids = [nid(p[:, 0], p[:, 1], p[:, 2]) for p in v]
        # signed volume of (e_p0, e_p0+e_p1, e_p0+e_p1+e_p2) is the parity of perm; the reference
        # asserts Orientation() <= 0, i.e. (P1-P0).((P2-P0)x(P3-P0)) >= 0
        parity = np.linalg.det(eye[list(perm)].astype(float))
        if parity < 0:
            ids[2], ids[3] = ids[3], ids[2]
        tets[:, q, :] = np.stack(ids, 1)
    tets = tets.reshape(-1, 4)

    # boundary triangles: faces of tets lying on the box planes, owning tet's outward order
    fv = tets[:, _FACE_VERTS]                       # (nT,4,3)
    coords = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], 1)
    fc = coords[fv]                                 # (nT,4,3 verts,3)
    tris, ents = [], []
    for ax in range(3):
        for side, val in ((0, 0), (1, dims[ax])):
            on = np.all(fc[..., ax] == val, axis=2)
            tris.append(fv[on])
            ents.append(np.full(on.sum(), 2 * ax + side + 1))
    return nodes, tets.astype(np.int32), np.concatenate(tris).astype(np.int32), np.concatenate(ents).astype(np.int32)


def mesh_tables(nodes, tets, tris=None, tri_entity=None, periodic_wrap=None):
    """Flatten a tet mesh into the tables of ``Context.mesh_upload`` (numpy, vectorised).

    ``periodic_wrap``: optional int array mapping node id -> representative node id; faces that
    coincide after the mapping are periodic partners (exact for translation-invariant meshes).
    """
    nodes = np.asarray(nodes, float)
    tets = np.asarray(tets, np.int64)
    nT = len(tets)
    P = nodes[tets]                                              # (nT,4,3)
    centroid = (((P[:, 0] + P[:, 1]) + P[:, 2]) + P[:, 3]) / 4.0  # primitives.cpp:125
    a, b, c = P[:, 1] - P[:, 0], P[:, 2] - P[:, 0], P[:, 3] - P[:, 0]
    orientation = -np.einsum("ij,ij->i", a, np.cross(b, c))      # primitives.cpp:130-139
    if np.any(orientation > 0):
        raise ValueError("positively oriented tetrahedra (mesh.cpp:152 asserts Orientation() <= 0)")
    volume = np.abs(orientation) / 6.0
    fv = tets[:, _FACE_VERTS]                                    # (nT,4,3)
    F = nodes[fv]                                                # (nT,4,3,3)
    fcen = ((F[:, :, 0] + F[:, :, 1]) + F[:, :, 2]) / 3.0         # primitives.cpp:97
    u, w = F[:, :, 1] - F[:, :, 0], F[:, :, 2] - F[:, :, 0]
    nrm = np.cross(u, w)
    ln = np.sqrt(nrm[..., 0] * nrm[..., 0] + nrm[..., 1] * nrm[..., 1] + nrm[..., 2] * nrm[..., 2])
    normal = nrm / ln[..., None]
    area = ln / 2.0

    # adjacency: faces with the same vertex set (after periodic identification) face each other
    key_nodes = fv if periodic_wrap is None else np.asarray(periodic_wrap)[fv]
    ks = np.sort(key_nodes.reshape(-1, 3), axis=1)
    nN = int(ks.max()) + 1
    key = (ks[:, 0] * nN + ks[:, 1]) * nN + ks[:, 2]
    o = np.argsort(key, kind="stable")
    sk = key[o]
    same = sk[1:] == sk[:-1]
    nbr = np.full(4 * nT, -1, np.int64)
    i0 = o[:-1][same]
    i1 = o[1:][same]
    nbr[i0] = i1 // 4
    nbr[i1] = i0 // 4
    nbr = nbr.reshape(nT, 4)

    entity = np.full((nT, 4), -1, np.int32)
    if tris is not None and len(tris):
        tk = np.sort(np.asarray(tris, np.int64), axis=1)
        tkey = (tk[:, 0] * nN + tk[:, 1]) * nN + tk[:, 2]
        fk = np.sort(fv.reshape(-1, 3), axis=1)
        fkey = (fk[:, 0] * nN + fk[:, 1]) * nN + fk[:, 2]
        fo = np.argsort(fkey, kind="stable")
        pos = np.searchsorted(fkey[fo], tkey)
        ok = (pos < len(fkey)) & (fkey[fo][np.minimum(pos, len(fkey) - 1)] == tkey)
        if not ok.all():
            raise ValueError("boundary triangle matches no tet face")
        # a boundary triangle is owned by exactly one tet face in the unwrapped mesh
        entity.reshape(-1)[fo[pos]] = np.asarray(tri_entity, np.int32)
    return MeshTables(nbr=nbr.astype(np.int32), area=area, volume=volume, normal=normal, entity=entity,
                      tetCentroid=centroid, faceCentroid=fcen)


def brick_order(nx, ny, nz, brick):
    """Locality permutation for the Kuhn box: hexes grouped in brick[0]*brick[1]*brick[2] bricks,
    the 6 tets of a hex kept together.  Returns (order, tets_per_brick)."""
    bx, by, bz = brick
    if nx % bx or ny % by or nz % bz:
        raise ValueError("brick dims must divide the box dims")
    hk, hj, hi = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    hi, hj, hk = hi.ravel(), hj.ravel(), hk.ravel()
    keys = np.stack([hi % bx, hj % by, hk % bz, hi // bx, hj // by, hk // bz], 0)   # last key is primary
    ho = np.lexsort(keys)
    order = (ho[:, None] * 6 + np.arange(6)[None, :]).reshape(-1)
    return order.astype(np.int32), 6 * bx * by * bz


def pencil_order(nx, ny, nz, a, depth=1):
    """Locality permutation for the Kuhn box that sweeps a x a-hex pencils along z, back and forth,
    the pencils themselves in serpentine order: the rows in flight always border the rows just
    processed, so only the four side faces of a pencil miss the cache (scripts/prototypes/
    l2_order_model.py).  Returns (order, tets_per_brick) with bricks of a*a*depth hexes."""
    if nx % a or ny % a or nz % depth:
        raise ValueError("pencil dims must divide the box dims")
    hk, hj, hi = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    hi, hj, hk = hi.ravel(), hj.ravel(), hk.ravel()
    pi, pj = hi // a, hj // a
    npj = ny // a
    pen = pi * npj + np.where(pi % 2 == 0, pj, npj - 1 - pj)
    kk = np.where(pen % 2 == 0, hk, nz - 1 - hk)
    ho = np.lexsort(np.stack([hi % a, hj % a, kk, pen], 0))                          # last key is primary
    order = (ho[:, None] * 6 + np.arange(6)[None, :]).reshape(-1)
    return order.astype(np.int32), 6 * a * a * depth


def periodic_kuhn_tables(nx, ny, nz, lengths=(1.0, 1.0, 1.0), brick=None):
    """Fully periodic Kuhn box as MeshTables (entities 1-6 are periodic pairs {1,2},{3,4},{5,6})."""
    nodes, tets, tris, ents = kuhn_box(nx, ny, nz, lengths)
    gx, gy, gz = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    wrap = (((gx % nx) * (ny + 1) + (gy % ny)) * (nz + 1) + (gz % nz)).ravel()
    mt = mesh_tables(nodes, tets, tris, ents, periodic_wrap=wrap)
    mt.periodic = [(1, 2), (3, 4), (5, 6)]
    if brick is not None:
        mt.order, mt.brickTets = brick_order(nx, ny, nz, brick)
    return mt


def write_msh(path, nodes, tets, tris, tri_entity):
    """MSH 2.2 ASCII file as SURVEY.md §8d specifies the C4 input: node tags 1..n in order, boundary
    triangles first (the owning tet's outward vertex order, elementary tags = entities), then all
    tetrahedra in one volume entity.  This is the file the reference's ``Mesh(std::string)`` reads."""
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % len(nodes))
        np.savetxt(f, np.column_stack([np.arange(1, len(nodes) + 1), nodes]), fmt="%d %.17g %.17g %.17g")
        f.write("$EndNodes\n$Elements\n%d\n" % (len(tris) + len(tets)))
        nt = len(tris)
        ids = np.arange(1, nt + 1)
        np.savetxt(f, np.column_stack([ids, np.full(nt, 2), np.full(nt, 2), tri_entity, tri_entity, tris + 1]), fmt="%d")
        ids = np.arange(nt + 1, nt + len(tets) + 1)
        one = np.ones(len(tets), np.int64)
        np.savetxt(f, np.column_stack([ids, 4 * one, 2 * one, one, one, tets + 1]), fmt="%d")
        f.write("$EndElements\n")
