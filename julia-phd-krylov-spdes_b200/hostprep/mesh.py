"""Synthetic inputs for the oracle and the benchmarks (TEST INFRASTRUCTURE).

The reference builds its meshes with Triangle (Fem/Mesh.jl:21-31) and its
partitions with the ``mpmetis`` executable (Fem/Mesh.jl:169-195); neither is
available here, so SURVEY.md section 8(d) prescribes seeded synthetic stand-ins:
a structured (optionally jittered) triangulation of the unit square, a
recursive-coordinate-bisection element partition and a random-Fourier-feature
lognormal coefficient.  Array layouts follow the reference's in-memory ones
(``cells`` 3 x nel, ``points`` 2 x nnode, ``cell_neighbors`` 3 x nel with -1 on
the boundary) but 0-based.  ``save_mesh``/``save_partition`` write the
reference's on-disk formats (Fem/Mesh.jl:49-56, 221-229) so a machine with
Julia can replay the same inputs through the real code.
"""
import numpy as np


def structured_mesh(nx, jitter=0.0, seed=481456):
    """Unit-square P1 mesh, nx x nx nodes, 2 triangles per cell, alternating diagonal.

    Returns cells (3, nel) int64 0-based CCW, points (2, nnode), point_markers
    (nnode,) with 1 on the boundary (Dirichlet, Fem/BoundaryConditions.jl:45),
    cell_neighbors (3, nel) 0-based with -1 on the boundary; neighbour j is the
    element across the edge opposite vertex j (Triangle's convention; the
    reference only tests membership, EllipticPdeDomainDecomposition.jl:131-146).
    """
    h = 1.0 / (nx - 1)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(nx), indexing="xy")
    x = ix.astype(np.float64) * h
    y = iy.astype(np.float64) * h
    interior = (ix > 0) & (ix < nx - 1) & (iy > 0) & (iy < nx - 1)
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        dx = rng.uniform(-jitter * h, jitter * h, size=x.shape)
        dy = rng.uniform(-jitter * h, jitter * h, size=x.shape)
        x = np.where(interior, x + dx, x)
        y = np.where(interior, y + dy, y)
    points = np.stack([x.ravel(), y.ravel()])
    markers = (~interior).ravel().astype(np.int64)

    cx, cy = np.meshgrid(np.arange(nx - 1), np.arange(nx - 1), indexing="xy")
    cx = cx.ravel()
    cy = cy.ravel()
    n00 = cy * nx + cx
    n10 = n00 + 1
    n01 = n00 + nx
    n11 = n01 + 1
    even = ((cx + cy) % 2) == 0
    # even cells: diagonal n00-n11 ; odd cells: diagonal n10-n01
    t0 = np.where(even[None, :], np.stack([n00, n10, n11]), np.stack([n00, n10, n01]))
    t1 = np.where(even[None, :], np.stack([n00, n11, n01]), np.stack([n10, n11, n01]))
    nq = cx.size
    cells = np.empty((3, 2 * nq), dtype=np.int64)
    cells[:, 0::2] = t0
    cells[:, 1::2] = t1
    neighbors = cell_neighbors_from_cells(cells, points.shape[1])
    return cells, points, markers, neighbors


def cell_neighbors_from_cells(cells, nnode):
    """Edge-adjacency of a conforming triangulation: (3, nel), -1 on the boundary."""
    nel = cells.shape[1]
    a = cells[[1, 2, 0], :]
    b = cells[[2, 0, 1], :]
    lo = np.minimum(a, b).ravel()
    hi = np.maximum(a, b).ravel()
    key = lo * np.int64(nnode) + hi
    order = np.argsort(key, kind="stable")
    ks = key[order]
    same = ks[1:] == ks[:-1]
    elem = np.tile(np.arange(nel, dtype=np.int64), 3)  # flat index j*nel+iel -> iel
    nb = np.full(3 * nel, -1, dtype=np.int64)
    i0 = order[:-1][same]
    i1 = order[1:][same]
    nb[i0] = elem[i1]
    nb[i1] = elem[i0]
    return nb.reshape(3, nel)


def rcb_partition(cells, points, ndom):
    """Recursive coordinate bisection of element centroids into ndom parts.

    Stand-in for ``mesh_partition`` (Fem/Mesh.jl:169-195, mpmetis -contig).
    Returns epart (nel,) and npart (nnode,) 0-based; npart[node] is the part of
    the lowest-numbered element containing the node.
    """
    nel = cells.shape[1]
    cen = points[:, cells].mean(axis=1)  # (2, nel)
    epart = np.zeros(nel, dtype=np.int64)

    def split(idx, first, count):
        if count == 1:
            epart[idx] = first
            return
        c = cen[:, idx]
        ext = c.max(axis=1) - c.min(axis=1)
        ax = 0 if ext[0] >= ext[1] else 1
        nl = count // 2
        k = int(round(idx.size * nl / count))
        order = np.argsort(c[ax], kind="stable")
        split(idx[order[:k]], first, nl)
        split(idx[order[k:]], first + nl, count - nl)

    split(np.arange(nel, dtype=np.int64), 0, ndom)
    nnode = points.shape[1]
    first_el = np.full(nnode, nel, dtype=np.int64)
    for j in range(3):
        np.minimum.at(first_el, cells[j], np.arange(nel, dtype=np.int64))
    npart = epart[first_el]
    return epart, npart


def lognormal_field(points, ireal, sig2=1.0, L=0.1, m=256, median=False):
    """g(x) for a = exp(g): random-Fourier-feature Gaussian field with the
    squared-exponential covariance sig2*exp(-|d|^2/L^2) of Fem/Covariances.jl:23-27
    (Example07_PcgSchurStochasticEllipticPde.jl:17-19).  ``median=True`` gives the
    xi = 0 realization g = 0 used for the constant preconditioner (Example07:89).
    """
    nnode = points.shape[1]
    if median:
        return np.zeros(nnode)
    rng = np.random.default_rng(123456 + int(ireal))
    omega = rng.normal(0.0, np.sqrt(2.0) / L, size=(m, 2))
    phi = rng.uniform(0.0, 2.0 * np.pi, size=m)
    g = np.zeros(nnode)
    step = 1 << 16
    for s in range(0, nnode, step):
        arg = omega @ points[:, s:s + step] + phi[:, None]
        g[s:s + step] = np.cos(arg).sum(axis=0)
    return np.sqrt(2.0 * sig2 / m) * g


def save_mesh(prefix, tentative_nnode, cells, points, markers, neighbors):
    """Reference wire format (Fem/Mesh.jl:49-56): bare .npy payloads named *.npz;
    cells nel x 3 0-based, points nnode x 2, point_markers nnode x 1,
    cell_neighbors nel x 3 in the reference's in-memory values (1-based, -1 boundary)."""
    def w(name, arr):
        with open(f"{prefix}/DoF{tentative_nnode}.{name}.npz", "wb") as f:
            np.save(f, arr)
    w("cells", np.ascontiguousarray(cells.T))
    w("points", np.ascontiguousarray(points.T))
    w("point_markers", np.ascontiguousarray(markers.reshape(-1, 1)))
    nb = np.where(neighbors >= 0, neighbors + 1, -1)
    w("cell_neighbors", np.ascontiguousarray(nb.T))


def load_mesh(prefix, tentative_nnode):
    """Inverse of save_mesh (Fem/Mesh.jl:84-91), back to this package's 0-based layout."""
    def r(name):
        with open(f"{prefix}/DoF{tentative_nnode}.{name}.npz", "rb") as f:
            return np.load(f)
    cells = np.ascontiguousarray(r("cells").T).astype(np.int64)
    points = np.ascontiguousarray(r("points").T)
    markers = r("point_markers").ravel().astype(np.int64)
    nb = np.ascontiguousarray(r("cell_neighbors").T).astype(np.int64)
    nb = np.where(nb >= 1, nb - 1, -1)
    return cells, points, markers, nb


def save_partition(prefix, tentative_nnode, ndom, epart, npart):
    """Reference wire format (Fem/Mesh.jl:221-229): 0-based epart / npart."""
    for name, arr in (("epart", epart), ("npart", npart)):
        with open(f"{prefix}/DoF{tentative_nnode}-ndom{ndom}.{name}.npz", "wb") as f:
            np.save(f, np.asarray(arr, dtype=np.int64))


def load_partition(prefix, tentative_nnode, ndom):
    out = []
    for name in ("epart", "npart"):
        with open(f"{prefix}/DoF{tentative_nnode}-ndom{ndom}.{name}.npz", "rb") as f:
            out.append(np.load(f).astype(np.int64))
    return out[0], out[1]
