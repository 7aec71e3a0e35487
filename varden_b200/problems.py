"""
varden_b200/problems.py -- host-side synthetic problem set-up (pure numpy; no oracle, no GPU).

Geometry of one single-level VARDEN run (domain, physical BCs, box list as boxarray_maxsize would chop it:
initialize.f90:198-215) and the synthetic initial data of SURVEY 8(d): the Rayleigh-Taylor-type density
profile of src/initdata.f90:195-200,261-274 plus a deterministic, not discretely divergence-free velocity.
Used by bench.py, the tests and (through oracle/oracle.py) the CPU oracle, so every arm sees identical inputs.

Array convention: one numpy array per box, shape (n0, n1, n2, ncomp), order='F', covering lo-ng .. hi+ng
(+1 in the face direction); 2-D uses n2 == 1.  This is exactly the memory `dataptr(mf,i)` points at in the reference.
"""
import numpy as np

PERIODIC, INTERIOR, INLET, OUTLET, SYMMETRY, SLIP_WALL, NO_SLIP_WALL = -1, 0, 11, 12, 13, 14, 15


def chop_domain(dlo, dhi, dim, max_grid_size):
    """boxarray_maxsize: chop each direction into the fewest equal-ish pieces <= max_grid_size."""
    cuts = []
    for d in range(3):
        if d >= dim:
            cuts.append([(0, 0)])
            continue
        n = dhi[d] - dlo[d] + 1
        npieces = (n + max_grid_size - 1) // max_grid_size
        base, rem = divmod(n, npieces)
        segs, s = [], dlo[d]
        for p in range(npieces):
            ln = base + (1 if p < rem else 0)
            segs.append((s, s + ln - 1))
            s += ln
        cuts.append(segs)
    boxes = []
    for kz in cuts[2]:
        for jy in cuts[1]:
            for ix in cuts[0]:
                boxes.append(([ix[0], jy[0], kz[0]], [ix[1], jy[1], kz[1]]))
    return boxes


class Geom:
    """One AMR level: domain, physical BCs, dx and the box list."""

    def __init__(self, dim, n_cell, phys_bc, prob_lo=(0., 0., 0.), prob_hi=(1., 1., 1.), max_grid_size=256, boxes=None):
        self.dim = dim
        self.n_cell = [int(n_cell[d]) if d < dim else 1 for d in range(3)]
        self.dlo = [0, 0, 0]
        self.dhi = [self.n_cell[d] - 1 if d < dim else 0 for d in range(3)]
        self.phys_bc = np.zeros((3, 2), dtype=np.int32)
        self.phys_bc[:dim, :] = np.asarray(phys_bc, dtype=np.int32).reshape(-1, 2)[:dim]
        self.dx = [(prob_hi[d] - prob_lo[d]) / self.n_cell[d] if d < dim else 0.0 for d in range(3)]
        self.prob_lo = list(prob_lo)
        if boxes is None:
            boxes = chop_domain(self.dlo, self.dhi, dim, max_grid_size)
        self.boxes = boxes
        self._blo = np.ascontiguousarray([b[0] for b in boxes], dtype=np.int32)
        self._bhi = np.ascontiguousarray([b[1] for b in boxes], dtype=np.int32)

    @property
    def nboxes(self):
        return len(self.boxes)

    @property
    def ncells(self):
        return int(np.prod([self.n_cell[d] for d in range(self.dim)]))

    def box_shape(self, ib, ng, face_dir=-1):
        lo, hi = self.boxes[ib]
        return tuple((hi[d] - lo[d] + 1 + 2 * ng + (1 if d == face_dir else 0)) if d < self.dim else 1 for d in range(3))

    def subset(self, box_ids):
        """the same level restricted to some of its boxes (one rank's share)"""
        g = Geom.__new__(Geom)
        g.__dict__.update(self.__dict__)
        g.boxes = [self.boxes[i] for i in box_ids]
        g._blo = np.ascontiguousarray([b[0] for b in g.boxes], dtype=np.int32)
        g._bhi = np.ascontiguousarray([b[1] for b in g.boxes], dtype=np.int32)
        return g


def mf_alloc(geom, ng, ncomp, face_dir=-1, val=0.0):
    return [np.full(geom.box_shape(ib, ng, face_dir) + (ncomp,), val, dtype=np.float64, order='F') for ib in range(geom.nboxes)]


def valid(geom, a, ib, ng, face_dir=-1):
    """view of the valid region of one box array"""
    sl = []
    for d in range(3):
        sl.append(slice(ng, a.shape[d] - ng) if d < geom.dim else slice(None))
    return a[tuple(sl)]


def _h(x):
    return 0.02 * np.sin(4.0 * np.pi * x) + 0.01 * np.sin(8.0 * np.pi * x)


def rt_problem(n, dim=3, max_grid_size=256, ratio=2.0, grav=-9.8, nscal=2, seeded_velocity=True, phys_bc=None,
               prob_hi=(1., 1., 1.), box_ids=None):
    """
    Density-stratified Rayleigh-Taylor-type state: periodic in x(,y), no-slip walls in the last direction
    (exec/test/inputs_RayleighTaylor_3d:32-37).  rho = mid + amp*tanh((z - 1/2 - h(x) - h(y))/0.01); ratio 2 reproduces
    initdata.f90:270 exactly (1.5 + 0.5 tanh).  VALID cells only are initialised; ghost cells are the caller's job
    (varden.f90:291-300: fill_boundary + multifab_physbc).  Returns (geom, state dict, dt); with box_ids the returned
    geom holds only those boxes (same domain).
    """
    if np.isscalar(n):
        n = [n] * dim
    if phys_bc is None:
        phys_bc = [[PERIODIC, PERIODIC]] * (dim - 1) + [[NO_SLIP_WALL, NO_SLIP_WALL]]
    geom = Geom(dim, n, phys_bc, prob_hi=prob_hi, max_grid_size=max_grid_size)
    if box_ids is not None:          # one rank's share: only these boxes are materialised
        geom = geom.subset(box_ids)
    mid, amp = 0.5 * (float(ratio) + 1.0), 0.5 * (float(ratio) - 1.0)
    st = dict(uold=mf_alloc(geom, 3, dim), sold=mf_alloc(geom, 3, nscal), gp=mf_alloc(geom, 1, dim),
              ext_vel_force=mf_alloc(geom, 1, dim), ext_scal_force=mf_alloc(geom, 1, nscal))
    for ib, (lo, hi) in enumerate(geom.boxes):
        ax = [(np.arange(lo[d], hi[d] + 1) + 0.5) * geom.dx[d] if d < dim else np.zeros(1) for d in range(3)]
        X, Y, Z = np.meshgrid(ax[0], ax[1], ax[2], indexing='ij', sparse=True)
        if dim == 3:
            rho = mid + amp * np.tanh((Z - 0.5 - _h(X) - _h(Y)) / 0.01)
        else:
            rho = mid + amp * np.tanh((Y - 0.5 - _h(X)) / 0.01) + 0.0 * Z
        valid(geom, st["sold"][ib], ib, 3)[..., 0] = rho
        valid(geom, st["sold"][ib], ib, 3)[..., 1] = 0.0
        if seeded_velocity:
            u = valid(geom, st["uold"][ib], ib, 3)
            if dim == 3:
                u[..., 0] = 0.1 * np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y) * np.sin(np.pi * Z)
                u[..., 1] = -0.1 * np.cos(2 * np.pi * X) * np.sin(2 * np.pi * Y) * np.sin(np.pi * Z)
                u[..., 2] = 0.05 * np.sin(2 * np.pi * X) * np.sin(2 * np.pi * Y) * np.sin(2 * np.pi * Z)
            else:
                u[..., 0] = 0.1 * np.sin(2 * np.pi * X) * np.sin(np.pi * Y) + 0.0 * Z
                u[..., 1] = 0.05 * np.sin(2 * np.pi * X) * np.sin(2 * np.pi * Y) + 0.0 * Z
        st["ext_vel_force"][ib][..., dim - 1] = grav      # varden.f90:428-429
    dt = 0.45 * geom.dx[0] / 0.1                           # fixed_dt: CFL ~ 0.45 on |u| = 0.1
    return geom, st, dt


def bubble_problem(n=64, max_grid_size=32, grav=-9.8, nscal=2, densfact=2.0, cflfac=0.9, init_shrink=0.1, box_ids=None):
    """
    BASELINE config 1: the reference's own CPU-runnable case exec/test/inputs_2d-regt with max_levs = 1 -- 2-D 64^2, max_grid_size 32
    (4 boxes), prob_type = 1 (initdata.f90:136-160): u = 0, rho = 1 + (densfact-1)/2 (1 - tanh(30 (r - 0.1))) around (0.5, 0.5),
    tracer = rho; no-slip walls on every side (bcx/bcy = 15); grav = -9.8.  The deck is viscous (visc_coef = 0.001): visc_solve stays the
    reference's and the path is compared inviscid.  dt as the first step takes it: init_shrink * cflfac * the forcing bound
    sqrt(2 dx / |F|) of estdt.f90:165-172 (u = 0, gp = 0).  VALID cells only; ghost cells are the caller's job (varden.f90:291-300).
    """
    dim = 2
    phys_bc = [[NO_SLIP_WALL, NO_SLIP_WALL], [NO_SLIP_WALL, NO_SLIP_WALL]]
    geom = Geom(dim, [n, n], phys_bc, prob_hi=(1., 1., 1.), max_grid_size=max_grid_size)
    if box_ids is not None:
        geom = geom.subset(box_ids)
    st = dict(uold=mf_alloc(geom, 3, dim), sold=mf_alloc(geom, 3, nscal), gp=mf_alloc(geom, 1, dim),
              ext_vel_force=mf_alloc(geom, 1, dim), ext_scal_force=mf_alloc(geom, 1, nscal))
    for ib, (lo, hi) in enumerate(geom.boxes):
        ax = [(np.arange(lo[d], hi[d] + 1) + 0.5) * geom.dx[d] if d < dim else np.zeros(1) for d in range(3)]
        X, Y, Z = np.meshgrid(ax[0], ax[1], ax[2], indexing='ij', sparse=True)
        dist = np.sqrt((X - 0.5) ** 2 + (Y - 0.5) ** 2) + 0.0 * Z
        rho = 1.0 + 0.5 * (densfact - 1.0) * (1.0 - np.tanh(30.0 * (dist - 0.1)))
        valid(geom, st["sold"][ib], ib, 3)[..., 0] = rho
        valid(geom, st["sold"][ib], ib, 3)[..., 1] = rho
        st["ext_vel_force"][ib][..., dim - 1] = grav
    dt = init_shrink * cflfac * np.sqrt(2.0 * geom.dx[1] / abs(grav))
    return geom, st, dt
