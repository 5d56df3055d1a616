"""Problem data for the five in-scope test cases: point sets, targets, normalisation constants.

Host-side (numpy) restatement of the "Data Creation" sections of the reference scripts:

* grid + one permutation split into PDE / Vel / Pres / Test      cavity_steady.py:83-96
* exact / reference fields and spread normalisation               cavity_steady.py:100-120,
                                                                  colliding_flow.py:71-73,104-118,
                                                                  poiseuille_flow.py:113-132
* per-edge boundary sampling and (noisy) Dirichlet targets         cavity_steady.py:124-147
* noisy fit targets                                                cavity_steady.py:150-153
* Poisson point sets                                               poisson_misto.py:47-60

Differences forced by the environment, all explicit:

* random numbers come from ``numpy.random.default_rng(seed)`` -- TensorFlow's Philox streams are
  not reproducible without TensorFlow;
* the FEM fields of DataGeneration/ (``navier-stokes_cavity_steady.h5`` ...) are not in the
  reference repository, so ``synthetic_cavity_field`` supplies arrays OF THE SAME SHAPE (values at
  the (n1+1)(n2+1) grid vertices, x fastest, pressure mean-subtracted, cavity_steady.py:100-109);
  a real file can be passed instead (``fem_fields=``) and is read with ``h5lite``;
* ``n_pde`` larger than the grid (the BASELINE sizes: 1e4 .. 4e6) cannot be drawn from the
  101x101 grid, so those collocation sets are uniform random in the domain; grid-sized requests
  follow the reference's permutation split exactly.

Every array is rounded to float32-representable values (the device path is FP32) so the product and
the float64 oracle see bit-identical inputs.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from .options import SimulationOptions

EDGES = ("BOT", "DX", "TOP", "SX")


def _f32(a) -> np.ndarray:
    return np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)


@dataclass
class ProblemData:
    name: str
    dim: int
    hidden: List[int]
    out_dim: int
    options: SimulationOptions
    consts: Dict[str, float] = field(default_factory=dict)
    x_pde: Optional[np.ndarray] = None
    bnd_pts: Dict[str, np.ndarray] = field(default_factory=dict)       # edge (and "IC") -> [n, d]
    bnd_val: List[Dict[str, np.ndarray]] = field(default_factory=lambda: [{}, {}])
    x_vel: Optional[np.ndarray] = None
    x_pres: Optional[np.ndarray] = None
    x_test: Optional[np.ndarray] = None
    sol_noise: List[np.ndarray] = field(default_factory=list)          # [u, v, p] fit targets
    sol_test: List[np.ndarray] = field(default_factory=list)           # [u, v, p] exact (normalised)
    extra: Dict[str, np.ndarray] = field(default_factory=dict)         # Poisson: f, g, x_BC, ...

    @property
    def norm_vel(self) -> float:
        return self.consts["norm_vel"]

    @property
    def norm_pre(self) -> float:
        return self.consts["norm_pre"]

    def n_params(self) -> int:
        sizes = [self.dim] + list(self.hidden) + [self.out_dim]
        return sum(sizes[i] * sizes[i + 1] + sizes[i + 1] for i in range(len(sizes) - 1))


# --------------------------------------------------------------------------------------------
# shared helpers
# --------------------------------------------------------------------------------------------

def spread(vec) -> float:
    """cavity_steady.py:113"""
    return float(np.max(vec) - np.min(vec))


def build_grid(lx, ux, ly, uy, n1, n2, times: Optional[np.ndarray] = None) -> np.ndarray:
    """All mesh vertices, x fastest (cavity_steady.py:88-91); (t, x, y) rows when ``times`` is
    given (cavity_unsteady.py:94-95)."""
    x_vec = np.linspace(lx, ux, n1 + 1)
    y_vec = np.linspace(ly, uy, n2 + 1)
    xx, yy = np.meshgrid(x_vec, y_vec)  # rows = y, cols = x  -> ravel gives x fastest
    g2 = np.stack([xx.ravel(), yy.ravel()], axis=1)
    if times is None:
        return g2
    return np.concatenate([np.concatenate([np.full((g2.shape[0], 1), t), g2], axis=1) for t in times], axis=0)


def split_indices(n_grid: int, n_pts: Dict[str, int], rng: np.random.Generator) -> Dict[str, np.ndarray]:
    """One permutation cut at cumulative counts; remainder dropped; overshoot silently truncates the
    last subsets (cavity_steady.py:93-96)."""
    keys = ("PDE", "Vel", "Pres", "Test")
    perm = rng.permutation(n_grid)
    parts = np.split(perm, np.cumsum([n_pts[k] for k in keys]))[:-1]
    return {k: v for k, v in zip(keys, parts)}


def sample_edges(rng, n_bc: int, dim: int, lx, ux, ly, uy, T: Optional[float] = None,
                 float32_sampling: bool = False) -> Dict[str, np.ndarray]:
    """tf.random.uniform([n_BC, dim], minval, maxval) per edge with min == max on the fixed
    coordinate (cavity_steady.py:124-130; cavity_unsteady.py:130,137-140 prepends t ~ U(0,T)).
    The Cavity scripts sample in TensorFlow's default float32, the others in float64."""
    corners = {"BOT": ([lx, ly], [ux, ly]), "DX": ([ux, ly], [ux, uy]),
               "TOP": ([lx, uy], [ux, uy]), "SX": ([lx, ly], [lx, uy])}
    out = {}
    for edge in EDGES:
        lo, hi = corners[edge]
        if dim == 3:
            lo, hi = [0.0] + lo, [T] + hi
        lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
        u = rng.random((n_bc, dim))
        pts = lo + u * (hi - lo)
        out[edge] = pts.astype(np.float32).astype(np.float64) if float32_sampling else pts
    return out


def generate_noise(rng, n: int, factor: float) -> np.ndarray:
    """cavity_steady.py:141-143"""
    return rng.standard_normal(n) * factor


def _boundary_targets(data: ProblemData, bnd_raw: List[Dict[str, object]], rng, n_bc: int) -> None:
    """value/norm_vel (scalar) or value(points)/norm_vel (callable) + noise; noise is drawn for
    component 0 then component 1 edge by edge (cavity_steady.py:132-147)."""
    nv = data.norm_vel
    zero = np.zeros(n_bc)
    for comp in (0, 1):
        for edge, value in bnd_raw[comp].items():
            if isinstance(value, (int, float)):
                data.bnd_val[comp][edge] = zero + value / nv
            else:
                data.bnd_val[comp][edge] = zero + value(data.bnd_pts[edge]) / nv
    for edge in bnd_raw[0].keys():
        data.bnd_val[0][edge] = data.bnd_val[0][edge] + generate_noise(rng, n_bc, data.options.noise_factor_bnd)
        data.bnd_val[1][edge] = data.bnd_val[1][edge] + generate_noise(rng, n_bc, data.options.noise_factor_bnd)


def _fit_targets(data: ProblemData, grid: np.ndarray, fields_norm: Sequence[np.ndarray],
                 idx: Dict[str, np.ndarray], rng) -> None:
    """cavity_steady.py:150-153 and exact_value (:200)."""
    o = data.options
    data.x_vel, data.x_pres, data.x_test = grid[idx["Vel"]], grid[idx["Pres"]], grid[idx["Test"]]
    u_n = fields_norm[0][idx["Vel"]] + generate_noise(rng, len(idx["Vel"]), o.noise_factor_fit)
    v_n = fields_norm[1][idx["Vel"]] + generate_noise(rng, len(idx["Vel"]), o.noise_factor_fit)
    p_n = fields_norm[2][idx["Pres"]] + generate_noise(rng, len(idx["Pres"]), o.noise_factor_fit)
    data.sol_noise = [u_n, v_n, p_n]
    data.sol_test = [f[idx["Test"]] for f in fields_norm]


def _split_counts(n_pts: Dict[str, int], n_grid: int) -> Dict[str, int]:
    """Counts handed to the permutation split.  A collocation request that fits the grid is taken
    from it like the reference does; a larger one (BASELINE sizes) is sampled uniformly instead and
    takes no grid vertices, so the Vel / Pres / Test subsets keep their requested sizes."""
    return {**n_pts, "PDE": n_pts["PDE"] if n_pts["PDE"] <= n_grid else 0}


def _collocation(grid, idx, n_pde, rng, lo, hi) -> np.ndarray:
    if n_pde <= grid.shape[0]:
        return grid[idx["PDE"]]
    lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    return lo + rng.random((n_pde, len(lo))) * (hi - lo)


def _round_all(data: ProblemData) -> ProblemData:
    for name in ("x_pde", "x_vel", "x_pres", "x_test"):
        v = getattr(data, name)
        if v is not None:
            setattr(data, name, _f32(v))
    data.bnd_pts = {k: _f32(v) for k, v in data.bnd_pts.items()}
    data.bnd_val = [{k: _f32(v) for k, v in d.items()} for d in data.bnd_val]
    data.sol_noise = [_f32(v) for v in data.sol_noise]
    data.sol_test = [_f32(v) for v in data.sol_test]
    data.extra = {k: _f32(v) for k, v in data.extra.items()}
    return data


def _opts(options: Optional[SimulationOptions], **counts) -> SimulationOptions:
    if options is not None:
        return options
    o = SimulationOptions()
    o.n_pts.update({k: v for k, v in counts.items() if k in o.n_pts})
    o.noise_factor_bnd = counts.get("noise_bnd", 0.0)
    o.noise_factor_fit = counts.get("noise_fit", 0.0)
    o.epochs = counts.get("epochs", 10000)
    return o


# --------------------------------------------------------------------------------------------
# synthetic stand-in for the absent FEM data
# --------------------------------------------------------------------------------------------

def synthetic_cavity_field(grid_xy: np.ndarray, U: float, t: Optional[np.ndarray] = None, T: float = 1.0):
    """Smooth divergence-free lid-driven-cavity-like (u, v, p) on grid vertices.

    Stream function psi = U * 16 x^2 (1-x)^2 * y^2 (y-1):  u = psi_y vanishes on the three walls
    and equals U*16x^2(1-x)^2 on the lid y=1; v = -psi_x vanishes on all four edges.  Pressure is a
    smooth field, mean-subtracted like cavity_steady.py:105.  With ``t`` the fields ramp up as
    1-exp(-5 t/T) (an impulsively started lid)."""
    x, y = grid_xy[:, 0], grid_xy[:, 1]
    fx, dfx = 16 * x ** 2 * (1 - x) ** 2, 16 * (2 * x * (1 - x) ** 2 - 2 * x ** 2 * (1 - x))
    gy, dgy = y ** 2 * (y - 1), 3 * y ** 2 - 2 * y
    ramp = 1.0 if t is None else (1.0 - np.exp(-5.0 * t / T))
    u = U * fx * dgy * ramp
    v = -U * dfx * gy * ramp
    p = U * (np.cos(np.pi * x) * np.sin(np.pi * y) + 0.25 * x * y) * ramp
    return u, v, p - np.mean(p)


def load_fem_fields(path: str):
    """FEniCS XDMF companion file: VisualisationVector/0 = velocity [n,2|3], /1 = pressure
    (cavity_steady.py:100-105)."""
    from .h5lite import H5File
    f = H5File(path)
    vel, pre = f["VisualisationVector/0"], f["VisualisationVector/1"].reshape(-1)
    return vel[:, 0], vel[:, 1], pre - np.mean(pre)


def load_unsteady_fem_fields(folder: str, n_times: int, pattern: str = "navier-stokes_SI_cavity_unsteady_{:05d}.h5"):
    """The per-time-step FEniCS files of DataGeneration/fluid_solver_unsteady.py as the unsteady script reads them
    (cavity_unsteady.py:103-113): file ``x`` holds VisualisationVector/0 = velocity [n, 2|3] and /1 = pressure [n] on
    the grid vertices at time step ``x``; the pressure of EVERY step has its own mean subtracted (:109); the steps are
    concatenated in time order, matching the (t, y, x) ordering of ``dom_grid`` (:95)."""
    import os
    from .h5lite import H5File
    u, v, p = [], [], []
    for step in range(int(n_times)):
        path = os.path.join(folder, pattern.format(step))
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: time step {step} of {n_times} is missing")
        f = H5File(path)
        vel, pre = f["VisualisationVector/0"], np.asarray(f["VisualisationVector/1"]).reshape(-1)
        u.append(np.asarray(vel[:, 0], dtype=np.float64))
        v.append(np.asarray(vel[:, 1], dtype=np.float64))
        p.append(pre.astype(np.float64) - np.mean(pre))
    return np.concatenate(u), np.concatenate(v), np.concatenate(p)


def read_gmsh_nodes(path: str) -> np.ndarray:
    """Node coordinates ``[n, 3]`` (ordered by node tag) of an ASCII gmsh 4.x mesh -- the ``$Nodes`` section of
    Examples/Coronary_Flow/coroParam.msh.  dolfin keeps this order: the arrays of sol_pinn.h5
    (coronary_flow_steady.py:297-301, written at ``Mesh/0/mesh/geometry``) vanish on the no-slip nodes found here."""
    with open(path) as fh:
        lines = fh.read().split("\n")
    if not lines[1].startswith("4."):
        raise ValueError(f"{path}: gmsh format {lines[1]!r} is not 4.x ASCII")
    i = lines.index("$Nodes") + 1
    n_blocks, n_nodes = (int(t) for t in lines[i].split()[:2])
    xyz = np.zeros((n_nodes, 3))
    i += 1
    for _ in range(n_blocks):
        n = int(lines[i].split()[3])
        tags = [int(t) for t in lines[i + 1:i + 1 + n]]
        for k, tag in enumerate(tags):
            xyz[tag - 1] = [float(c) for c in lines[i + 1 + n + k].split()]
        i += 1 + 2 * n
    return xyz


def load_coronary_geometry(source) -> Dict[str, np.ndarray]:
    """``source``: a dict, an ``.npz`` (tests/golden/coronary_geometry.npz layout: nodes, bpoints_xy, bpoints_label,
    u, v, p) or a tuple ``(fem_h5, bpoints_npy)`` naming the DataGeneration outputs the script reads
    (coronary_flow_steady.py:92-117,134: ``Mesh/0/mesh/geometry``, ``VisualisationVector/0,1`` and the
    ``[x, y, 0, label]`` rows of bpoints.npy)."""
    if isinstance(source, dict):
        return source
    if isinstance(source, (tuple, list)):
        from .h5lite import H5File
        f = H5File(source[0])
        vel, pre = f["VisualisationVector/0"], f["VisualisationVector/1"].reshape(-1)
        b = np.load(source[1])
        return {"nodes": f["Mesh/0/mesh/geometry"][:, :2], "bpoints_xy": b[:, :2], "bpoints_label": b[:, 3].astype(np.int8),
                "u": vel[:, 0], "v": vel[:, 1], "p": pre}
    with np.load(source) as z:
        return {k: z[k] for k in z.files}


# --------------------------------------------------------------------------------------------
# the test cases
# --------------------------------------------------------------------------------------------

def cavity_steady(options: Optional[SimulationOptions] = None, seed: int = 1, fem_fields=None,
                  lid_velocity: float = 500.0, **counts) -> ProblemData:
    """Examples/Cavity_Steady/cavity_steady.py:60-153 (U=500, nu=1: fluid_solver_steady.py:9-10)."""
    o = _opts(options, **counts)
    rng = np.random.default_rng(seed)
    d = ProblemData("cavity_steady", 2, [32, 32, 32], 3, o)
    grid = build_grid(0, 1, 0, 1, 100, 100)
    idx = split_indices(grid.shape[0], _split_counts(o.n_pts, grid.shape[0]), rng)
    u_ex, v_ex, p_ex = fem_fields if fem_fields is not None else synthetic_cavity_field(grid, lid_velocity)
    d.consts = {"norm_vel": max(spread(u_ex), spread(v_ex)), "norm_pre": spread(p_ex)}
    fields_norm = [u_ex / d.norm_vel, v_ex / d.norm_vel, p_ex / d.norm_pre]
    d.x_pde = _collocation(grid, idx, o.n_pts["PDE"], rng, [0, 0], [1, 1])
    d.bnd_pts = sample_edges(rng, o.n_pts["BC"], 2, 0, 1, 0, 1, float32_sampling=True)
    bnd_raw = [{"BOT": 0, "DX": 0, "TOP": lid_velocity, "SX": 0}, {"BOT": 0, "DX": 0, "TOP": 0, "SX": 0}]
    _boundary_targets(d, bnd_raw, rng, o.n_pts["BC"])
    _fit_targets(d, grid, fields_norm, idx, rng)
    return _round_all(d)


def cavity_unsteady(options: Optional[SimulationOptions] = None, seed: int = 1, hidden=(32, 32, 32),
                    lid_velocity: float = 1.0, T: float = 1e-2, dt: float = 1e-4, n_times: Optional[int] = None,
                    fem_fields=None, **counts) -> ProblemData:
    """Examples/Cavity_Unsteady/cavity_unsteady.py:60-163; rows are (t, x, y).  ``hidden`` lets the
    BASELINE config 5 ask for the 8x128 network.  ``fem_fields``: the folder of the per-time-step FEniCS files
    (``../../DataGeneration/data/UnsteadyCase``, cavity_unsteady.py:104-105; read by ``load_unsteady_fem_fields``) or
    a ready ``(u, v, p)`` triple over the (t, y, x) grid; without it a smooth synthetic field stands in."""
    o = _opts(options, **counts)
    rng = np.random.default_rng(seed)
    d = ProblemData("cavity_unsteady", 3, list(hidden), 3, o)
    n_times = int(T / dt) if n_times is None else n_times
    time_vec = np.arange(0.0, T, dt)[:n_times]
    grid2 = build_grid(0, 1, 0, 1, 100, 100)
    grid = build_grid(0, 1, 0, 1, 100, 100, times=time_vec)
    idx = split_indices(grid.shape[0], _split_counts(o.n_pts, grid.shape[0]), rng)
    n2 = grid2.shape[0]
    if fem_fields is None:
        u_ex, v_ex, p_ex = synthetic_cavity_field(grid[:, 1:], lid_velocity, t=grid[:, 0], T=T)
        # per-time-step mean subtraction of the pressure (cavity_unsteady.py:109)
        p_ex = (p_ex.reshape(-1, n2) - p_ex.reshape(-1, n2).mean(axis=1, keepdims=True)).reshape(-1)
    else:
        u_ex, v_ex, p_ex = (load_unsteady_fem_fields(fem_fields, len(time_vec)) if isinstance(fem_fields, str)
                            else tuple(np.asarray(a, dtype=np.float64).reshape(-1) for a in fem_fields))
        if not (u_ex.shape[0] == v_ex.shape[0] == p_ex.shape[0] == grid.shape[0]):
            raise ValueError(f"FEM fields hold {u_ex.shape[0]} values, the space-time grid {grid.shape[0]} "
                             f"({len(time_vec)} time steps x {n2} vertices)")
    d.consts = {"norm_vel": max(spread(u_ex), spread(v_ex)), "norm_pre": spread(p_ex), "T": T}
    fields_norm = [u_ex / d.norm_vel, v_ex / d.norm_vel, p_ex / d.norm_pre]
    d.x_pde = _collocation(grid, idx, o.n_pts["PDE"], rng, [0, 0, 0], [T, 1, 1])
    d.bnd_pts = sample_edges(rng, o.n_pts["BC"], 3, 0, 1, 0, 1, T=T, float32_sampling=True)
    # starting_sampling([0, Le_x, Le_y], [0, Ue_x, Ue_y]) (cavity_unsteady.py:131-132)
    ic = rng.random((o.n_pts["IC"], 3)) * np.array([0.0, 1.0, 1.0])
    d.bnd_pts["IC"] = ic.astype(np.float32).astype(np.float64)
    bnd_raw = [{"BOT": 0, "DX": 0, "TOP": lid_velocity, "SX": 0}, {"BOT": 0, "DX": 0, "TOP": 0, "SX": 0}]
    _boundary_targets(d, bnd_raw, rng, o.n_pts["BC"])
    _fit_targets(d, grid, fields_norm, idx, rng)
    return _round_all(d)


def colliding_flow(options: Optional[SimulationOptions] = None, seed: int = 1, **counts) -> ProblemData:
    """Examples/Colliding_Flow/colliding_flow.py:60-150."""
    o = _opts(options, **counts)
    rng = np.random.default_rng(seed)
    d = ProblemData("colliding_flow", 2, [32, 32, 32], 3, o)
    p_f = lambda x: 60 * x[:, 0] ** 2 * x[:, 1] - 20 * x[:, 1] ** 3
    u_f = lambda x: 20 * x[:, 0] * x[:, 1] ** 3
    v_f = lambda x: 5 * x[:, 0] ** 4 - 5 * x[:, 1] ** 4
    grid = build_grid(-1, 1, -1, 1, 100, 100)
    idx = split_indices(grid.shape[0], _split_counts(o.n_pts, grid.shape[0]), rng)
    u_ex, v_ex, p_ex = u_f(grid), v_f(grid), p_f(grid)
    d.consts = {"norm_vel": max(spread(u_ex), spread(v_ex)), "norm_pre": spread(p_ex)}
    fields_norm = [u_ex / d.norm_vel, v_ex / d.norm_vel, p_ex / d.norm_pre]
    d.x_pde = _collocation(grid, idx, o.n_pts["PDE"], rng, [-1, -1], [1, 1])
    d.bnd_pts = sample_edges(rng, o.n_pts["BC"], 2, -1, 1, -1, 1)
    bnd_raw = [{e: u_f for e in EDGES}, {e: v_f for e in EDGES}]
    _boundary_targets(d, bnd_raw, rng, o.n_pts["BC"])
    _fit_targets(d, grid, fields_norm, idx, rng)
    return _round_all(d)


def poiseuille_flow(options: Optional[SimulationOptions] = None, seed: int = 1, **counts) -> ProblemData:
    """Examples/Poiseuille_Flow/poiseuille_flow.py:60-164."""
    o = _opts(options, **counts)
    rng = np.random.default_rng(seed)
    d = ProblemData("poiseuille_flow", 2, [32, 32, 32], 3, o)
    lx, ux, ly, uy = 0.0, 1.0, 0.0, 0.1
    delta, L = (uy - ly) / 2, ux - lx
    rho, mu, P_str, P_end = 3100.0, 890.0, 1e6, 0.0
    P_x = P_end - P_str
    p_f = lambda x: (P_end - P_str) / L * x[:, 0] + P_str
    u_f = lambda x: -P_x * x[:, 1] * (2 - x[:, 1] / delta) * delta / (2 * mu)
    v_f = lambda x: 0 * x[:, 0]
    grid = build_grid(lx, ux, ly, uy, 100, 25)
    idx = split_indices(grid.shape[0], _split_counts(o.n_pts, grid.shape[0]), rng)
    u_ex, v_ex, p_ex = u_f(grid), v_f(grid), p_f(grid)
    d.consts = {"norm_vel": max(spread(u_ex), spread(v_ex)), "norm_pre": spread(p_ex), "rho": rho, "mu": mu}
    fields_norm = [u_ex / d.norm_vel, v_ex / d.norm_vel, p_ex / d.norm_pre]
    d.x_pde = _collocation(grid, idx, o.n_pts["PDE"], rng, [lx, ly], [ux, uy])
    d.bnd_pts = sample_edges(rng, o.n_pts["BC"], 2, lx, ux, ly, uy)
    # poiseuille_flow.py:84-92,117: DX carries the outflow stress target P_end, SX the inflow profile
    bnd_raw = [{"BOT": 0, "DX": P_end, "TOP": 0, "SX": u_f}, {"BOT": 0, "DX": 0, "TOP": 0, "SX": 0}]
    _boundary_targets(d, bnd_raw, rng, o.n_pts["BC"])
    _fit_targets(d, grid, fields_norm, idx, rng)
    return _round_all(d)


CORONARY_EDGES = ("NOSL", "INF", "OUT1", "OUT2")   # bpoints.npy labels 0..3 (DataGeneration/coronary.py:64)


def coronary_flow(geometry, options: Optional[SimulationOptions] = None, seed: int = 1, norm_vel: Optional[float] = None,
                  norm_pre: Optional[float] = None, **counts) -> ProblemData:
    """Examples/Coronary_Flow/coronary_flow_steady.py:60-150: collocation / fit / test points are a permutation split of
    the mesh nodes, boundary points come labelled from ``bpoints.npy`` (all of them are used; the BC count of the
    options file only switches the boundary terms on), the inlet carries the parabolic profile u_inf / v_inf.
    ``norm_vel`` / ``norm_pre`` override the spreads of the supplied fields (tests replay Test_Case_#123 with the FEM
    spreads recovered from sol_pinn.h5 = trained model x norm)."""
    g = load_coronary_geometry(geometry)
    o = _opts(options, **counts)
    rng = np.random.default_rng(seed)
    d = ProblemData("coronary_flow", 2, [32, 32, 32], 3, o)
    H, U, x0, y0 = np.sqrt(0.4 ** 2 + 0.1 ** 2), 20.0, -1.4, -0.8
    mu, rho = 1e-2, 1.06e3
    ni = 1e4 * mu / rho
    cos_t, sin_t = np.cos(np.arctan(1 / 4)), np.sin(np.arctan(1 / 4))
    r_in = lambda x: np.sqrt((x[:, 0] - x0) ** 2 + (x[:, 1] - y0) ** 2) / H
    u_inf = lambda x: U * cos_t * r_in(x) * (1 - r_in(x))
    v_inf = lambda x: U * sin_t * r_in(x) * (1 - r_in(x))
    grid = np.asarray(g["nodes"], dtype=np.float64)[:, :2]
    split = {k: o.n_pts[k] for k in ("PDE", "Vel", "Pres", "Test")}      # key_subset (:100)
    idx = split_indices(grid.shape[0], _split_counts(split, grid.shape[0]), rng)
    u_ex, v_ex, p_ex = (np.asarray(g[k], dtype=np.float64) for k in ("u", "v", "p"))   # no mean subtraction here (:112)
    d.consts = {"norm_vel": max(spread(u_ex), spread(v_ex)) if norm_vel is None else float(norm_vel),
                "norm_pre": spread(p_ex) if norm_pre is None else float(norm_pre), "ni": ni, "rho": rho, "mu": mu}
    fields_norm = [u_ex / d.norm_vel, v_ex / d.norm_vel, p_ex / d.norm_pre]
    if o.n_pts["PDE"] > grid.shape[0]:
        raise ValueError("Coronary_Flow draws its collocation points from the mesh nodes")
    d.x_pde = grid[idx["PDE"]]
    bxy, lab = np.asarray(g["bpoints_xy"], dtype=np.float64), np.asarray(g["bpoints_label"]).astype(int)
    d.bnd_pts = {e: bxy[lab == i] for i, e in enumerate(CORONARY_EDGES)}
    bnd_raw = [{"NOSL": 0, "INF": u_inf, "OUT1": 0, "OUT2": 0}, {"NOSL": 0, "INF": v_inf, "OUT1": 0, "OUT2": 0}]
    nv = d.norm_vel
    for comp in (0, 1):      # :141-146
        for edge, value in bnd_raw[comp].items():
            zero = np.zeros(len(d.bnd_pts[edge]))
            d.bnd_val[comp][edge] = zero + (value / nv if isinstance(value, (int, float)) else value(d.bnd_pts[edge]) / nv)
    for edge in bnd_raw[0].keys():   # :154-156
        d.bnd_val[0][edge] = d.bnd_val[0][edge] + generate_noise(rng, len(d.bnd_pts[edge]), o.noise_factor_bnd)
        d.bnd_val[1][edge] = d.bnd_val[1][edge] + generate_noise(rng, len(d.bnd_pts[edge]), o.noise_factor_bnd)
    _fit_targets(d, grid, fields_norm, idx, rng)
    return _round_all(d)


def poisson(mixed: bool = True, seed: int = 1, num_pde: int = 200, num_bc: int = 20,
            num_test: int = 1000) -> ProblemData:
    """Examples/Poisson_Problem/poisson_misto.py:21-60 (mixed=True) / poisson.py:20-56."""
    rng = np.random.default_rng(seed)
    o = SimulationOptions()
    o.n_pts.update({"PDE": num_pde, "BC": num_bc, "Test": num_test})
    d = ProblemData("poisson_misto" if mixed else "poisson", 2, [20, 20, 20], 1, o)
    W1 = W2 = 2 * np.pi
    uni = lambda n, lo, hi: np.asarray(lo) + rng.random((n, 2)) * (np.asarray(hi, dtype=np.float64) - np.asarray(lo))
    d.x_pde = uni(num_pde, [0, 0], [W1, W2])
    x0 = uni(num_bc, [0, 0], [0, W2]); x1 = uni(num_bc, [W1, 0], [W1, W2])
    y0 = uni(num_bc, [0, 0], [W1, 0]); y1 = uni(num_bc, [0, W2], [W1, W2])
    d.x_test = uni(num_test, [0, 0], [W1, W2])
    d = _round_all(d)
    d.extra["f"] = _f32(2 * np.sin(d.x_pde[:, 0]) * np.sin(d.x_pde[:, 1]))
    d.extra["u_test"] = _f32(np.sin(d.x_test[:, 0]) * np.sin(d.x_test[:, 1]))
    if mixed:
        d.extra["x_BC_D"] = _f32(np.concatenate([y0, y1], axis=0))
        d.extra["x_BC_N"] = _f32(np.concatenate([x0, x1], axis=0))
        d.extra["g"] = _f32(np.sin(d.extra["x_BC_N"][:, 1]))  # poisson_misto.py:29
    else:
        d.extra["x_BC"] = _f32(np.concatenate([x0, x1, y0, y1], axis=0))
    d.consts = {"norm_vel": 1.0, "norm_pre": 1.0}
    return d


def colliding_flow_pressmean(seed: int = 1, num_pde: int = 1000, num_bc: int = 100, num_col: int = 0, num_test: int = 1000,
                             num_pres: int = 100, press_mode: str = "Mean", collocation: bool = False,
                             use_noise: bool = False) -> ProblemData:
    """Examples/Colliding_Flow/colliding_flow_pressmean.py:35-97,127-133: Stokes flow on (-1, 1)^2 with a 2-20x3-3
    network, all point sets uniform random, normalisation by the maxima of |u|, |v|, |p| on the boundary points,
    pressure fixed by ``press_mode``: "Mean" (ns.Loss over |mean p|), "Collocation" (fit p on x_pres) or "None"."""
    if press_mode not in ("Mean", "Collocation", "None"):
        raise ValueError("press_mode must be 'Mean', 'Collocation' or 'None'")
    rng = np.random.default_rng(seed)
    o = SimulationOptions(epochs=5000)
    o.n_pts.update({"PDE": num_pde, "BC": num_bc, "Vel": num_col, "Pres": num_pres, "Test": num_test})
    d = ProblemData("colliding_flow_pressmean", 2, [20, 20, 20], 3, o)
    a, b = -1.0, 1.0
    p_exact = lambda x: 60 * x[:, 0] ** 2 * x[:, 1] - 20 * x[:, 1] ** 3
    u_exact = lambda x: 20 * x[:, 0] * x[:, 1] ** 3
    v_exact = lambda x: 5 * x[:, 0] ** 4 - 5 * x[:, 1] ** 4
    uni = lambda n, lo, hi: np.asarray(lo) + rng.random((n, 2)) * (np.asarray(hi, dtype=np.float64) - np.asarray(lo))
    d.x_pde = uni(num_pde, [a, a], [b, b])
    d.x_vel = uni(num_col, [a, a], [b, b])                           # x_col
    edges = [uni(num_bc, [a, a], [a, b]), uni(num_bc, [b, a], [b, b]), uni(num_bc, [a, a], [b, a]), uni(num_bc, [a, b], [b, b])]
    d.x_test = uni(num_test, [a, a], [b, b])
    d.x_pres = uni(num_pres, [a, a], [b, b])
    d = _round_all(d)
    x_bcd = _f32(np.concatenate(edges, axis=0))                      # x0, x1, y0, y1 (:72-78)
    vel_max = max(np.max(np.abs(u_exact(x_bcd))), np.max(np.abs(v_exact(x_bcd))))
    p_max = np.max(np.abs(p_exact(x_bcd)))
    d.consts = {"norm_vel": float(vel_max), "norm_pre": float(p_max), "press_mode": press_mode,
                "collocation": bool(collocation)}
    noise = (lambda: 1e-1 * rng.standard_normal(x_bcd.shape[0])) if use_noise else (lambda: 0.0)   # :127-133
    d.extra = {"x_BCD": x_bcd,
               "bcd_u": _f32((u_exact(x_bcd) + noise()) / vel_max), "bcd_v": _f32((v_exact(x_bcd) + noise()) / vel_max),
               "col_u": _f32(u_exact(d.x_vel) / vel_max), "col_v": _f32(v_exact(d.x_vel) / vel_max),
               "col_p": _f32(p_exact(d.x_pres) / p_max)}
    d.sol_test = [_f32(u_exact(d.x_test) / vel_max), _f32(v_exact(d.x_test) / vel_max), _f32(p_exact(d.x_test) / p_max)]
    return d


BUILDERS: Dict[str, Callable[..., ProblemData]] = {
    "poisson": lambda **kw: poisson(mixed=False, **kw),
    "poisson_misto": lambda **kw: poisson(mixed=True, **kw),
    "poiseuille_flow": poiseuille_flow,
    "colliding_flow": colliding_flow,
    "cavity_steady": cavity_steady,
    "cavity_unsteady": cavity_unsteady,
    "coronary_flow": coronary_flow,
    "colliding_flow_pressmean": colliding_flow_pressmean,
}

# BASELINE.json configs: name -> builder kwargs
BASELINE_CONFIGS: Dict[str, Dict] = {
    "Poisson_Problem": dict(builder="poisson_misto"),
    "Poiseuille_Flow": dict(builder="poiseuille_flow", PDE=10_000, BC=250, Vel=10, Pres=0, Test=1000),
    "Colliding_Flow": dict(builder="colliding_flow", PDE=100_000, BC=100, Vel=5, Pres=1, Test=1000),
    "Cavity_Steady": dict(builder="cavity_steady", PDE=1_000_000, BC=1000, Vel=100, Pres=1, Test=1000,
                          noise_bnd=0.01, noise_fit=0.01),
    "Cavity_Unsteady": dict(builder="cavity_unsteady", PDE=4_000_000, BC=1000, IC=1000, Vel=1, Pres=1, Test=1000,
                            noise_bnd=0.05, noise_fit=0.05, hidden=(128,) * 8),
}


def build_baseline_config(name: str, seed: int = 1, **override) -> ProblemData:
    cfg = dict(BASELINE_CONFIGS[name])
    cfg.update(override)
    builder = cfg.pop("builder")
    if builder.startswith("poisson"):      # the Poisson scripts hard-code their counts (poisson_misto.py:49-53)
        cfg = {{"PDE": "num_pde", "BC": "num_bc", "Test": "num_test"}.get(k, k): v for k, v in cfg.items()}
    return BUILDERS[builder](seed=seed, **cfg)
