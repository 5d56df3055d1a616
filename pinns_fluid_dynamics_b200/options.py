"""`simulation_options.txt` reader.

Mirrors the positional parser every example script of the reference carries
(Examples/Cavity_Steady/cavity_steady.py:37-58, identical in poiseuille_flow.py:37-58,
colliding_flow.py:37-58, cavity_unsteady.py:37-58):

* the file is read with ``readlines()[0:-1:2]`` -- i.e. the LAST line is dropped and every
  second remaining line is kept, so labels are ignored and values are taken by position;
* index 1 -> epochs, 2 -> noise_factor_fit, 3 -> noise_factor_bnd (note: the label above
  index 2 says "NOISE ON BOUNDARY" in the Cavity/Colliding files -- the swap is the
  reference's behaviour and is kept), 4..9 -> point counts
  PDE / BC (per edge) / IC / Vel / Pres / Test.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict

POINT_KEYS = ("PDE", "BC", "IC", "Vel", "Pres", "Test")


@dataclass
class SimulationOptions:
    epochs: int = 10000
    noise_factor_fit: float = 0.0
    noise_factor_bnd: float = 0.0
    n_pts: Dict[str, int] = field(default_factory=lambda: {k: 0 for k in POINT_KEYS})

    # flags exactly as the scripts derive them (cavity_steady.py:54-58)
    @property
    def use_collloss(self) -> bool:
        return bool(self.n_pts["PDE"])

    @property
    def use_boundary(self) -> bool:
        return bool(self.n_pts["BC"])

    @property
    def use_initialc(self) -> bool:
        return bool(self.n_pts["IC"])

    @property
    def fit_velocity(self) -> bool:
        return bool(self.n_pts["Vel"])

    @property
    def fit_pressure(self) -> bool:
        return bool(self.n_pts["Pres"])


def parse_option_lines(lines) -> SimulationOptions:
    """Positional parse of an already-read list of lines (``readlines()`` output)."""
    kept = list(lines)[0:-1:2]
    if len(kept) < 10:
        raise ValueError(
            f"simulation options need 10 value slots after slicing, found {len(kept)}")
    opt = SimulationOptions()
    opt.epochs = int(kept[1])
    opt.noise_factor_fit = float(kept[2])
    opt.noise_factor_bnd = float(kept[3])
    opt.n_pts = {k: int(kept[4 + i]) for i, k in enumerate(POINT_KEYS)}
    return opt


def read_simulation_options(path: str) -> SimulationOptions:
    with open(path) as fh:
        return parse_option_lines(fh.readlines())


def write_simulation_options(path: str, opt: SimulationOptions) -> None:
    """Write a file in the reference's 20-line layout (no trailing newline)."""
    rows = [
        "### Put this file into the folder of the given problem ###",
        "TRAINING EPOCHS", str(opt.epochs),
        "NOISE ON BOUNDARY", repr(opt.noise_factor_fit) if opt.noise_factor_fit else "0",
        "NOISE ON FITTING", repr(opt.noise_factor_bnd) if opt.noise_factor_bnd else "0",
        "POINTS PDE", str(opt.n_pts["PDE"]),
        "POINTS BOUNDARY CONDITIONS", str(opt.n_pts["BC"]),
        "POINTS INITIAL CONDITIONS", str(opt.n_pts["IC"]),
        "POINTS VELOCITY FITTING", str(opt.n_pts["Vel"]),
        "POINTS PRESSURE FITTING", str(opt.n_pts["Pres"]),
        "POINT TEST EVALUATION", str(opt.n_pts["Test"]),
        "### End of the File ###",
    ]
    with open(path, "w") as fh:
        fh.write("\n".join(rows))


def recap_lines(problem_name: str, opt: SimulationOptions, fit_velocity: bool = True, fit_pressure: bool = True,
                with_initial_conditions: bool = True):
    """The "Final Recap" rows of the scripts (cavity_steady.py:366-375).  ``fit_velocity`` / ``fit_pressure`` default to
    True because the scripts' booleans are shadowed by the lambdas of the same name (:198-199, quirk Q3), so the
    reference always prints the requested counts; coronary_flow_steady.py has no initial-condition row."""
    rows = ["Problem Name    -> {}".format(problem_name),
            "Training Epochs -> {} epochs".format(opt.epochs),
            "Pyhsical PDE Losses  -> {} points".format(opt.n_pts["PDE"]),
            "Boundary Conditions  -> {} points".format(opt.n_pts["BC"])]
    if with_initial_conditions:
        rows.append("Initial  Conditions  -> {} points".format(opt.n_pts["IC"]))
    rows += ["Fitting Velocity  -> {} points".format(opt.n_pts["Vel"] if fit_velocity else 0),
             "Fitting Pressure  -> {} points".format(opt.n_pts["Pres"] if fit_pressure else 0),
             "Noise on Boundary -> {} times a gaussian N(0,1)".format(opt.noise_factor_bnd),
             "Noise on Domain   -> {} times a gaussian N(0,1)".format(opt.noise_factor_fit)]
    return rows


def write_recap(path: str, problem_name: str, opt: SimulationOptions, **kw) -> None:
    """``Test_Options.txt`` (cavity_steady.py:377-383): one row per line, each newline-terminated."""
    with open(path, "w") as fh:
        for row in recap_lines(problem_name, opt, **kw):
            fh.write(row + "\n")
