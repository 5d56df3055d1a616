"""simulation_options.txt parser vs the reference's five checked-in option files (fixtures are the
files verbatim, tests/golden/simulation_options.json)."""
import json
import os

import pytest

from pinns_fluid_dynamics_b200.options import (SimulationOptions, parse_option_lines, read_simulation_options,
                                               write_simulation_options)

GOLD = os.path.join(os.path.dirname(__file__), "golden", "simulation_options.json")
# epochs, noise_fit, noise_bnd, PDE, BC, IC, Vel, Pres, Test  
EXPECTED = {
    "Poiseuille_Flow": (10000, 0.0, 0.0, 1000, 100, 100, 10, 0, 1000),
    "Colliding_Flow": (10000, 0.0, 0.0, 1000, 100, 100, 5, 1, 10000),
    "Cavity_Steady": (10000, 0.01, 0.01, 1000, 1000, 1000, 100, 1, 1000),
    "Cavity_Unsteady": (10000, 0.05, 0.05, 10000, 1000, 1000, 1, 1, 10000),
    "Coronary_Flow": (30000, 0.01, 0.01, 3000, 800, 0, 50, 0, 1000),
}


@pytest.mark.parametrize("case", sorted(EXPECTED))
def test_checked_in_option_files(case):
    text = json.load(open(GOLD))[case]
    opt = parse_option_lines(text.splitlines(keepends=True))
    e = EXPECTED[case]
    assert (opt.epochs, opt.noise_factor_fit, opt.noise_factor_bnd) == e[:3]
    assert tuple(opt.n_pts[k] for k in ("PDE", "BC", "IC", "Vel", "Pres", "Test")) == e[3:]


def test_positional_not_label_based():
    """Q5: the value under the label 'NOISE ON BOUNDARY' is read into noise_factor_fit."""
    lines = ["hdr\n", "TRAINING EPOCHS\n", "7\n", "NOISE ON BOUNDARY\n", "0.25\n", "NOISE ON FITTING\n", "0.5\n"]
    for lab, v in zip("ABCDEF", (1, 2, 3, 4, 5, 6)):
        lines += [f"{lab}\n", f"{v}\n"]
    lines += ["### End of the File ###"]
    opt = parse_option_lines(lines)
    assert opt.epochs == 7 and opt.noise_factor_fit == 0.25 and opt.noise_factor_bnd == 0.5
    assert opt.n_pts == {"PDE": 1, "BC": 2, "IC": 3, "Vel": 4, "Pres": 5, "Test": 6}
    assert opt.use_collloss and opt.fit_pressure


def test_last_line_is_dropped_and_flags(tmp_path):
    o = SimulationOptions(epochs=12, noise_factor_fit=0.1, noise_factor_bnd=0.2,
                          n_pts={"PDE": 10, "BC": 0, "IC": 0, "Vel": 3, "Pres": 0, "Test": 9})
    p = tmp_path / "simulation_options.txt"
    write_simulation_options(str(p), o)
    assert len(p.read_text().split("\n")) == 20 and not p.read_text().endswith("\n")
    r = read_simulation_options(str(p))
    assert (r.epochs, r.noise_factor_fit, r.noise_factor_bnd, r.n_pts) == (12, 0.1, 0.2, o.n_pts)
    assert r.use_collloss and not r.use_boundary and r.fit_velocity and not r.fit_pressure


def test_too_short_file_raises():
    with pytest.raises(ValueError):
        parse_option_lines(["a\n", "b\n", "1\n"])
