"""Regenerate the golden fixtures from the reference checkout (run in the build container only):

    python tests/golden/make_golden.py /root/reference

Writes, next to this script:
  weights_<case>.npz   trained float64 weights of the saved runs (Weights.h5 via the product's h5lite)
  history_<case>.json  head + tail of every History_Loss.json (term names, weights, log values)
  test_options.json    every Test_Options.txt recap, verbatim
  coronary_geometry.npz  Coronary_Flow inputs: mesh nodes (coroParam.msh via the product's gmsh reader), labelled
                       boundary points (DataGeneration/data/Coronary/bpoints.npy) and the u/v/p arrays of sol_pinn.h5
                       (same node order) as the stand-in for the absent FEM solution
The GPU box has no /root/reference; tests read only these fixtures.
"""
import glob
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from pinns_fluid_dynamics_b200.h5lite import load_keras_dense_weights  # noqa: E402

HEAD = 14   # entries 0..100 of round 1 plus the first three of round 2


def trim(h):
    n = len(h["log"]["iter"])
    keep = list(range(min(HEAD, n))) + ([n - 1] if n > HEAD else [])
    pick = lambda a: [a[i] for i in keep]
    out = {"n_entries": n, "kept": keep,
           "log": {k: pick(v) for k, v in h["log"].items()},
           "log_rounds": h["log_rounds"], "losses": {}, "losses_test": {}}
    for grp in ("losses", "losses_test"):
        for name, d in h[grp].items():
            out[grp][name] = {"weight": d["weight"], "non_negative": d["non_negative"],
                              "display_sqrt": d["display_sqrt"], "log": pick(d["log"])}
    # full-length check of loss_global == sum_t weight_t * log_t, recorded for the record
    tot = np.zeros(n)
    for d in h["losses"].values():
        tot += d["weight"] * np.asarray(d["log"])
    lg = np.asarray(h["log"]["loss_global"])
    out["max_rel_dev_sum_vs_global"] = float(np.max(np.abs(tot - lg) / np.maximum(np.abs(lg), 1e-300)))
    out["iter_stride_round1"] = sorted(set(np.diff(h["log"]["iter"][:11]).tolist()))
    return out


def main(ref):
    cases = {}
    for p in sorted(glob.glob(os.path.join(ref, "Examples", "*", "Test_Case_*"))):
        case = os.path.basename(os.path.dirname(p)).lower()
        cases[case] = p
    for case, p in cases.items():
        w = os.path.join(p, "Weights.h5")
        if os.path.exists(w):
            arrs = load_keras_dense_weights(w)
            np.savez_compressed(os.path.join(HERE, f"weights_{case}.npz"), **{f"v{i}": a for i, a in enumerate(arrs)})
        hfile = os.path.join(p, "History_Loss.json")
        if os.path.exists(hfile):
            with open(hfile) as fh:
                h = json.load(fh)
            with open(os.path.join(HERE, f"history_{case}.json"), "w") as fh:
                json.dump(trim(h), fh, indent=1)
    for p in sorted(glob.glob(os.path.join(ref, "Examples_Old", "*", "Images", "*history_loss.json"))):
        case = "old_" + os.path.basename(p).replace("_history_loss.json", "").replace(" ", "_").replace("-", "").lower()
        with open(p) as fh:
            h = json.load(fh)
        with open(os.path.join(HERE, f"history_{case}.json"), "w") as fh:
            json.dump(trim(h), fh, indent=1)
    # the five checked-in option files, verbatim
    opts = {}
    for p in sorted(glob.glob(os.path.join(ref, "Examples", "*", "simulation_options.txt"))):
        with open(p) as fh:
            opts[os.path.basename(os.path.dirname(p))] = fh.read()
    with open(os.path.join(HERE, "simulation_options.json"), "w") as fh:
        json.dump(opts, fh, indent=1)
    recaps = {}
    for p in sorted(glob.glob(os.path.join(ref, "Examples", "*", "Test_Case_*", "Test_Options.txt"))):
        with open(p) as fh:
            recaps[os.path.basename(os.path.dirname(os.path.dirname(p)))] = fh.read()
    with open(os.path.join(HERE, "test_options.json"), "w") as fh:
        json.dump(recaps, fh, indent=1)
    # Coronary_Flow geometry and fields
    from pinns_fluid_dynamics_b200.h5lite import H5File
    from pinns_fluid_dynamics_b200.problems import read_gmsh_nodes
    cor = os.path.join(ref, "Examples", "Coronary_Flow")
    nodes = read_gmsh_nodes(os.path.join(cor, "coroParam.msh"))[:, :2]
    bpts = np.load(os.path.join(ref, "DataGeneration", "data", "Coronary", "bpoints.npy"))
    sol = H5File(os.path.join(cor, "sol_pinn.h5"))
    np.savez_compressed(os.path.join(HERE, "coronary_geometry.npz"), nodes=nodes.astype(np.float32),
                        bpoints_xy=bpts[:, :2].astype(np.float32), bpoints_label=bpts[:, 3].astype(np.int8),
                        u=sol["u_pinn"].astype(np.float32), v=sol["v_pinn"].astype(np.float32),
                        p=sol["p_pinn"].astype(np.float32))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
