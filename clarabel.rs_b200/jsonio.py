"""Problem files in the reference's JSON interchange format.

Reads and writes what ``DefaultSolver::save_to_file`` / ``load_from_file`` produce
(/root/reference/src/solver/implementations/default/json.rs:11-95, data file
examples/data/hs35.json): ``{"P": csc, "q": [...], "A": csc, "b": [...], "cones": [...], "settings": {...}}`` with
``csc = {"m","n","colptr","rowval","nzval"}`` and the cones in serde's externally tagged enum form
(``{"NonnegativeConeT": 4}``, ``{"PowerConeT": 0.5}``, ``{"ExponentialConeT": []}`` ...).  A problem saved by the
reference loads here and vice versa, which is the cheapest cross-check against the real solver a user can run.

Host-side convenience only: nothing on the solve path imports this module.
"""
from __future__ import annotations

import json

import numpy as np
import scipy.sparse as sp

_TAGS = {"ZeroConeT": "zero", "NonnegativeConeT": "nonneg", "SecondOrderConeT": "soc", "PSDTriangleConeT": "psd",
         "ExponentialConeT": "exp", "PowerConeT": "pow", "GenPowerConeT": "genpow"}
_RTAGS = {v: k for k, v in _TAGS.items()}
# settings the C ABI knows (cipm_settings); the reference's other fields (verbose, direct_solve_method, ...) are
# printing / backend selection and are ignored on load
_SETTING_KEYS = ["max_iter", "time_limit", "max_step_fraction", "tol_gap_abs", "tol_gap_rel", "tol_feas",
                 "tol_infeas_abs", "tol_infeas_rel", "tol_ktratio", "reduced_tol_gap_abs", "reduced_tol_gap_rel",
                 "reduced_tol_feas", "reduced_tol_infeas_abs", "reduced_tol_infeas_rel", "reduced_tol_ktratio",
                 "equilibrate_enable", "equilibrate_max_iter", "equilibrate_min_scaling", "equilibrate_max_scaling",
                 "min_terminate_step_length", "static_regularization_enable", "static_regularization_constant",
                 "static_regularization_proportional", "dynamic_regularization_enable",
                 "dynamic_regularization_eps", "dynamic_regularization_delta", "iterative_refinement_enable",
                 "iterative_refinement_reltol", "iterative_refinement_abstol", "iterative_refinement_max_iter",
                 "iterative_refinement_stop_ratio", "linesearch_backtrack_step", "min_switch_step_length"]


def _csc_from(d):
    return sp.csc_matrix((np.asarray(d["nzval"], dtype=float), np.asarray(d["rowval"], dtype=np.int64),
                          np.asarray(d["colptr"], dtype=np.int64)), shape=(int(d["m"]), int(d["n"])))


def _csc_to(M):
    M = sp.csc_matrix(M)
    M.sort_indices()
    return {"m": int(M.shape[0]), "n": int(M.shape[1]), "colptr": M.indptr.astype(int).tolist(),
            "rowval": M.indices.astype(int).tolist(), "nzval": M.data.astype(float).tolist()}


def _cone_from(c):
    if isinstance(c, str):                      # unit variant written as a bare string
        tag, val = c, None
    else:
        (tag, val), = c.items()
    if tag not in _TAGS:
        raise ValueError(f"unknown cone tag {tag!r}")
    kind = _TAGS[tag]
    if kind == "exp":
        return (kind, 3)
    if kind == "pow":
        return (kind, float(val))
    if kind == "genpow":                        # serde tuple variant: [[alpha...], dim2]
        return (kind, ([float(a) for a in val[0]], int(val[1])))
    return (kind, int(val))


def _cone_to(kind, val):
    if kind == "exp":
        return {_RTAGS[kind]: []}
    if kind == "pow":
        return {_RTAGS[kind]: float(val)}
    if kind == "genpow":
        return {_RTAGS[kind]: [[float(a) for a in val[0]], int(val[1])]}
    return {_RTAGS[kind]: int(val)}


def load_problem(path):
    """-> dict(P, q, A, b, cones, settings) with scipy CSC P (as stored: the reference writes the upper triangle),
    A, numpy q, b, cones as the (kind, value) list CudaSolver takes, settings as a plain dict of the fields the
    backend knows."""
    with open(path) as f:
        d = json.load(f)
    st = {k: v for k, v in d.get("settings", {}).items() if k in _SETTING_KEYS and v is not None}
    for k, v in list(st.items()):
        if isinstance(v, bool):
            st[k] = int(v)
    if st.get("time_limit", 0) >= 1.7976931348623157e308:   # sanitised infinity (json.rs:97-110)
        st["time_limit"] = float("inf")
    return dict(P=_csc_from(d["P"]), q=np.asarray(d["q"], dtype=float), A=_csc_from(d["A"]),
                b=np.asarray(d["b"], dtype=float), cones=[_cone_from(c) for c in d["cones"]], settings=st)


def save_problem(path, P, q, A, b, cones, settings=None):
    """Write a problem in the same format (P is stored as its upper triangle, like the reference does after
    problemdata.rs:79-81)."""
    st = dict(settings or {})
    if st.get("time_limit") == float("inf"):
        st["time_limit"] = 1.7976931348623157e308
    d = {"P": _csc_to(sp.triu(sp.csc_matrix(P), format="csc")), "q": np.asarray(q, dtype=float).tolist(),
         "A": _csc_to(A), "b": np.asarray(b, dtype=float).tolist(),
         "cones": [_cone_to(k, v) for k, v in cones], "settings": st}
    with open(path, "w") as f:
        json.dump(d, f)
