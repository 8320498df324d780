#!/usr/bin/env python
"""Mine the reference's notebooks for RECORDED outputs of the un-vendored xmps package (test
infrastructure; run in the build container only -- /root/reference does not exist on the GPU box).

xmps (``Map.left_fixed_point`` / ``right_fixed_point``) cannot be run here, and the reference's tests hold
no vectors for it, but three notebook cells kept what it printed:

* ``Time Evo.ipynb`` cell 23: ``Map(A, B).left_fixed_point()[1]`` for a random left-canonical ``A`` and
  ``B = exp(-i Z dt) . A`` -- a 2 x 2 matrix of unit Frobenius norm whose largest entry is real positive
  (LAPACK zgeev's eigenvector convention);
* cell 24: the same for ``Map(merge(A, A), merge(B, B))`` -- identical to 8 digits: the two-site map has
  the same fixed point (its eigenvalue is the square);
* ``scripts/opt.ipynb`` cell 11: ``Map(A, A_).left_fixed_point()`` from another xmps version -- eigenvalue
  returned as a length-1 array (ARPACK ``eigs(k=1)``; the reference indexes ``x[0]`` at
  qmps/loschmidts/time_evo.py:113), vector of norm 1.203 in no recognisable gauge.

The inputs of those cells were random and are not recorded, so the values themselves cannot be reproduced;
what they pin is the normalisation, the phase convention and the merge identity.  Output:
tests/golden/ref_notebook_outputs.json.
"""
import json
import os
import re
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "ref_notebook_outputs.json")


def cell_output(nb_path, index):
    nb = json.load(open(os.path.join(REF, nb_path)))
    cell = nb["cells"][index]
    text = []
    for o in cell.get("outputs", []):
        if "text" in o:
            text.append("".join(o["text"]))
        elif "data" in o and "text/plain" in o["data"]:
            text.append("".join(o["data"]["text/plain"]))
    return "".join(cell["source"]), "\n".join(text)


_NUM = r"[-+]?\d+\.?\d*(?:e[-+]?\d+)?"


def complex_numbers(text):
    """every 'a+bj' literal of a numpy repr, in order"""
    return [complex(float(a), float(b)) for a, b in re.findall(rf"({_NUM})\s*([-+]\s*\d+\.?\d*(?:e[-+]?\d+)?)j", text.replace(" ", ""))]


def main():
    out = {"generated_by": "oracle/make_golden_notebooks.py", "cells": {}}
    src, txt = cell_output("Time Evo.ipynb", 22)
    out["cells"]["time_evo_22_source"] = src
    for idx, key in ((23, "time_evo_23_left_fixed_point"), (24, "time_evo_24_left_fixed_point_merged")):
        src, txt = cell_output("Time Evo.ipynb", idx)
        z = complex_numbers(txt)
        assert len(z) == 4, (idx, txt)
        out["cells"][key] = {"source": src, "re": [v.real for v in z], "im": [v.imag for v in z], "shape": [2, 2]}
    src, txt = cell_output("scripts/opt.ipynb", 11)
    z = complex_numbers(txt)
    assert len(z) == 5, txt
    out["cells"]["opt_11_left_fixed_point"] = {"source": src, "eta_shape": [1], "eta": [z[0].real, z[0].imag],
                                               "re": [v.real for v in z[1:]], "im": [v.imag for v in z[1:]],
                                               "shape": [2, 2]}
    m = np.array(out["cells"]["time_evo_23_left_fixed_point"]["re"]) + 1j * np.array(out["cells"]["time_evo_23_left_fixed_point"]["im"])
    out["derived"] = {"time_evo_23_frobenius_norm": float(np.linalg.norm(m)),
                      "opt_11_frobenius_norm": float(np.linalg.norm(np.array(out["cells"]["opt_11_left_fixed_point"]["re"])
                                                                    + 1j * np.array(out["cells"]["opt_11_left_fixed_point"]["im"])))}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, ensure_ascii=False)
    print("wrote", OUT, out["derived"])


if __name__ == "__main__":
    sys.exit(main())
