"""Regenerates tests/golden/*.json from the CPU oracle (no reference import possible: MadNLP is Julia and there is no
Julia in this image).  The 2x2 system + solution are copied from the reference's own test
(lib/MadNLPTests/src/MadNLPTests.jl:24-51); the HS15 vectors restate test_kkt_system (MadNLPTests.jl:53-110) on
lib/MadNLPTests/src/Instances/hs15.jl with LAPACK dsytrf/dsytrs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import madnlp_oracle as o  # noqa: E402

out = {"kat_2x2": {"row": [0, 1, 1], "col": [0, 0, 1], "val": [1.0, 0.1, 2.0], "b": [1.0, 3.0],
                   "x": [0.8542713567839195, 1.4572864321608041], "inertia": [2, 0, 0],
                   "source": "lib/MadNLPTests/src/MadNLPTests.jl:24-51"}}
cb = o.HS15Model.callback()
for name, K, dense in (("sparse", o.SparseKKTSystem, False), ("condensed", o.SparseCondensedKKTSystem, False),
                       ("dense_condensed", o.DenseCondensedKKTSystem, True)):
    kkt = K(cb, o.DenseLDLInertiaSolver) if name == "sparse" else K(cb)
    x, y, inertia = o.test_kkt_system(kkt, o.HS15Model, dense=dense)
    entry = {"solve_kkt_of_ones": x.full().tolist(), "K_times_x": y.full().tolist(), "inertia": list(inertia)}
    if name != "dense_condensed":
        entry.update(colptr=kkt.aug_colptr.tolist(), rowval=kkt.aug_rowval.tolist(), nzval=kkt.aug_nz.tolist())
    else:
        entry.update(aug_lower=np.tril(kkt.aug_com).tolist())
    out["hs15_" + name] = entry
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hs15_kkt.json"), "w"), indent=1)
print("written")
