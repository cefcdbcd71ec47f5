"""Generate judo_b200/models/<task>.json from the reference MJCF files (run in the authoring container).

The GPU box has no /root/reference, so the constant tables are committed.  Re-run after changing
judo_b200/mjcf.py:   python tools/gen_model_tables.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from judo_b200.mjcf import compile_mjcf  # noqa: E402

XML = "/root/reference/judo/models/xml"
OUT = os.path.join(os.path.dirname(__file__), "..", "judo_b200", "models")

for task, fname in [("cartpole", "cartpole.xml"), ("cylinder_push", "cylinder_push.xml"), ("leap_cube", "leap_cube.xml"),
                    ("leap_cube_down", "leap_cube_palm_down.xml"), ("fr3_pick", "fr3_pick.xml")]:
    m = compile_mjcf(os.path.join(XML, fname))
    with open(os.path.join(OUT, f"{task}.json"), "w") as f:
        json.dump(m, f, indent=1)
    print(task, {k: m[k] for k in ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "nsensordata", "meaninertia")})
