#!/usr/bin/env python
"""Derive the constant model tables shipped under smplsim_b200/assets/ from the
reference's MJCF files (run in the build container, where /root/reference exists).

    python tools/extract_model.py [/root/reference]

smpl  <- smpl_sim/data/assets/mjcf/smpl_humanoid.xml   (24 bodies; HumanoidEnv fallback model,
         smpl_sim/envs/humanoid_env.py:249-254)
smplx <- smpl_humanoid.xml at the reference root       (52 bodies)
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from smplsim_b200 import mjcf, model  # noqa: E402

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
for name, rel in (("smpl_humanoid", "smpl_sim/data/assets/mjcf/smpl_humanoid.xml"), ("smplx_humanoid", "smpl_humanoid.xml")):
    with open(os.path.join(ref, rel)) as f:
        p = mjcf.parse_mjcf(f.read())
    model.save_asset(p, name)
    m = model.build_model(p)
    print(f"{name}: nbody={m.nbody} nq={m.nq} nv={m.nv} nu={m.nu} ngeom={m.ngeom} mass={m.total_mass:.4f} kg")
