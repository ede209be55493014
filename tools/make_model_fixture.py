"""Regenerates ``assets/pupper_v3.xml`` from the reference's model file.

The reference ships its robot model as ``test/test_pupper_model.xml`` with STL visual meshes.  This repo
carries a canonical re-emission of the compiled model (explicit attributes, no default classes, no mesh
assets -- visual geoms become non-colliding placeholders so geom ids are preserved) produced by
``pupperv3_mjx_b200.mjcf.to_xml``.  Run in the build container (the only place /root/reference exists):

    python tools/make_model_fixture.py [/root/reference/test/test_pupper_model.xml]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pupperv3_mjx_b200 import mjcf  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/test/test_pupper_model.xml"
dst = os.path.join(os.path.dirname(__file__), "..", "assets", "pupper_v3.xml")
m = mjcf.compile_model(src)
xml = mjcf.to_xml(m)
with open(dst, "w") as f:
    f.write(xml + "\n")
m2 = mjcf.compile_model(dst)
import numpy as np  # noqa: E402
for k in ("body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia", "dof_invweight0",
          "body_invweight0", "geom_friction", "sphere_pos", "sphere_radius", "site_pos", "jnt_range",
          "plane_sphere_solimp", "sphere_sphere_solimp", "actuator_gainprm", "actuator_biasprm"):
    assert np.allclose(getattr(m, k), getattr(m2, k), rtol=1e-14, atol=1e-16), k
assert m.geom_names == m2.geom_names and m.site_names == m2.site_names and m.body_names == m2.body_names
print("wrote", os.path.abspath(dst), "ngeom", m2.ngeom)
