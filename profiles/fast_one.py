import importlib.util, sys
import numpy as np
sys.path.insert(0, "tests")
import harness as H
spec = importlib.util.spec_from_file_location("vbkkt", "linear-programming-vanderbei_b200/__init__.py")
vb = importlib.util.module_from_spec(spec); spec.loader.exec_module(vb)
lib = vb.load()
name = sys.argv[1]; mode = vb.MODE_FAST if (len(sys.argv) < 3 or sys.argv[2] == "fast") else vb.MODE_STRICT
if name.startswith("mcf"):
    sys.path.insert(0, "."); import bench
    lp, z = bench.mcf_workload(name)
else:
    lp = H.load_fixture(name); z = np.load(H.GOLDEN / "iterates" / f"{name}_it20.npz")
K = H.kkt_for(vb, lib, lp, mode=mode)
for _ in range(2):
    K.factor(z["E"], z["D"])
    K.solve(z["E"], z["D"], z["rhs_y"], z["rhs_x"])
print("done", K.last_passes)
