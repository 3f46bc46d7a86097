import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvio_fusion_b200 import _capi, backend, synth
lvb = _capi.load(); ctx = backend.Context(lvb)
sc = synth.make_icp_problem(120000, 1000000, kind="surf")
fa = backend.FeatureAssociation(ctx)
e0 = synth.relative_rpyxyz(sc["map_pose"], sc["frame_pose"])
fa.set_map(sc["map"], sc["cell_size"])
fa.scan_to_map(sc["mode"], sc["scan"], sc["frame_pose"], sc["map_pose"], e0, sc["weight"], -1.0, sc["huber_a"], sc["thr"])
