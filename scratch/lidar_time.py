import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvio_fusion_b200 import _capi, backend, synth
ctx = backend.Context(_capi.load()); lf = backend.LidarFeatures(ctx); sweep = synth.make_lidar_scan()
for _ in range(3): lf.extract(sweep)
t = time.perf_counter()
for _ in range(20): lf.extract(sweep)
print("lidar extract %.3f ms per sweep" % ((time.perf_counter() - t) * 50))
