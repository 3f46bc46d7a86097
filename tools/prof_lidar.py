import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvio_fusion_b200 import _capi, backend, synth
lvb = _capi.load(); ctx = backend.Context(lvb)
lf = backend.LidarFeatures(ctx)
sweep = synth.make_lidar_scan()
lf.extract(sweep)
rng = np.random.default_rng(0)
n = 5000; counts = np.full(n, 10); first = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32); m = int(first[-1])
smp = np.concatenate([np.full((m, 1), 0.01), rng.normal(0, 1, (m, 3)) + [0, 0, 9.81], rng.normal(0, 0.2, (m, 3))], axis=1)
backend.preintegrate(ctx, first, smp, rng.normal(0, 1, (n, 3)) + [0, 0, 9.81], rng.normal(0, 0.2, (n, 3)), np.zeros((n, 3)), np.zeros((n, 3)), np.array(synth.IMU_NOISE))
