import numpy as np, subprocess, tempfile, os, sys
sys.path.insert(0, "/root/repo")
from lvio_fusion_b200 import synth, backend
from oracle import binding
ctx=backend.Context(binding.load()); lf=backend.LidarFeatures(ctx)
def ref_run(scan):
    td=tempfile.mkdtemp()
    with open(td+"/in.bin","wb") as f:
        np.array([64,1800,0.427,24.9,60,5,30],dtype=np.float64).tofile(f); np.array([len(scan)],dtype=np.int32).tofile(f); scan[:,:3].astype(np.float32).tofile(f)
    subprocess.check_call(["/root/repo/oracle/_ref/ref_lidar",td+"/in.bin",td+"/out.bin"])
    raw=open(td+"/out.bin","rb").read()
    m=np.frombuffer(raw[:4],dtype=np.int32)[0]; off=4
    pts=np.frombuffer(raw[off:off+16*m],dtype=np.float32).reshape(m,4); off+=16*m
    rng=np.frombuffer(raw[off:off+4*m],dtype=np.float32); off+=4*m
    gnd=np.frombuffer(raw[off:off+m],dtype=np.uint8); off+=m
    col=np.frombuffer(raw[off:off+4*m],dtype=np.int32); off+=4*m
    sr=np.frombuffer(raw[off:off+256],dtype=np.int32); off+=256
    er=np.frombuffer(raw[off:off+256],dtype=np.int32); off+=256
    return pts,rng,gnd,col,sr,er
for seed in (11,12,13,14):
    scan=synth.make_lidar_scan(seed=seed)
    pts,rng,gnd,col,sr,er=ref_run(scan); s=lf.segment(scan)
    same=len(pts)==len(s["points"])
    print(seed, len(pts), len(s["points"]), "xyz", same and np.array_equal(pts[:,:3],s["points"][:,:3]), "range", same and np.array_equal(rng,s["range"]), "ground", same and np.array_equal(gnd,s["ground"]), "col", same and np.array_equal(col,s["col"]), "rings", np.array_equal(sr,s["start_ring"]) and np.array_equal(er,s["end_ring"]))
