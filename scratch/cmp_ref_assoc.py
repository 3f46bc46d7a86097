import numpy as np, subprocess, tempfile, os, sys
sys.path.insert(0, "/root/repo")
from lvio_fusion_b200 import synth, backend
from oracle import binding
orc=binding.load(); ctx=backend.Context(orc)
EXT=[0.01,-0.02,0.005,1.0,0.27,0.0,0.08]
def cfg_bytes(W):
    return np.array([64,W,0.427,24.9,60,0.1036,5,30,0.2,0],dtype=np.float64).tobytes()+np.array(EXT,dtype=np.float64).tobytes()
def get_cloud(raw,off):
    m=int(np.frombuffer(raw[off:off+4],dtype=np.int32)[0]); off+=4
    c=np.frombuffer(raw[off:off+16*m],dtype=np.float32).reshape(m,4).copy(); off+=16*m
    return c,off
for seed in (11,12):
    scan=synth.make_lidar_scan(seed=seed)
    td=tempfile.mkdtemp()
    with open(td+"/in.bin","wb") as f:
        f.write(cfg_bytes(1800)); np.array([len(scan)],dtype=np.int32).tofile(f); scan[:,:3].astype(np.float32).tofile(f)
    subprocess.check_call(["/root/repo/oracle/_ref/ref_assoc","extract",td+"/in.bin",td+"/out.bin"])
    raw=open(td+"/out.bin","rb").read()
    seg,off=get_cloud(raw,0); m=len(seg)
    curv=np.frombuffer(raw[off:off+4*m],dtype=np.float32).copy(); off+=4*m
    ground,off=get_cloud(raw,off); surf,off=get_cloud(raw,off)
    lf=backend.LidarFeatures(ctx, extrinsic=EXT)
    s=lf.segment(scan); g,sf=lf.extract(scan)
    print(seed,"seg", m, len(s["points"]), "points(with time)", np.array_equal(seg,s["points"]), "max dI", np.abs(seg[:,3]-s["points"][:,3]).max() if m==len(s["points"]) else None,
          "curv[5:-5]", np.array_equal(curv[5:-5], s["curvature"][5:-5]), "ground", ground.shape, g.shape, np.array_equal(ground,g), "surf", surf.shape, sf.shape, np.array_equal(surf,sf))

for kind,mode in (("ground",0),("surf",1)):
    sc=synth.make_icp_problem(1500, 12000, seed=7+mode, kind=kind)
    e0=synth.relative_rpyxyz(sc["map_pose"], sc["frame_pose"])
    scan=np.ascontiguousarray(sc["scan"][:,:4],dtype=np.float32); mp=np.ascontiguousarray(sc["map"][:,:4],dtype=np.float32)
    td=tempfile.mkdtemp()
    wv,wg,ws=71.8856, sc["weight"] if mode==0 else 1.0, sc["weight"] if mode==1 else 0.01
    with open(td+"/in.bin","wb") as f:
        f.write(np.array([64,1800,0.427,24.9,60,0.1036,5,30,0.2,0],dtype=np.float64).tobytes()+np.array([0,0,0,1,0,0,0],dtype=np.float64).tobytes())
        np.array([mode],dtype=np.int32).tofile(f); np.asarray(sc["frame_pose"],dtype=np.float64).tofile(f); np.asarray(sc["map_pose"],dtype=np.float64).tofile(f); np.asarray(e0,dtype=np.float64).tofile(f)
        np.array([wv,wg,ws],dtype=np.float64).tofile(f); np.array([120,0],dtype=np.int32).tofile(f)
        np.array([len(scan)],dtype=np.int32).tofile(f); scan.tofile(f); np.array([len(mp)],dtype=np.int32).tofile(f); mp.tofile(f)
    subprocess.check_call(["/root/repo/oracle/_ref/ref_assoc","scan2map",td+"/in.bin",td+"/out.bin"])
    o=np.fromfile(td+"/out.bin",dtype=np.float64); head=o[:4]; tab=o[4:].reshape(len(scan),5)
    fa=backend.FeatureAssociation(ctx); orc.icp_set_brute(fa.h,1); fa.set_map(sc["map"], sc["cell_size"])
    acc,r,J=fa.evaluate(sc["mode"], sc["scan"], sc["frame_pose"], sc["map_pose"], e0, sc["weight"], sc["thr"])
    a_ref=tab[:,0]>0
    print(kind, "blocks", head, "acc equal", np.array_equal(a_ref, acc.astype(bool)), acc.sum(), "thr", sc["thr"], "r", np.abs(tab[a_ref,1]-r[acc.astype(bool)]).max(), "J", np.abs(tab[a_ref,2:]-J[acc.astype(bool)]).max(), "prior", head[2], 120*wv, "huber", head[3], sc["huber_a"])
