// oracle/factors.h -- CPU restatement of the reference's autodiff cost functors
// (TEST INFRASTRUCTURE ONLY).  The reference has no golden vectors; the visual, lidar-plane and pose factors of this file
// are pinned against the reference's own functor headers compiled in place (oracle/ref_harness.cpp, oracle/ref_compat,
// tests/golden/ref_factors.npz, tests/test_ref_golden.py); everything else in oracle/ stays "parity unpinned".
//
// Every functor below is templated on the scalar so that it can be run on doubles (the
// residual) or oracle::Dual<N> (the Jacobian an AutoDiffCostFunction would return).
//
//   visual : /root/reference/src/lvio_fusion/include/lvio_fusion/ceres/visual_error.hpp:10-137
//   lidar  : /root/reference/src/lvio_fusion/include/lvio_fusion/ceres/lidar_error.hpp:10-110
//   priors : /root/reference/src/lvio_fusion/include/lvio_fusion/ceres/pose_error.hpp:10-86,135-190
#pragma once
#include "geometry.h"

namespace oracle {

// One pinhole camera: intrinsics + T_body_cam (sensor.h:21-24 / camera.h:79-80).
// ABI layout (11 doubles): fx fy cx cy  qx qy qz qw tx ty tz
struct Camera {
    double fx, fy, cx, cy;
    double ext[7];
};
inline Camera load_camera(const double* c) {
    Camera k; k.fx = c[0]; k.fy = c[1]; k.cx = c[2]; k.cy = c[3];
    for (int i = 0; i < 7; ++i) k.ext[i] = c[4 + i];
    return k;
}

template <class T> struct Pix { T u, v; };

// visual_error.hpp:10-23 Reprojection: world point -> pixel of `cam` mounted on body pose Twc
template <class T> inline Pix<T> reproject(const Vec3<T>& pw, const Rigid<T>& Twc, const Camera& cam) {
    const Vec3<T> p_body = apply(inverse(Twc), pw);
    const Rigid<T> ext = load_rigid<T>(cam.ext);
    const Vec3<T> pc = apply(inverse(ext), p_body);
    const T xp = pc.x / pc.z;
    const T yp = pc.y / pc.z;
    return Pix<T>{xp * cam.fx + cam.cx, yp * cam.fy + cam.cy};
}

// visual_error.hpp:25-33 Pixel2Robot: pixel + inverse depth in `cam` -> body frame
template <class T> inline Vec3<T> pixel_to_body(double u, double v, const T& inv_d, const Camera& cam) {
    const T d = T(1) / inv_d;
    const Vec3<T> ps(T((u - cam.cx) / cam.fx) * d, T((v - cam.cy) / cam.fy) * d, d);
    const Rigid<T> ext = load_rigid<T>(cam.ext);
    return apply(ext, ps);
}

// visual_error.hpp:35-46 Robot2Pixel
template <class T> inline Pix<T> body_to_pixel(const Vec3<T>& pb, const Camera& cam) {
    const Rigid<T> ext = load_rigid<T>(cam.ext);
    const Vec3<T> pc = apply(inverse(ext), pb);
    const T xp = pc.x / pc.z;
    const T yp = pc.y / pc.z;
    return Pix<T>{xp * cam.fx + cam.cx, yp * cam.fy + cam.cy};
}

// ---- a1 TwoFrameReprojectionError  visual_error.hpp:78-107  <2,1,7,7>
// consts[5] = first_ob.x first_ob.y ob.x ob.y weight ;  cam0 = "left_", cam1 = "right_"
// (Create(first_ob, ob, Camera::Get(0), Camera::Get(1), w), backend.cpp:138)
template <class T>
inline void two_frame(const double* c, const Camera& cam0, const Camera& cam1,
                      const T& inv_d, const T* Twc1, const T* Twc2, T* res) {
    const Vec3<T> pb = pixel_to_body<T>(c[0], c[1], inv_d, cam1);
    const Vec3<T> pw = apply(load_rigid<T>(Twc1), pb);
    const Pix<T> px = reproject(pw, load_rigid<T>(Twc2), cam0);
    res[0] = T(c[4]) * (px.u - T(c[2]));
    res[1] = T(c[4]) * (px.v - T(c[3]));
}

// ---- a2 PoseOnlyReprojectionError  visual_error.hpp:48-76  <2,7>
// consts[6] = ob.x ob.y pw.x pw.y pw.z weight ; camera = Camera::Get() = cam0 (backend.cpp:129)
template <class T>
inline void pose_only(const double* c, const Camera& cam0, const T* Twc, T* res) {
    const Vec3<T> pw = Vec3<T>(T(c[2]), T(c[3]), T(c[4]));
    const Pix<T> px = reproject(pw, load_rigid<T>(Twc), cam0);
    res[0] = T(c[5]) * (px.u - T(c[0]));
    res[1] = T(c[5]) * (px.v - T(c[1]));
}

// ---- a3 TwoCameraReprojectionError  visual_error.hpp:109-137  <2,1>
// consts[5] = left_ob.x left_ob.y right_ob.x right_ob.y weight   (weight = 5*w_visual, backend.cpp:123)
template <class T>
inline void two_camera(const double* c, const Camera& cam0, const Camera& cam1, const T& inv_d, T* res) {
    const Vec3<T> pb = pixel_to_body<T>(c[2], c[3], inv_d, cam1);
    const Pix<T> px = body_to_pixel(pb, cam0);
    res[0] = T(c[4]) * (px.u - T(c[0]));
    res[1] = T(c[4]) * (px.v - T(c[1]));
}

// ---- a5 LidarPlaneError  lidar_error.hpp:10-40 : ctor normal (no degeneracy guard, NaN if collinear)
inline V3d plane_normal(const V3d& pa, const V3d& pb, const V3d& pc) {
    V3d n = cross(pa - pb, pa - pc);
    const double len = std::sqrt(n.x * n.x + n.y * n.y + n.z * n.z);  // Eigen normalize(): v / norm
    return V3d(n.x / len, n.y / len, n.z / len);
}
template <class T> inline T plane_distance(const V3d& p, const V3d& pa, const V3d& n, const Rigid<T>& Twc2) {
    const Vec3<T> lp = apply(Twc2, Vec3<T>(T(p.x), T(p.y), T(p.z)));
    const Vec3<T> d = lp - Vec3<T>(T(pa.x), T(pa.y), T(pa.z));
    return dot(d, Vec3<T>(T(n.x), T(n.y), T(n.z)));
}

// ---- a5 LidarPlaneErrorRPZ (mode 0) lidar_error.hpp:42-75 / LidarPlaneErrorYXY (mode 1) :77-110
// consts[10] = p(3) pa(3) n(3) weight ; shared: Twc1[7] (map pose), rpyxyz[6] (live array),
// free scalars: mode 0 -> (pitch=rpyxyz[1], roll=[2], z=[5]); mode 1 -> (yaw=[0], x=[3], y=[4])
template <class T>
inline T lidar_plane(const double* c, int mode, const double* Twc1, const double* rpyxyz, const T* free3) {
    T e[6];
    for (int i = 0; i < 6; ++i) e[i] = T(rpyxyz[i]);
    if (mode == 0) { e[1] = free3[0]; e[2] = free3[1]; e[5] = free3[2]; }
    else           { e[0] = free3[0]; e[3] = free3[1]; e[4] = free3[2]; }
    const Rigid<T> rel = from_rpyxyz(e);
    const Rigid<T> Twc2 = compose(load_rigid<T>(Twc1), rel);
    const T d = plane_distance(V3d(c[0], c[1], c[2]), V3d(c[3], c[4], c[5]), V3d(c[6], c[7], c[8]), Twc2);
    return T(c[9]) * d;
}

// ---- a6 PoseErrorRPZ pose_error.hpp:135-162 (residual order roll, pitch, z) / PoseErrorYXY :164-190
// target[3] is captured from the rpyxyz array at construction: mode 0 -> (p_, r_, z_), mode 1 -> (Y_, x_, y_)
template <class T>
inline void icp_prior(int mode, const double* target3, double weight, const T* free3, T* res) {
    if (mode == 0) {
        res[0] = T(weight) * (free3[1] - T(target3[1]));  // roll
        res[1] = T(weight) * (free3[0] - T(target3[0]));  // pitch
        res[2] = T(weight) * (free3[2] - T(target3[2]));  // z
    } else {
        for (int i = 0; i < 3; ++i) res[i] = T(weight) * (free3[i] - T(target3[i]));
    }
}

// ---- a6 PoseGraphError pose_error.hpp:10-53 <6,7,7> ; consts[8] = rpyxyz_(6) weight v
template <class T>
inline void pose_graph(const double* c, const T* Twc1, const T* Twc2, T* res) {
    const Rigid<T> rel = compose(inverse(load_rigid<T>(Twc1)), load_rigid<T>(Twc2));
    T e[6];
    to_rpyxyz(rel, e);
    const double w = c[6], v = c[7];
    res[0] = T(v * w) * (T(c[0]) - e[0]);
    res[1] = T(v * w) * (T(c[1]) - e[1]);
    res[2] = T(v * w) * (T(c[2]) - e[2]);
    res[3] = T(w) * (T(c[3]) - e[3]);
    res[4] = T(10 * w) * (T(c[4]) - e[4]);
    res[5] = T(10 * w) * (T(c[5]) - e[5]);
}
// PoseGraphError ctor: rpyxyz_ = SE3ToRpyxyz(last_pose^-1 * pose) using Sophus (unit quaternion product)
inline void pose_graph_target(const double* last_pose, const double* pose, double* rpyxyz6) {
    // Sophus::SE3d::inverse()/operator*: conj for unit q, t' = -(R^T t); product renormalises
    // only when far from unit [upstream].  With unit inputs this equals compose(inverse()).
    const Rigid<double> rel = compose(inverse(load_rigid<double>(last_pose)), load_rigid<double>(pose));
    to_rpyxyz(rel, rpyxyz6);
}

// ---- a6 PoseError pose_error.hpp:55-86 <6,7> ; consts[9] = pose_(7) weight v
template <class T>
inline void pose_prior(const double* c, const T* pose, T* res) {
    const Rigid<T> rel = compose(inverse(load_rigid<T>(c)), load_rigid<T>(pose));
    T e[6];
    to_rpyxyz(rel, e);
    const double w = c[7], v = c[8];
    res[0] = T(v * w) * e[0];
    res[1] = T(v * w) * e[1];
    res[2] = T(v * w) * e[2];
    res[3] = T(w) * e[3];
    res[4] = T(w) * e[4];
    res[5] = T(w) * e[5];
}

// ------------------------------------------------------------------------------------
// Autodiff wrappers: residual + row-major Jacobian [n_res x sum(block sizes)], columns in
// AddResidualBlock order (what Ceres hands back per block, concatenated).
// ------------------------------------------------------------------------------------
inline void two_frame_eval(const double* c, const Camera& cam0, const Camera& cam1,
                           double rho, const double* T1, const double* T2, double* r, double* J /*2x15*/) {
    typedef Dual<15> D;
    D inv_d = D::seed(rho, 0), a[7], b[7], res[2];
    for (int i = 0; i < 7; ++i) { a[i] = D::seed(T1[i], 1 + i); b[i] = D::seed(T2[i], 8 + i); }
    two_frame<D>(c, cam0, cam1, inv_d, a, b, res);
    for (int k = 0; k < 2; ++k) { r[k] = res[k].v; if (J) for (int i = 0; i < 15; ++i) J[k * 15 + i] = res[k].d[i]; }
}
inline void pose_only_eval(const double* c, const Camera& cam0, const double* Twc, double* r, double* J /*2x7*/) {
    typedef Dual<7> D;
    D a[7], res[2];
    for (int i = 0; i < 7; ++i) a[i] = D::seed(Twc[i], i);
    pose_only<D>(c, cam0, a, res);
    for (int k = 0; k < 2; ++k) { r[k] = res[k].v; if (J) for (int i = 0; i < 7; ++i) J[k * 7 + i] = res[k].d[i]; }
}
inline void two_camera_eval(const double* c, const Camera& cam0, const Camera& cam1, double rho, double* r, double* J /*2x1*/) {
    typedef Dual<1> D;
    D inv_d = D::seed(rho, 0), res[2];
    two_camera<D>(c, cam0, cam1, inv_d, res);
    for (int k = 0; k < 2; ++k) { r[k] = res[k].v; if (J) J[k] = res[k].d[0]; }
}
inline void lidar_plane_eval(const double* c, int mode, const double* Twc1, const double* rpyxyz,
                             const double* free3, double* r, double* J /*1x3*/) {
    typedef Dual<3> D;
    D f[3] = {D::seed(free3[0], 0), D::seed(free3[1], 1), D::seed(free3[2], 2)};
    const D res = lidar_plane<D>(c, mode, Twc1, rpyxyz, f);
    r[0] = res.v; if (J) for (int i = 0; i < 3; ++i) J[i] = res.d[i];
}
inline void pose_graph_eval(const double* c, const double* T1, const double* T2, double* r, double* J /*6x14*/) {
    typedef Dual<14> D;
    D a[7], b[7], res[6];
    for (int i = 0; i < 7; ++i) { a[i] = D::seed(T1[i], i); b[i] = D::seed(T2[i], 7 + i); }
    pose_graph<D>(c, a, b, res);
    for (int k = 0; k < 6; ++k) { r[k] = res[k].v; if (J) for (int i = 0; i < 14; ++i) J[k * 14 + i] = res[k].d[i]; }
}
inline void pose_prior_eval(const double* c, const double* pose, double* r, double* J /*6x7*/) {
    typedef Dual<7> D;
    D a[7], res[6];
    for (int i = 0; i < 7; ++i) a[i] = D::seed(pose[i], i);
    pose_prior<D>(c, a, res);
    for (int k = 0; k < 6; ++k) { r[k] = res[k].v; if (J) for (int i = 0; i < 7; ++i) J[k * 7 + i] = res[k].d[i]; }
}

}  // namespace oracle
