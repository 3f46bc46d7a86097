// include/lvio_b200/factors.h -- the reference's cost-function classes on the hot path, re-expressed as
// device factor records.  Same class names and Create() argument order as
// /root/reference/src/lvio_fusion/include/lvio_fusion/ceres/{visual_error,imu_error,pose_error}.hpp so that
// Backend::BuildProblem (src/backend.cpp:96-183) and imu::FullBA (src/tools.cpp:92-171) read unchanged.
//
// The value types are duck-typed: anything with .data() returning double* works (Eigen::Vector2d / Vector3d,
// Sophus::SE3d, Eigen::Quaterniond::coeffs(), or the PODs in types.h when Eigen is not available).  A camera is
// any pointer-like object with fx, fy, cx, cy and extrinsic.data() (include/lvio_fusion/visual/camera.h:79-80);
// the rig actually used on the device is the one registered with lvb::Runtime::set_cameras (Camera::Get(0/1)).
#pragma once
#include <type_traits>
#include "ceres_shim.h"
#include "ceres_autodiff.h"

namespace lvio_fusion {
// The classes below carry the reference's names but are different types from the functors in the reference's ceres/*.hpp,
// and translation units that stay on those headers (pose_graph.cpp, navsat.cpp, relocator.cpp: off-path solves on the host
// LM) end up in the same binary as the ones switched to this header.  The inline namespace keeps the two sets apart for
// the linker (lvio_fusion::lvb_device::PoseGraphError vs lvio_fusion::PoseGraphError) while unqualified and
// lvio_fusion::-qualified lookup in a translation unit that includes only this header still finds them.
inline namespace lvb_device {

namespace detail {
template <class V> inline const double* ptr(const V& v) { return v.data(); }
// ceres::QuaternionRotatePoint on a stored (x, y, z, w) quaternion: normalises first (base.hpp:26-31)
inline void rotate_xyzw(const double* q, const double* p, double* out) {
    const double s = 1.0 / std::sqrt(q[3] * q[3] + q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    const double x = q[0] * s, y = q[1] * s, z = q[2] * s, w = q[3] * s;
    double u0 = y * p[2] - z * p[1], u1 = z * p[0] - x * p[2], u2 = x * p[1] - y * p[0];
    u0 += u0; u1 += u1; u2 += u2;
    out[0] = p[0] + w * u0 + (y * u2 - z * u1); out[1] = p[1] + w * u1 + (z * u0 - x * u2); out[2] = p[2] + w * u2 + (x * u1 - y * u0);
}
// SE3Inverse + SE3TransformPoint (base.hpp:40-55,70-77): p in the frame of T = [q | t]
inline void inverse_transform(const double* T, const double* p, double* out) {
    const double qi[4] = {-T[0], -T[1], -T[2], T[3]}, nt[3] = {-T[4], -T[5], -T[6]};
    double ti[3], r[3];
    rotate_xyzw(qi, nt, ti); rotate_xyzw(qi, p, r);
    out[0] = r[0] + ti[0]; out[1] = r[1] + ti[1]; out[2] = r[2] + ti[2];
}
}  // namespace detail

// visual_error.hpp:48-76   AutoDiffCostFunction<PoseOnlyReprojectionError, 2, 7>
class PoseOnlyReprojectionError {
public:
    // The functor itself, as compute_reprojection_error builds and calls it on the host for the outlier test after the
    // solve (backend.cpp:185-190,232: one scalar evaluation per feature; the batched form is lvb_ba_reprojection_errors).
    template <class V2, class V3, class CameraPtr>
    PoseOnlyReprojectionError(const V2& ob, const V3& pw, CameraPtr camera, double weight) : weight_(weight) {
        const double* o = detail::ptr(ob); const double* p = detail::ptr(pw); const double* e = camera->extrinsic.data();
        ob_[0] = o[0]; ob_[1] = o[1]; pw_[0] = p[0]; pw_[1] = p[1]; pw_[2] = p[2];
        for (int i = 0; i < 7; ++i) ext_[i] = e[i];
        fx_ = camera->fx; fy_ = camera->fy; cx_ = camera->cx; cy_ = camera->cy;
    }
    bool operator()(const double* Twc, double* residuals) const {       // Reprojection(), visual_error.hpp:10-24
        double pb[3], pc[3];
        detail::inverse_transform(Twc, pw_, pb); detail::inverse_transform(ext_, pb, pc);
        residuals[0] = weight_ * (fx_ * (pc[0] / pc[2]) + cx_ - ob_[0]);
        residuals[1] = weight_ * (fy_ * (pc[1] / pc[2]) + cy_ - ob_[1]);
        return true;
    }
    template <class V2, class V3, class CameraPtr>
    static ceres::CostFunction* Create(const V2& ob, const V3& pw, CameraPtr /*camera == Camera::Get(0)*/, double weight) {
        const double* o = detail::ptr(ob); const double* p = detail::ptr(pw);
        return new lvb::DeviceCost(LVB_POSE_ONLY, 2, {7}, {o[0], o[1], p[0], p[1], p[2], weight});
    }
private:
    double ob_[2], pw_[3], ext_[7], fx_, fy_, cx_, cy_, weight_;
};

// visual_error.hpp:78-107  AutoDiffCostFunction<TwoFrameReprojectionError, 2, 1, 7, 7>
class TwoFrameReprojectionError {
public:
    template <class V2, class CameraPtr>
    static ceres::CostFunction* Create(const V2& first_ob, const V2& ob, CameraPtr /*left*/, CameraPtr /*right*/, double weight) {
        const double* f = detail::ptr(first_ob); const double* o = detail::ptr(ob);
        return new lvb::DeviceCost(LVB_TWO_FRAME, 2, {1, 7, 7}, {f[0], f[1], o[0], o[1], weight});
    }
};

// visual_error.hpp:109-137 AutoDiffCostFunction<TwoCameraReprojectionError, 2, 1>
class TwoCameraReprojectionError {
public:
    template <class V2, class CameraPtr>
    static ceres::CostFunction* Create(const V2& left_ob, const V2& right_ob, CameraPtr /*left*/, CameraPtr /*right*/, double weight) {
        const double* l = detail::ptr(left_ob); const double* r = detail::ptr(right_ob);
        return new lvb::DeviceCost(LVB_TWO_CAMERA, 2, {1}, {l[0], l[1], r[0], r[1], weight});
    }
};

// imu_error.hpp:12-122     SizedCostFunction<15, 7,3,3,3, 7,3,3,3>
// `preintegration` is pointer-like with the members of imu::Preintegration (imu/preintegration.h:66-80):
// delta_p, delta_v, linearized_ba, linearized_bg (.data() -> 3 doubles), delta_q (.coeffs().data() -> xyzw),
// sum_dt, jacobian / covariance (.data() -> 225 doubles; Eigen is column-major, hence the transpose flag).
class ImuError {
public:
    template <class PreintegrationPtr>
    static ceres::CostFunction* Create(PreintegrationPtr pre, bool matrices_are_column_major = true) {
        std::vector<double> c(469);
        c[467] = c[468] = -1.0;
        const double* dp = pre->delta_p.data(); const double* dq = pre->delta_q.coeffs().data(); const double* dv = pre->delta_v.data();
        const double* ba = pre->linearized_ba.data(); const double* bg = pre->linearized_bg.data();
        for (int i = 0; i < 3; ++i) { c[i] = dp[i]; c[7 + i] = dv[i]; c[10 + i] = ba[i]; c[13 + i] = bg[i]; }
        for (int i = 0; i < 4; ++i) c[3 + i] = dq[i];
        c[16] = pre->sum_dt;
        const double* J = pre->jacobian.data(); const double* C = pre->covariance.data();
        for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) {
            const int src = matrices_are_column_major ? j * 15 + i : i * 15 + j;
            c[17 + i * 15 + j] = J[src]; c[242 + i * 15 + j] = C[src];
        }
        return new lvb::DeviceCost(LVB_IMU, 15, {7, 3, 3, 3, 7, 3, 3, 3}, std::move(c));
    }
};

// imu_error.hpp:124-229    SizedCostFunction<15, 7,3,3,3, 7,3>  (imu::FullBA, tools.cpp:138-162: ba / bg are one shared block)
class ImuInitError {
public:
    template <class PreintegrationPtr>
    static ceres::CostFunction* Create(PreintegrationPtr pre, double prior_a, double prior_g, bool matrices_are_column_major = true) {
        lvb::DeviceCost* base = static_cast<lvb::DeviceCost*>(ImuError::Create(pre, matrices_are_column_major));
        std::vector<double> c = base->consts();
        delete base;
        c[467] = prior_a; c[468] = prior_g;
        return new lvb::DeviceCost(LVB_IMU, 15, {7, 3, 3, 3, 7, 3}, std::move(c));
    }
};

// imu_error.hpp:231-274    NumericDiffCostFunction<ImuInitGError, FORWARD, 15, 3, 3, 3, 3, 4>: the gravity-direction factor of
// imu::InertialOptimization (tools.cpp:35-90) -- velocities, one shared bias pair and the quaternion Rwg, poses fixed.  An
// off-path problem (3 + 3 + 4 + 3 N unknowns, DENSE_QR): evaluated on the host and solved by the host LM.  The raw residual is the
// caller's own Preintegration::Evaluate(..., Rg) (preintegration.cpp:167-188); the whitening is the ImuInitError one.
template <class PreintegrationPtr, class SE3>
class ImuInitGFunctor {
public:
    ImuInitGFunctor(PreintegrationPtr pre, const SE3& current_pose, const SE3& last_pose, double prior_a, double prior_g)
        : pre_(pre), current_pose_(current_pose), last_pose_(last_pose) {
        double cov[225];
        const double* C = pre->covariance.data();                       // Eigen: column-major
        for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) cov[i * 15 + j] = C[j * 15 + i];
        ok_ = lvb::host::imu_sqrt_information(cov, prior_a, prior_g, U_);
    }
    bool operator()(const double* vi, const double* bai, const double* bgi, const double* vj, const double* rg, double* residuals) const {
        if (!ok_) return false;
        typedef typename std::decay<decltype(pre_->delta_p)>::type V3;
        typedef typename std::decay<decltype(pre_->delta_q)>::type Q;
        const Q Qi(last_pose_.rotationMatrix()), Qj(current_pose_.rotationMatrix());
        const V3 Pi = last_pose_.translation(), Pj = current_pose_.translation();
        const Q Rg(rg[3], rg[0], rg[1], rg[2]);
        const auto raw = pre_->Evaluate(Pi, Qi, V3(vi[0], vi[1], vi[2]), V3(bai[0], bai[1], bai[2]), V3(bgi[0], bgi[1], bgi[2]),
                                        Pj, Qj, V3(vj[0], vj[1], vj[2]), V3(0, 0, 0), V3(0, 0, 0), Rg);
        const double* r = raw.data();
        for (int i = 0; i < 15; ++i) { double s = 0; for (int k = 0; k < 15; ++k) s += U_[i * 15 + k] * r[k]; residuals[i] = s; }
        return true;
    }
private:
    PreintegrationPtr pre_;
    SE3 current_pose_, last_pose_;
    double U_[225];
    bool ok_;
};
class ImuInitGError {
public:
    template <class PreintegrationPtr, class SE3>
    static ceres::CostFunction* Create(PreintegrationPtr pre, const SE3& current_pose, const SE3& last_pose, double prior_a, double prior_g) {
        typedef ImuInitGFunctor<PreintegrationPtr, SE3> F;
        return new ceres::NumericDiffCostFunction<F, ceres::FORWARD, 15, 3, 3, 3, 3, 4>(new F(pre, current_pose, last_pose, prior_a, prior_g));
    }
};

// pose_error.hpp:10-53     AutoDiffCostFunction<PoseGraphError, 6, 7, 7>; rpyxyz_ = SE3ToRpyxyz(last^-1 * pose)
class PoseGraphError {
public:
    template <class SE3>
    static ceres::CostFunction* Create(const SE3& last_pose, const SE3& pose, double weight = 1, double v = 1) {
        double e[6]; relative_rpyxyz(detail::ptr(last_pose), detail::ptr(pose), e);
        return make(e, weight, v);
    }
    // second constructor of the reference (pose_error.hpp:20-23, navsat.cpp:300): the relative pose is given
    template <class SE3>
    static ceres::CostFunction* Create(const SE3& relative_i_j, double weight = 1, double v = 1) {
        double e[6]; lvb::host::se3_to_rpyxyz(detail::ptr(relative_i_j), e);
        return make(e, weight, v);
    }
    static ceres::CostFunction* make(const double* e, double weight, double v) {
        lvb::DeviceCost* c = new lvb::DeviceCost(LVB_POSE_GRAPH, 6, {7, 7}, {e[0], e[1], e[2], e[3], e[4], e[5], weight, v});
        lvb::host::PoseGraphFunctor* f = new lvb::host::PoseGraphFunctor();      // host evaluator: only the off-path solves use it
        for (int i = 0; i < 6; ++i) f->e[i] = e[i];
        f->w = weight; f->v = v;
        c->set_host_evaluator(new ceres::AutoDiffCostFunction<lvb::host::PoseGraphFunctor, 6, 7, 7>(f));
        return c;
    }
    // base.hpp:40-55,70-77,94-141 on doubles (unit quaternions assumed for the stored poses)
    static void relative_rpyxyz(const double* a, const double* b, double* e);
};

// pose_error.hpp:55-86     AutoDiffCostFunction<PoseError, 6, 7>
class PoseError {
public:
    template <class SE3>
    static ceres::CostFunction* Create(const SE3& pose, double weight = 1, double v = 1) {
        const double* p = detail::ptr(pose);
        lvb::DeviceCost* c = new lvb::DeviceCost(LVB_POSE_PRIOR, 6, {7}, {p[0], p[1], p[2], p[3], p[4], p[5], p[6], weight, v});
        lvb::host::PoseFunctor* f = new lvb::host::PoseFunctor();
        for (int i = 0; i < 7; ++i) f->pose[i] = p[i];
        f->w = weight; f->v = v;
        c->set_host_evaluator(new ceres::AutoDiffCostFunction<lvb::host::PoseFunctor, 6, 7>(f));
        return c;
    }
};

inline void PoseGraphError::relative_rpyxyz(const double* a, const double* b, double* e) {
    auto rot = detail::rotate_xyzw;
    const double qi[4] = {-a[0], -a[1], -a[2], a[3]}, nt[3] = {-a[4], -a[5], -a[6]};
    double ti[3]; rot(qi, nt, ti);
    // q = qi (x) qb (Hamilton, stored xyzw), t = R(qi) tb + ti
    const double zw = qi[3], zx = qi[0], zy = qi[1], zz = qi[2], ww = b[3], wx = b[0], wy = b[1], wz = b[2];
    const double q0 = zw * ww - zx * wx - zy * wy - zz * wz, q1 = zw * wx + zx * ww + zy * wz - zz * wy;
    const double q2 = zw * wy - zx * wz + zy * ww + zz * wx, q3 = zw * wz + zx * wy - zy * wx + zz * ww;
    double t[3]; rot(qi, b + 4, t);
    e[0] = std::atan2(2 * (q1 * q2 + q0 * q3), 1 - 2 * (q2 * q2 + q3 * q3));
    e[1] = std::asin(2 * (q0 * q2 - q1 * q3));
    e[2] = std::atan2(2 * (q2 * q3 + q0 * q1), 1 - 2 * (q1 * q1 + q2 * q2));
    e[3] = t[0] + ti[0]; e[4] = t[1] + ti[1]; e[5] = t[2] + ti[2];
}

}  // inline namespace lvb_device
}  // namespace lvio_fusion
