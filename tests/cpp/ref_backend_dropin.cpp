// tests/cpp/ref_backend_dropin.cpp -- TEST INFRASTRUCTURE ONLY: drop-in proof for SURVEY 8(a1/a2/a9), built in the
// development container (the reference tree is needed to compile it), prebuilt binaries travel in oracle/_ref/.
//
// The REFERENCE's own translation units src/backend.cpp, src/tools.cpp, src/frame.cpp, src/landmark.cpp, src/map.cpp,
// src/manager.cpp and src/preintegration.cpp are compiled where they lie (third-party headers replaced by the stand-ins of
// oracle/ref_compat + tests/cpp/compat_backend/shared) and linked with this file, which only plays the frontend: it
// fills the reference's Map with a synthetic stereo + IMU window through the reference's own classes (Frame, Feature,
// Landmark, Camera, Imu, Preintegration) and then calls the reference's Backend::BuildProblem / adapt::Solve /
// imu::RecoverBias (the body of Backend::Optimize, src/backend.cpp:192-216), imu::FullBA (src/tools.cpp:92-171) and
// compute_reprojection_error (src/backend.cpp:185-190).
//
// Three builds of the same source (tests/test_zz_ref_backend_dropin.py):
//   -DDROPIN_PRODUCT, lvb_* bound to liblvio_b200.so   the reference drives the CUDA path through include/lvio_b200/*.h
//   -DDROPIN_PRODUCT, lvb_* renamed to orc_*           same shim, the CPU oracle behind the C ABI (CPU check of the host side)
//   (neither)                                          recording ceres + the reference's OWN factor headers: what the
//                                                      reference's functors say the cost of that problem is
#include "lvio_fusion/common.h"
#include <cstdio>
#include <cstring>
#define private public
#include "lvio_fusion/backend.h"
#include "lvio_fusion/adapt/environment.h"
#undef private
#include "lvio_fusion/imu/tools.h"
#include "lvio_fusion/manager.h"
#include "lvio_fusion/map.h"
#include "lvio_fusion/utility.h"
#include "lvio_fusion/visual/feature.h"
#include "lvio_fusion/visual/landmark.h"

const double epsilon = 1e-3;          // src/estimator.cpp:9-10
const int num_threads = 1;

namespace lvio_fusion {
double compute_reprojection_error(Vector2d ob, Vector3d pw, SE3d pose, Camera::Ptr camera);      // src/backend.cpp:185
// declared by the headers above, defined in translation units the harness does not link (never reached)
Matrix3d normalize_R(const Matrix3d&) { std::abort(); }
// src/utility.cpp (needs OpenCV as a whole): the rotation that takes the z axis onto `vec`, about z x vec -- what
// Initializer::EstimateVelAndRwg / Initialize ask for (initializer.cpp:27,42).  Rodrigues' formula.
Matrix3d get_R_from_vector(Vector3d vec) {
    vec.normalize();
    Vector3d axis = Vector3d::UnitZ().cross(vec);
    const double s = axis.norm(), c = vec.z();
    if (s < 1e-12) return Matrix3d::Identity();
    axis = axis / s;
    const double ang = std::atan2(s, c);
    Matrix3d K; K << 0, -axis.z(), axis.y(), axis.z(), 0, -axis.x(), -axis.y(), axis.x(), 0;
    return Matrix3d(Matrix3d::Identity() + std::sin(ang) * K + (1 - std::cos(ang)) * (K * K));
}
}  // namespace lvio_fusion
#ifndef DROPIN_PRODUCT
namespace ceres { void Solve(const Solver::Options&, Problem*, Solver::Summary*) {} }
#endif

using namespace lvio_fusion;

static unsigned g_seed = 20240917u;
static double urand() { g_seed = g_seed * 1664525u + 1013904223u; return (double)(g_seed >> 8) / 16777216.0; }
static double nrand() { double s = 0; for (int i = 0; i < 12; ++i) s += urand(); return s - 6.0; }

// ---- a smooth trajectory: body x forward, gentle yaw
static Vector3d traj_p(double t) { return Vector3d(2.0 * t, 0.5 * std::sin(0.4 * t), 0.1 * std::cos(0.3 * t)); }
static Vector3d traj_v(double t) { return Vector3d(2.0, 0.2 * std::cos(0.4 * t), -0.03 * std::sin(0.3 * t)); }
static Vector3d traj_a(double t) { return Vector3d(0.0, -0.08 * std::sin(0.4 * t), -0.009 * std::cos(0.3 * t)); }
static double traj_yaw(double t) { return 0.1 * std::sin(0.3 * t); }
static double traj_yaw_rate(double t) { return 0.03 * std::cos(0.3 * t); }
static Quaterniond yaw_q(double y) { return Quaterniond(std::cos(0.5 * y), 0, 0, std::sin(0.5 * y)); }
static SE3d traj_pose(double t) { return SE3d(yaw_q(traj_yaw(t)), traj_p(t)); }

static SE3d perturb(const SE3d& T, double rot, double trans) {
    Quaterniond dq(1, rot * nrand(), rot * nrand(), rot * nrand()); dq.normalize();
    return SE3d(T.unit_quaternion() * dq, Vector3d(T.translation() + Vector3d(trans * nrand(), trans * nrand(), trans * nrand())));
}

#ifndef DROPIN_PRODUCT
// cost of a recorded problem, evaluated by the reference's own functors: sum over blocks of rho(|r|^2) / 2,
// rho = Huber(a) as ceres::HuberLoss defines it (s <= a^2 ? s : 2 a sqrt(s) - a^2)
static double recorded_cost(const ceres::Problem& problem) {
    double total = 0;
    for (const ceres::ResidualBlock* rb : problem.residual_blocks) {
        std::vector<const double*> p(rb->blocks.begin(), rb->blocks.end());
        double r[32];
        const int n = rb->cost->num_residuals();
        if (n <= 0 || n > 32 || !rb->cost->Evaluate(p.data(), r, nullptr)) { fprintf(stderr, "cannot evaluate a block\n"); std::abort(); }
        double s = 0; for (int i = 0; i < n; ++i) s += r[i] * r[i];
        const double a = rb->loss ? rb->loss->huber_a() : 0.0;
        total += 0.5 * ((a > 0 && s > a * a) ? 2 * a * std::sqrt(s) - a * a : s);
    }
    return total;
}
#endif

struct Window { Frames all; Frames active; std::vector<SE3d> truth; double start = 0; };

// mode: 0 visual + IMU (initialised), 1 visual only with two starved keyframes (the weak-constraint branch, backend.cpp:164-177),
// 3 visual + IMU samples, IMU not initialised yet (biases unknown, set to zero)
static Window make_window(int mode) {
    const double fx = 718.856, fy = 718.856, cx = 607.1928, cy = 185.2157, base = 0.537;
    Matrix3d Rbc; Rbc << 0, 0, 1, -1, 0, 0, 0, -1, 0;                 // camera z = body x
    Quaterniond qbc(Rbc);
    Quaterniond tilt(1, 0.01, -0.015, 0.02); tilt.normalize();
    const Quaterniond qe = qbc * tilt;
    const Vector3d t0(0.27, 0.26, 0.08);
    Camera::Create(fx, fy, cx, cy, SE3d(qe, t0));
    Camera::Create(fx, fy, cx, cy, SE3d(qe, Vector3d(t0 + qe * Vector3d(base, 0, 0))));
    Camera::baseline = base;
    const bool with_imu = mode == 0 || mode == 3;
    if (with_imu) { Imu::Create(SE3d(), 0.08, 0.00004, 0.004, 2.0e-6, 9.81007); Imu::Get()->initialized = mode == 0; }

    const int n_before = 2, n_active = 8;
    const double dt_kf = 0.4;
    Window w;
    std::vector<Frame::Ptr> frames;
    Frame::Ptr last;
    for (int k = -n_before; k < n_active; ++k) {
        Frame::Ptr f = Frame::Create();
        f->time = 10.0 + dt_kf * k;
        f->pose = traj_pose(f->time);
        f->last_keyframe = last;
        f->Vw = traj_v(f->time);
        if (with_imu) {
            const Bias true_bias(0.02, -0.01, 0.015, 0.001, -0.002, 0.0015);
            f->good_imu = mode == 0;
            f->bias = mode == 0 ? true_bias : Bias();                     // mode 3: nothing known yet, the initializer estimates it
            if (last) {                                                   // 100 Hz samples between the two keyframes
                f->preintegration = imu::Preintegration::Create(f->bias);
                const int ns = 40; const double h = dt_kf / ns;
                auto meas = [&](double t, Vector3d& acc, Vector3d& gyr) {
                    const Quaterniond q = yaw_q(traj_yaw(t));
                    acc = q.conjugate() * Vector3d(traj_a(t) + imu::g) + true_bias.linearized_ba + Vector3d(0.02 * nrand(), 0.02 * nrand(), 0.02 * nrand());
                    gyr = Vector3d(0, 0, traj_yaw_rate(t)) + true_bias.linearized_bg + Vector3d(0.001 * nrand(), 0.001 * nrand(), 0.001 * nrand());
                };
                Vector3d acc0, gyr0; meas(last->time, acc0, gyr0);
                for (int s = 1; s <= ns; ++s) { Vector3d acc, gyr; meas(last->time + h * s, acc, gyr); f->preintegration->Append(h, acc, gyr, acc0, gyr0); }
            }
        }
        lvio_fusion::Map::Instance().InsertKeyFrame(f);
        frames.push_back(f); last = f;
    }
    // landmarks: born in frame b (stereo pair), tracked in the following keyframes while in view
    const int n_landmarks = 420;
    for (int l = 0; l < n_landmarks; ++l) {
        const int b = (int)(urand() * (n_before + n_active - 1));          // index into frames
        Frame::Ptr fb = frames[b];
        const double depth = (l % 7 == 0) ? 30.0 + 30.0 * urand() : 4.0 + 18.0 * urand();      // every 7th beyond baseline * 50: WeakError
        const Vector3d pc_right(depth * (urand() - 0.5) * 1.2, depth * (urand() - 0.5) * 0.4, depth);
        const Vector3d pw = Camera::Get(1)->Sensor2World(pc_right, fb->pose);
        auto px = [&](int cam, const SE3d& pose, Vector2d& out) {
            const Vector3d pc = Camera::Get(cam)->World2Sensor(pw, pose);
            if (pc.z() < 1.0) return false;
            out = Camera::Get(cam)->Sensor2Pixel(pc);
            return out.x() > 5 && out.x() < 1236 && out.y() > 5 && out.y() < 371;
        };
        Vector2d pr, pl;
        if (!px(1, fb->pose, pr) || !px(0, fb->pose, pl)) continue;
        if (mode == 1 && (b == n_before + 3 || b == n_before + 4)) continue;                   // starve two keyframes of new points
        visual::Landmark::Ptr lm = visual::Landmark::Create(1.0 / depth);
        auto noisy = [&](const Vector2d& p) { return cv::KeyPoint(cv::Point2f((float)(p.x() + 0.3 * nrand()), (float)(p.y() + 0.3 * nrand())), 1.f); };
        visual::Feature::Ptr left = visual::Feature::Create(fb, noisy(pl), lm);
        lm->AddObservation(left); fb->AddFeature(left);
        visual::Feature::Ptr right = visual::Feature::Create(fb, noisy(pr), lm);
        right->is_on_left_image = false;
        lm->AddObservation(right); fb->AddFeature(right);
        lvio_fusion::Map::Instance().InsertLandmark(lm);
        const int track = 2 + (int)(urand() * 5);
        for (int k = b + 1; k < (int)frames.size() && k <= b + track; ++k) {
            if (mode == 1 && (k == n_before + 3 || k == n_before + 4) && urand() < 0.93) continue;
            Vector2d p;
            if (!px(0, frames[k]->pose, p)) break;
            visual::Feature::Ptr ob = visual::Feature::Create(frames[k], noisy(p), lm);
            lm->AddObservation(ob); frames[k]->AddFeature(ob);
        }
    }
    // what the frontend hands over is an estimate: perturb poses, velocities and inverse depths
    for (auto& f : frames) w.truth.push_back(f->pose);
    for (size_t i = 0; i < frames.size(); ++i) {
        frames[i]->pose = perturb(frames[i]->pose, i < (size_t)n_before ? 0.0 : 0.004, i < (size_t)n_before ? 0.0 : 0.06);
        if (with_imu) frames[i]->Vw = Vector3d(frames[i]->Vw + Vector3d(0.05 * nrand(), 0.05 * nrand(), 0.05 * nrand()));
    }
    for (auto& pl : lvio_fusion::Map::Instance().landmarks) pl.second->inv_depth *= 1.0 + 0.05 * nrand();
    w.all = lvio_fusion::Map::Instance().keyframes;
    w.start = frames[n_before]->time;
    w.active = lvio_fusion::Map::Instance().GetKeyFrames(w.start);
    return w;
}

static void print_state(const char* tag, const Window& w) {
    // absolute error of the active keyframes, and the error of the relative pose between neighbours (gauge-free: FullBA has
    // no anchor, the whole map may move in yaw and translation)
    double e_t = 0, e_r = 0, g_t = 0, g_r = 0; int n = 0, m = 0, i = 0;
    const SE3d* prev_est = nullptr; const SE3d* prev_true = nullptr;
    for (auto& kv : w.all) {
        const SE3d& est = kv.second->pose; const SE3d& tru = w.truth[i++];
        const SE3d d = tru.inverse() * est;
        if (kv.first >= w.start) { e_t += d.translation().squaredNorm(); e_r += d.unit_quaternion().vec().squaredNorm() * 4; ++n; }
        if (prev_est) {
            const SE3d dd = (prev_true->inverse() * tru).inverse() * (prev_est->inverse() * est);
            g_t += dd.translation().squaredNorm(); g_r += dd.unit_quaternion().vec().squaredNorm() * 4; ++m;
        }
        prev_est = &est; prev_true = &tru;
    }
    printf("%s rmse_t %.9e rmse_r %.9e rel_t %.9e rel_r %.9e\n", tag, std::sqrt(e_t / n), std::sqrt(e_r / n), std::sqrt(g_t / m), std::sqrt(g_r / m));
}

static double reprojection_sum(const Frames& kfs, int* count) {                 // the outlier test of backend.cpp:224-238, summed
    double s = 0; *count = 0;
    for (auto& kv : kfs)
        for (auto& pf : kv.second->features_left) {
            auto lm = pf.second->landmark.lock();
            if (lm->FirstFrame().lock() == kv.second) continue;
            s += compute_reprojection_error(cv2eigen(pf.second->keypoint.pt), lm->ToWorld(), kv.second->pose, Camera::Get());
            ++*count;
        }
    return s;
}

int main(int argc, char** argv) {
    const int mode = argc > 1 ? std::atoi(argv[1]) : 0;          // 0 window with IMU, 1 visual only (weak constraints), 2 imu::FullBA, 3 Initializer, 4 Environment::Optimize
    const char* dump = argc > 2 ? argv[2] : nullptr;
    Window w = make_window(mode == 2 || mode == 4 ? 0 : mode);
    if (mode == 3) {
        // the visual-only map is not gravity aligned: tilt everything (poses and velocities; landmarks hang on their first keyframe)
        const Quaterniond tilt = Quaterniond(std::cos(0.09), std::sin(0.09) * 0.8, std::sin(0.09) * 0.6, 0);        // 0.18 rad about (0.8, 0.6, 0)
        lvio_fusion::Map::Instance().ApplyGravityRotation(tilt.toRotationMatrix());
        w.all = lvio_fusion::Map::Instance().keyframes;
    }
#ifdef DROPIN_PRODUCT
    {   // Estimator would register the rig once
        double cam[2][11];
        for (int c = 0; c < 2; ++c) {
            cam[c][0] = Camera::Get(c)->fx; cam[c][1] = Camera::Get(c)->fy; cam[c][2] = Camera::Get(c)->cx; cam[c][3] = Camera::Get(c)->cy;
            std::memcpy(cam[c] + 4, Camera::Get(c)->extrinsic.data(), 7 * sizeof(double));
        }
        lvb::Runtime::get().set_cameras(cam[0], cam[1]);
    }
#endif
    int n_rep = 0;
    const double rep0 = reprojection_sum(w.active, &n_rep);
    printf("mode %d keyframes %zu active %zu landmarks %zu reproj_n %d reproj_sum %.12e\n", mode, w.all.size(), w.active.size(), lvio_fusion::Map::Instance().landmarks.size(), n_rep, rep0);
    print_state("before", w);

    if (mode == 4) {
        // Environment::Optimize (src/environment.cpp:18-113; the weight-adaptation environment re-solves ONE keyframe): a copy of the
        // frame, PoseOnlyReprojectionError for every left feature, one ImuError whose other seven blocks are constant, DENSE_QR.
        // No estimator behind it here (mapping == nullptr: the lidar branch is skipped).
#ifdef DROPIN_PRODUCT
        alignas(Estimator) static unsigned char est_storage[sizeof(Estimator)] = {0};       // all-zero shared_ptr members are empty ones
        Environment::estimator_ = Estimator::Ptr(reinterpret_cast<Estimator*>(est_storage), [](Estimator*) {});
        Environment::u_ = std::uniform_real_distribution<double>(w.start, w.start + 1.0);
        Environment* env = new Environment();
        env->frames_ = w.active;
        env->state_ = env->frames_.begin(); ++env->state_; ++env->state_; ++env->state_;     // the fourth active keyframe
        Frame::Ptr frame = env->state_->second;
        auto frame_reproj = [&](const SE3d& pose) { double s = 0; for (auto& pf : frame->features_left) s += compute_reprojection_error(cv2eigen(pf.second->keypoint.pt), pf.second->landmark.lock()->ToWorld(), pose, Camera::Get()); return s; };
        const SE3d before = frame->pose;
        const SE3d result = env->Optimize();
        const SE3d moved = before.inverse() * result;
        printf("env features %zu reproj_before %.9e reproj_after %.9e moved_t %.6e moved_r %.6e frame_untouched %d\n", frame->features_left.size(), frame_reproj(before), frame_reproj(result),
               moved.translation().norm(), 2 * moved.unit_quaternion().vec().norm(), (int)(std::memcmp(frame->pose.data(), before.data(), 7 * sizeof(double)) == 0));
#else
        printf("environment skipped in the recording build\n");
#endif
    } else if (mode == 3) {
        // Initializer::Initialize(frames, prior_a, prior_g) (src/initializer.cpp:32-55): velocities and gravity direction from the
        // preintegrated velocities, imu::InertialOptimization (NumericDiff ImuInitGError on the host LM), the map rotated onto
        // gravity, imu::FullBA (device path), Imu::initialized = true.  Every frame handed over has a predecessor and a preintegration.
        auto gravity_tilt = [&]() {       // angle between the estimated and the true vertical, worst keyframe
            double worst = 0; int i = 0;
            for (auto& kv : w.all) { const Vector3d up = (kv.second->pose.unit_quaternion() * w.truth[i++].unit_quaternion().conjugate()) * Vector3d::UnitZ(); worst = std::max(worst, std::acos(std::min(1.0, up.z()))); }
            return worst;
        };
#ifdef DROPIN_PRODUCT
        Frames frames(++w.all.begin(), w.all.end());
        printf("init tilt_before %.6e\n", gravity_tilt());
        Initializer initializer;
        const bool ok = initializer.Initialize(frames, 1e4, 1e2);
        print_state("after", w);
        const Bias b = frames.begin()->second->bias;
        printf("init ok %d initialized %d tilt_after %.6e bias %.6e %.6e %.6e %.6e %.6e %.6e\n", (int)ok, (int)Imu::Get()->initialized, gravity_tilt(),
               b.linearized_ba[0], b.linearized_ba[1], b.linearized_ba[2], b.linearized_bg[0], b.linearized_bg[1], b.linearized_bg[2]);
#else
        printf("initializer skipped in the recording build\n");
#endif
    } else if (mode == 2) {
        // imu::FullBA (src/tools.cpp:92-171): every keyframe of the map, ImuInitError with one shared ba / bg block
#ifdef DROPIN_PRODUCT
        imu::FullBA(w.all, 1e4, 1e2);                     // the priors of Initializer::Initialize (src/initializer.cpp:62)
        print_state("after", w);
        const Bias b = w.all.begin()->second->bias;
        printf("fullba bias %.9e %.9e %.9e %.9e %.9e %.9e\n", b.linearized_ba[0], b.linearized_ba[1], b.linearized_ba[2], b.linearized_bg[0], b.linearized_bg[1], b.linearized_bg[2]);
#else
        printf("fullba skipped in the recording build (tools.cpp keeps its problem local)\n");
#endif
    } else {
        // storage for a Backend without running its constructor (which starts the two worker threads): BuildProblem only
        // reads global_end_
        alignas(Backend) static unsigned char storage[sizeof(Backend)];
        Backend* backend = reinterpret_cast<Backend*>(storage);
        backend->global_end_ = 0;
        adapt::Problem problem;
        const double global_end = backend->BuildProblem(w.active, problem);
        printf("types visual %d weak %d lidar %d navsat %d pose %d imu %d other %d frames %d global_end %.6f\n",
               problem.num_types[ProblemType::VisualError], problem.num_types[ProblemType::WeakError], problem.num_types[ProblemType::LidarError],
               problem.num_types[ProblemType::NavsatError], problem.num_types[ProblemType::PoseError], problem.num_types[ProblemType::ImuError],
               problem.num_types[ProblemType::Other], problem.num_frames, global_end);
#ifdef DROPIN_PRODUCT
        int kinds[LVB_NUM_KINDS] = {0};
        for (auto& rb : problem.residual_blocks()) kinds[static_cast<const lvb::DeviceCost*>(rb->cost)->kind()]++;
        printf("kinds two_frame %d pose_only %d two_camera %d imu %d pose_graph %d pose_prior %d blocks %d\n", kinds[LVB_TWO_FRAME], kinds[LVB_POSE_ONLY],
               kinds[LVB_TWO_CAMERA], kinds[LVB_IMU], kinds[LVB_POSE_GRAPH], kinds[LVB_POSE_PRIOR], problem.NumParameterBlocks());
        ceres::Solver::Options options;                          // backend.cpp:205-211
        options.linear_solver_type = ceres::SPARSE_SCHUR;
        options.max_solver_time_in_seconds = 1e9;                // the reference caps by wall clock ((end - start) / n): not reproducible, lifted here
        options.num_threads = num_threads;
        ceres::Solver::Summary summary;
        adapt::Solve(options, &problem, &summary);
        if (Imu::Num() && Imu::Get()->initialized) imu::RecoverBias(w.active);
        printf("solve initial_cost %.12e final_cost %.12e steps %d term %d msg %s\n", summary.initial_cost, summary.final_cost,
               summary.num_successful_steps, (int)summary.termination_type, summary.message.c_str());
        print_state("after", w);
        const double rep1 = reprojection_sum(w.active, &n_rep);
        printf("after reproj_n %d reproj_sum %.12e\n", n_rep, rep1);
#else
        printf("blocks %zu residual_blocks %zu\n", problem.parameter_blocks.size(), problem.residual_blocks.size());
        printf("reference initial_cost %.12e\n", recorded_cost(problem));
#endif
    }
    if (dump) {                                                  // final state, for comparing the builds
        FILE* f = fopen(dump, "wb");
        for (auto& kv : w.all) { fwrite(kv.second->pose.data(), sizeof(double), 7, f); fwrite(kv.second->Vw.data(), sizeof(double), 3, f); }
        std::map<unsigned long, visual::Landmark::Ptr> ordered(lvio_fusion::Map::Instance().landmarks.begin(), lvio_fusion::Map::Instance().landmarks.end());
        for (auto& kv : ordered) fwrite(&kv.second->inv_depth, sizeof(double), 1, f);
        fclose(f);
    }
    return 0;
}
