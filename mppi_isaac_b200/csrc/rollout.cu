// rollout.cu -- K2: batched articulated rigid-body rollout (replaces gym.simulate()/IsaacGymWrapper.step
// on the MPPI path: mppiisaac/planner/isaacgym_wrapper.py:524-572 apply_robot_cmd, :639-655 step).
//
// Mapping: ONE THREAD PER ROLLOUT, 32-thread CTAs.  The recursion over bodies and over time is strictly
// serial (SURVEY.md section 5 "T stays sequential"); all parallelism is across the K samples.  The model and
// parameter blocks arrive as __grid_constant__ kernel parameters, i.e. they live in the constant bank and
// feed FFMA operands directly -- every lane reads the same constant at the same time, which is the access
// pattern the constant cache is built for (no shared-memory staging instruction is needed at all).
// All K-indexed global arrays are [..][K] with k innermost, so every load/store of a warp is one fully
// coalesced 128-byte line.
//
// Dynamics formulation: articulated-body algorithm in WORLD coordinates (spatial vectors taken about the
// world origin), so articulated inertias are summed into the parent without any 6x6 frame transform.
// This is deliberately a different formulation from the CPU oracle (body-coordinate ABA with dense 6x6
// Pluecker transforms); the two must agree to float32 round-off.
//
// Per substep h = dt/substeps:
//   1. kinematics + world spatial inertia + velocity-product terms
//   2. ABA with the PD drive and joint damping treated implicitly (added to the joint-space diagonal),
//      one re-solve with constant saturated torques for joints whose drive torque exceeds URDF effort
//   3. semi-implicit Euler, velocity limit, position limits
// After the last substep of a model step the observed rows are written to obs[R][T][K].
#include "common.cuh"

namespace {

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

struct S3 { float xx, yy, zz, xy, xz, yz; };          // symmetric 3x3
struct M3 { float m00, m01, m02, m10, m11, m12, m20, m21, m22; };  // general 3x3, row major

__device__ __forceinline__ V3 mul(const S3& s, V3 v) {
    return mk(s.xx * v.x + s.xy * v.y + s.xz * v.z, s.xy * v.x + s.yy * v.y + s.yz * v.z, s.xz * v.x + s.yz * v.y + s.zz * v.z);
}
__device__ __forceinline__ V3 mul(const M3& m, V3 v) {
    return mk(m.m00 * v.x + m.m01 * v.y + m.m02 * v.z, m.m10 * v.x + m.m11 * v.y + m.m12 * v.z, m.m20 * v.x + m.m21 * v.y + m.m22 * v.z);
}
__device__ __forceinline__ V3 mulT(const M3& m, V3 v) {
    return mk(m.m00 * v.x + m.m10 * v.y + m.m20 * v.z, m.m01 * v.x + m.m11 * v.y + m.m21 * v.z, m.m02 * v.x + m.m12 * v.y + m.m22 * v.z);
}
__device__ __forceinline__ V3 col(const M3& m, int c) {
    return c == 0 ? mk(m.m00, m.m10, m.m20) : (c == 1 ? mk(m.m01, m.m11, m.m21) : mk(m.m02, m.m12, m.m22));
}

struct Quat { float x, y, z, w; };
__device__ __forceinline__ Quat qmul(Quat a, Quat b) {
    Quat o;
    o.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    o.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    o.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    o.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    return o;
}
__device__ __forceinline__ M3 quat_to_R(Quat q) {
    M3 R;
    R.m00 = 1 - 2 * (q.y * q.y + q.z * q.z); R.m01 = 2 * (q.x * q.y - q.z * q.w); R.m02 = 2 * (q.x * q.z + q.y * q.w);
    R.m10 = 2 * (q.x * q.y + q.z * q.w); R.m11 = 1 - 2 * (q.x * q.x + q.z * q.z); R.m12 = 2 * (q.y * q.z - q.x * q.w);
    R.m20 = 2 * (q.x * q.z - q.y * q.w); R.m21 = 2 * (q.y * q.z + q.x * q.w); R.m22 = 1 - 2 * (q.x * q.x + q.y * q.y);
    return R;
}

// spatial (6x6 symmetric) inertia about the world origin: [[A, B], [B^T, C]]
struct SpI { S3 A; M3 B; S3 C; };
struct V6 { V3 n, f; };   // also used for motion vectors (n = angular, f = linear)

__device__ __forceinline__ V6 mul(const SpI& I, const V6& s) {
    V6 o;
    o.n = mul(I.A, s.n) + mul(I.B, s.f);
    o.f = mulT(I.B, s.n) + mul(I.C, s.f);
    return o;
}
__device__ __forceinline__ float dot6(const V6& a, const V6& b) { return dot(a.n, b.n) + dot(a.f, b.f); }
__device__ __forceinline__ void add_to(SpI& p, const SpI& c) {
    p.A.xx += c.A.xx; p.A.yy += c.A.yy; p.A.zz += c.A.zz; p.A.xy += c.A.xy; p.A.xz += c.A.xz; p.A.yz += c.A.yz;
    p.B.m00 += c.B.m00; p.B.m01 += c.B.m01; p.B.m02 += c.B.m02; p.B.m10 += c.B.m10; p.B.m11 += c.B.m11; p.B.m12 += c.B.m12;
    p.B.m20 += c.B.m20; p.B.m21 += c.B.m21; p.B.m22 += c.B.m22;
    p.C.xx += c.C.xx; p.C.yy += c.C.yy; p.C.zz += c.C.zz; p.C.xy += c.C.xy; p.C.xz += c.C.xz; p.C.yz += c.C.yz;
}
// I -= s * U U^T
__device__ __forceinline__ void rank1_sub(SpI& I, const V6& U, float s) {
    V3 a = s * U.n, b = s * U.f;
    I.A.xx -= a.x * U.n.x; I.A.yy -= a.y * U.n.y; I.A.zz -= a.z * U.n.z;
    I.A.xy -= a.x * U.n.y; I.A.xz -= a.x * U.n.z; I.A.yz -= a.y * U.n.z;
    I.B.m00 -= a.x * U.f.x; I.B.m01 -= a.x * U.f.y; I.B.m02 -= a.x * U.f.z;
    I.B.m10 -= a.y * U.f.x; I.B.m11 -= a.y * U.f.y; I.B.m12 -= a.y * U.f.z;
    I.B.m20 -= a.z * U.f.x; I.B.m21 -= a.z * U.f.y; I.B.m22 -= a.z * U.f.z;
    I.C.xx -= b.x * U.f.x; I.C.yy -= b.y * U.f.y; I.C.zz -= b.z * U.f.z;
    I.C.xy -= b.x * U.f.y; I.C.xz -= b.x * U.f.z; I.C.yz -= b.y * U.f.z;
}

// topology policy: FAN >= 0 -> parent(i) = min(i-1, FAN) known at compile time (serial chain when FAN >= NB-1,
// two-finger gripper when FAN = NB-3); FAN < 0 -> runtime parents from the model block (loops stay rolled).
template <int NB_, int FAN_> struct Topo {
    static constexpr int NB = NB_;
    static constexpr bool kStatic = FAN_ >= 0;
    __device__ __forceinline__ static int parent(const MppibModel& m, int i) {
        if (kStatic) return (i - 1 <= FAN_) ? i - 1 : FAN_;
        return m.parent[i];
    }
    __device__ __forceinline__ static int nb(const MppibModel& m) { return kStatic ? NB_ : m.nb; }
};

template <class TP>
__global__ void __launch_bounds__(32)
rollout_kernel(const __grid_constant__ MppibModel m, const __grid_constant__ MppibParams p,
               const float* __restrict__ state0, float* __restrict__ state, const float* __restrict__ actions,
               int t0, int nsteps, float* __restrict__ obs) {
    constexpr int NB = TP::NB;
    const int K = p.K, T = p.T, nu = m.nu;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int nb = TP::nb(m);
    const float h = p.dt / (float)p.substeps;

    float q[NB], qd[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        if (i < nb) {
            q[i] = state0 ? state0[i] : state[(size_t)i * K + k];
            qd[i] = state0 ? state0[nb + i] : state[(size_t)(nb + i) * K + k];
        }
    }
    const Quat bq = {m.base_quat[0], m.base_quat[1], m.base_quat[2], m.base_quat[3]};
    const M3 Rb = quat_to_R(bq);
    const V3 ob = mk(m.base_pos[0], m.base_pos[1], m.base_pos[2]);
    // gravity enters as a fictitious base acceleration a0 = [0; -g]
    const V3 a0f = m.gravity_on ? mk(-m.gravity[0], -m.gravity[1], -m.gravity[2]) : mk(0.f, 0.f, 0.f);

    M3 R[NB]; V3 o[NB]; V6 S[NB], V[NB];

    auto kinematics = [&]() {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            if (i >= nb) continue;
            const int par = TP::parent(m, i);
            const M3& Rp = par >= 0 ? R[par > 0 ? par : 0] : Rb;
            const V3 op = par >= 0 ? o[par > 0 ? par : 0] : ob;
            const float* tr = m.tree_R[i];
            // Rt = Rp * tree_R
            M3 Rt;
            {
                V3 c0 = mul(Rp, mk(tr[0], tr[3], tr[6])), c1 = mul(Rp, mk(tr[1], tr[4], tr[7])), c2 = mul(Rp, mk(tr[2], tr[5], tr[8]));
                Rt.m00 = c0.x; Rt.m10 = c0.y; Rt.m20 = c0.z; Rt.m01 = c1.x; Rt.m11 = c1.y; Rt.m21 = c1.z; Rt.m02 = c2.x; Rt.m12 = c2.y; Rt.m22 = c2.z;
            }
            V3 oi = op + mul(Rp, mk(m.tree_p[i][0], m.tree_p[i][1], m.tree_p[i][2]));
            const V3 axis = mk(Rt.m02, Rt.m12, Rt.m22);
            V6 Vp;
            if (par >= 0) Vp = V[par > 0 ? par : 0]; else { Vp.n = mk(0, 0, 0); Vp.f = mk(0, 0, 0); }
            if (m.jtype[i] == MPPIB_JOINT_REVOLUTE) {
                float sq, cq; sincosf(q[i], &sq, &cq);
                M3 Ri = Rt;   // Rt * Rz(q)
                Ri.m00 = Rt.m00 * cq + Rt.m01 * sq; Ri.m01 = Rt.m01 * cq - Rt.m00 * sq;
                Ri.m10 = Rt.m10 * cq + Rt.m11 * sq; Ri.m11 = Rt.m11 * cq - Rt.m10 * sq;
                Ri.m20 = Rt.m20 * cq + Rt.m21 * sq; Ri.m21 = Rt.m21 * cq - Rt.m20 * sq;
                R[i] = Ri; o[i] = oi;
                S[i].n = axis; S[i].f = cross(oi, axis);
            } else {
                R[i] = Rt; o[i] = oi + q[i] * axis;
                S[i].n = mk(0, 0, 0); S[i].f = axis;
            }
            V[i].n = Vp.n + qd[i] * S[i].n;
            V[i].f = Vp.f + qd[i] * S[i].f;
        }
    };

    const int nloop = nsteps > 0 ? nsteps : 1;   // nsteps == 0: observe the current state into slot t0
    for (int t = t0; t < t0 + nloop; ++t) {
        // apply_robot_cmd: command -> per-DOF targets (diff-drive IK folded into the cmd map)
        float target[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            if (i >= nb || nsteps == 0) continue;
            float u0 = p.u_scale * actions[((size_t)t * nu + m.cmd_i0[i]) * K + k];
            float u1 = p.u_scale * actions[((size_t)t * nu + m.cmd_i1[i]) * K + k];
            target[i] = m.cmd_c0[i] * u0 + m.cmd_c1[i] * u1;
        }
        const int nsub = nsteps > 0 ? p.substeps : 0;
        for (int sub = 0; sub < nsub; ++sub) {
            kinematics();
            // per-body world inertia, velocity-product acceleration c and bias force pb
            SpI Ib[NB]; V6 c[NB], pb[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                if (i >= nb) continue;
                const M3& Ri = R[i];
                const float mass = m.mass[i];
                // centre of mass (world) and first moment
                V3 cw = o[i];
                if (mass > 0.f) cw = cw + (1.0f / mass) * mul(Ri, mk(m.mcom[i][0], m.mcom[i][1], m.mcom[i][2]));
                V3 hw = mass * cw;
                // I_O = R I_o,body R^T shifted from the body origin to the world origin:
                //   I_world_origin = R (I_o - m(|c|^2 1 - c c^T)) R^T + m(|cw|^2 1 - cw cw^T)
                // computed as R I_o R^T + m[(|cw|^2-|cb|^2) 1 - (cw cw^T - cb cb^T)], cb = R c_body
                S3 Io = {m.inertia[i][0], m.inertia[i][1], m.inertia[i][2], m.inertia[i][3], m.inertia[i][4], m.inertia[i][5]};
                V3 r0 = mk(Ri.m00, Ri.m01, Ri.m02), r1 = mk(Ri.m10, Ri.m11, Ri.m12), r2 = mk(Ri.m20, Ri.m21, Ri.m22);
                V3 t0v = mul(Io, r0), t1v = mul(Io, r1), t2v = mul(Io, r2);   // rows of R Io  (Io symmetric)
                S3 A;
                A.xx = dot(r0, t0v); A.yy = dot(r1, t1v); A.zz = dot(r2, t2v);
                A.xy = dot(r0, t1v); A.xz = dot(r0, t2v); A.yz = dot(r1, t2v);
                V3 cb = cw - o[i];
                float d2 = mass * (dot(cw, cw) - dot(cb, cb));
                A.xx += d2 - mass * (cw.x * cw.x - cb.x * cb.x); A.yy += d2 - mass * (cw.y * cw.y - cb.y * cb.y);
                A.zz += d2 - mass * (cw.z * cw.z - cb.z * cb.z);
                A.xy -= mass * (cw.x * cw.y - cb.x * cb.y); A.xz -= mass * (cw.x * cw.z - cb.x * cb.z);
                A.yz -= mass * (cw.y * cw.z - cb.y * cb.z);
                SpI I;
                I.A = A;
                I.B.m00 = 0; I.B.m01 = -hw.z; I.B.m02 = hw.y; I.B.m10 = hw.z; I.B.m11 = 0; I.B.m12 = -hw.x;
                I.B.m20 = -hw.y; I.B.m21 = hw.x; I.B.m22 = 0;
                I.C.xx = mass; I.C.yy = mass; I.C.zz = mass; I.C.xy = 0; I.C.xz = 0; I.C.yz = 0;
                Ib[i] = I;
                // bias force  V x* (I V)
                V3 w = V[i].n, v = V[i].f;
                V3 nn = mul(A, w) + cross(hw, v);
                V3 ff = mass * v - cross(hw, w);
                pb[i].n = cross(w, nn) + cross(v, ff);
                pb[i].f = cross(w, ff);
                // c = V x (S qd)
                V3 sw = qd[i] * S[i].n, sv = qd[i] * S[i].f;
                c[i].n = cross(w, sw);
                c[i].f = cross(w, sv) + cross(v, sw);
            }
            float tau[NB], dimp[NB], qdd[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                if (i >= nb) continue;
                const float kd = m.kd[i], b = m.damping[i];
                if (m.drive_mode == MPPIB_DRIVE_VELOCITY) {
                    tau[i] = kd * (target[i] - qd[i]) - b * qd[i];
                } else {
                    float e = fminf(fmaxf(target[i], -m.effort[i]), m.effort[i]);
                    tau[i] = e - (kd + b) * qd[i];
                }
                dimp[i] = m.armature[i] + h * (kd + b);
            }
#pragma unroll 1
            for (int solve = 0; solve < 2; ++solve) {
                SpI IA[NB]; V6 pA[NB], U[NB]; float invD[NB], uu[NB];
#pragma unroll
                for (int i = 0; i < NB; ++i) { if (i < nb) { IA[i] = Ib[i]; pA[i] = pb[i]; } }
#pragma unroll
                for (int i = NB - 1; i >= 0; --i) {
                    if (i >= nb) continue;
                    U[i] = mul(IA[i], S[i]);
                    float D = dot6(S[i], U[i]) + dimp[i];
                    invD[i] = 1.0f / D;
                    uu[i] = tau[i] - dot6(S[i], pA[i]);
                    const int par = TP::parent(m, i);
                    if (par >= 0) {
                        SpI Ia = IA[i];
                        rank1_sub(Ia, U[i], invD[i]);
                        V6 Iac = mul(Ia, c[i]);
                        float s = uu[i] * invD[i];
                        V6 pa;
                        pa.n = pA[i].n + Iac.n + s * U[i].n;
                        pa.f = pA[i].f + Iac.f + s * U[i].f;
                        const int pi = par > 0 ? par : 0;
                        add_to(IA[pi], Ia);
                        pA[pi].n = pA[pi].n + pa.n; pA[pi].f = pA[pi].f + pa.f;
                    }
                }
                V6 acc[NB];
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    if (i >= nb) continue;
                    const int par = TP::parent(m, i);
                    V6 ap;
                    if (par >= 0) ap = acc[par > 0 ? par : 0]; else { ap.n = mk(0, 0, 0); ap.f = a0f; }
                    ap.n = ap.n + c[i].n; ap.f = ap.f + c[i].f;
                    qdd[i] = (uu[i] - dot6(U[i], ap)) * invD[i];
                    acc[i].n = ap.n + qdd[i] * S[i].n;
                    acc[i].f = ap.f + qdd[i] * S[i].f;
                }
                if (solve == 1 || m.drive_mode != MPPIB_DRIVE_VELOCITY) break;
                bool any = false;
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    if (i >= nb) continue;
                    float td = m.kd[i] * (target[i] - (qd[i] + h * qdd[i]));
                    if (fabsf(td) > m.effort[i]) {
                        any = true;
                        tau[i] = (td > 0.f ? m.effort[i] : -m.effort[i]) - m.damping[i] * qd[i];
                        dimp[i] = m.armature[i] + h * m.damping[i];
                    }
                }
                if (!any) break;
            }
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                if (i >= nb) continue;
                float v = qd[i] + h * qdd[i];
                v = fminf(fmaxf(v, -m.qd_max[i]), m.qd_max[i]);
                float x = q[i] + h * v;
                if (x < m.q_lo[i]) { x = m.q_lo[i]; if (v < 0.f) v = 0.f; }
                if (x > m.q_hi[i]) { x = m.q_hi[i]; if (v > 0.f) v = 0.f; }
                q[i] = x; qd[i] = v;
            }
        }
        if (obs == nullptr) continue;
        // ---------------------------------------------------------------------------------- observe
        kinematics();
        // per-thread scratch for the dynamically indexed link lookup (kept out of the hot arrays)
        float bo[NB][3], bw[NB][3], bv[NB][3], bqv[NB][4], bR[NB][9];
        {
            Quat qw[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                if (i >= nb) continue;
                const int par = TP::parent(m, i);
                Quat qp = par >= 0 ? qw[par > 0 ? par : 0] : bq;
                Quat qt = {m.tree_quat[i][0], m.tree_quat[i][1], m.tree_quat[i][2], m.tree_quat[i][3]};
                Quat r = qmul(qp, qt);
                if (m.jtype[i] == MPPIB_JOINT_REVOLUTE) {
                    float sh, ch; sincosf(0.5f * q[i], &sh, &ch);
                    Quat qz = {0.f, 0.f, sh, ch};
                    r = qmul(r, qz);
                }
                qw[i] = r;
                bo[i][0] = o[i].x; bo[i][1] = o[i].y; bo[i][2] = o[i].z;
                bw[i][0] = V[i].n.x; bw[i][1] = V[i].n.y; bw[i][2] = V[i].n.z;
                bv[i][0] = V[i].f.x; bv[i][1] = V[i].f.y; bv[i][2] = V[i].f.z;
                bqv[i][0] = r.x; bqv[i][1] = r.y; bqv[i][2] = r.z; bqv[i][3] = r.w;
                bR[i][0] = R[i].m00; bR[i][1] = R[i].m01; bR[i][2] = R[i].m02; bR[i][3] = R[i].m10; bR[i][4] = R[i].m11;
                bR[i][5] = R[i].m12; bR[i][6] = R[i].m20; bR[i][7] = R[i].m21; bR[i][8] = R[i].m22;
            }
        }
        const size_t TK = (size_t)T * K;
        float* dst = obs + (size_t)t * K + k;
        int row = 0;
        for (int oi = 0; oi < p.nobs; ++oi) {
            const int kind = p.obs[oi].kind, idx = p.obs[oi].index;
            if (kind == MPPIB_OBS_LINK_STATE) {
                const int b = m.link_body[idx];
                M3 Rl; V3 ol, w, vO; Quat qb;
                if (b >= 0) {
                    Rl.m00 = bR[b][0]; Rl.m01 = bR[b][1]; Rl.m02 = bR[b][2]; Rl.m10 = bR[b][3]; Rl.m11 = bR[b][4]; Rl.m12 = bR[b][5];
                    Rl.m20 = bR[b][6]; Rl.m21 = bR[b][7]; Rl.m22 = bR[b][8];
                    ol = mk(bo[b][0], bo[b][1], bo[b][2]); w = mk(bw[b][0], bw[b][1], bw[b][2]); vO = mk(bv[b][0], bv[b][1], bv[b][2]);
                    qb.x = bqv[b][0]; qb.y = bqv[b][1]; qb.z = bqv[b][2]; qb.w = bqv[b][3];
                } else { Rl = Rb; ol = ob; w = mk(0, 0, 0); vO = mk(0, 0, 0); qb = bq; }
                V3 pos = ol + mul(Rl, mk(m.link_p[idx][0], m.link_p[idx][1], m.link_p[idx][2]));
                Quat ql = {m.link_quat[idx][0], m.link_quat[idx][1], m.link_quat[idx][2], m.link_quat[idx][3]};
                Quat qo = qmul(qb, ql);
                V3 vel = vO + cross(w, pos);   // spatial velocity about the world origin -> velocity of the link origin
                dst[(size_t)(row + 0) * TK] = pos.x; dst[(size_t)(row + 1) * TK] = pos.y; dst[(size_t)(row + 2) * TK] = pos.z;
                dst[(size_t)(row + 3) * TK] = qo.x; dst[(size_t)(row + 4) * TK] = qo.y; dst[(size_t)(row + 5) * TK] = qo.z;
                dst[(size_t)(row + 6) * TK] = qo.w;
                dst[(size_t)(row + 7) * TK] = vel.x; dst[(size_t)(row + 8) * TK] = vel.y; dst[(size_t)(row + 9) * TK] = vel.z;
                dst[(size_t)(row + 10) * TK] = w.x; dst[(size_t)(row + 11) * TK] = w.y; dst[(size_t)(row + 12) * TK] = w.z;
                row += 13;
            } else if (kind == MPPIB_OBS_DOF_STATE) {
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    if (i >= nb) continue;
                    dst[(size_t)(row + 2 * i) * TK] = q[i];
                    dst[(size_t)(row + 2 * i + 1) * TK] = qd[i];
                }
                row += 2 * nb;
            } else {
                const int w = kind == MPPIB_OBS_CONTACT ? 3 : 13;
                for (int r = 0; r < w; ++r) dst[(size_t)(row + r) * TK] = 0.f;
                row += w;
            }
        }
    }
    if (state != nullptr) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            if (i >= nb) continue;
            state[(size_t)i * K + k] = q[i];
            state[(size_t)(nb + i) * K + k] = qd[i];
        }
    }
}

template <class TP>
int launch_t(MppibContext* c, const float* state0, float* state, const float* actions, int t0, int nsteps, float* obs, cudaStream_t s) {
    const int K = c->params.K;
    dim3 grid((K + 31) / 32), block(32);
    rollout_kernel<TP><<<grid, block, 0, s>>>(c->model, c->params, state0, state, actions, t0, nsteps, obs);
    MPPIB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int launch_rollout(MppibContext* c, const float* state0, float* state, const float* actions, int t0, int nsteps,
                   float* obs, cudaStream_t s) {
    const MppibModel& m = c->model;
    bool chain = true, fan2 = m.nb >= 3;
    for (int i = 0; i < m.nb; ++i) {
        if (m.parent[i] != i - 1) chain = false;
        int want = (i - 1 <= m.nb - 3) ? i - 1 : m.nb - 3;
        if (m.parent[i] != want) fan2 = false;
    }
    if (chain && m.nb == 3) return launch_t<Topo<3, 16>>(c, state0, state, actions, t0, nsteps, obs, s);
    if (chain && m.nb == 7) return launch_t<Topo<7, 16>>(c, state0, state, actions, t0, nsteps, obs, s);
    if (fan2 && m.nb == 9) return launch_t<Topo<9, 6>>(c, state0, state, actions, t0, nsteps, obs, s);
    return launch_t<Topo<MPPIB_MAX_BODIES, -1>>(c, state0, state, actions, t0, nsteps, obs, s);
}
