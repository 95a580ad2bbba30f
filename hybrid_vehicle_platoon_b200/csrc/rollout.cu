// rollout.cu -- batched PlatoonEnv.step on sm_100a (compile this file with -fmad=false).
//
// One thread per (scenario, vehicle): stage cost on the pre-step state (env.py:126-180), optional
// gear derivation (env.py:198-204 -> models.py:494-515), ten explicit-Euler sub-steps of the
// friction/gear hybrid dynamics (models.py:236-257, 114-125, 30-51), safe-distance flag and the
// reference's exception predicates as an error code.
//
// Parity: every floating-point operation is issued in the reference's order with IEEE
// round-to-nearest and NO fused multiply-add (nvcc -fmad=false for this translation unit), so the
// result is bit-identical to numpy's wherever numpy itself does not fuse.
//
// Division: the IEEE quotients a/m, (v-v0)/(v1-v0), (v-v2)/(v3-v2) have denominators that are
// constant per vehicle / per gear, so they are computed with Markstein's sequence
// q = a*y; r = fma(-b,q,a); q = fma(r,y,q); r = fma(-b,q,a); q = fma(r,y,q) with y = RN(1/b)
// (5 FP64 instructions instead of the ~20-instruction generic division).  With y correctly
// rounded and q faithful after the first correction, the second correction returns the correctly
// rounded quotient (Markstein 1990), i.e. exactly what numpy's division returns; 1.2e8 random
// quotients of this kernel's operand ranges were checked against the hardware division on the CPU.
//
// Memory: state rows [scenario][2n] are contiguous, so consecutive threads read consecutive
// 16-byte (p,v) pairs, 8-byte inputs/masses and 4-byte gears: every global access is fully
// coalesced and each byte is touched once (the predecessor's state comes from the same lines via
// L1).  HBM is the intended bound; the FP64 pipe (IEEE divisions in the sub-steps) competes with it.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "hvp_internal.h"

namespace hvp {

__device__ __forceinline__ double cost2(double e0, double e1, double q0, double q1, bool quad) {
    if (quad) return (e0 * q0) * e0 + (e1 * q1) * e1;  // x.T @ Q @ x with Q = diag(q0, q1)
    return fabs(q0 * e0) + fabs(q1 * e1);              // ||Q x||_1 (env.py:122-124)
}

// correctly rounded a / b given y = RN(1 / b)
__device__ __forceinline__ double div_by(double a, double b, double y) {
    double q = a * y;
    double r = fma(-b, q, a);
    q = fma(r, y, q);
    r = fma(-b, q, a);
    return fma(r, y, q);
}

// Ten explicit-Euler sub-steps with the gear held (models.py:236-257, Q5).  FAST = u finite and c_fric a power of two:
// `v + 0*u` is v bit for bit (v > 0 inside a gear's range) and c_fric*(v*v)/m == (v*v)/(m/c_fric), so the loop needs
// 12 FP64 instructions per sub-step instead of 16; the flat-part test compares the HIGH WORDS of the doubles on the
// integer pipe (conservative: anything near the edges takes the exact path below).  r02h ncu: in the bench
// distribution 65 % of the warp-iterations have a lane or two on the sloped part of a traction curve (the PWA gear
// regions reach 0.06-0.16 m/s beyond the flat part), so that path is a real share of the kernel.
template <bool FAST>
__device__ __forceinline__ int32_t substeps(const RolloutParams& P, double& p, double& v, double ui, double zu, double m,
                                            double ym, double mf, double ymf, double Bu_flat, int j, int i, int v1h,
                                            int v2h) {
    const double DT = 1.0 / 10;
    const double mug = P.mug;
    asm volatile("" : "+d"(mf), "+d"(ymf));           // loop invariants stay in registers (not re-derived per sub-step)
#pragma unroll 1
    for (int s = 0; s < 10; ++s) {
        double Bu = Bu_flat;                                                 // B(x) u, models.py:109-112
        const int vh = __double2hiint(v);
        if (!(vh > v1h && vh < v2h)) {                // conservative: strictly inside by the high words
            // rare path.  It works on an opaque copy of v, so that its FP64 comparisons cannot be hoisted into the
            // hot loop (the compiler did exactly that: the integer test bought nothing)
            double vs = v;
            asm volatile("" : "+d"(vs));
            const double v1 = P.tr_v[j][1], v2 = P.tr_v[j][2];
            if (!(vs >= v1 && vs <= v2)) {
                const double v0 = P.tr_v[j][0], v3 = P.tr_v[j][3];
                if (!(vs > v0 && vs < v3))
                    return ((vs < P.tr_v[0][0] || vs > P.tr_v[5][3]) ? 1 : 3) | (i << 8) | (s << 16);
                const double t0 = P.tr_t[j][0], t1 = P.tr_t[j][1], t2 = P.tr_t[j][2];
                double Bv;
                if (vs < v1) Bv = div_by(div_by(vs - v0, v1 - v0, P.inv_rise[j]) * (t1 - t0) + t0, m, ym);
                else Bv = div_by(t1 - div_by(vs - v2, v3 - v2, P.inv_fall[j]) * (t1 - t2), m, ym);
                Bu = Bv * ui;
            }
        }
        const double vv = FAST ? v * v : (P.fric_pow2 ? v * v : P.c_fric * (v * v));
        const double Av = -div_by(vv, mf, ymf) - mug;                        // models.py:99-107
        const double va = FAST ? v : v + zu;
        const double pn = p + DT * va;
        const double vn = v + DT * (Av + Bu);
        p = pn; v = vn;
    }
    return 0;
}

__global__ void __launch_bounds__(ROLLOUT_BLOCK, 8)
rollout_kernel(const __grid_constant__ RolloutParams P, int64_t batch, const double* __restrict__ x,
               const double* __restrict__ u, const int32_t* __restrict__ gear,
               const double* __restrict__ mass, const double* __restrict__ leader,
               double* __restrict__ x_out, double* __restrict__ cost, uint8_t* __restrict__ viol,
               int32_t* __restrict__ err) {
    __shared__ double s_follow[ROLLOUT_BLOCK];
    __shared__ double s_effort[ROLLOUT_BLOCK];
    __shared__ double s_lead[ROLLOUT_BLOCK];       // one per scenario of the CTA
    __shared__ int32_t s_err[ROLLOUT_BLOCK];
    __shared__ int32_t s_scen_err[ROLLOUT_BLOCK];
    __shared__ uint8_t s_gap[ROLLOUT_BLOCK];

    const int n = P.n;
    const int S = P.scen_per_block;
    const int tid = threadIdx.x;
    const int ls = tid / n;                    // local scenario
    const int i = tid - ls * n;                // vehicle
    const int64_t scen = (int64_t)blockIdx.x * S + ls;
    const bool active = (ls < S) && (scen < batch);
    const bool quad = P.flags & 1, real_ref = P.flags & 2, mass_per = P.flags & 4;

    double p = 0, v = 0, ui = 0, m = P.default_mass;
    int g = 0;
    int32_t myerr = 0;
    if (active) {
        const int64_t pair = scen * n + i;
        const double2 pv = reinterpret_cast<const double2*>(x)[pair];
        p = pv.x; v = pv.y;
        ui = u[pair];
        if (mass) m = mass_per ? mass[pair] : mass[i];
        if (gear) {
            g = gear[pair];
        } else {                                // PwaGearVehicle.get_gear_from_velocity (models.py:494-515)
            g = 1;
            if (v >= P.lim[4]) g = 6;
            else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (v >= P.lim[k] && v < P.lim[k + 1]) g = k + 2;
            }
        }
        // ---- stage cost terms on the pre-step state (env.py:126-180) ----
        double follow = 0.0;
        uint8_t gapflag = 0;
        if (i > 0) {
            const double2 pr = reinterpret_cast<const double2*>(x)[pair - 1];
            const double sp0 = (P.t0 != 0.0) ? (-P.t0 * v + (-P.d0)) : -P.d0;
            follow = cost2(p - pr.x - sp0, v - pr.y - 0.0, 1.0, 0.1, quad);
            gapflag = (pr.x - p < P.d_safe) ? 1 : 0;          // env.py:173-176
        }
        s_follow[tid] = follow;
        s_effort[tid] = quad ? (ui * 1.0) * ui : fabs(1.0 * ui);
        if (!real_ref) {
            if (i == P.leader_index) {
                const double l0 = leader[2 * scen], l1 = leader[2 * scen + 1];
                s_lead[ls] = cost2(p - l0, v - l1, 1.0, 0.1, quad);
            }
        } else if (i == 0) {
            const double l0 = leader[2 * scen], l1 = leader[2 * scen + 1];
            const double sp0 = -P.t0 * v - P.d0;
            s_lead[ls] = cost2(p - l0 - sp0, v - l1 - 0.0, 1.0, 0.1, quad);
            if (l0 - p < P.d_safe) gapflag = 1;                // env.py:165-172
        }
        s_gap[tid] = gapflag;

        // ---- ten explicit Euler sub-steps with the gear held (models.py:236-257, Q5) ----
        const double vlo = P.tr_v[0][0], vhi = P.tr_v[5][3];
        if (g < 1 || g > 6) {
            // Vehicle.step's own velocity check fires first (models.py:119-122)
            myerr = (v < vlo || v > vhi) ? (1 | (i << 8)) : (2 | (i << 8));
        } else {
            const int j = g - 1;
            const double v1 = P.tr_v[j][1], v2 = P.tr_v[j][2];     // flat part of the traction curve
            const double ym = __drcp_rn(m);                  // RN(1/m), once per vehicle-step
            const double Bv_flat = div_by(P.tr_t[j][1], m, ym);
            // (c_fric v^2)/m with c_fric a power of two equals v^2/(m/c_fric) bit for bit (scaling
            // by 2^k commutes with rounding), which saves the multiplication by c_fric
            const double mf = P.fric_pow2 ? m * P.inv_c_fric : m, ymf = P.fric_pow2 ? ym * P.c_fric : ym;
            const double zu = 0.0 * ui;                      // the reference adds 0*u to v (models.py:124)
            double Bu_flat = Bv_flat * ui;                   // loop invariant on the flat part of the curve
            asm volatile("" : "+d"(Bu_flat));                // ... kept as a value, not re-multiplied per sub-step
            const int v1h = __double2hiint(v1), v2h = __double2hiint(v2);
            if (zu == 0.0 && P.fric_pow2) myerr = substeps<true>(P, p, v, ui, zu, m, ym, mf, ymf, Bu_flat, j, i, v1h, v2h);
            else myerr = substeps<false>(P, p, v, ui, zu, m, ym, mf, ymf, Bu_flat, j, i, v1h, v2h);
        }
        s_err[tid] = myerr;
    }
    __syncthreads();
    // ---- per-scenario reduction in the reference's summation order ----
    if (tid < S) {
        const int64_t sc = (int64_t)blockIdx.x * S + tid;
        if (sc < batch) {
            const int base = tid * n;
            double s1 = 0.0, s2 = 0.0;
            int32_t e = 0;
            uint8_t vf = 0;
            for (int k = 0; k < n; ++k) {
                if (k > 0) s1 += s_follow[base + k];
                s2 += s_effort[base + k];
                vf |= s_gap[base + k];
                const int32_t ek = s_err[base + k];
                if (ek) {
                    // first exception in (substep, vehicle) order
                    const int32_t key = ((ek >> 16) << 16) | (((ek >> 8) & 0xff) << 8) | (ek & 0xff);
                    const int32_t cur = ((e >> 16) << 16) | (((e >> 8) & 0xff) << 8) | (e & 0xff);
                    if (e == 0 || key < cur) e = ek;
                }
            }
            double c = s_lead[tid];
            c += s1;
            c += s2;
            cost[sc] = c;                   // the Q_du = 0 variation term adds exactly 0
            viol[sc] = vf;
            err[sc] = e;
            s_scen_err[tid] = e;
        }
    }
    __syncthreads();
    if (active) {
        const int64_t pair = scen * n + i;
        double2 o;
        if (s_scen_err[ls]) { o.x = nan(""); o.y = nan(""); }
        else { o.x = p; o.y = v; }
        reinterpret_cast<double2*>(x_out)[pair] = o;
    }
}

cudaError_t launch_rollout(const RolloutParams& P, int64_t batch, const double* x, const double* u,
                           const int32_t* gear, const double* mass, const double* leader, double* x_out,
                           double* cost, uint8_t* viol, int32_t* err, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    const int S = P.scen_per_block;
    const unsigned grid = (unsigned)((batch + S - 1) / S);
    rollout_kernel<<<grid, ROLLOUT_BLOCK, 0, stream>>>(P, batch, x, u, gear, mass, leader, x_out, cost, viol,
                                                      err);
    return cudaGetLastError();
}

}  // namespace hvp
