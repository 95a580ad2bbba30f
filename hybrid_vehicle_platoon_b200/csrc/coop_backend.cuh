// coop_backend.cuh -- the tiny SPMD vocabulary the cooperative MIQP solver (coop_core.cuh) is
// written in: a GROUP of G lanes owns one problem; "D" is one double per lane, "I" one int per
// lane, "Bm" one predicate per lane; plain double/int/bool are group-uniform values.
//
//   DevBK<G>  : the product.  D = double in a register, group collectives are warp shuffles
//               restricted to the group's lanes (G = 8 -> four problems per warp).
//   HostBK<G> : test harness only (tests/host_harness): D = array of G doubles, collectives are
//               loops that reproduce the butterfly order of the device shuffles, so the host run
//               is an emulation of the device algorithm, used to check it against the oracle
//               where no GPU exists.  The product never runs this backend.
#pragma once
#include <math.h>
#include <stdint.h>

namespace hvp {

#if defined(__CUDACC__)
template <int G_>
struct DevBK {
    static constexpr int G = G_;
    using D = double;
    using I = int;
    using Bm = bool;
    unsigned gmask;
    int ln;
    __device__ __forceinline__ DevBK() {
        const int lane = threadIdx.x & 31;
        ln = lane % G;
        gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - ln));
    }
    __device__ __forceinline__ I lane() const { return ln; }
    __device__ __forceinline__ D splat(double v) const { return v; }
    __device__ __forceinline__ I splati(int v) const { return v; }
    __device__ __forceinline__ D todouble(I v) const { return (double)v; }
    __device__ __forceinline__ D sel(Bm c, D a, D b) const { return c ? a : b; }
    __device__ __forceinline__ I seli(Bm c, I a, I b) const { return c ? a : b; }
    // quotients through a MUFU-seeded reciprocal (~1 ulp): the IEEE division sequence is ~25 instructions with a slow
    // path, and this kernel is a single dependent chain per group
    static __device__ __forceinline__ double rcp(double v) {
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
        r = r * (2.0 - v * r);
        r = r * (2.0 - v * r);
        return r;
    }
    __device__ __forceinline__ D div(D a, D b) const { return a * rcp(b); }
    __device__ __forceinline__ double sdiv(double a, double b) const { return a * rcp(b); }
    __device__ __forceinline__ D dmax(D a, D b) const { return fmax(a, b); }
    __device__ __forceinline__ D dmin(D a, D b) const { return fmin(a, b); }
    __device__ __forceinline__ double bcast(D v, int src) const { return __shfl_sync(gmask, v, src, G); }
    __device__ __forceinline__ int bcasti(I v, int src) const { return __shfl_sync(gmask, v, src, G); }
    __device__ __forceinline__ D shfl(D v, I src) const { return __shfl_sync(gmask, v, src, G); }
    __device__ __forceinline__ D up1(D v, double fill) const {
        const D t = __shfl_up_sync(gmask, v, 1, G);
        return ln == 0 ? fill : t;
    }
    __device__ __forceinline__ D dn1(D v) const {
        const D t = __shfl_down_sync(gmask, v, 1, G);
        return ln == G - 1 ? 0.0 : t;
    }
    __device__ __forceinline__ I dn1i(I v) const {
        const I t = __shfl_down_sync(gmask, v, 1, G);
        return ln == G - 1 ? 0 : t;
    }
    __device__ __forceinline__ double gsum(D v) const {
#pragma unroll
        for (int o = G / 2; o; o >>= 1) v += __shfl_xor_sync(gmask, v, o, G);
        return v;
    }
    __device__ __forceinline__ D scan_excl(D v) const {       // exclusive prefix sum over lanes
        D s = v;
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
            const D t = __shfl_up_sync(gmask, s, o, G);
            if (ln >= o) s += t;
        }
        return up1(s, 0.0);
    }
    // (max value, its id); ties -> smaller id.  All lanes return the same pair.
    __device__ __forceinline__ void gmax_arg(D v, I id, double& bv, int& bi) const {
#pragma unroll
        for (int o = G / 2; o; o >>= 1) {
            const D ov = __shfl_xor_sync(gmask, v, o, G);
            const I oi = __shfl_xor_sync(gmask, id, o, G);
            if (ov > v || (ov == v && oi < id)) { v = ov; id = oi; }
        }
        bv = v; bi = id;
    }
    __device__ __forceinline__ void gmin_arg(D v, I id, double& bv, int& bi) const {
#pragma unroll
        for (int o = G / 2; o; o >>= 1) {
            const D ov = __shfl_xor_sync(gmask, v, o, G);
            const I oi = __shfl_xor_sync(gmask, id, o, G);
            if (ov < v || (ov == v && oi < id)) { v = ov; id = oi; }
        }
        bv = v; bi = id;
    }
    __device__ __forceinline__ Bm bit128(uint64_t lo, uint64_t hi, I id) const {
        return id < 64 ? (lo >> id) & 1u : (hi >> (id - 64)) & 1u;
    }
    __device__ __forceinline__ Bm bit32(uint32_t m, I id) const { return (m >> id) & 1u; }
    __device__ __forceinline__ I bits3(uint64_t pk, I idx) const { return (int)((pk >> (3 * idx)) & 7u); }
    __device__ __forceinline__ D lookup(const double* tbl, I idx) const { return tbl[idx]; }
    __device__ __forceinline__ D ld(const double* p, I idx, Bm ok) const { return ok ? p[idx] : 0.0; }
    __device__ __forceinline__ void st(double* p, I idx, D v, Bm ok) const { if (ok) p[idx] = v; }
    __device__ __forceinline__ void sti(int32_t* p, I idx, I v, Bm ok) const { if (ok) p[idx] = v; }
};
#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------
// Host emulation (tests only).
// ---------------------------------------------------------------------------------------------
template <int G>
struct HVec {
    double v[G];
};
template <int G>
struct HInt {
    int v[G];
};
template <int G>
struct HMask {
    bool v[G];
};

#define HVP_HOP(op)                                                                                  \
    template <int G> inline HVec<G> operator op(const HVec<G>& a, const HVec<G>& b) {                \
        HVec<G> r; for (int i = 0; i < G; ++i) r.v[i] = a.v[i] op b.v[i]; return r; }                \
    template <int G> inline HVec<G> operator op(const HVec<G>& a, double b) {                        \
        HVec<G> r; for (int i = 0; i < G; ++i) r.v[i] = a.v[i] op b; return r; }                     \
    template <int G> inline HVec<G> operator op(double a, const HVec<G>& b) {                        \
        HVec<G> r; for (int i = 0; i < G; ++i) r.v[i] = a op b.v[i]; return r; }
HVP_HOP(+) HVP_HOP(-) HVP_HOP(*) HVP_HOP(/)
#undef HVP_HOP
template <int G> inline HVec<G> operator-(const HVec<G>& a) { HVec<G> r; for (int i = 0; i < G; ++i) r.v[i] = -a.v[i]; return r; }
template <int G> inline HVec<G>& operator+=(HVec<G>& a, const HVec<G>& b) { for (int i = 0; i < G; ++i) a.v[i] += b.v[i]; return a; }
template <int G> inline HVec<G>& operator-=(HVec<G>& a, const HVec<G>& b) { for (int i = 0; i < G; ++i) a.v[i] -= b.v[i]; return a; }
#define HVP_HCMP(op)                                                                                 \
    template <int G> inline HMask<G> operator op(const HVec<G>& a, const HVec<G>& b) {               \
        HMask<G> r; for (int i = 0; i < G; ++i) r.v[i] = a.v[i] op b.v[i]; return r; }               \
    template <int G> inline HMask<G> operator op(const HVec<G>& a, double b) {                       \
        HMask<G> r; for (int i = 0; i < G; ++i) r.v[i] = a.v[i] op b; return r; }                    \
    template <int G> inline HMask<G> operator op(const HInt<G>& a, const HInt<G>& b) {               \
        HMask<G> r; for (int i = 0; i < G; ++i) r.v[i] = a.v[i] op b.v[i]; return r; }               \
    template <int G> inline HMask<G> operator op(const HInt<G>& a, int b) {                          \
        HMask<G> r; for (int i = 0; i < G; ++i) r.v[i] = a.v[i] op b; return r; }
HVP_HCMP(<) HVP_HCMP(>) HVP_HCMP(<=) HVP_HCMP(>=) HVP_HCMP(==) HVP_HCMP(!=)
#undef HVP_HCMP
template <int G> inline HMask<G> operator&(const HMask<G>& a, const HMask<G>& b) { HMask<G> r; for (int i = 0; i < G; ++i) r.v[i] = a.v[i] && b.v[i]; return r; }
template <int G> inline HMask<G> operator|(const HMask<G>& a, const HMask<G>& b) { HMask<G> r; for (int i = 0; i < G; ++i) r.v[i] = a.v[i] || b.v[i]; return r; }
template <int G> inline HMask<G> operator&(const HMask<G>& a, bool b) { HMask<G> r; for (int i = 0; i < G; ++i) r.v[i] = a.v[i] && b; return r; }
template <int G> inline HMask<G> operator|(const HMask<G>& a, bool b) { HMask<G> r; for (int i = 0; i < G; ++i) r.v[i] = a.v[i] || b; return r; }
template <int G> inline HMask<G> operator!(const HMask<G>& a) { HMask<G> r; for (int i = 0; i < G; ++i) r.v[i] = !a.v[i]; return r; }
template <int G> inline HInt<G> operator+(const HInt<G>& a, int b) { HInt<G> r; for (int i = 0; i < G; ++i) r.v[i] = a.v[i] + b; return r; }
template <int G> inline HInt<G> operator-(const HInt<G>& a, int b) { HInt<G> r; for (int i = 0; i < G; ++i) r.v[i] = a.v[i] - b; return r; }
template <int G> inline HInt<G> operator*(const HInt<G>& a, int b) { HInt<G> r; for (int i = 0; i < G; ++i) r.v[i] = a.v[i] * b; return r; }

template <int G_>
struct HostBK {
    static constexpr int G = G_;
    using D = HVec<G>;
    using I = HInt<G>;
    using Bm = HMask<G>;
    I lane() const { I r; for (int i = 0; i < G; ++i) r.v[i] = i; return r; }
    D splat(double v) const { D r; for (int i = 0; i < G; ++i) r.v[i] = v; return r; }
    I splati(int v) const { I r; for (int i = 0; i < G; ++i) r.v[i] = v; return r; }
    D todouble(const I& a) const { D r; for (int i = 0; i < G; ++i) r.v[i] = (double)a.v[i]; return r; }
    D sel(const Bm& c, const D& a, const D& b) const { D r; for (int i = 0; i < G; ++i) r.v[i] = c.v[i] ? a.v[i] : b.v[i]; return r; }
    I seli(const Bm& c, const I& a, const I& b) const { I r; for (int i = 0; i < G; ++i) r.v[i] = c.v[i] ? a.v[i] : b.v[i]; return r; }
    D div(const D& a, const D& b) const { D r; for (int i = 0; i < G; ++i) r.v[i] = a.v[i] / b.v[i]; return r; }
    double sdiv(double a, double b) const { return a / b; }
    D dmax(const D& a, const D& b) const { D r; for (int i = 0; i < G; ++i) r.v[i] = fmax(a.v[i], b.v[i]); return r; }
    D dmin(const D& a, const D& b) const { D r; for (int i = 0; i < G; ++i) r.v[i] = fmin(a.v[i], b.v[i]); return r; }
    double bcast(const D& v, int src) const { return v.v[src]; }
    int bcasti(const I& v, int src) const { return v.v[src]; }
    D shfl(const D& v, const I& src) const { D r; for (int i = 0; i < G; ++i) r.v[i] = v.v[((src.v[i] % G) + G) % G]; return r; }
    D up1(const D& v, double fill) const { D r; r.v[0] = fill; for (int i = 1; i < G; ++i) r.v[i] = v.v[i - 1]; return r; }
    D dn1(const D& v) const { D r; r.v[G - 1] = 0.0; for (int i = 0; i + 1 < G; ++i) r.v[i] = v.v[i + 1]; return r; }
    I dn1i(const I& v) const { I r; r.v[G - 1] = 0; for (int i = 0; i + 1 < G; ++i) r.v[i] = v.v[i + 1]; return r; }
    double gsum(D v) const {
        for (int o = G / 2; o; o >>= 1) { D t = v; for (int i = 0; i < G; ++i) v.v[i] = t.v[i] + t.v[i ^ o]; }
        return v.v[0];
    }
    D scan_excl(const D& v) const {
        D s = v;
        for (int o = 1; o < G; o <<= 1) { D t = s; for (int i = o; i < G; ++i) s.v[i] = t.v[i] + t.v[i - o]; }
        return up1(s, 0.0);
    }
    void gmax_arg(D v, I id, double& bv, int& bi) const {
        for (int o = G / 2; o; o >>= 1) {
            D tv = v; I ti = id;
            for (int i = 0; i < G; ++i) {
                const double ov = tv.v[i ^ o]; const int oi = ti.v[i ^ o];
                if (ov > tv.v[i] || (ov == tv.v[i] && oi < ti.v[i])) { v.v[i] = ov; id.v[i] = oi; }
            }
        }
        bv = v.v[0]; bi = id.v[0];
    }
    void gmin_arg(D v, I id, double& bv, int& bi) const {
        for (int o = G / 2; o; o >>= 1) {
            D tv = v; I ti = id;
            for (int i = 0; i < G; ++i) {
                const double ov = tv.v[i ^ o]; const int oi = ti.v[i ^ o];
                if (ov < tv.v[i] || (ov == tv.v[i] && oi < ti.v[i])) { v.v[i] = ov; id.v[i] = oi; }
            }
        }
        bv = v.v[0]; bi = id.v[0];
    }
    Bm bit128(uint64_t lo, uint64_t hi, const I& id) const {
        Bm r; for (int i = 0; i < G; ++i) r.v[i] = id.v[i] < 64 ? (lo >> id.v[i]) & 1u : (hi >> (id.v[i] - 64)) & 1u; return r;
    }
    Bm bit32(uint32_t m, const I& id) const { Bm r; for (int i = 0; i < G; ++i) r.v[i] = (m >> id.v[i]) & 1u; return r; }
    I bits3(uint64_t pk, const I& idx) const { I r; for (int i = 0; i < G; ++i) r.v[i] = (int)((pk >> (3 * idx.v[i])) & 7u); return r; }
    D lookup(const double* tbl, const I& idx) const { D r; for (int i = 0; i < G; ++i) r.v[i] = tbl[idx.v[i]]; return r; }
    D ld(const double* p, const I& idx, const Bm& ok) const { D r; for (int i = 0; i < G; ++i) r.v[i] = ok.v[i] ? p[idx.v[i]] : 0.0; return r; }
    void st(double* p, const I& idx, const D& v, const Bm& ok) const { for (int i = 0; i < G; ++i) if (ok.v[i]) p[idx.v[i]] = v.v[i]; }
    void sti(int32_t* p, const I& idx, const I& v, const Bm& ok) const { for (int i = 0; i < G; ++i) if (ok.v[i]) p[idx.v[i]] = v.v[i]; }
};

}  // namespace hvp
