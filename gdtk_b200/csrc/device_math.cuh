// device_math.cuh -- per-face and per-cell arithmetic of the explicit update (FP64).
//
// Written for sm_100a.  The formulas follow Eilmer 4 (gdtk-uq/gdtk, src/eilmer) and keep
// its evaluation order so that the FMA-free build (namespace eb_strict, -fmad=false) is
// bit-comparable with the reference's x86-64 arithmetic; the throughput build (eb_fast)
// lets ptxas contract a*b+c.  Nothing here is translated D: cells are read-only SoA
// values in registers, faces are computed by the thread that owns them.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "eb200_internal.h"

#ifndef EB_NS
#error "EB_NS must be defined (eb_strict or eb_fast)"
#endif

namespace EB_NS {

template <int NSP>
struct Prim {                       // FlowState of one cell or one side of a face
    double rho, u, p, T, a, vx, vy, vz;
    double massf[NSP], rho_s[NSP];
};

template <int DIM, int NSP>
struct Layout {
    static constexpr int NCQ = (DIM == 3 ? 5 : 4) + (NSP > 1 ? NSP : 0);
    static constexpr int NPRIM = 8 + (NSP > 1 ? 2 * NSP : 0);
    static constexpr int iMass = 0, iXMom = 1, iYMom = 2, iZMom = 3;
    static constexpr int iEnergy = (DIM == 3 ? 4 : 3);
    static constexpr int iSpecies = iEnergy + 1;
};

__device__ __forceinline__ double ldg(const double* p) { return __ldg(p); }

// Division and square root.  The FMA-free build uses IEEE division / sqrt in the reference's
// operation order.  The throughput build (EB_FAST_MATH) replaces a/b by a * rcp(b) with a
// branch-free refined reciprocal (MUFU.RCP64H + 3 DFMA, <= 2 ulp) and lets several
// quotients share one reciprocal; parity with the reference is then within 1e-10, not bitwise.
#ifdef EB_FAST_MATH
__device__ __forceinline__ double eb_rcp(double b)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    // one third-order step: r (1 + e + e^2), e = 1 - b r; the seed is good to ~2^-23
    const double e = fma(-b, r, 1.0);
    return fma(r, fma(e, e, e), r);
}
__device__ __forceinline__ double eb_div(double a, double b) { return a * eb_rcp(b); }
__device__ __forceinline__ double eb_sqrt(double x)
{
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    // two Newton steps on r ~ 1/sqrt(x), then s = x*r with one correction
    double h = 0.5 * r;
    double e = fma(-x * r, h, 0.5); r = fma(r, e, r); h = 0.5 * r;
    e = fma(-x * r, h, 0.5); r = fma(r, e, r);
    double sq = x * r;
    sq = fma(fma(-sq, sq, x), 0.5 * r, sq);
    return (x > 0.0) ? sq : ((x == 0.0) ? 0.0 : sqrt(x));
}
#else
__device__ __forceinline__ double eb_div(double a, double b) { return a / b; }
__device__ __forceinline__ double eb_sqrt(double x) { return sqrt(x); }
#endif

// ---------------------------------------------------------------------------------------
// Face frames

struct Frame {                      // general (n, t1, t2)
    double nx, ny, nz, t1x, t1y, t1z, t2x, t2y, t2z;
};

template <int DIM>
__device__ __forceinline__ void load_frame(Frame& f, const double* __restrict__ g, long long total, long long cf)
{
    f.nx = ldg(g + cf); f.ny = ldg(g + total + cf);
    f.t1x = ldg(g + 3 * total + cf); f.t1y = ldg(g + 4 * total + cf);
    if (DIM == 3) {
        f.nz = ldg(g + 2 * total + cf); f.t1z = ldg(g + 5 * total + cf);
        f.t2x = ldg(g + 6 * total + cf); f.t2y = ldg(g + 7 * total + cf); f.t2z = ldg(g + 8 * total + cf);
    } else {
        f.nz = 0.0; f.t1z = 0.0; f.t2x = 0.0; f.t2y = 0.0; f.t2z = 1.0;
    }
}

// reference src/geom/elements/vector3.d:403-412 / :417-424
template <int DIM>
__device__ __forceinline__ void to_local(const Frame& f, double& x, double& y, double& z)
{
    if (DIM == 3) {
        double a = x * f.nx + y * f.ny + z * f.nz;
        double b = x * f.t1x + y * f.t1y + z * f.t1z;
        double c = x * f.t2x + y * f.t2y + z * f.t2z;
        x = a; y = b; z = c;
    } else {
        // z = 0, n.z = t1.z = 0, t2 = (0,0,1): the dropped terms are exact +0.0.  They are still added because
        // (-0) + (+0) = +0: the sign of a zero velocity component must be the reference's -- efm's
        // erf approximation jumps by 1e-9 between sn = -0 and sn = +0 (copysign, fluxcalc.d:1311)
        double a = x * f.nx + y * f.ny + 0.0;
        double b = x * f.t1x + y * f.t1y + 0.0;
        x = a; y = b;
    }
}
template <int DIM>
__device__ __forceinline__ void to_global(const Frame& f, double& x, double& y, double& z)
{
    if (DIM == 3) {
        double a = x * f.nx + y * f.t1x + z * f.t2x;
        double b = x * f.ny + y * f.t1y + z * f.t2y;
        double c = x * f.nz + y * f.t1z + z * f.t2z;
        x = a; y = b; z = c;
    } else {
        double a = x * f.nx + y * f.t1x;
        double b = x * f.ny + y * f.t1y;
        x = a; y = b;
    }
}

// axis-aligned frames: a signed permutation, exact in floating point
__device__ __forceinline__ double pick(int p, double x, double y, double z) { return p == 0 ? x : (p == 1 ? y : z); }
__device__ __forceinline__ void axis_to_local(const EbAxisFrame& f, double& x, double& y, double& z)
{
    double a = pick(f.perm[0], x, y, z), b = pick(f.perm[1], x, y, z), c = pick(f.perm[2], x, y, z);
    x = f.neg[0] ? -a : a; y = f.neg[1] ? -b : b; z = f.neg[2] ? -c : c;
}
__device__ __forceinline__ void axis_to_global(const EbAxisFrame& f, double& x, double& y, double& z)
{
    double a = f.neg[0] ? -x : x, b = f.neg[1] ? -y : y, c = f.neg[2] ? -z : z;
    double gx = 0.0, gy = 0.0, gz = 0.0;
    if (f.perm[0] == 0) gx = a; else if (f.perm[0] == 1) gy = a; else gz = a;
    if (f.perm[1] == 0) gx = b; else if (f.perm[1] == 1) gy = b; else gz = b;
    if (f.perm[2] == 0) gx = c; else if (f.perm[2] == 1) gy = c; else gz = c;
    x = gx; y = gy; z = gz;
}

// ---------------------------------------------------------------------------------------
// Thermally perfect gas: CEA curves + Newton iteration
// reference src/gas/thermo/cea_thermo_curves.d:56-181, therm_perf_gas_mix_eos.d:61-160,
// src/nm/newton.d:70-132.  Returns false where the reference throws.

__device__ __forceinline__ bool cea_coeffs(const EbCurve& c, double T, double a[9])
{
    const int nb = c.nbreaks;
    if (T < (c.T_breaks[1] - 0.5 * c.T_blends[0])) {
#pragma unroll
        for (int j = 0; j < 9; ++j) a[j] = c.coeffs[0][j];
        return true;
    }
    if (T > (c.T_breaks[nb - 2] + 0.5 * c.T_blends[c.nseg - 2])) {
#pragma unroll
        for (int j = 0; j < 9; ++j) a[j] = c.coeffs[c.nseg - 1][j];
        return true;
    }
    for (int i = 1; i < nb - 1; ++i) {
        double lo = c.T_breaks[i] - 0.5 * c.T_blends[i - 1];
        double hi = c.T_breaks[i] + 0.5 * c.T_blends[i - 1];
        if (T >= lo && T <= hi) {
            double wB = (1. / c.T_blends[i - 1]) * (T - lo);
            double wA = 1.0 - wB;
#pragma unroll
            for (int j = 0; j < 9; ++j) a[j] = wA * c.coeffs[i - 1][j] + wB * c.coeffs[i][j];
            return true;
        }
        if (T > hi && T < (c.T_breaks[i + 1] - 0.5 * c.T_blends[i])) {
#pragma unroll
            for (int j = 0; j < 9; ++j) a[j] = c.coeffs[i][j];
            return true;
        }
    }
    return false;
}

__device__ __forceinline__ bool cea_Cp(const EbCurve& c, double T, double& out)
{
    if (T < c.T_low) { out = c.Cp_low; return true; }
    if (T > c.T_high) { out = c.Cp_high; return true; }
    double a[9];
    if (!cea_coeffs(c, T, a)) return false;
    double Cp_on_R = a[0] / (T * T) + a[1] / T + a[2] + a[3] * T;
    Cp_on_R += a[4] * T * T + a[5] * T * T * T + a[6] * T * T * T * T;
    out = c.R * Cp_on_R;
    return true;
}

// logT is log(T): the reference recomputes it per species with identical result.
__device__ __forceinline__ bool cea_h(const EbCurve& c, double T, double logT, double& out)
{
    if (T < c.T_low) { out = c.h_low - c.Cp_low * (c.T_low - T); return true; }
    if (T > c.T_high) { out = c.h_high + c.Cp_high * (T - c.T_high); return true; }
    double a[9];
    if (!cea_coeffs(c, T, a)) return false;
    double h_on_RT = -a[0] / T + a[1] * logT + a[2] * T + a[3] * T * T / 2.0;
    h_on_RT += a[4] * T * T * T / 3.0 + a[5] * T * T * T * T / 4.0 + a[6] * T * T * T * T * T / 5.0 + a[7];
    out = c.R * h_on_RT;
    return true;
}

template <int NSP>
__device__ __forceinline__ bool tpg_energy(const EbGas* __restrict__ g, const double* massf, double T, double& u)
{
    double logT = log(T);
    double result = 0.0;
    bool ok = true;
#pragma unroll
    for (int i = 0; i < NSP; ++i) {
        double h;
        ok &= cea_h(g->curves[i], T, logT, h);
        double e = h - g->Rsp[i] * T;
        result += massf[i] * e;
    }
    u = result;
    return ok;
}

template <int NSP>
__device__ __forceinline__ bool tpg_dzdT(const EbGas* __restrict__ g, const double* massf, double T, double& out)
{
    double result = 0.0;
    bool ok = true;
#pragma unroll
    for (int i = 0; i < NSP; ++i) {
        double cp;
        ok &= cea_Cp(g->curves[i], T, cp);
        result += massf[i] * (cp - g->Rsp[i]);
    }
    out = -1.0 * result;
    return ok;
}

// nm.newton.solve; Q.T / Q.u are written by every function evaluation like the reference's zeroFn.
// rc: 0 root found, 1 NumericalMethodException, 2 GasModelException inside f, 3 (evaluator) use the other evaluator
// Eval: bool energy(T, u) and bool energy_and_slope(T, u, dfdT) with f = e_tgt - u(T), dfdT = -Cv(T)
template <class Eval>
__device__ __forceinline__ int newton_rtsafe(Eval& ev, double e_tgt, double x0, double xMin, double xMax, double tol,
                                             double& root, double& Qu)
{
    double xL = xMin, xH = xMax, u;
    bool ok = ev.energy(xL, u); Qu = u;
    double fL = e_tgt - u;
    ok &= ev.energy(xH, u); Qu = u;
    double fH = e_tgt - u;
    if (!ok) return ev.failure_code();
    if ((fL > 0.0 && fH > 0.0) || (fL < 0.0 && fH < 0.0)) return 1;
    if (fL == 0.0) { root = xMin; return 0; }
    if (fH == 0.0) { root = xMax; return 0; }
    if (fL < 0.0) { xL = xMin; xH = xMax; } else { xH = xMin; xL = xMax; }
    double rts = x0;
    double dxold = (xMax - xMin);
    double dx = dxold;
    double f0, df0;
    ok &= ev.energy_and_slope(rts, u, df0); Qu = u;
    f0 = e_tgt - u;
    if (!ok) return ev.failure_code();
    for (int j = 0; j < 30; ++j) {
        if ((((rts - xH) * df0 - f0) * ((rts - xL) * df0 - f0) > 0.0) || (fabs(2.0 * f0) > fabs(dxold * df0))) {
            dxold = dx;
            dx = 0.5 * (xH - xL);
            rts = xL + dx;
            if (xL == rts) { root = rts; return 0; }
        } else {
            dxold = dx;
            dx = eb_div(f0, df0);
            double tmp = rts;
            rts -= dx;
            if (tmp == rts) { root = rts; return 0; }
        }
        if (fabs(dx) < tol) { root = rts; return 0; }
        ok &= ev.energy_and_slope(rts, u, df0); Qu = u;
        f0 = e_tgt - u;
        if (!ok) return ev.failure_code();
        if (f0 < 0.0) xL = rts; else xH = rts;
    }
    return 1;
}

// species by species, in the reference's order of operations
template <int NSP>
struct SpeciesEval {
    const EbGas* __restrict__ g;
    const double* massf;
    __device__ __forceinline__ bool energy(double T, double& u) const { return tpg_energy<NSP>(g, massf, T, u); }
    __device__ __forceinline__ bool energy_and_slope(double T, double& u, double& df) const
    {
        bool ok = tpg_energy<NSP>(g, massf, T, u);
        ok &= tpg_dzdT<NSP>(g, massf, T, df);
        return ok;
    }
    __device__ __forceinline__ int failure_code() const { return 2; }
};

template <int NSP>
__device__ __forceinline__ int tpg_newton(const EbGas* __restrict__ g, const double* massf, double e_tgt, double x0,
                                          double xMin, double xMax, double tol, double& root, double& Qu)
{
    SpeciesEval<NSP> ev{ g, massf };
    return newton_rtsafe(ev, e_tgt, x0, xMin, xMax, tol, root, Qu);
}

#ifdef EB_FAST_MATH
// Throughput build: the mixture as one polynomial.  With common break points the NASA/CEA forms
// (cea_thermo_curves.d:56-181) summed over the species give, with A_k = sum_i massf_i R_i a_ik,
//   u(T)  = -A0/T + A1 ln T + (A2 - Rmix) T + A3 T^2/2 + A4 T^3/3 + A5 T^4/4 + A6 T^5/5 + A7
//   Cv(T) =  A0/T^2 + A1/T + (A2 - Rmix) + A3 T + A4 T^2 + A5 T^3 + A6 T^4
// inside a segment; in a blend zone the coefficients of the two segments are mixed with the
// reference's weights.  Outside [T_low, T_high] (linear extrapolation there) the evaluator
// says so and the caller takes the species-by-species route.
template <int NSP>
struct MixEval {
    const EbGas* __restrict__ g;
    const double* massf;
    double A[8], B[8];          // lower / upper segment of the cached region
    double Rmix;
    int region;                 // 2 s: inside segment s; 2 i - 1: blend zone around break i; -1: nothing cached
    bool left;                  // an evaluation fell outside the curves
    __device__ __forceinline__ void init(const EbGas* __restrict__ g_, const double* massf_)
    {
        g = g_; massf = massf_; region = -1; left = false;
        double r = 0.0;
#pragma unroll
        for (int i = 0; i < NSP; ++i) r += massf[i] * g->Rsp[i];
        Rmix = r;
    }
    __device__ __forceinline__ void mix(int seg, double* out) const
    {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double a = 0.0;
#pragma unroll
            for (int i = 0; i < NSP; ++i) a += massf[i] * g->RA[seg][k][i];
            out[k] = a;
        }
    }
    // region of T as cea_coeffs finds it, wB = weight of the upper segment
    __device__ __forceinline__ int classify(double T, double& wB) const
    {
        const EbCurve& c = g->curves[0];
        const int nb = c.nbreaks;
        wB = 0.0;
        if (T < c.T_low || T > c.T_high) return -1;
        if (T < (c.T_breaks[1] - 0.5 * c.T_blends[0])) return 0;
        if (T > (c.T_breaks[nb - 2] + 0.5 * c.T_blends[c.nseg - 2])) return 2 * (c.nseg - 1);
        for (int i = 1; i < nb - 1; ++i) {
            const double lo = c.T_breaks[i] - 0.5 * c.T_blends[i - 1], hi = c.T_breaks[i] + 0.5 * c.T_blends[i - 1];
            if (T >= lo && T <= hi) { wB = (1. / c.T_blends[i - 1]) * (T - lo); return 2 * i - 1; }
            if (T > hi && T < (c.T_breaks[i + 1] - 0.5 * c.T_blends[i])) return 2 * i;
        }
        return -1;
    }
    __device__ __forceinline__ bool energy_and_slope(double T, double& u, double& df)
    {
        double wB;
        const int r = classify(T, wB);
        if (r < 0) { left = true; u = 0.0; df = -1.0; return false; }
        if (r != region) {
            region = r;
            if (r & 1) { mix((r - 1) >> 1, A); mix((r + 1) >> 1, B); }
            else mix(r >> 1, A);
        }
        double a[8];
        if (r & 1) {
            const double wA = 1.0 - wB;
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = wA * A[k] + wB * B[k];
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = A[k];
        }
        const double rT = eb_rcp(T), lnT = log(T), a2 = a[2] - Rmix;
        u = -a[0] * rT + a[1] * lnT + a[7] + T * (a2 + T * (0.5 * a[3] + T * ((1.0 / 3.0) * a[4] + T * (0.25 * a[5] + T * (0.2 * a[6])))));
        df = -(rT * (a[0] * rT + a[1]) + a2 + T * (a[3] + T * (a[4] + T * (a[5] + T * a[6]))));
        return true;
    }
    __device__ __forceinline__ bool energy(double T, double& u) { double df; return energy_and_slope(T, u, df); }
    __device__ __forceinline__ int failure_code() const { return left ? 3 : 2; }
};
#endif

// gmodel.update_thermo_from_rhou: T is the starting guess on entry (TPG).  Returns false where the
// reference throws GasModelException.
template <int GASM, int NSP>
__device__ __forceinline__ bool thermo_from_rhou(const EbGas* __restrict__ g, Prim<NSP>& Q)
{
    if (GASM == EB200_GAS_IDEAL) {          // src/gas/ideal_gas.d:98-106
        if (Q.u <= 0.0 || Q.rho <= 0.0) return false;
        Q.T = Q.u * g->Cvinv;
        Q.p = Q.rho * g->Rgas * Q.T;
        return true;
    } else {                                // therm_perf_gas_mix_eos.d:68-160 + perf_gas_mix_eos.d:43-49
        double Tsave = Q.T, e_tgt = Q.u;
        double T1 = fmax(Q.T - 0.5 * 1000.0, 10.0);
        double T2 = T1 + 1000.0;
        double root, Qu = Q.u;
        int rc = 3;
#ifdef EB_FAST_MATH
        if (g->uniform_curves) {
            MixEval<NSP> ev;
            ev.init(g, Q.massf);
            rc = newton_rtsafe(ev, e_tgt, Tsave, T1, T2, 1.0e-6, root, Qu);
            // (the second, wide bracket [10, 100000] K leaves the curves at both ends: species route)
        }
#endif
        if (rc == 3 || rc == 1) {
            Qu = Q.u;
            rc = tpg_newton<NSP>(g, Q.massf, e_tgt, Tsave, T1, T2, 1.0e-6, root, Qu);
            if (rc == 1) rc = tpg_newton<NSP>(g, Q.massf, e_tgt, Tsave, 10.0, 100000.0, 1.0e-6, root, Qu);
        }
        if (rc == 2) { Q.u = Qu; return false; }
        if (rc == 1) { Q.T = Tsave; tpg_energy<NSP>(g, Q.massf, Tsave, Q.u); return false; }
        Q.T = root; Q.u = Qu;
        double Rmix = 0.0;
#pragma unroll
        for (int i = 0; i < NSP; ++i) Rmix += Q.massf[i] * g->Rsp[i];
        Q.p = Q.rho * Rmix * Q.T;
        return true;
    }
}

template <int GASM, int NSP>
__device__ __forceinline__ bool thermo_from_rhoT(const EbGas* __restrict__ g, Prim<NSP>& Q)
{
    if (GASM == EB200_GAS_IDEAL) {          // ideal_gas.d:107-115
        if (Q.T <= 0.0 || Q.rho <= 0.0) return false;
        Q.p = Q.rho * g->Rgas * Q.T;
        Q.u = g->Cv * Q.T;
        return true;
    } else {
        bool ok = tpg_energy<NSP>(g, Q.massf, Q.T, Q.u);
        double Rmix = 0.0;
#pragma unroll
        for (int i = 0; i < NSP; ++i) Rmix += Q.massf[i] * g->Rsp[i];
        Q.p = Q.rho * Rmix * Q.T;
        return ok;
    }
}

// ideal_gas.d:89-97 / therm_perf_gas.d:230-234 update_thermo_from_pT
template <int GASM, int NSP>
__device__ __forceinline__ bool thermo_from_pT(const EbGas* __restrict__ g, Prim<NSP>& Q)
{
    if (GASM == EB200_GAS_IDEAL) {
        if (Q.T <= 0.0 || Q.p <= 0.0) return false;
        Q.rho = eb_div(Q.p, (Q.T * g->Rgas));
        Q.u = g->Cv * Q.T;
        return true;
    } else {
        double Rmix = 0.0;
#pragma unroll
        for (int i = 0; i < NSP; ++i) Rmix += Q.massf[i] * g->Rsp[i];
        const double denom = Rmix * Q.T;
        Q.rho = eb_div(Q.p, denom);
        return tpg_energy<NSP>(g, Q.massf, Q.T, Q.u);
    }
}

// ideal_gas.d:116-124 / therm_perf_gas.d:245-249 update_thermo_from_rhop
template <int GASM, int NSP>
__device__ __forceinline__ bool thermo_from_rhop(const EbGas* __restrict__ g, Prim<NSP>& Q)
{
    if (GASM == EB200_GAS_IDEAL) {
        if (Q.p <= 0.0 || Q.rho <= 0.0) return false;
        Q.T = eb_div(Q.p, (Q.rho * g->Rgas));
        Q.u = g->Cv * Q.T;
        return true;
    } else {
        double Rmix = 0.0;
#pragma unroll
        for (int i = 0; i < NSP; ++i) Rmix += Q.massf[i] * g->Rsp[i];
        Q.T = eb_div(Q.p, (Rmix * Q.rho));
        return tpg_energy<NSP>(g, Q.massf, Q.T, Q.u);
    }
}

template <int GASM, int NSP>
__device__ __forceinline__ bool sound_speed(const EbGas* __restrict__ g, Prim<NSP>& Q)
{
    if (GASM == EB200_GAS_IDEAL) {          // ideal_gas.d:136-143
        if (Q.T <= 0.0) return false;
        Q.a = eb_sqrt(g->gamma * g->Rgas * Q.T);
        return true;
    } else {                                // therm_perf_gas.d:394-430, gas_model.d:205
#ifdef EB_FAST_MATH
        if (g->uniform_curves) {
            MixEval<NSP> ev;
            ev.init(g, Q.massf);
            double u, df;
            if (ev.energy_and_slope(Q.T, u, df)) {
                const double Cv = -df, Cp = Cv + ev.Rmix;
                Q.a = eb_sqrt(Cp * eb_rcp(Cv) * (ev.Rmix * Q.T));
                return true;
            }
        }
#endif
        double Cp = 0.0, Cv = 0.0, R = 0.0;
        bool ok = true;
        double cps[NSP];
#pragma unroll
        for (int i = 0; i < NSP; ++i) { ok &= cea_Cp(g->curves[i], Q.T, cps[i]); Cp += Q.massf[i] * cps[i]; }
#pragma unroll
        for (int i = 0; i < NSP; ++i) Cv += Q.massf[i] * (cps[i] - g->Rsp[i]);
#pragma unroll
        for (int i = 0; i < NSP; ++i) R += Q.massf[i] * g->Rsp[i];
        double gam = Cp / Cv;
        Q.a = sqrt(gam * (R * Q.T));
        return ok;
    }
}

// gas_model.d:373-408 scale_mass_fractions(massf, 0.0, 0.1)
template <int NSP>
__device__ __forceinline__ bool scale_mass_fractions(double* massf)
{
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < NSP; ++i) { massf[i] = massf[i] >= 0.0 ? massf[i] : 0.0; sum += massf[i]; }
    if (fabs(sum - 1.0) > 0.1) return false;
    if (fabs(sum - 1.0) > 0.0) {
#pragma unroll
        for (int i = 0; i < NSP; ++i) massf[i] /= sum;
    }
    return true;
}

// ---------------------------------------------------------------------------------------
// Reconstruction: onedinterp.d:357-384 interp_l2r2_scalar, limiters.d:43-51 clip_to_limits

__device__ __forceinline__ double clip_to_limits(double q, double A, double B)
{
    const double lower = (A <= B) ? A : B;
    const double upper = (A > B) ? A : B;
    const double qc = (q > lower) ? q : lower;
    return (qc <= upper) ? qc : upper;
}

// lmr: van Albada's epsilon scaled by the local magnitude and the cell spacing (lmr/onedinterp.d:147-151)
__device__ __forceinline__ void interp_scalar(const EbWeights& w, bool limiter, bool clip, double eps,
                                              double qL1, double qL0, double qR0, double qR1,
                                              double& qL, double& qR, bool lmr = false)
{
    double delLminus = (qL0 - qL1) * w.two_over_L0L1;
    double del = (qR0 - qL0) * w.two_over_R0L0;
    double delRplus = (qR1 - qR0) * w.two_over_R1R0;
    double sL = 1.0, sR = 1.0;
    if (lmr) {
        const double qqL = fmax(1e-12, fabs(qL0));
        const double qqR = fmax(1e-12, fabs(qR0));
        const double qq = fmax(qqL, qqR);
        eps = qq * eps * w.two_over_R0L0;
    }
    if (limiter) {
        sL = eb_div(delLminus * del + fabs(delLminus * del) + eps, delLminus * delLminus + del * del + eps);
        sR = eb_div(del * delRplus + fabs(del * delRplus) + eps, del * del + delRplus * delRplus + eps);
    }
    qL = qL0 + sL * w.aL0 * (del * w.two_L0_plus_L1 + delLminus * w.lenR0);
    qR = qR0 - sR * w.aR0 * (delRplus * w.lenL0 + del * w.two_R0_plus_R1);
    if (clip) {
        qL = clip_to_limits(qL, qL0, qR0);
        qR = clip_to_limits(qR, qL0, qR0);
    }
}

// onedinterp.d:338-354 l2r2_prepare
__device__ __forceinline__ void l2r2_prepare(EbWeights& w, double lenL1, double lenL0, double lenR0, double lenR1)
{
    w.lenL0 = lenL0; w.lenR0 = lenR0;
    w.aL0 = eb_div(0.5 * lenL0, (lenL1 + 2.0 * lenL0 + lenR0));
    w.aR0 = eb_div(0.5 * lenR0, (lenL0 + 2.0 * lenR0 + lenR1));
    w.two_over_L0L1 = eb_div(2.0, (lenL0 + lenL1));
    w.two_over_R0L0 = eb_div(2.0, (lenR0 + lenL0));
    w.two_over_R1R0 = eb_div(2.0, (lenR1 + lenR0));
    w.two_L0_plus_L1 = (2.0 * lenL0 + lenL1);
    w.two_R0_plus_R1 = (2.0 * lenR0 + lenR1);
}

// ---------------------------------------------------------------------------------------
// Flux calculators (local frame: x = face normal).  reference src/eilmer/fluxcalc.d

#define EB_UNPACK_LR                                                                  \
    const double rL = L.rho, pL = L.p, pLrL = eb_div(pL, rL);                                \
    const double uL = L.vx, vL = L.vy, wL = (DIM == 3 ? L.vz : 0.0);                  \
    const double eL = L.u, aL = L.a;                                                  \
    const double keL = 0.5 * (uL * uL + vL * vL + wL * wL);                           \
    const double HL = eL + pLrL + keL;                                                \
    const double rR = R.rho, pR = R.p, pRrR = eb_div(pR, rR);                                \
    const double uR = R.vx, vR = R.vy, wR = (DIM == 3 ? R.vz : 0.0);                  \
    const double eR = R.u, aR = R.a;                                                  \
    const double keR = 0.5 * (uR * uR + vR * vR + wR * wR);                           \
    const double HR = eR + pRrR + keR;

// fluxcalc.d:474-647
template <int DIM, int NSP>
__device__ __forceinline__ void flux_ausmdv(const Prim<NSP>& L, const Prim<NSP>& R, bool entropy_fix, double* F, bool smooth_am = false)
{
    typedef Layout<DIM, NSP> Lay;
    EB_UNPACK_LR
    double am = fmax(aL, aR);
    if (smooth_am) {            // lmr/fluxcalc.d:553-561: smooth maximum instead of max(aL, aR)
        const double da = aL - aR;
        const double scale = 0.5 * (aL + aR);
        const double eps = 1e-6 * scale + 1e-12;
        am = 0.5 * (aL + aR) + 0.5 * sqrt(da * da + eps * eps);
    }
    double duL = 0.5 * (uL + fabs(uL));
    double duR = 0.5 * (uR - fabs(uR));
    double pLplus, uLplus, pRminus, uRminus;
#ifdef EB_FAST_MATH
    // shared reciprocals: 1/(pLrL+pRrR) for both alphas, 1/am for both Mach numbers and the 1/(4 am) terms
    const double rs = eb_rcp(pLrL + pRrR), ram = eb_rcp(am), qam = 0.25 * ram;
    double alphaL = 2.0 * pLrL * rs;
    double alphaR = 2.0 * pRrR * rs;
    double ML = uL * ram;
    double MR = uR * ram;
    if (fabs(ML) <= 1.0) {
        pLplus = pL * (ML + 1.0) * (ML + 1.0) * (2.0 - ML) * 0.25;
        uLplus = alphaL * ((uL + am) * (uL + am) * qam - duL) + duL;
    } else {
        pLplus = (uL > 0.0) ? pL : 0.0;       // pL*duL/uL
        uLplus = duL;
    }
    if (fabs(MR) <= 1.0) {
        pRminus = pR * (MR - 1.0) * (MR - 1.0) * (2.0 + MR) * 0.25;
        uRminus = alphaR * (-(uR - am) * (uR - am) * qam - duR) + duR;
    } else {
        pRminus = (uR < 0.0) ? pR : 0.0;      // pR*duR/uR
        uRminus = duR;
    }
#else
    double alphaL = 2.0 * pLrL / (pLrL + pRrR);
    double alphaR = 2.0 * pRrR / (pLrL + pRrR);
    double ML = uL / am;
    double MR = uR / am;
    if (fabs(ML) <= 1.0) {
        pLplus = pL * (ML + 1.0) * (ML + 1.0) * (2.0 - ML) * 0.25;
        uLplus = alphaL * ((uL + am) * (uL + am) / (4.0 * am) - duL) + duL;
    } else {
        pLplus = pL * duL / uL;
        uLplus = duL;
    }
    if (fabs(MR) <= 1.0) {
        pRminus = pR * (MR - 1.0) * (MR - 1.0) * (2.0 + MR) * 0.25;
        uRminus = alphaR * (-(uR - am) * (uR - am) / (4.0 * am) - duR) + duR;
    } else {
        pRminus = pR * duR / uR;
        uRminus = duR;
    }
#endif
    double ru_half = uLplus * rL + uRminus * rR;
    double p_half = pLplus + pRminus;
    double dp = pL - pR;
    dp = eb_div(10.0 * fabs(dp), fmin(pL, pR));
    double s = 0.5 * fmin(1.0, dp);
    double ru2_AUSMV = uLplus * rL * uL + uRminus * rR * uR;
    double ru2_AUSMD = 0.5 * (ru_half * (uL + uR) - fabs(ru_half) * (uR - uL));
    double ru2_half = (0.5 + s) * ru2_AUSMV + (0.5 - s) * ru2_AUSMD;
    F[Lay::iMass] = ru_half;
    F[Lay::iXMom] = (ru2_half + p_half);
    if (ru_half >= 0.0) {
        F[Lay::iYMom] = (ru_half * vL);
        if (DIM == 3) F[Lay::iZMom] = (ru_half * wL);
        F[Lay::iEnergy] = ru_half * HL;
        if (NSP > 1) {
#pragma unroll
            for (int i = 0; i < NSP; ++i) F[Lay::iSpecies + i] = ru_half * L.massf[i];
        }
    } else {
        F[Lay::iYMom] = (ru_half * vR);
        if (DIM == 3) F[Lay::iZMom] = (ru_half * wR);
        F[Lay::iEnergy] = ru_half * HR;
        if (NSP > 1) {
#pragma unroll
            for (int i = 0; i < NSP; ++i) F[Lay::iSpecies + i] = ru_half * R.massf[i];
        }
    }
    if (entropy_fix) {
        const double C_EFIX = 0.125;
        bool caseA = ((uL - aL) < 0.0) && ((uR - aR) > 0.0);
        bool caseB = ((uL + aL) < 0.0) && ((uR + aR) > 0.0);
        double d_ua = 0.0;
        if (caseA && !caseB) d_ua = C_EFIX * ((uR - aR) - (uL - aL));
        if (caseB && !caseA) d_ua = C_EFIX * ((uR + aR) - (uL + aL));
        if (d_ua != 0.0) {
            F[Lay::iMass] -= d_ua * (rR - rL);
            F[Lay::iXMom] -= d_ua * (rR * uR - rL * uL);
            F[Lay::iYMom] -= d_ua * (rR * vR - rL * vL);
            if (DIM == 3) F[Lay::iZMom] -= d_ua * (rR * wR - rL * wL);
            F[Lay::iEnergy] -= d_ua * (rR * HR - rL * HL);
            if (NSP > 1) {
#pragma unroll
                for (int i = 0; i < NSP; ++i) F[Lay::iSpecies + i] -= d_ua * (rR * R.massf[i] - rL * L.massf[i]);
            }
        }
    }
}

__device__ __forceinline__ double sgn_d(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : x); }

// fluxcalc.d:819-914 (VARIANT 0) and :917-1025 (VARIANT 2)
template <int DIM, int NSP, int VARIANT>
__device__ __forceinline__ void flux_ldfss(const Prim<NSP>& L, const Prim<NSP>& R, double* F)
{
    typedef Layout<DIM, NSP> Lay;
    EB_UNPACK_LR
    double am = 0.5 * (aL + aR);
    double ML, MR;
    if (VARIANT == 0) { ML = eb_div(uL, aL); MR = eb_div(uR, aR); } else { ML = eb_div(uL, am); MR = eb_div(uR, am); }
    double MpL = 0.25 * ((ML + 1.0) * (ML + 1.0));
    double MmR = -0.25 * ((MR - 1.0) * (MR - 1.0));
    double alphaL = 0.5 * (1.0 + sgn_d(ML));
    double alphaR = 0.5 * (1.0 - sgn_d(MR));
    double betaL = -fmax(0.0, 1.0 - floor(fabs(ML)));
    double betaR = -fmax(0.0, 1.0 - floor(fabs(MR)));
    double PL = 0.25 * ((ML + 1.0) * (ML + 1.0)) * (2.0 - ML);
    double PR = 0.25 * ((MR - 1.0) * (MR - 1.0)) * (2.0 + MR);
    double DL = alphaL * (1.0 + betaL) - betaL * PL;
    double DR = alphaR * (1.0 + betaR) - betaR * PR;
    double sq = eb_sqrt(0.5 * (ML * ML + MR * MR)) - 1.0;
    double Mhalf = 0.25 * betaL * betaR * (sq * sq);
    double cL, cR;                  // (a * rho * C) of each side
    if (VARIANT == 0) {
        double CL = alphaL * (1.0 + betaL) * ML - betaL * MpL - Mhalf;
        double CR = alphaR * (1.0 + betaR) * MR - betaR * MmR + Mhalf;
        cL = aL * rL * CL; cR = aR * rR * CR;
    } else {
        const double delta = 2.0;
        double MhalfL = Mhalf * (1.0 - (eb_div(pL - pR, pL + pR) + delta * eb_div(fabs(pL - pR), pL)));
        double MhalfR = Mhalf * (1.0 + (eb_div(pL - pR, pL + pR) - delta * eb_div(fabs(pL - pR), pR)));
        double CL = alphaL * (1.0 + betaL) * ML - betaL * MpL - MhalfL;
        double CR = alphaR * (1.0 + betaR) * MR - betaR * MmR + MhalfR;
        cL = am * rL * CL; cR = am * rR * CR;
    }
    double ru_half = cL + cR;
    double ru2_half = cL * uL + cR * uR;
    double p_half = DL * pL + DR * pR;
    F[Lay::iMass] = ru_half;
    F[Lay::iXMom] = (ru2_half + p_half);
    F[Lay::iYMom] = (cL * vL + cR * vR);
    if (DIM == 3) F[Lay::iZMom] = (cL * wL + cR * wR);
    F[Lay::iEnergy] = (cL * HL + cR * HR);
    if (NSP > 1) {
#pragma unroll
        for (int i = 0; i < NSP; ++i) F[Lay::iSpecies + i] = (ru_half * (ru_half >= 0.0 ? L.massf[i] : R.massf[i]));
    }
}

// fluxcalc.d:1028-1128
template <int DIM, int NSP>
__device__ __forceinline__ void flux_hanel(const Prim<NSP>& L, const Prim<NSP>& R, double* F)
{
    typedef Layout<DIM, NSP> Lay;
    EB_UNPACK_LR
    double pLplus, uLplus;
    if (fabs(uL) <= aL) {
        uLplus = eb_div(1.0, 4.0 * aL) * (uL + aL) * (uL + aL);
        pLplus = pL * uLplus * (eb_div(1.0, aL) * (2.0 - eb_div(uL, aL)));
    } else {
        uLplus = 0.5 * (uL + fabs(uL));
        pLplus = pL * uLplus * eb_div(1.0, uL);
    }
    double pRminus, uRminus;
    if (fabs(uR) <= aR) {
        uRminus = eb_div(-1.0, 4.0 * aR) * (uR - aR) * (uR - aR);
        pRminus = pR * uRminus * (eb_div(1.0, aR) * (-2.0 - eb_div(uR, aR)));
    } else {
        uRminus = 0.5 * (uR - fabs(uR));
        pRminus = pR * uRminus * eb_div(1.0, uR);
    }
    double p_half = pLplus + pRminus;
    F[Lay::iMass] = (uLplus * rL + uRminus * rR);
    F[Lay::iXMom] = (uLplus * rL * uL + uRminus * rR * uR + p_half);
    F[Lay::iYMom] = (uLplus * rL * vL + uRminus * rR * vR);
    if (DIM == 3) F[Lay::iZMom] = (uLplus * rL * wL + uRminus * rR * wR);
    F[Lay::iEnergy] = (uLplus * rL * HL + uRminus * rR * HR);
    if (NSP > 1) {
#pragma unroll
        for (int i = 0; i < NSP; ++i) F[Lay::iSpecies + i] = (uLplus * rL * L.massf[i] + uRminus * rR * R.massf[i]);
    }
}

// fluxcalc.d:1437-1477 split functions of AUSM+up
__device__ __forceinline__ double M1plus(double M) { return 0.5 * (M + fabs(M)); }
__device__ __forceinline__ double M1minus(double M) { return 0.5 * (M - fabs(M)); }
__device__ __forceinline__ double M2plus(double M) { return 0.25 * (M + 1.0) * (M + 1.0); }
__device__ __forceinline__ double M2minus(double M) { return -0.25 * (M - 1.0) * (M - 1.0); }

// fluxcalc.d:1415-1602
template <int DIM, int NSP>
__device__ __forceinline__ void flux_ausm_plus_up(const Prim<NSP>& L, const Prim<NSP>& R, double M_inf, double* F)
{
    typedef Layout<DIM, NSP> Lay;
    const double rL = L.rho, pL = L.p, uL = L.vx, vL = L.vy, wL = (DIM == 3 ? L.vz : 0.0);
    const double eL = L.u, aL = L.a;
    const double keL = 0.5 * (uL * uL + vL * vL + wL * wL);
    const double HL = eL + eb_div(pL, rL) + keL;
    const double rR = R.rho, pR = R.p, uR = R.vx, vR = R.vy, wR = (DIM == 3 ? R.vz : 0.0);
    const double eR = R.u, aR = R.a;
    const double keR = 0.5 * (uR * uR + vR * vR + wR * wR);
    const double HR = eR + eb_div(pR, rR) + keR;
    double a_half = 0.5 * (aR + aL);
    double ML = eb_div(uL, a_half);
    double MR = eb_div(uR, a_half);
    double MbarSq = eb_div(uL * uL + uR * uR, 2.0 * a_half * a_half);
    double M0Sq = fmin(1.0, fmax(MbarSq, M_inf * M_inf));
    double sqM0 = eb_sqrt(M0Sq);
    double fa = sqM0 * (2.0 - sqM0);
    double alpha = 0.1875 * (-4.0 + 5 * fa * fa);
    const double beta = 0.125;
    double M4plus_ML, P5plus_ML, M4minus_MR, P5minus_MR;
    if (fabs(ML) >= 1.0) {
        M4plus_ML = M1plus(ML);
        P5plus_ML = eb_div(1.0, ML) * M1plus(ML);
    } else {
        double M2p = M2plus(ML), M2m = M2minus(ML);
        M4plus_ML = M2p * (1.0 - 16.0 * beta * M2m);
        P5plus_ML = M2p * ((2.0 - ML) - 16.0 * alpha * ML * M2m);
    }
    if (fabs(MR) >= 1.0) {
        M4minus_MR = M1minus(MR);
        P5minus_MR = eb_div(1.0, MR) * M1minus(MR);
    } else {
        double M2p = M2plus(MR), M2m = M2minus(MR);
        M4minus_MR = M2m * (1.0 + 16.0 * beta * M2p);
        P5minus_MR = M2m * ((-2.0 - MR) + 16.0 * alpha * MR * M2p);
    }
    const double KP = 0.25, KU = 0.75, SIGMA = 1.0;
    double r_half = 0.5 * (rL + rR);
    double Mp = eb_div(eb_div(-KP, fa) * fmax((1.0 - SIGMA * MbarSq), 0.0) * (pR - pL), r_half * a_half * a_half);
    double Pu = -KU * P5plus_ML * P5minus_MR * (rL + rR) * fa * a_half * (uR - uL);
    double M_half = M4plus_ML + M4minus_MR + Mp;
    double ru_half = a_half * M_half;
    if (M_half > 0.0) ru_half *= rL; else ru_half *= rR;
    double p_half = P5plus_ML * pL + P5minus_MR * pR + Pu;
    double ru2_half = (ru_half >= 0.0) ? ru_half * uL : ru_half * uR;
    F[Lay::iMass] = ru_half;
    F[Lay::iXMom] = (ru2_half + p_half);
    if (ru_half >= 0.0) {
        F[Lay::iYMom] = (ru_half * vL);
        if (DIM == 3) F[Lay::iZMom] = (ru_half * wL);
        F[Lay::iEnergy] = ru_half * HL;
        if (NSP > 1) {
#pragma unroll
            for (int i = 0; i < NSP; ++i) F[Lay::iSpecies + i] = ru_half * L.massf[i];
        }
    } else {
        F[Lay::iYMom] = (ru_half * vR);
        if (DIM == 3) F[Lay::iZMom] = (ru_half * wR);
        F[Lay::iEnergy] = ru_half * HR;
        if (NSP > 1) {
#pragma unroll
            for (int i = 0; i < NSP; ++i) F[Lay::iSpecies + i] = ru_half * R.massf[i];
        }
    }
}

// gamma = Cp / Cv of a state (gas_model.d:205)
template <int GASM, int NSP>
__device__ double gas_gamma(const EbGas* __restrict__ g, const Prim<NSP>& Q)
{
    if (GASM == EB200_GAS_IDEAL) return g->gamma_CpCv;
    double Cp = 0.0, Cv = 0.0;
    double cps[NSP];
    for (int i = 0; i < NSP; ++i) { cea_Cp(g->curves[i], Q.T, cps[i]); Cp += Q.massf[i] * cps[i]; }
    for (int i = 0; i < NSP; ++i) Cv += Q.massf[i] * (cps[i] - g->Rsp[i]);
    return Cp / Cv;
}

// fluxcalc.d:1929-2120 (gamma(Q) = Cp/Cv as gas_model.d:205; the species terms :2055-2107 when NSP > 1)
template <int DIM, int NSP>
__device__ __forceinline__ void flux_roe(const EbGas* __restrict__ gas, const Prim<NSP>& L, const Prim<NSP>& R, double gL, double gR, double* F)
{
    typedef Layout<DIM, NSP> Lay;
    EB_UNPACK_LR
    (void)aL; (void)aR;
    const double sL = eb_sqrt(rL), sR = eb_sqrt(rR), sden = sL + sR;   // recomputed in the reference; same values
    double ghat = eb_div((sL * gL + sR * gR), sden);
    double rhat = eb_sqrt(rL * rR);
    double uhat = eb_div((sL * uL + sR * uR), sden);
    double vhat = eb_div((sL * vL + sR * vR), sden);
    double what = eb_div((sL * wL + sR * wR), sden);
    double Hhat = eb_div((sL * HL + sR * HR), sden);
    double tkehat = eb_div((sL * 0.0 + sR * 0.0), sden);
    double kehat = 0.5 * (uhat * uhat + vhat * vhat + what * what);
    double ahat2 = (ghat - 1.0) * (Hhat - kehat - tkehat);
    double ahat = eb_sqrt(ahat2);
    double dr = rR - rL, dp = pR - pL, du = uR - uL, dv = vR - vL, dw = wR - wL;
    double lam0 = uhat, lam1 = uhat + ahat, lam2 = uhat - ahat;
    const double phi = 0.5;
    double V = eb_sqrt(uhat * uhat + vhat * vhat + what * what);
    double lref = phi * (V + ahat);
    lam0 = (fabs(lam0) >= 2 * lref) ? fabs(lam0) : eb_div(lam0 * lam0, 4 * lref) + lref;
    lam1 = (fabs(lam1) >= 2 * lref) ? fabs(lam1) : eb_div(lam1 * lam1, 4 * lref) + lref;
    lam2 = (fabs(lam2) >= 2 * lref) ? fabs(lam2) : eb_div(lam2 * lam2, 4 * lref) + lref;
    const double a0 = fabs(lam0), a1 = fabs(lam1), a2 = fabs(lam2);
    const double w0 = (dr - eb_div(dp, ahat2));
    const double w1 = eb_div(dp + rhat * ahat * du, 2.0 * ahat2);
    const double w2 = eb_div(dp - rhat * ahat * du, 2.0 * ahat2);
    double FL, FR;
    FL = rL * uL; FR = rR * uR;
    F[Lay::iMass] = 0.5 * (FL + FR - (a0 * w0) - (a1 * w1) - (a2 * w2));
    FL = pL + rL * uL * uL; FR = pR + rR * uR * uR;
    F[Lay::iXMom] = 0.5 * (FL + FR - (a0 * w0 * uhat) - (a1 * w1 * (uhat + ahat)) - (a2 * w2 * (uhat - ahat)));
    FL = rL * uL * vL; FR = rR * uR * vR;
    F[Lay::iYMom] = 0.5 * (FL + FR - (a0 * (w0 * vhat + rhat * dv)) - (a1 * w1 * vhat) - (a2 * w2 * vhat));
    if (DIM == 3) {
        FL = rL * uL * wL; FR = rR * uR * wR;
        F[Lay::iZMom] = 0.5 * (FL + FR - (a0 * (w0 * what + rhat * dw)) - (a1 * w1 * what) - (a2 * w2 * what));
    }
    const double dtke = 0.0;
    double theta = 0.0;
    if (NSP > 1) {                      // Walters et al. (1992) eq. 33b; internal_energy(Q, isp) = h_i(T) - R_i T (therm_perf_gas.d:443-448)
        const double That = eb_div((sL * L.T + sR * R.T), sden);
        const double logTL = log(L.T), logTR = log(R.T);
#pragma unroll
        for (int i = 0; i < NSP; ++i) {
            const double dmassf = R.massf[i] - L.massf[i];
            double hL, hR;
            cea_h(gas->curves[i], L.T, logTL, hL); cea_h(gas->curves[i], R.T, logTR, hR);
            const double eiL = hL - gas->Rsp[i] * L.T;
            const double eiR = hR - gas->Rsp[i] * R.T;
            const double eihat = eb_div((sL * eiL + sR * eiR), sden);
            const double psihat = eb_div(gas->Rsp[i] * That, (ghat - 1.0)) - eihat + kehat;
            theta += dmassf * psihat;
        }
    }
    FL = rL * uL * HL; FR = rR * uR * HR;
    F[Lay::iEnergy] = 0.5 * (FL + FR
                             - (a0 * (w0 * (kehat + tkehat) + rhat * (vhat * dv + what * dw + dtke - theta)))
                             - (a1 * w1 * (Hhat + uhat * ahat))
                             - (a2 * w2 * (Hhat - uhat * ahat)));
    if (NSP > 1) {                      // :2093-2107
#pragma unroll
        for (int i = 0; i < NSP; ++i) {
            const double massfhat = eb_div((sL * L.massf[i] + sR * R.massf[i]), sden);
            const double dmassf = R.massf[i] - L.massf[i];
            FL = rL * uL * L.massf[i]; FR = rR * uR * R.massf[i];
            F[Lay::iSpecies + i] = 0.5 * (FL + FR - (a0 * (w0 * massfhat + rhat * dmassf)) - (a1 * w1 * massfhat) - (a2 * w2 * massfhat));
        }
    }
}


// fluxcalc.d:650-816 hllc (Toro's HLLC with Einfeldt's wave speeds); single temperature, no turbulence, factor = 1
template <int DIM, int NSP>
__device__ __forceinline__ void flux_hllc(const Prim<NSP>& L, const Prim<NSP>& R, double gL, double gR, double* F)
{
    typedef Layout<DIM, NSP> Lay;
    const double rL = L.rho, pL = L.p, uL = L.vx, vL = L.vy, wL = (DIM == 3 ? L.vz : 0.0), eL = L.u, aL = L.a;
    const double keL = 0.5 * (uL * uL + vL * vL + wL * wL);
    const double EL = rL * eL + rL * keL;
    const double rR = R.rho, pR = R.p, uR = R.vx, vR = R.vy, wR = (DIM == 3 ? R.vz : 0.0), eR = R.u, aR = R.a;
    const double keR = 0.5 * (uR * uR + vR * vR + wR * wR);
    const double ER = rR * eR + rR * keR;
    const double sL = eb_sqrt(rL), sR = eb_sqrt(rR), sden = sL + sR;          // recomputed in the reference; same values
    const double uhat = eb_div((sL * uL + sR * uR), sden);
    const double ghat = eb_div((sL * gL + sR * gR), sden);
    const double ahat2 = (eb_div((sL * aL * aL + sR * aR * aR), sden)) +
                         0.5 * (ghat - 1.0) * (eb_div(sden, eb_sqrt(sden))) * (uR - uL) * (uR - uL);
    const double ahat = eb_sqrt(ahat2);
    const double SL = fmin(uL - aL, uhat - ahat);
    const double SR = fmax(uR + aR, uhat + ahat);
    const double S_star = eb_div((pR - pL + rL * uL * (SL - uL) - rR * uR * (SR - uR)), (rL * (SL - uL) - rR * (SR - uR)));
    bool star_region;
    double coeff, r, p, u, v, w, E, S;
    if (S_star > 0.0) {
        r = rL; p = pL; u = uL; v = vL; w = wL; E = EL; S = SL;
        if (SL > 0.0) { star_region = false; coeff = 0.0; }
        else { star_region = true; coeff = eb_div(rL * (SL - uL), (SL - S_star)); }
    } else {
        r = rR; p = pR; u = uR; v = vR; w = wR; E = ER; S = SR;
        if (SR < 0.0) { star_region = false; coeff = 0.0; }
        else { star_region = true; coeff = eb_div(rR * (SR - uR), (SR - S_star)); }
    }
    const double F_mass = r * u;
    const double ru_half = star_region ? F_mass + S * (coeff - r) : F_mass;
    F[Lay::iMass] = ru_half;
    const double F_momx = r * u * u + p, F_momy = r * u * v, F_momz = r * u * w;
    if (star_region) {
        F[Lay::iXMom] = (F_momx + S * (coeff * S_star - r * u));
        F[Lay::iYMom] = (F_momy + S * (coeff * v - r * v));
        if (DIM == 3) F[Lay::iZMom] = (F_momz + S * (coeff * w - r * w));
    } else {
        F[Lay::iXMom] = F_momx;
        F[Lay::iYMom] = F_momy;
        if (DIM == 3) F[Lay::iZMom] = F_momz;
    }
    const double F_totenergy = u * (E + p);
    if (star_region) {
        const double U_star_totenergy = coeff * (eb_div(E, r) + (S_star - u) * (S_star + eb_div(p, (r * (S - u)))));
        F[Lay::iEnergy] = (F_totenergy + S * (U_star_totenergy - E));
    } else F[Lay::iEnergy] = (F_totenergy);
    if (NSP > 1) {
#pragma unroll
        for (int i = 0; i < NSP; ++i) F[Lay::iSpecies + i] = (ru_half * ((ru_half >= 0.0) ? L.massf[i] : R.massf[i]));
    }
}

// fluxcalc.d:1779-1926 hlle2 (HLL with Einfeldt's wave speeds).  In the subsonic branch the reference adds the
// z-momentum flux to the y-momentum entry (:1901) and leaves the z entry at zero: reproduced.
template <int DIM, int NSP>
__device__ __forceinline__ void flux_hlle2(const Prim<NSP>& L, const Prim<NSP>& R, double gL, double gR, double* F)
{
    typedef Layout<DIM, NSP> Lay;
    EB_UNPACK_LR
    const double sL = eb_sqrt(rL), sR = eb_sqrt(rR), sden = sL + sR;
    const double uhat = eb_div((sL * uL + sR * uR), sden);
    const double ghat = eb_div((sL * gL + sR * gR), sden);
    const double ahat2 = (eb_div((sL * aL * aL + sR * aR * aR), sden)) +
                         0.5 * (ghat - 1.0) * (eb_div(sden, eb_sqrt(sden))) * (uR - uL) * (uR - uL);
    const double ahat = eb_sqrt(ahat2);
    const double SLm = fmin(uL - aL, uhat - ahat);
    const double SRp = fmax(uR + aR, uhat + ahat);
    if (SLm >= 0) {
        F[Lay::iMass] = (rL * uL);
        F[Lay::iXMom] = (rL * uL * uL + pL);
        F[Lay::iYMom] = (rL * uL * vL);
        if (DIM == 3) F[Lay::iZMom] = (rL * uL * wL);
        F[Lay::iEnergy] = (rL * uL * HL);
        if (NSP > 1) {
#pragma unroll
            for (int i = 0; i < NSP; ++i) F[Lay::iSpecies + i] = (rL * uL * L.massf[i]);
        }
    } else if (SRp <= 0) {
        F[Lay::iMass] = (rR * uR);
        F[Lay::iXMom] = (rR * uR * uR + pR);
        F[Lay::iYMom] = (rR * uR * vR);
        if (DIM == 3) F[Lay::iZMom] = (rR * uR * wR);
        F[Lay::iEnergy] = (rR * uR * HR);
        if (NSP > 1) {
#pragma unroll
            for (int i = 0; i < NSP; ++i) F[Lay::iSpecies + i] = (rR * uR * R.massf[i]);
        }
    } else {
#ifdef EB_FAST_MATH
        const double rdS = eb_rcp(SRp - SLm);
#define EB_HLLE_DIV(x) ((x) * rdS)
#else
#define EB_HLLE_DIV(x) ((x) / (SRp - SLm))
#endif
        const double ru_half = EB_HLLE_DIV(SRp * rL * uL - SLm * rR * uR + SLm * SRp * (rR - rL));
        F[Lay::iMass] = ru_half;
        F[Lay::iXMom] = EB_HLLE_DIV(SRp * (rL * uL * uL + pL) - SLm * (rR * uR * uR + pR) + SLm * SRp * (rR * uR - rL * uL));
        double fy = EB_HLLE_DIV(SRp * (rL * uL * vL) - SLm * (rR * uR * vR) + SLm * SRp * (rR * vR - rL * vL));
        if (DIM == 3) {
            fy += EB_HLLE_DIV(SRp * (rL * uL * wL) - SLm * (rR * uR * wR) + SLm * SRp * (rR * wR - rL * wL));
            F[Lay::iZMom] = 0.0;
        }
        F[Lay::iYMom] = fy;
        F[Lay::iEnergy] = EB_HLLE_DIV(SRp * (rL * uL * HL) - SLm * (rR * uR * HR) + SLm * SRp * (rR * HR - rL * HL));
#undef EB_HLLE_DIV
        if (NSP > 1) {
#pragma unroll
            for (int i = 0; i < NSP; ++i) F[Lay::iSpecies + i] = (ru_half * ((ru_half >= 0.0) ? L.massf[i] : R.massf[i]));
        }
    }
}

// fluxcalc.d:1280-1312 exxef: exp(-x^2) and erf(x) by a polynomial approximation
__device__ __forceinline__ void exxef(double sn, double& exx, double& ef)
{
    const double Pc = 0.327591100, A1 = 0.254829592, A2 = -0.284496736, A3 = 1.421413741, A4 = -1.453152027, A5 = 1.061405429;
    double ef1;
    if (fabs(sn) > 5.0) { exx = 0.138879e-10; ef1 = 1.0; }
    else {
        const double snsq = sn * sn;
        exx = exp(-snsq);
        const double y = eb_div(1.0, 1.0 + Pc * fabs(sn));
        ef1 = 1.0 - y * (A1 + y * (A2 + y * (A3 + y * (A4 + A5 * y)))) * exx;
    }
    ef = copysign(ef1, sn);
}

// gmodel.Cv(Q): ideal_gas.d (a constant), therm_perf_gas.d:417-424
template <int GASM, int NSP>
__device__ __forceinline__ double gas_Cv(const EbGas* __restrict__ g, const Prim<NSP>& Q)
{
    if (GASM == EB200_GAS_IDEAL) return g->Cv;
    double cv = 0.0;
#pragma unroll
    for (int i = 0; i < NSP; ++i) { double c; cea_Cp(g->curves[i], Q.T, c); cv += Q.massf[i] * (c - g->Rsp[i]); }
    return cv;
}

// fluxcalc.d:1131-1277 efmflx, the equilibrium flux method of Macrossan & Pullin (factor = 1).
// Reads T of both states.  Note rtL = Rgas*tL but rtR = presR/rhoR, as in the reference.
template <int DIM, int NSP, int GASM>
__device__ __forceinline__ void flux_efm(const EbGas* __restrict__ g, const Prim<NSP>& L, const Prim<NSP>& R, double* F)
{
    typedef Layout<DIM, NSP> Lay;
    const double dtwspi = 0.282094792;
    const double rhoL = L.rho, presL = L.p, eL = L.u, tL = L.T, vnL = L.vx, vpL = L.vy, vqL = (DIM == 3) ? L.vz : 0.0;
    const double rhoR = R.rho, presR = R.p, eR = R.u, tR = R.T, vnR = R.vx, vpR = R.vy, vqR = (DIM == 3) ? R.vz : 0.0;
    double hL = eL + eb_div(presL, rhoL); hL += 0.0;
    double hR = eR + eb_div(presR, rhoR); hR += 0.0;
    const double cvL = gas_Cv<GASM, NSP>(g, L), RgasL = eb_div(presL, (rhoL * tL));
    const double cvR = gas_Cv<GASM, NSP>(g, R), RgasR = eb_div(presR, (rhoR * tR));
    const double rLsqrt = eb_sqrt(rhoL), rRsqrt = eb_sqrt(rhoR);
    const double alpha = eb_div(rLsqrt, (rLsqrt + rRsqrt));
    const double cv = alpha * cvL + (1.0 - alpha) * cvR;
    const double Rgas = alpha * RgasL + (1.0 - alpha) * RgasR;
    const double cp = cv + Rgas;
    const double gam = eb_div(cp, cv);
    const double con = eb_div(0.5 * (gam + 1.0), (gam - 1.0));
    double exL, efL, exR, efR;
    const double rtL = Rgas * tL;
    const double cmpL = eb_sqrt(2.0 * rtL);
    const double hvsqL = 0.5 * (vnL * vnL + vpL * vpL + vqL * vqL);
    const double snL = eb_div(vnL, (1.0 * cmpL));
    exxef(snL, exL, efL);
    const double wL = 0.5 * (1.0 + efL);
    const double dL = exL * dtwspi;
    const double rtR = eb_div(presR, rhoR);
    const double cmpR = eb_sqrt(2.0 * rtR);
    const double hvsqR = 0.5 * (vnR * vnR + vpR * vpR + vqR * vqR);
    const double snR = eb_div(vnR, (1.0 * cmpR));
    exxef(snR, exR, efR);
    const double wR = 0.5 * (1.0 - efR);
    const double dR = -exR * dtwspi;
    const double fmsL = (wL * rhoL * vnL) + (dL * cmpL * rhoL);
    const double fmsR = (wR * rhoR * vnR) + (dR * cmpR * rhoR);
    const double mass_flux = 1.0 * (fmsL + fmsR);
    F[Lay::iMass] = mass_flux;
    F[Lay::iXMom] = 1.0 * (fmsL * vnL + fmsR * vnR + wL * presL + wR * presR);
    F[Lay::iYMom] = 1.0 * (fmsL * vpL + fmsR * vpR);
    if (DIM == 3) F[Lay::iZMom] = 1.0 * (fmsL * vqL + fmsR * vqR);
    F[Lay::iEnergy] = 1.0 * ((wL * rhoL * vnL) * (hvsqL + hL) + (wR * rhoR * vnR) * (hvsqR + hR) +
                             (dL * cmpL * rhoL) * (hvsqL + con * rtL) + (dR * cmpR * rhoR) * (hvsqR + con * rtR));
    if (NSP > 1) {
#pragma unroll
        for (int i = 0; i < NSP; ++i) F[Lay::iSpecies + i] = mass_flux * ((mass_flux > 0.0) ? L.massf[i] : R.massf[i]);
    }
}

// One place that knows which basic calculators an adaptive one blends (fluxcalc.d:1315-1412): `shock` where the
// detector marks the face (IFace.fs.S = 1; alpha is 0 or 1 on this path), `smooth` elsewhere.
template <int FLUX>
struct FluxPair {
    static constexpr bool adaptive = (FLUX == EB200_FLUX_ADAPTIVE_HANEL_AUSMDV || FLUX == EB200_FLUX_ADAPTIVE_HANEL_AUSM_PLUS_UP ||
                                      FLUX == EB200_FLUX_ADAPTIVE_LDFSS0_LDFSS2 || FLUX == EB200_FLUX_ADAPTIVE_EFM_AUSMDV);
    static constexpr int shock = (FLUX == EB200_FLUX_ADAPTIVE_HANEL_AUSMDV || FLUX == EB200_FLUX_ADAPTIVE_HANEL_AUSM_PLUS_UP) ? EB200_FLUX_HANEL
                               : (FLUX == EB200_FLUX_ADAPTIVE_LDFSS0_LDFSS2) ? EB200_FLUX_LDFSS0
                               : (FLUX == EB200_FLUX_ADAPTIVE_EFM_AUSMDV) ? EB200_FLUX_EFM : FLUX;
    static constexpr int smooth = (FLUX == EB200_FLUX_ADAPTIVE_HANEL_AUSMDV || FLUX == EB200_FLUX_ADAPTIVE_EFM_AUSMDV) ? EB200_FLUX_AUSMDV
                                : (FLUX == EB200_FLUX_ADAPTIVE_HANEL_AUSM_PLUS_UP) ? EB200_FLUX_AUSM_PLUS_UP
                                : (FLUX == EB200_FLUX_ADAPTIVE_LDFSS0_LDFSS2) ? EB200_FLUX_LDFSS2 : FLUX;
};

template <int DIM, int NSP, int GASM, int BASE>
__device__ __forceinline__ void basic_flux(const EbParams& P, const EbGas* __restrict__ gas, const Prim<NSP>& L, const Prim<NSP>& R, double* F)
{
    if (BASE == EB200_FLUX_AUSMDV) flux_ausmdv<DIM, NSP>(L, R, P.entropy_fix != 0, F, P.lmr != 0);
    else if (BASE == EB200_FLUX_HANEL) flux_hanel<DIM, NSP>(L, R, F);
    else if (BASE == EB200_FLUX_LDFSS0) flux_ldfss<DIM, NSP, 0>(L, R, F);
    else if (BASE == EB200_FLUX_LDFSS2) flux_ldfss<DIM, NSP, 2>(L, R, F);
    else if (BASE == EB200_FLUX_AUSM_PLUS_UP) flux_ausm_plus_up<DIM, NSP>(L, R, P.M_inf, F);
    else if (BASE == EB200_FLUX_ROE) flux_roe<DIM, NSP>(gas, L, R, gas_gamma<GASM, NSP>(gas, L), gas_gamma<GASM, NSP>(gas, R), F);
    else if (BASE == EB200_FLUX_HLLC) flux_hllc<DIM, NSP>(L, R, gas->gamma_CpCv, gas->gamma_CpCv, F);
    else if (BASE == EB200_FLUX_HLLE2) flux_hlle2<DIM, NSP>(L, R, gas->gamma_CpCv, gas->gamma_CpCv, F);
    else flux_efm<DIM, NSP, GASM>(gas, L, R, F);
}

// The flux calculator FLUX for states in the face frame, in the reference's order of operations
template <int DIM, int NSP, int GASM, int FLUX>
__device__ __forceinline__ void flux_in_face_frame(const EbParams& P, const EbGas* __restrict__ gas, const Prim<NSP>& L,
                                                   const Prim<NSP>& R, double alpha, double* F)
{
    if (FluxPair<FLUX>::adaptive && alpha > 0.0) basic_flux<DIM, NSP, GASM, FluxPair<FLUX>::shock>(P, gas, L, R, F);
    else basic_flux<DIM, NSP, GASM, FluxPair<FLUX>::smooth>(P, gas, L, R, F);
}

// ---------------------------------------------------------------------------------------
// decode_conserved (fvcell.d:586-821) for one cell.  U[] in/out (species rescale and the
// low-temperature fix write it), Q.T on entry = previous temperature (Newton start).
// returns 0 ok, 1 = the reference would throw (step failed)

template <int DIM, int GASM, int NSP>
__device__ __forceinline__ int decode_cell(const EbParams& P, const EbGas* __restrict__ g, double* U, Prim<NSP>& Q,
                                           bool& U_modified)
{
    typedef Layout<DIM, NSP> Lay;
    U_modified = false;
    double rho = U[Lay::iMass];
    if (!(rho > 0.0)) return 1;
    Q.rho = rho;
    double dinv = eb_div(1.0, rho);
    Q.vx = U[Lay::iXMom] * dinv; Q.vy = U[Lay::iYMom] * dinv;
    Q.vz = (DIM == 3) ? U[Lay::iZMom] * dinv : 0.0;
    double u = U[Lay::iEnergy] * dinv;
    double ke = 0.5 * (Q.vx * Q.vx + Q.vy * Q.vy + Q.vz * Q.vz);
    u -= ke;
    Q.u = u;
    if (NSP > 1) {
        double rhos_sum = 0.0;
#pragma unroll
        for (int i = 0; i < NSP; ++i) {
            if (U[Lay::iSpecies + i] < 0.0) { U[Lay::iSpecies + i] = 0.0; U_modified = true; }
            rhos_sum += U[Lay::iSpecies + i];
        }
        if (fabs(rhos_sum - rho) > 0.1) return 1;
        if (fabs(rhos_sum - rho) > 0.0) {
            double scale = rho / rhos_sum;
#pragma unroll
            for (int i = 0; i < NSP; ++i) U[Lay::iSpecies + i] *= scale;
            U_modified = true;
        }
#pragma unroll
        for (int i = 0; i < NSP; ++i) { Q.massf[i] = U[Lay::iSpecies + i] * dinv; Q.rho_s[i] = U[Lay::iSpecies + i]; }
    } else {
        Q.massf[0] = 1.0; Q.rho_s[0] = rho;
    }
    if (!thermo_from_rhou<GASM, NSP>(g, Q)) {
        if (P.ignore_low_T && (rho > 0.0)) {
            Q.T = P.low_T;
            if (!thermo_from_rhoT<GASM, NSP>(g, Q)) return 1;
            // encode_conserved (fvcell.d:511-583)
            U[Lay::iMass] = Q.rho;
            U[Lay::iXMom] = Q.rho * Q.vx; U[Lay::iYMom] = Q.rho * Q.vy;
            if (DIM == 3) U[Lay::iZMom] = Q.rho * Q.vz;
            double ke2 = 0.5 * (Q.vx * Q.vx + Q.vy * Q.vy + Q.vz * Q.vz);
            U[Lay::iEnergy] = Q.rho * (Q.u + ke2);
            if (NSP > 1) {
#pragma unroll
                for (int i = 0; i < NSP; ++i) U[Lay::iSpecies + i] = Q.rho * Q.massf[i];
            }
            U_modified = true;
        } else return 1;
    }
    if (Q.T <= 0.0) return 1;
    if (!sound_speed<GASM, NSP>(g, Q)) return 1;
    return 0;
}

// flowstate.d:363-390 check_data + gas_state.d:226-268 check_values
template <int NSP>
__device__ __forceinline__ bool check_data(const EbParams& P, const Prim<NSP>& Q)
{
    bool ok = true;
    if (!isfinite(Q.rho) || Q.rho < 0.0) ok = false;
    if (!isfinite(Q.T) || Q.T < 0.0) ok = false;
    if (!isfinite(Q.p)) ok = false;
    if (!isfinite(Q.a)) ok = false;
    double fsum = 0.0;
#pragma unroll
    for (int i = 0; i < NSP; ++i) fsum += Q.massf[i];
    if (fsum < 0.99 || fsum > 1.01 || !isfinite(fsum)) ok = false;
    if (fabs(Q.vx) > P.max_velocity || fabs(Q.vy) > P.max_velocity || fabs(Q.vz) > P.max_velocity) ok = false;
    if (Q.T < P.min_temp) ok = false;
    if (Q.T > P.max_temp) ok = false;
    return ok;
}

template <int NSP>
__device__ __forceinline__ void load_prim(Prim<NSP>& Q, const double* __restrict__ prim, long long total, long long c)
{
    Q.rho = ldg(prim + c); Q.u = ldg(prim + total + c); Q.p = ldg(prim + 2 * total + c);
    Q.T = ldg(prim + 3 * total + c); Q.a = ldg(prim + 4 * total + c);
    Q.vx = ldg(prim + 5 * total + c); Q.vy = ldg(prim + 6 * total + c); Q.vz = ldg(prim + 7 * total + c);
    if (NSP > 1) {
#pragma unroll
        for (int i = 0; i < NSP; ++i) { Q.massf[i] = ldg(prim + (8 + i) * total + c); Q.rho_s[i] = ldg(prim + (8 + NSP + i) * total + c); }
    } else { Q.massf[0] = 1.0; Q.rho_s[0] = Q.rho; }
}

template <int NSP>
__device__ __forceinline__ void store_prim(const Prim<NSP>& Q, double* __restrict__ prim, long long total, long long c)
{
    prim[c] = Q.rho; prim[total + c] = Q.u; prim[2 * total + c] = Q.p; prim[3 * total + c] = Q.T;
    prim[4 * total + c] = Q.a; prim[5 * total + c] = Q.vx; prim[6 * total + c] = Q.vy; prim[7 * total + c] = Q.vz;
    if (NSP > 1) {
#pragma unroll
        for (int i = 0; i < NSP; ++i) { prim[(8 + i) * total + c] = Q.massf[i]; prim[(8 + NSP + i) * total + c] = Q.rho_s[i]; }
    }
}

}  // namespace EB_NS
