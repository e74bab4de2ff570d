// flux_kernel_v2.cuh -- the tuned fused per-stage kernel (ideal gas, second-order l2r2
// reconstruction with the van Albada limiter: the default configuration of the reference).
//
// Same mathematics and evaluation order as flux_kernel.cuh (which stays as the generic
// path for other gas models / options); what differs is how data reaches the FP64 pipe:
//   * the (i, j) tile of the current k-plane, with its two-cell halo, is staged in shared
//     memory with cp.async (LDGSTS), double buffered: the copy of plane k+1 overlaps the
//     face computations of plane k; stencil values are then LDS with immediate offsets
//     instead of global loads with 64-bit address arithmetic;
//   * the three index directions are unrolled at compile time, so that on the uniform-
//     Cartesian path the face frame is a free renaming of the velocity components;
//   * west/south fluxes are exchanged through shared memory (double buffered, one barrier
//     per plane); the faces on the far edge of the tile are a second trip through the same
//     code for warps 0 and 1;
//   * __launch_bounds__(256, 2): 128 registers, two CTAs (16 warps) per SM.
#pragma once
#include "device_math.cuh"

#ifndef EB_V2_TY
#define EB_V2_TY 8
#endif
#ifndef EB_V2_MIN_CTAS
#define EB_V2_MIN_CTAS 2
#endif

namespace EB_NS {

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }

// TMA (cp.async.bulk.tensor) + mbarrier, used to stage the tile of a k-plane: one thread issues three
// box loads per plane, all threads wait on the barrier's phase.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_4d(void* dst, const void* tmap, unsigned long long* bar, int x0, int x1, int x2, int x3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// van Albada-limited l2r2 reconstruction of one scalar with precomputed weights; LIM/CLIP are
// compile-time copies of config.apply_limiter / config.extrema_clipping (onedinterp.d:357-384)
template <bool CLIP>
__device__ __forceinline__ void interp_v2(const EbWeights& w, double eps, double qL1, double qL0, double qR0, double qR1,
                                          double& qL, double& qR)
{
    double delLminus = (qL0 - qL1) * w.two_over_L0L1;
    double del = (qR0 - qL0) * w.two_over_R0L0;
    double delRplus = (qR1 - qR0) * w.two_over_R1R0;
#ifdef EB_FAST_MATH
    // one reciprocal for both limiter values: sL = nL/dL, sR = nR/dR = (nL*dR, nR*dL) / (dL*dR)
    const double d2 = del * del;
    const double nL = delLminus * del + fabs(delLminus * del) + eps, dL = delLminus * delLminus + d2 + eps;
    const double nR = del * delRplus + fabs(del * delRplus) + eps, dR = d2 + delRplus * delRplus + eps;
    const double rr = eb_rcp(dL * dR);
    double sL = nL * dR * rr;
    double sR = nR * dL * rr;
#else
    double sL = (delLminus * del + fabs(delLminus * del) + eps) / (delLminus * delLminus + del * del + eps);
    double sR = (del * delRplus + fabs(del * delRplus) + eps) / (del * del + delRplus * delRplus + eps);
#endif
    qL = qL0 + sL * w.aL0 * (del * w.two_L0_plus_L1 + delLminus * w.lenR0);
    qR = qR0 - sR * w.aR0 * (delRplus * w.lenL0 + del * w.two_R0_plus_R1);
    if (CLIP) {
        qL = clip_to_limits(qL, qL0, qR0);
        qR = clip_to_limits(qR, qL0, qR0);
    }
}

// Stencil values of one face: [m] = L1, L0, R0, R1; velocities already in the face frame
// (vn, vt1, vt2) on the Cartesian path, global (x, y, z) on the general path.
struct Stencil {
    double rho[4], u[4], v0[4], v1[4], v2[4];
    double aL, aR;
};

__device__ __forceinline__ double pick3(int d, double a0, double a1, double a2) { return d == 0 ? a0 : (d == 1 ? a1 : a2); }

#ifdef EB_FAST_MATH
// Component-form flux calculators for the uniform-Cartesian path of the throughput build:
// velocities stay in (x, y, z); `d` names the face-normal component.  The tangential
// momentum fluxes of these schemes have one common form, evaluated for all three components;
// the normal component is then overwritten.  Sums are not in the reference's (n, t1, t2) order
// (kinetic energy), so this is for the 1e-10 tolerance build only.
// min/max without fmin/fmax's NaN bookkeeping (DSETP + 2 FSEL instead of ~8 instructions); a NaN in a
// reconstructed state ends the step in decode_conserved either way
__device__ __forceinline__ double qmin(double a, double b) { return (a < b) ? a : b; }
__device__ __forceinline__ double qmax(double a, double b) { return (a > b) ? a : b; }

template <int DIM>
__device__ __forceinline__ void set_normal(int d, double fn, double* F)
{
    typedef Layout<DIM, 1> Lay;
    F[Lay::iXMom] = (d == 0) ? fn : F[Lay::iXMom];
    F[Lay::iYMom] = (d == 1) ? fn : F[Lay::iYMom];
    if (DIM == 3) F[Lay::iZMom] = (d == 2) ? fn : F[Lay::iZMom];
}

// pLrL, pRrR: p/rho of the two states, supplied by the caller (pL * rcp(rL), or (gamma-1) * u for the ideal gas)
template <int DIM, int FLUX>
__device__ __forceinline__ void flux_components_pr(const Prim<1>& L, const Prim<1>& R, int d, bool entropy_fix, double M_inf,
                                                   double pLrL, double pRrR, double* F)
{
    typedef Layout<DIM, 1> Lay;
    const double rL = L.rho, pL = L.p, rR = R.rho, pR = R.p, aL = L.a, aR = R.a;
    const double uL = (DIM == 3) ? pick3(d, L.vx, L.vy, L.vz) : ((d == 0) ? L.vx : L.vy);
    const double uR = (DIM == 3) ? pick3(d, R.vx, R.vy, R.vz) : ((d == 0) ? R.vx : R.vy);
    const double keL = 0.5 * (L.vx * L.vx + L.vy * L.vy + ((DIM == 3) ? L.vz * L.vz : 0.0));
    const double keR = 0.5 * (R.vx * R.vx + R.vy * R.vy + ((DIM == 3) ? R.vz * R.vz : 0.0));
    const double HL = L.u + pLrL + keL, HR = R.u + pRrR + keR;
    if (FLUX == EB200_FLUX_AUSMDV) {                 // fluxcalc.d:474-647
        const double am = qmax(aL, aR);
        const double duL = 0.5 * (uL + fabs(uL)), duR = 0.5 * (uR - fabs(uR));
        const double rs = eb_rcp(pLrL + pRrR), ram = eb_rcp(am), qam = 0.25 * ram;
        const double alphaL = 2.0 * pLrL * rs, alphaR = 2.0 * pRrR * rs;
        const double ML = uL * ram, MR = uR * ram;
        double pLplus, uLplus, pRminus, uRminus;
        if (fabs(ML) <= 1.0) {
            pLplus = pL * (ML + 1.0) * (ML + 1.0) * (2.0 - ML) * 0.25;
            uLplus = alphaL * ((uL + am) * (uL + am) * qam - duL) + duL;
        } else { pLplus = (uL > 0.0) ? pL : 0.0; uLplus = duL; }
        if (fabs(MR) <= 1.0) {
            pRminus = pR * (MR - 1.0) * (MR - 1.0) * (2.0 + MR) * 0.25;
            uRminus = alphaR * (-(uR - am) * (uR - am) * qam - duR) + duR;
        } else { pRminus = (uR < 0.0) ? pR : 0.0; uRminus = duR; }
        const double ru_half = uLplus * rL + uRminus * rR;
        const double p_half = pLplus + pRminus;
        const double dp = 10.0 * fabs(pL - pR) * eb_rcp(qmin(pL, pR));
        const double sw = 0.5 * qmin(1.0, dp);
        const double ru2_AUSMV = uLplus * rL * uL + uRminus * rR * uR;
        const double ru2_AUSMD = 0.5 * (ru_half * (uL + uR) - fabs(ru_half) * (uR - uL));
        const double ru2_half = (0.5 + sw) * ru2_AUSMV + (0.5 - sw) * ru2_AUSMD;
        const bool fromL = (ru_half >= 0.0);
        F[Lay::iMass] = ru_half;
        F[Lay::iXMom] = ru_half * (fromL ? L.vx : R.vx);
        F[Lay::iYMom] = ru_half * (fromL ? L.vy : R.vy);
        if (DIM == 3) F[Lay::iZMom] = ru_half * (fromL ? L.vz : R.vz);
        set_normal<DIM>(d, ru2_half + p_half, F);
        F[Lay::iEnergy] = ru_half * (fromL ? HL : HR);
        if (entropy_fix) {
            const bool caseA = ((uL - aL) < 0.0) && ((uR - aR) > 0.0);
            const bool caseB = ((uL + aL) < 0.0) && ((uR + aR) > 0.0);
            double d_ua = 0.0;
            if (caseA && !caseB) d_ua = 0.125 * ((uR - aR) - (uL - aL));
            if (caseB && !caseA) d_ua = 0.125 * ((uR + aR) - (uL + aL));
            if (d_ua != 0.0) {
                F[Lay::iMass] -= d_ua * (rR - rL);
                F[Lay::iXMom] -= d_ua * (rR * R.vx - rL * L.vx);
                F[Lay::iYMom] -= d_ua * (rR * R.vy - rL * L.vy);
                if (DIM == 3) F[Lay::iZMom] -= d_ua * (rR * R.vz - rL * L.vz);
                F[Lay::iEnergy] -= d_ua * (rR * HR - rL * HL);
            }
        }
    } else if (FLUX == EB200_FLUX_HANEL) {           // fluxcalc.d:1028-1128
        double pLplus, uLplus, pRminus, uRminus;
        if (fabs(uL) <= aL) {
            const double raL = eb_rcp(aL);
            uLplus = 0.25 * raL * (uL + aL) * (uL + aL);
            pLplus = pL * uLplus * (raL * (2.0 - uL * raL));
        } else { uLplus = 0.5 * (uL + fabs(uL)); pLplus = (uL > 0.0) ? pL : 0.0; }
        if (fabs(uR) <= aR) {
            const double raR = eb_rcp(aR);
            uRminus = -0.25 * raR * (uR - aR) * (uR - aR);
            pRminus = pR * uRminus * (raR * (-2.0 - uR * raR));
        } else { uRminus = 0.5 * (uR - fabs(uR)); pRminus = (uR < 0.0) ? pR : 0.0; }
        const double mL = uLplus * rL, mR = uRminus * rR;
        F[Lay::iMass] = mL + mR;
        F[Lay::iXMom] = mL * L.vx + mR * R.vx;
        F[Lay::iYMom] = mL * L.vy + mR * R.vy;
        if (DIM == 3) F[Lay::iZMom] = mL * L.vz + mR * R.vz;
        set_normal<DIM>(d, mL * uL + mR * uR + (pLplus + pRminus), F);
        F[Lay::iEnergy] = mL * HL + mR * HR;
    } else if (FLUX == EB200_FLUX_LDFSS0 || FLUX == EB200_FLUX_LDFSS2) {   // fluxcalc.d:819-1025
        const double am = 0.5 * (aL + aR);
        double ML, MR;
        if (FLUX == EB200_FLUX_LDFSS0) { ML = uL * eb_rcp(aL); MR = uR * eb_rcp(aR); }
        else { const double ram = eb_rcp(am); ML = uL * ram; MR = uR * ram; }
        const double MpL = 0.25 * ((ML + 1.0) * (ML + 1.0));
        const double MmR = -0.25 * ((MR - 1.0) * (MR - 1.0));
        const double alphaL = 0.5 * (1.0 + sgn_d(ML)), alphaR = 0.5 * (1.0 - sgn_d(MR));
        const double betaL = -qmax(0.0, 1.0 - floor(fabs(ML))), betaR = -qmax(0.0, 1.0 - floor(fabs(MR)));
        const double PL = MpL * (2.0 - ML), PR = 0.25 * ((MR - 1.0) * (MR - 1.0)) * (2.0 + MR);
        const double DL = alphaL * (1.0 + betaL) - betaL * PL, DR = alphaR * (1.0 + betaR) - betaR * PR;
        const double sq = eb_sqrt(0.5 * (ML * ML + MR * MR)) - 1.0;
        const double Mhalf = 0.25 * betaL * betaR * (sq * sq);
        double cL, cR;
        if (FLUX == EB200_FLUX_LDFSS0) {
            cL = aL * rL * (alphaL * (1.0 + betaL) * ML - betaL * MpL - Mhalf);
            cR = aR * rR * (alphaR * (1.0 + betaR) * MR - betaR * MmR + Mhalf);
        } else {
            const double t = (pL - pR) * eb_rcp(pL + pR), ad = fabs(pL - pR);
            const double MhalfL = Mhalf * (1.0 - (t + 2.0 * (ad * eb_rcp(pL))));
            const double MhalfR = Mhalf * (1.0 + (t - 2.0 * (ad * eb_rcp(pR))));
            cL = am * rL * (alphaL * (1.0 + betaL) * ML - betaL * MpL - MhalfL);
            cR = am * rR * (alphaR * (1.0 + betaR) * MR - betaR * MmR + MhalfR);
        }
        F[Lay::iMass] = cL + cR;
        F[Lay::iXMom] = cL * L.vx + cR * R.vx;
        F[Lay::iYMom] = cL * L.vy + cR * R.vy;
        if (DIM == 3) F[Lay::iZMom] = cL * L.vz + cR * R.vz;
        set_normal<DIM>(d, (cL * uL + cR * uR) + (DL * pL + DR * pR), F);
        F[Lay::iEnergy] = cL * HL + cR * HR;
    } else {                                          // ausm_plus_up, fluxcalc.d:1415-1602
        const double a_half = 0.5 * (aR + aL), rah = eb_rcp(a_half);
        const double ML = uL * rah, MR = uR * rah;
        const double MbarSq = (uL * uL + uR * uR) * (0.5 * rah * rah);
        const double M0Sq = qmin(1.0, qmax(MbarSq, M_inf * M_inf));
        const double sqM0 = eb_sqrt(M0Sq);
        const double fa = sqM0 * (2.0 - sqM0);
        const double alpha = 0.1875 * (-4.0 + 5 * fa * fa);
        const double beta = 0.125;
        double M4p, P5p, M4m, P5m;
        if (fabs(ML) >= 1.0) { M4p = M1plus(ML); P5p = (ML > 0.0) ? 1.0 : 0.0; }
        else { const double a2 = M2plus(ML), b2 = M2minus(ML); M4p = a2 * (1.0 - 16.0 * beta * b2); P5p = a2 * ((2.0 - ML) - 16.0 * alpha * ML * b2); }
        if (fabs(MR) >= 1.0) { M4m = M1minus(MR); P5m = (MR < 0.0) ? 1.0 : 0.0; }
        else { const double a2 = M2plus(MR), b2 = M2minus(MR); M4m = b2 * (1.0 + 16.0 * beta * a2); P5m = b2 * ((-2.0 - MR) + 16.0 * alpha * MR * a2); }
        const double r_half = 0.5 * (rL + rR);
        const double Mp = -0.25 * eb_rcp(fa) * qmax((1.0 - MbarSq), 0.0) * (pR - pL) * eb_rcp(r_half * a_half * a_half);
        const double Pu = -0.75 * P5p * P5m * (rL + rR) * fa * a_half * (uR - uL);
        const double M_half = M4p + M4m + Mp;
        const double ru_half = a_half * M_half * ((M_half > 0.0) ? rL : rR);
        const double p_half = P5p * pL + P5m * pR + Pu;
        const bool fromL = (ru_half >= 0.0);
        F[Lay::iMass] = ru_half;
        F[Lay::iXMom] = ru_half * (fromL ? L.vx : R.vx);
        F[Lay::iYMom] = ru_half * (fromL ? L.vy : R.vy);
        if (DIM == 3) F[Lay::iZMom] = ru_half * (fromL ? L.vz : R.vz);
        set_normal<DIM>(d, ru_half * (fromL ? uL : uR) + p_half, F);
        F[Lay::iEnergy] = ru_half * (fromL ? HL : HR);
    }
}

template <int DIM, int FLUX>
__device__ __forceinline__ void flux_components(const Prim<1>& L, const Prim<1>& R, int d, bool entropy_fix, double M_inf, double* F)
{
    flux_components_pr<DIM, FLUX>(L, R, d, entropy_fix, M_inf, L.p * eb_rcp(L.rho), R.p * eb_rcp(R.rho), F);
}
#endif

#ifdef EB_FAST_MATH
// ---- throughput build, uniform-Cartesian blocks ---------------------------------------------------
// Reconstruction on raw differences with the constants of EbBlockDesc::uq (same limiter value as
// interp_v2: numerator and denominator are both scaled by 1/c^2, and so is epsilon).  Extrema
// clipping (limiters.d:43-51) is only *detected* here: inc*(inc - del) > 0 says the reconstructed
// value left the interval of its two neighbours; the caller then redoes the face with the
// clipping reconstruction.  With van Albada on uniform spacing that only happens at local
// extrema, where the limiter value is ~epsilon.
__device__ __forceinline__ void interp_uniform(const double* __restrict__ K, double eps2, double qL1, double qL0,
                                               double qR0, double qR1, double& qL, double& qR, int& outside)
{
    const double a = qL0 - qL1, b = qR0 - qL0, c = qR1 - qR0;
    const double ab = a * b, bc = b * c, b2 = b * b;
    const double nL = ab + fabs(ab) + eps2, dL = a * a + b2 + eps2;
    const double nR = bc + fabs(bc) + eps2, dR = c * c + b2 + eps2;
    const double rr = eb_rcp(dL * dR);
    const double iL = (nL * dR * rr) * (b * K[0] + a * K[1]);
    const double iR = (nR * dL * rr) * (c * K[2] + b * K[3]);
    qL = qL0 + iL;
    qR = qR0 - iR;
    outside = max(outside, max(__double2hiint(iL * (iL - b)), __double2hiint(iR * (iR - b))));
}

// Stencil velocities are (normal, t1, t2) by the way they were loaded (a renaming of the
// components, CartFrame conventions); F comes back as (mass, normal, t1, t2, energy).
template <int DIM, int FLUX, bool CLIP>
__device__ __forceinline__ void face_core_uniform(const EbParams& P, const EbGas* __restrict__ gas,
                                                  const double* __restrict__ K, const Stencil& s, double alpha,
                                                  const double* __restrict__ prim_fallback, long long cL0, long long cR0, double* F)
{
    Prim<1> L, R;
    const double eps2 = P.eps_va * K[4];
    const double gm1 = gas->Rgas * gas->Cvinv;
    int outside = 0;
    interp_uniform(K, eps2, s.v0[0], s.v0[1], s.v0[2], s.v0[3], L.vx, R.vx, outside);
    interp_uniform(K, eps2, s.v1[0], s.v1[1], s.v1[2], s.v1[3], L.vy, R.vy, outside);
    if (DIM == 3) interp_uniform(K, eps2, s.v2[0], s.v2[1], s.v2[2], s.v2[3], L.vz, R.vz, outside);
    else { L.vz = 0.0; R.vz = 0.0; }
    interp_uniform(K, eps2, s.rho[0], s.rho[1], s.rho[2], s.rho[3], L.rho, R.rho, outside);
    interp_uniform(K, eps2, s.u[0], s.u[1], s.u[2], s.u[3], L.u, R.u, outside);
    if (CLIP && outside > 0) {     // rare (local extrema): clip_to_limits, limiters.d:43-51
        L.vx = clip_to_limits(L.vx, s.v0[1], s.v0[2]); R.vx = clip_to_limits(R.vx, s.v0[1], s.v0[2]);
        L.vy = clip_to_limits(L.vy, s.v1[1], s.v1[2]); R.vy = clip_to_limits(R.vy, s.v1[1], s.v1[2]);
        if (DIM == 3) { L.vz = clip_to_limits(L.vz, s.v2[1], s.v2[2]); R.vz = clip_to_limits(R.vz, s.v2[1], s.v2[2]); }
        L.rho = clip_to_limits(L.rho, s.rho[1], s.rho[2]); R.rho = clip_to_limits(R.rho, s.rho[1], s.rho[2]);
        L.u = clip_to_limits(L.u, s.u[1], s.u[2]); R.u = clip_to_limits(R.u, s.u[1], s.u[2]);
    }
    L.a = s.aL; R.a = s.aR;
    L.p = L.rho * L.u * gm1;
    R.p = R.rho * R.u * gm1;
    // all four positive?  Compare the high words as integers (sign bit set or zero -> not positive; denormals
    // count as not positive and take the fall-back)
    if (min(min(__double2hiint(L.u), __double2hiint(L.rho)), min(__double2hiint(R.u), __double2hiint(R.rho))) <= 0) {
        // rare: first-order fall-back of a side whose reconstructed state is not physical (onedinterp.d:45-74)
        const long long total = P.total;
        if (L.u <= 0.0 || L.rho <= 0.0) {
            L.rho = s.rho[1]; L.u = s.u[1]; L.vx = s.v0[1]; L.vy = s.v1[1]; L.vz = s.v2[1];
            L.p = ldg(prim_fallback + 2 * total + cL0);
        }
        if (R.u <= 0.0 || R.rho <= 0.0) {
            R.rho = s.rho[2]; R.u = s.u[2]; R.vx = s.v0[2]; R.vy = s.v1[2]; R.vz = s.v2[2];
            R.p = ldg(prim_fallback + 2 * total + cR0);
        }
    }
    constexpr int SHOCK = FluxPair<FLUX>::shock, SMOOTH = FluxPair<FLUX>::smooth;
    L.massf[0] = 1.0; R.massf[0] = 1.0;
    if (SHOCK == EB200_FLUX_EFM || SMOOTH == EB200_FLUX_EFM) { L.T = L.u * gas->Cvinv; R.T = R.u * gas->Cvinv; }   // efm reads T
    if (FluxPair<FLUX>::adaptive && alpha > 0.0) {
        if constexpr (SHOCK < EB200_FLUX_ROE) flux_components<DIM, SHOCK>(L, R, 0, false, P.M_inf, F);
        else basic_flux<DIM, 1, EB200_GAS_IDEAL, SHOCK>(P, gas, L, R, F);
    } else {
        if constexpr (SMOOTH < EB200_FLUX_ROE) flux_components<DIM, SMOOTH>(L, R, 0, P.entropy_fix != 0, P.M_inf, F);
        else basic_flux<DIM, 1, EB200_GAS_IDEAL, SMOOTH>(P, gas, L, R, F);
    }
}
#endif

// Reconstruction + thermo + flux.  ROT: general-metric path, the stencil velocities are global and
// rotate with `fr`, F comes back in the face frame.  !ROT: uniform-Cartesian path, the face frame is
// a renaming of the components ((n,t1,t2) = (e_d, e_d+1, e_d+2) in 3D; 2D: i-face (x,-y), j-face
// (y,x)): the components are reconstructed under their own names (each reconstruction is
// independent, so this is bit-identical to reconstructing in the face frame), only the two
// reconstructed velocities are renamed by d, and F comes back with momentum in (x,y,z) order.
template <int DIM, int FLUX, bool CLIP, bool ROT>
__device__ __forceinline__ void face_core(const EbParams& P, const EbGas* __restrict__ gas, const EbWeights& w,
                                          Stencil& s, const Frame& fr, int d, double alpha,
                                          const double* __restrict__ prim_fallback, long long cL0, long long cR0, double* F)
{
    typedef Layout<DIM, 1> Lay;
    Prim<1> L, R;
    if (ROT && P.local_frame) {
#pragma unroll
        for (int m = 0; m < 4; ++m) to_local<DIM>(fr, s.v0[m], s.v1[m], s.v2[m]);
    }
    const double eps = P.eps_va;
    interp_v2<CLIP>(w, eps, s.v0[0], s.v0[1], s.v0[2], s.v0[3], L.vx, R.vx);
    interp_v2<CLIP>(w, eps, s.v1[0], s.v1[1], s.v1[2], s.v1[3], L.vy, R.vy);
    if (DIM == 3) interp_v2<CLIP>(w, eps, s.v2[0], s.v2[1], s.v2[2], s.v2[3], L.vz, R.vz);
    else { L.vz = 0.0; R.vz = 0.0; }
    interp_v2<CLIP>(w, eps, s.rho[0], s.rho[1], s.rho[2], s.rho[3], L.rho, R.rho);
    interp_v2<CLIP>(w, eps, s.u[0], s.u[1], s.u[2], s.u[3], L.u, R.u);
    L.a = s.aL; R.a = s.aR;
    L.massf[0] = 1.0; R.massf[0] = 1.0;
    // ideal gas update_thermo_from_rhou with fall-back to the cell state (onedinterp.d:45-74)
    if (L.u <= 0.0 || L.rho <= 0.0) {
        const long long total = P.total;
        L.rho = s.rho[1]; L.u = s.u[1]; L.vx = s.v0[1]; L.vy = s.v1[1]; L.vz = s.v2[1];
        L.p = ldg(prim_fallback + 2 * total + cL0); L.T = ldg(prim_fallback + 3 * total + cL0);
    } else { L.T = L.u * gas->Cvinv; L.p = L.rho * gas->Rgas * L.T; }
    if (R.u <= 0.0 || R.rho <= 0.0) {
        const long long total = P.total;
        R.rho = s.rho[2]; R.u = s.u[2]; R.vx = s.v0[2]; R.vy = s.v1[2]; R.vz = s.v2[2];
        R.p = ldg(prim_fallback + 2 * total + cR0); R.T = ldg(prim_fallback + 3 * total + cR0);
    } else { R.T = R.u * gas->Cvinv; R.p = R.rho * gas->Rgas * R.T; }
    if (ROT) {
#ifdef EB_FAST_MATH
        // the reference's trip back to the global frame and into the face frame again is the identity
        if (!P.local_frame) { to_local<DIM>(fr, L.vx, L.vy, L.vz); to_local<DIM>(fr, R.vx, R.vy, R.vz); }
#else
        if (P.local_frame) {   // reference: back to global (onedinterp.d:979-987), then into the face frame again
            to_global<DIM>(fr, L.vx, L.vy, L.vz); to_global<DIM>(fr, R.vx, R.vy, R.vz);
        }
        to_local<DIM>(fr, L.vx, L.vy, L.vz); to_local<DIM>(fr, R.vx, R.vy, R.vz);
#endif
    }
#ifdef EB_FAST_MATH
    constexpr bool has_component_form = (FluxPair<FLUX>::shock < EB200_FLUX_ROE) && (FluxPair<FLUX>::smooth < EB200_FLUX_ROE);
    if constexpr (!ROT && has_component_form) {
        // 2D i-faces have t1 = -y; the component form never looks at the sign of a tangential component
        if (FluxPair<FLUX>::adaptive && alpha > 0.0) flux_components<DIM, FluxPair<FLUX>::shock>(L, R, d, false, P.M_inf, F);
        else flux_components<DIM, FluxPair<FLUX>::smooth>(L, R, d, P.entropy_fix != 0, P.M_inf, F);
        return;
    }
#endif
    if (ROT) {
    } else if (DIM == 3) {
        const double lx = L.vx, ly = L.vy, lz = L.vz, rx = R.vx, ry = R.vy, rz = R.vz;
        L.vx = pick3(d, lx, ly, lz); L.vy = pick3(d, ly, lz, lx); L.vz = pick3(d, lz, lx, ly);
        R.vx = pick3(d, rx, ry, rz); R.vy = pick3(d, ry, rz, rx); R.vz = pick3(d, rz, rx, ry);
    } else {
        const double lx = L.vx, ly = L.vy, rx = R.vx, ry = R.vy;
        L.vx = (d == 0) ? lx : ly; L.vy = (d == 0) ? -ly : lx;
        R.vx = (d == 0) ? rx : ry; R.vy = (d == 0) ? -ry : rx;
    }
    flux_in_face_frame<DIM, 1, EB200_GAS_IDEAL, FLUX>(P, gas, L, R, alpha, F);
    if (!ROT) {              // momentum flux back to (x, y, z) order
        if (DIM == 3) {
            const double f0 = F[Lay::iXMom], f1 = F[Lay::iYMom], f2 = F[Lay::iZMom];
            F[Lay::iXMom] = pick3(d, f0, f2, f1); F[Lay::iYMom] = pick3(d, f1, f0, f2); F[Lay::iZMom] = pick3(d, f2, f1, f0);
        } else {
            const double f0 = F[Lay::iXMom], f1 = F[Lay::iYMom];
            F[Lay::iXMom] = (d == 0) ? f0 : f1; F[Lay::iYMom] = (d == 0) ? -f1 : f0;
        }
    }
}

// Shared-memory tile of one k-plane: NF fields x ROWS x COLS doubles (halo of 2 on each side).
template <int DIM, int TY>
struct Tile {
    static constexpr int NF = 6;                           // rho, u, a, vx, vy, vz: prim fields (0,1), (4,5), (6,7) = three TMA boxes
    static constexpr int ROWS = TY + 4, COLS = EB_V2_COLS;
    static constexpr int FSZ = ROWS * COLS;                // doubles per field
    static constexpr int SIZE = NF * FSZ;
    static constexpr int F_RHO = 0, F_U = 1, F_A = 2, F_V = 3;
    __device__ static constexpr int prim_index(int f) { return (f < 2) ? f : f + 2; }
};

// Frame conventions of the uniform-Cartesian path (the frames Eilmer builds for a box grid:
// 3D quad_properties with the vertex cycles of sfluidblock.d:735-850, 2D fvinterface.d:301-332):
//   3D: face d has (n, t1, t2) = (e_d, e_{d+1}, e_{d+2});  2D: i-face (x, -y), j-face (y, x).
// Checked on the host when the block is classified.
template <int DIM, int D>
struct CartFrame {
    static constexpr int c0 = D, c1 = (DIM == 3) ? (D + 1) % 3 : 1 - D, c2 = (DIM == 3) ? (D + 2) % 3 : 2;
    static constexpr bool neg1 = (DIM == 2 && D == 0);     // local y = -v_y on 2D i-faces
};

template <int DIM, int FLUX, bool CART, bool CLIP, int TY>
__global__ void __launch_bounds__(32 * TY, EB_V2_MIN_CTAS)
flux_update_kernel_v2(const EbParams P, const EbGas* __restrict__ gas, const EbBlockDesc* __restrict__ descs, int nblocks,
                      const EbArena A, const EbStageArgs S)
{
    typedef Layout<DIM, 1> Lay;
    typedef Tile<DIM, TY> T;
    constexpr int NCQ = Lay::NCQ;
    constexpr int NT = 32 * TY;
    extern __shared__ __align__(128) double smem[];
    // layout: tile[2][SIZE] | fW[NCQ][TY][32] | fX[2][NCQ][TY] | fS[2][NCQ][TY+1][32] | fB[NCQ][TY][32] | ring[2][6][TY][32] | desc
    //   fW: west-face fluxes, exchanged inside a warp only (single buffer + __syncwarp); fX: the faces east of
    //   lane 31, written by another warp (double buffered like fS); fB: bottom-face flux, private to its thread;
    //   ring (3D): the thread's own cell values (rho, u, a, vx, vy, vz) of the two planes below, for the k-stencil.
    double* tile = smem;
    double* fWs = tile + 2 * T::SIZE;
    double* fXs = fWs + NCQ * TY * 32;
    double* fSs = fXs + 2 * NCQ * TY;
    double* fBs = fSs + 2 * NCQ * (TY + 1) * 32;
    double* ring = fBs + NCQ * TY * 32;
    constexpr int RING = (DIM == 3) ? 2 * 6 * TY * 32 : 0;
    EbBlockDesc& D = *reinterpret_cast<EbBlockDesc*>(ring + RING);
    __shared__ int s_blk;
    __shared__ __align__(8) unsigned long long s_bar[2];      // tile[buf] has landed (TMA staging)

    const int lane = threadIdx.x, wy = threadIdx.y;
    const int tid = wy * 32 + lane;
    const long long cta = S.tile_list ? (long long)S.tile_list[blockIdx.x] : (long long)blockIdx.x;
    const bool use_tma = (S.tmaps != nullptr);
    if (tid == 0) {
        int lo = 0, hi = nblocks - 1;
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (descs[mid].tile0 <= cta) lo = mid; else hi = mid - 1; }
        s_blk = lo;
        if (use_tma) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); }
    }
    __syncthreads();
    {
        const int* src = reinterpret_cast<const int*>(&descs[s_blk]);
        int* dst = reinterpret_cast<int*>(&D);
        for (int n = tid; n < (int)(sizeof(EbBlockDesc) / sizeof(int)); n += NT) dst[n] = src[n];
    }
    __syncthreads();
    if ((D.cartesian != 0) != CART || D.v3) return;      // v3: tiled for (and run by) flux_update_kernel_v3

    const long long t = cta - D.tile0;
    const int ti = (int)(t % D.tiles_i);
    const int tj = (int)((t / D.tiles_i) % D.tiles_j);
    const int tm = (int)(t / ((long long)D.tiles_i * D.tiles_j));
    const int i0 = ti * 32, j0 = tj * TY;
    const int i = i0 + lane, j = j0 + wy;
    const int nic = D.nic, njc = D.njc, nkc = D.nkc;
    const int NI = D.NI, NJ = D.NJ;
    const long long sj = D.stride[1], sk = D.stride[2];
    const long long total = P.total;
    const int k0 = (DIM == 3) ? tm * D.chunk_m : 0;
    const int k1 = (DIM == 3) ? min(nkc, k0 + D.chunk_m) : 1;

    const bool cell_ok = (i < nic) && (j < njc);
    // second trip, for two warps of the CTA: one does the faces east of lane 31 (row = lane), the other
    // the south faces of row TY.  The pair rotates from plane to plane so that the four schedulers of
    // the SM see the same load.
    const bool extraE_can = (lane < TY) && (i0 + 32 <= nic) && (j0 + lane < njc);
    const bool extraN_can = (j0 + TY <= njc) && (i < nic);

    // plane k (interior index; padded k + kg) -> tile[buf]: three TMA boxes issued by one thread, or (odd NI,
    // no driver entry point) cp.async by everybody, tile positions tid and tid + NT of ROWS*COLS
    auto stage_plane = [&](int k, int buf) {
        double* dst = tile + buf * T::SIZE;
        if (use_tma) {
            if (tid == 0) {
                const void* tm = reinterpret_cast<const char*>(S.tmaps) + (size_t)s_blk * 128;
                mbar_expect_tx(&s_bar[buf], (unsigned)(T::SIZE * sizeof(double)));
                tma_load_4d(dst, tm, &s_bar[buf], i0, j0, k + D.kg, 0);
                tma_load_4d(dst + 2 * T::FSZ, tm, &s_bar[buf], i0, j0, k + D.kg, 4);
                tma_load_4d(dst + 4 * T::FSZ, tm, &s_bar[buf], i0, j0, k + D.kg, 6);
            }
            return;
        }
        const long long plane = D.cell0 + (long long)(k + D.kg) * NJ * NI;
#pragma unroll
        for (int n = 0; n < 2; ++n) {
            const int p = tid + n * NT;
            const int r = p / T::COLS, cc = p % T::COLS;
            if ((p < T::ROWS * T::COLS) && (j0 + r < NJ) && (i0 + cc < NI)) {
                const double* src = S.prim_in + plane + (long long)(j0 + r) * NI + (i0 + cc);
#pragma unroll
                for (int f = 0; f < T::NF; ++f) cp_async8(dst + f * T::FSZ + p, src + (long long)T::prim_index(f) * total);
            }
        }
        cp_async_commit();
    };
    // m = k - k0 counts the planes of this CTA: plane m lives in tile[m & 1], its barrier is in phase (m >> 1) & 1
    auto wait_plane = [&](int m) {
        if (use_tma) mbar_wait(&s_bar[m & 1], (unsigned)((m >> 1) & 1));
        else cp_async_wait_all();
    };

    double acc[NCQ];
#pragma unroll
    for (int q = 0; q < NCQ; ++q) acc[q] = 0.0;
    bool fail = false;
    int n_invalid = 0;

    const int kend = (DIM == 3) ? k1 : 0;
    stage_plane(k0, 0);
    double* const myring = ring + wy * 32 + lane;           // [slot][field] at stride TY*32
    if (DIM == 3 && cell_ok) {
        // own cell values of the planes k0-2 (slot of even/odd plane parity) and k0-1
        const long long c0 = D.cell0 + ((long long)(k0 + D.kg) * NJ + (j + EB_NG)) * NI + (i + EB_NG);
#pragma unroll
        for (int m = 1; m <= 2; ++m) {
            double* r = myring + ((k0 - m) & 1) * 6 * TY * 32;
#pragma unroll
            for (int f = 0; f < 6; ++f) r[f * TY * 32] = ldg(S.prim_in + (long long)T::prim_index(f) * total + c0 - m * sk);
        }
    }
    wait_plane(0);
    __syncthreads();

    for (int k = k0; k <= kend; ++k) {
        const int buf = (k - k0) & 1;
        const bool plane_has_cells = (DIM == 3) ? (k < k1) : true;
        // prefetch the next plane (its own cells are needed for the top face of the chunk too)
        if (DIM == 3 && k < kend) stage_plane(k + 1, buf ^ 1);
        const double* tl = tile + buf * T::SIZE;
        double* fW = fWs;
        double* fX = fXs + buf * NCQ * TY;
        double* fS = fSs + buf * NCQ * (TY + 1) * 32;
        double* fB = fBs;
        const long long c = D.cell0 + ((long long)(k + D.kg) * NJ + (j + EB_NG)) * NI + (i + EB_NG);
        __syncwarp();          // fW is reused: every lane has read the previous plane's west fluxes

        // ---------------- the faces of this plane: one loop, ONE inlined copy of the face arithmetic -------
        // job 0: west face (d=0), job 1: south face (d=1), job 2: bottom face (d=2, 3D),
        // job 3: second trip for warp 0 (faces east of lane 31) and warp 1 (south faces of row TY).
        // Every flux goes straight to shared memory; the momentum components are put into their
        // global-frame slots by address, not by moving values around.
        const int xw = (TY >= 8) ? ((2 * (k - k0)) & (TY - 1)) : 0;   // warps xw, xw+1 do the second trip on this plane
#pragma unroll 1
        for (int job = 0; job < 4; ++job) {
            if (job == 2 && DIM != 3) continue;
            if (job == 3 && ((unsigned)(wy - xw) > 1u || !plane_has_cells)) break;
            const int d = (job < 3) ? job : (wy - xw);
            int row = wy + 2, col = lane + 2;            // tile coordinates of the plus-side cell of the face
            int fi = i, fj = j;                          // its interior indices
            bool active;
            if (job == 0) active = plane_has_cells && (i <= nic) && (j < njc);
            else if (job == 1) active = plane_has_cells && (i < nic) && (j <= njc);
            else if (job == 2) active = cell_ok;
            else if (d == 0) { row = lane + 2; col = 34; fi = i0 + 32; fj = j0 + lane; active = extraE_can; }
            else { row = TY + 2; fj = j0 + TY; active = extraN_can; }
            if (!active) continue;
            // where this face's flux goes: component q at out[q * qstride]
            double* out;
            int qstride;
            if (job == 3 && d == 0) { out = fX + (row - 2); qstride = TY; }
            else if (d == 0) { out = fW + wy * 32 + lane; qstride = TY * 32; }
            else if (d == 1) { out = fS + (row - 2) * 32 + (col - 2); qstride = (TY + 1) * 32; }
            else { out = fB + wy * 32 + lane; qstride = TY * 32; }
            const long long cf = (job < 3) ? c : c + (long long)(fj - j) * NI + (fi - i);
            const long long st = (d == 0) ? 1 : ((d == 1) ? sj : sk);
            int bcf = -1;
            if (D.outflow_flux_faces) {
                if (d == 0) { if (fi == 0) bcf = EB200_WEST; else if (fi == nic) bcf = EB200_EAST; }
                else if (d == 1) { if (fj == 0) bcf = EB200_SOUTH; else if (fj == njc) bcf = EB200_NORTH; }
                else { if (k == 0) bcf = EB200_BOTTOM; else if (k == nkc) bcf = EB200_TOP; }
            }
            if (bcf >= 0 && D.bc_kind[bcf] == EB200_BC_OUTFLOW_SIMPLE_FLUX) {
                const int hi = bcf & 1;
                Prim<1> fs;
                load_prim<1>(fs, S.prim_in, total, hi ? cf - st : cf);
                if (DIM == 2) fs.vz = 0.0;
                double nx, ny, nz;
                if (CART) { nx = D.nvec[d][0]; ny = D.nvec[d][1]; nz = D.nvec[d][2]; }
                else { nx = ldg(A.face[d] + cf); ny = ldg(A.face[d] + total + cf); nz = (DIM == 3) ? ldg(A.face[d] + 2 * total + cf) : 0.0; }
                double F[NCQ];
                outflow_flux<DIM, 1>(fs, hi ? 1 : -1, nx, ny, nz, F);
#pragma unroll
                for (int q = 0; q < NCQ; ++q) out[q * qstride] = F[q];
                continue;
            }
            const double* q0 = tl + row * T::COLS + col;     // own (R0) cell, field 0
            Stencil s;
            // ---- gather the stencil straight into its registers.  On the Cartesian path the face frame is a
            //      renaming of the velocity components (CartFrame conventions): pick the fields by address.
#ifdef EB_FAST_MATH
            constexpr bool UNIFORM = CART;       // velocities loaded as (normal, t1, t2): a renaming by address
#else
            constexpr bool UNIFORM = false;
#endif
            // 3D: (d, d+1, d+2) mod 3;  2D: (d, 1-d) -- the sign of t1 on 2D i-faces plays no role in
            // the component form of the flux
            const int c0 = UNIFORM ? d : 0;
            const int c1 = UNIFORM ? ((DIM == 3) ? ((d == 2) ? 0 : d + 1) : 1 - d) : 1;
            const int c2 = UNIFORM ? ((DIM == 3) ? ((d == 0) ? 2 : d - 1) : 2) : 2;
            if (d < 2) {
                const int so = (d == 0) ? 1 : T::COLS;       // tile stride along d
                const int o0 = (T::F_V + c0) * T::FSZ, o1 = (T::F_V + c1) * T::FSZ, o2 = (T::F_V + c2) * T::FSZ;
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const double* qm = q0 + (m - 2) * so;
                    s.rho[m] = qm[T::F_RHO * T::FSZ]; s.u[m] = qm[T::F_U * T::FSZ];
                    s.v0[m] = qm[o0]; s.v1[m] = qm[o1];
                    s.v2[m] = (DIM == 3) ? qm[o2] : 0.0;
                }
                s.aL = (q0 - so)[T::F_A * T::FSZ];
            } else if (UNIFORM) {
                // own column: plane k from the tile, planes k-2 and k-1 from the thread's ring, plane k+1 from
                // global memory (it is still on its way into the other tile buffer); d == 2: (z, x, y)
                const double* pin = S.prim_in;
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    if (m == 2) {
                        s.rho[m] = q0[T::F_RHO * T::FSZ]; s.u[m] = q0[T::F_U * T::FSZ];
                        s.v0[m] = q0[(T::F_V + 2) * T::FSZ]; s.v1[m] = q0[(T::F_V + 0) * T::FSZ]; s.v2[m] = q0[(T::F_V + 1) * T::FSZ];
                    } else if (m == 3) {
                        const double* pm = pin + cf + sk;
                        s.rho[m] = ldg(pm); s.u[m] = ldg(pm + total);
                        s.v0[m] = ldg(pm + 7 * total); s.v1[m] = ldg(pm + 5 * total); s.v2[m] = ldg(pm + 6 * total);
                    } else {
                        const double* r = myring + ((k + m) & 1) * 6 * TY * 32;      // plane k-2+m has the parity of k+m
                        s.rho[m] = r[T::F_RHO * TY * 32]; s.u[m] = r[T::F_U * TY * 32];
                        s.v0[m] = r[(T::F_V + 2) * TY * 32]; s.v1[m] = r[(T::F_V + 0) * TY * 32]; s.v2[m] = r[(T::F_V + 1) * TY * 32];
                    }
                }
                s.aL = (myring + ((k + 1) & 1) * 6 * TY * 32)[T::F_A * TY * 32];
            } else {
                // own column: plane k from the tile, planes k-2 and k-1 from the thread's ring, plane k+1 from global memory
                const double* pin = S.prim_in;
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    if (m == 2) {
                        s.rho[m] = q0[T::F_RHO * T::FSZ]; s.u[m] = q0[T::F_U * T::FSZ];
                        s.v0[m] = q0[(T::F_V + 0) * T::FSZ]; s.v1[m] = q0[(T::F_V + 1) * T::FSZ]; s.v2[m] = q0[(T::F_V + 2) * T::FSZ];
                    } else if (m == 3) {
                        const double* pm = pin + cf + sk;
                        s.rho[m] = ldg(pm); s.u[m] = ldg(pm + total);
                        s.v0[m] = ldg(pm + 5 * total); s.v1[m] = ldg(pm + 6 * total); s.v2[m] = ldg(pm + 7 * total);
                    } else {
                        const double* r = myring + ((k + m) & 1) * 6 * TY * 32;
                        s.rho[m] = r[T::F_RHO * TY * 32]; s.u[m] = r[T::F_U * TY * 32];
                        s.v0[m] = r[(T::F_V + 0) * TY * 32]; s.v1[m] = r[(T::F_V + 1) * TY * 32]; s.v2[m] = r[(T::F_V + 2) * TY * 32];
                    }
                }
                s.aL = (myring + ((k + 1) & 1) * 6 * TY * 32)[T::F_A * TY * 32];
            }
            s.aR = q0[T::F_A * T::FSZ];
            Frame fr;
            EbWeights wl;
            if (!CART) {
                load_frame<DIM>(fr, A.face[d], total, cf);
                const double* ln = A.len[d];
                l2r2_prepare(wl, ldg(ln + cf - 2 * st), ldg(ln + cf - st), ldg(ln + cf), ldg(ln + cf + st));
            }
            const EbWeights& w = CART ? D.w[d] : wl;
            double Fl[NCQ];
            const double alpha = FluxPair<FLUX>::adaptive ? A.Sf[d][cf] : 0.0;
#ifdef EB_FAST_MATH
            if (UNIFORM) {
                face_core_uniform<DIM, FLUX, CLIP>(P, gas, D.uq[d], s, alpha, S.prim_in, cf - st, cf, Fl);
                out[0] = Fl[Lay::iMass];
                out[(Lay::iXMom + c0) * qstride] = Fl[Lay::iXMom];
                out[(Lay::iXMom + c1) * qstride] = Fl[Lay::iYMom];
                if (DIM == 3) out[(Lay::iXMom + c2) * qstride] = Fl[Lay::iZMom];
                out[Lay::iEnergy * qstride] = Fl[Lay::iEnergy];
                continue;
            }
#endif
            face_core<DIM, FLUX, CLIP, !CART>(P, gas, w, s, fr, d, alpha, S.prim_in, cf - st, cf, Fl);
            if (!CART) {     // momentum flux back to the global frame (fluxcalc.d:169-175)
                double fx = Fl[Lay::iXMom], fy = Fl[Lay::iYMom], fz = (DIM == 3) ? Fl[Lay::iZMom] : 0.0;
                to_global<DIM>(fr, fx, fy, fz);
                Fl[Lay::iXMom] = fx; Fl[Lay::iYMom] = fy;
                if (DIM == 3) Fl[Lay::iZMom] = fz;
            }
#pragma unroll
            for (int q = 0; q < NCQ; ++q) out[q * qstride] = Fl[q];
        }

        if (DIM == 3 && cell_ok && plane_has_cells) {
            // this plane's own cell replaces plane k-2 in the ring (same parity); nobody else reads the slot
            const double* q0 = tl + (wy + 2) * T::COLS + (lane + 2);
            double* r = myring + (k & 1) * 6 * TY * 32;
#pragma unroll
            for (int f = 0; f < 6; ++f) r[f * TY * 32] = q0[f * T::FSZ];
        }
        if (DIM == 3 && k < kend) wait_plane(k + 1 - k0);
        __syncthreads();

        // finish the cell of the previous plane: its top face is this plane's bottom face
        if (DIM == 3 && k > k0 && cell_ok) {
            const long long cp = c - sk;
            const double areaT = CART ? D.area[2] : ldg(A.face[2] + 9 * total + c);
            const double vol_inv = CART ? D.vol_inv : eb_div(1.0, ldg(A.vol + cp));
            double dUdt[NCQ];
#pragma unroll
            for (int q = 0; q < NCQ; ++q) { double si = acc[q] - fB[(q * TY + wy) * 32 + lane] * areaT; dUdt[q] = vol_inv * si + 0.0; }
            long long p0, p1, p2;
            push_targets<DIM>(D, i, j, k - 1, cp, p0, p1, p2);
            finish_cell<DIM, EB200_GAS_IDEAL, 1>(P, gas, S, total, cp, dUdt, fail, n_invalid, nullptr, nullptr, p0, p1, p2);
        }

        if (plane_has_cells && cell_ok) {
            double aW, aE, aS, aN, aB = 0.0;
            if (CART) { aW = aE = D.area[0]; aS = aN = D.area[1]; if (DIM == 3) aB = D.area[2]; }
            else {
                aW = ldg(A.face[0] + 9 * total + c); aE = ldg(A.face[0] + 9 * total + c + 1);
                aS = ldg(A.face[1] + 9 * total + c); aN = ldg(A.face[1] + 9 * total + c + sj);
                if (DIM == 3) aB = ldg(A.face[2] + 9 * total + c);
            }
#pragma unroll
            for (int q = 0; q < NCQ; ++q) {
                const double FW = fW[(q * TY + wy) * 32 + lane];
                const double FE = (lane == 31) ? fX[q * TY + wy] : fW[(q * TY + wy) * 32 + lane + 1];
                const double FS = fS[(q * (TY + 1) + wy) * 32 + lane], FN = fS[(q * (TY + 1) + wy + 1) * 32 + lane];
                double si = FW * aW;          // 0 - F*(-A)
                si = si - FE * aE;
                si = si + FS * aS;
                si = si - FN * aN;
                if (DIM == 3) si = si + fB[(q * TY + wy) * 32 + lane] * aB;
                acc[q] = si;
            }
        }

        if (DIM == 2 && cell_ok) {
            const double vol = CART ? D.vol : ldg(A.vol + c);
            const double vol_inv = CART ? D.vol_inv : eb_div(1.0, vol);
            double Qy = 0.0;
            if (P.axisymmetric) {      // fvcell.d:1161-1165
                const double axy = CART ? D.areaxy : ldg(A.areaxy + c);
                Qy = ldg(S.prim_in + 2 * total + c) * axy / vol;
            }
            double dUdt[NCQ];
#pragma unroll
            for (int q = 0; q < NCQ; ++q) dUdt[q] = vol_inv * acc[q] + ((q == Lay::iYMom) ? Qy : 0.0);
            long long p0, p1, p2;
            push_targets<DIM>(D, i, j, 0, c, p0, p1, p2);
            finish_cell<DIM, EB200_GAS_IDEAL, 1>(P, gas, S, total, c, dUdt, fail, n_invalid, nullptr, nullptr, p0, p1, p2);
        }
    }

    unsigned any_fail = __ballot_sync(0xffffffffu, fail);
    int inv = n_invalid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) inv += __shfl_down_sync(0xffffffffu, inv, o);
    if (lane == 0) {
        if (any_fail) atomicOr(&S.status[0], 1);
        if (inv) atomicAdd(&S.status[S.stage], inv);
    }
}

template <int DIM, int TY>
constexpr size_t v2_smem_bytes()
{
    typedef Tile<DIM, TY> T;
    constexpr int NCQ = Layout<DIM, 1>::NCQ;
    return sizeof(double) * (2 * T::SIZE + NCQ * TY * 32 + 2 * NCQ * TY + 2 * NCQ * (TY + 1) * 32 + NCQ * TY * 32 +
                             ((DIM == 3) ? 2 * 6 * TY * 32 : 0)) + sizeof(EbBlockDesc);
}

template <int DIM, int FLUX, bool CART, bool CLIP>
void launch_one_v2(const EbParams& P, const EbGas* gas, const EbBlockDesc* desc, int nblocks, long long ncta,
                   const EbArena& A, const EbStageArgs& S, cudaStream_t st)
{
    constexpr int TY = EB_V2_TY;
    const size_t smem = v2_smem_bytes<DIM, TY>();
    auto kern = flux_update_kernel_v2<DIM, FLUX, CART, CLIP, TY>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured = true;
    }
    kern<<<(unsigned)ncta, dim3(32, TY), smem, st>>>(P, gas, desc, nblocks, A, S);
}

// ideal gas, interpolation_order = 2, apply_limiter = true
template <int FLUX>
void launch_flux_update_v2_impl(const EbParams& P, const EbGas* gas, const EbBlockDesc* desc, int nblocks,
                                long long ncta, const EbArena& A, const EbStageArgs& S, int which, cudaStream_t st)
{
#define EB_LAUNCH2(DIM)                                                                                          \
    do {                                                                                                         \
        if (P.extrema_clipping) {                                                                                \
            if (which & 1) launch_one_v2<DIM, FLUX, true, true>(P, gas, desc, nblocks, ncta, A, S, st);           \
            if (which & 2) launch_one_v2<DIM, FLUX, false, true>(P, gas, desc, nblocks, ncta, A, S, st);          \
        } else {                                                                                                 \
            if (which & 1) launch_one_v2<DIM, FLUX, true, false>(P, gas, desc, nblocks, ncta, A, S, st);          \
            if (which & 2) launch_one_v2<DIM, FLUX, false, false>(P, gas, desc, nblocks, ncta, A, S, st);         \
        }                                                                                                        \
    } while (0)
    if (P.dims == 3) EB_LAUNCH2(3); else EB_LAUNCH2(2);
#undef EB_LAUNCH2
}

}  // namespace EB_NS
