// flux_kernel.cuh -- the fused per-stage kernel of the explicit update.
//
// One launch per stage covers every local block (CTA -> (block, tile) through the block
// table).  A CTA owns a 32 x TY tile of cells in the (i, j) plane and marches along k
// (3D).  Per plane every thread computes the faces on the minus side of its cell
// (west, south, bottom): reconstruction -> interface flux, once per face.  The east flux
// comes from the neighbouring lane by warp shuffle, the north flux from the next warp
// through shared memory, the top flux is the bottom flux of the next plane kept in
// registers.  The faces on the far edge of the tile (east of lane 31, north of the last
// row) are computed by two warps in an extra pass.  The thread then forms the residual
// in the reference's summation order (W,E,S,N,B,T; fvcell.d:824-854), applies the stage
// update (simcore_gasdynamic_step.d:1250-1357), decodes the new conserved state to
// primitives (fvcell.d:586-821) and writes them to the other primitive buffer.
//
// Roofline: FP64 stencil work; algorithmic HBM traffic per cell and stage is
// read prim(8) + U0(ncq) [+ earlier dUdt], write prim(8) + dUdt or U (ncq).
#pragma once
#include "device_math.cuh"

#ifndef EB_TILE_Y
#define EB_TILE_Y 8          // rows of cells per CTA tile (CTA = 32 x EB_TILE_Y threads)
#endif
#ifndef EB_MIN_CTAS
#define EB_MIN_CTAS 1        // second argument of __launch_bounds__: resident CTAs per SM to aim for
#endif
#ifndef EB_MIN_CTAS_TPG
#define EB_MIN_CTAS_TPG 2    // thermally perfect gas: the Newton solves are latency-bound, two CTAs (128 registers,
                             // more spills) beat one with 255 registers by 23 %; three (80 registers) lose again
#endif

namespace EB_NS {

// bc/boundary_flux_effect.d:573-643 compute_outflow_flux (gvel = 0)
template <int DIM, int NSP>
__device__ __forceinline__ void outflow_flux(const Prim<NSP>& fs, int outsign, double nx, double ny, double nz, double* F)
{
    typedef Layout<DIM, NSP> Lay;
    double vn = fs.vx * nx + fs.vy * ny + fs.vz * nz;
    double mass_flux = fs.rho * vn;
    if ((outsign * mass_flux) > 0.0) {
        F[Lay::iMass] = mass_flux;
        F[Lay::iXMom] = fs.p * nx + fs.vx * mass_flux;
        F[Lay::iYMom] = fs.p * ny + fs.vy * mass_flux;
        if (DIM == 3) F[Lay::iZMom] = fs.p * nz + fs.vz * mass_flux;
        double utot = fs.u + 0.5 * (fs.vx * fs.vx + fs.vy * fs.vy + fs.vz * fs.vz);
        F[Lay::iEnergy] = mass_flux * utot + fs.p * vn;
        if (NSP > 1) {
#pragma unroll
            for (int i = 0; i < NSP; ++i) F[Lay::iSpecies + i] = mass_flux * fs.massf[i];
        }
    } else {
        F[Lay::iMass] = 0.0;
        F[Lay::iXMom] = nx * fs.p;
        F[Lay::iYMom] = ny * fs.p;
        if (DIM == 3) F[Lay::iZMom] = nz * fs.p;
        F[Lay::iEnergy] = fs.p * 0.0;
        if (NSP > 1) {
#pragma unroll
            for (int i = 0; i < NSP; ++i) F[Lay::iSpecies + i] = 0.0;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Faces next to a boundary WITHOUT ghost-cell data (EB200_BC_WALL_WITH_SLIP1, bc.lua:783-806): the wall face has no
// cells on one side, the next face in has one (sfluidblock.d:455-530, 548-611).  onedinterp.d:117-273 picks the
// stencil from those counts: l0r2 / l2r0 (linear extrapolation, or the cell copy when extrema_clipping is on,
// :1451-1456), l1r2 / l2r1 (:386-454); the wall face takes compute_flux_at_left_wall / _right_wall
// (fluxcalc.d:41-51, 187-385).  Few faces: out of line, written for clarity, the reference's order of operations.
enum { EB_ST_L2R2 = 0, EB_ST_L2R1 = 1, EB_ST_L1R2 = 2, EB_ST_L2R0 = 3, EB_ST_L0R2 = 4 };

struct EbOneSided {
    int mode;
    bool limiter, clip;
    double eps;
    EbWeights w;
    double w0, w1;
    __device__ double weight(double q0, double q1) const      // onedinterp.d:477-485 weight_scalar
    {
        double q = q0 * w0 + q1 * w1;
        if (clip) q = clip_to_limits(q, q0, q1);
        return q;
    }
    // one variable; a side the stencil does not produce keeps its value
    __device__ void scalar(double qL1, double qL0, double qR0, double qR1, double& qL, double& qR) const
    {
        if (mode == EB_ST_L2R2) { interp_scalar(w, limiter, clip, eps, qL1, qL0, qR0, qR1, qL, qR); return; }
        if (mode == EB_ST_L2R1) {           // :398-419
            const double delLminus = (qL0 - qL1) * w.two_over_L0L1;
            const double del = (qR0 - qL0) * w.two_over_R0L0;
            double sL = 1.0;
            if (limiter) sL = eb_div(delLminus * del + fabs(delLminus * del) + eps, delLminus * delLminus + del * del + eps);
            qL = qL0 + sL * w.aL0 * (del * w.two_L0_plus_L1 + delLminus * w.lenR0);
            if (limiter && (delLminus * del < 0.0)) qR = qR0; else qR = weight(qL0, qR0);
            if (clip) qL = clip_to_limits(qL, qL0, qR0);
        } else if (mode == EB_ST_L1R2) {    // :434-454
            const double del = (qR0 - qL0) * w.two_over_R0L0;
            const double delRplus = (qR1 - qR0) * w.two_over_R1R0;
            double sR = 1.0;
            if (limiter) sR = eb_div(del * delRplus + fabs(del * delRplus) + eps, del * del + delRplus * delRplus + eps);
            qR = qR0 - sR * w.aR0 * (delRplus * w.lenL0 + del * w.two_R0_plus_R1);
            if (limiter && (delRplus * del < 0.0)) qL = qL0; else qL = weight(qL0, qR0);
            if (clip) qR = clip_to_limits(qR, qL0, qR0);
        } else if (mode == EB_ST_L2R0) qL = weight(qL0, qL1);
        else qR = weight(qR0, qR1);
    }
};

// fluxcalc.d:187-385 for a state already in the face frame, gvel = 0.  side 0: gas on the right of the face.
template <int DIM, int GASM, int NSP>
__device__ void wall_flux_local(const EbGas* __restrict__ gas, const Prim<NSP>& fs, int side, double* F)
{
    typedef Layout<DIM, NSP> Lay;
    const double vstar = 0.0;
    const double a = fs.a, v = fs.vx;
    const double gm = gas_gamma<GASM, NSP>(gas, fs);
    const double rho = fs.rho, p = fs.p;
    double tmp;
    if (side == 0) {
        const double Jminus = v - 2.0 * a / (gm - 1.0);
        tmp = (vstar - Jminus) * (gm - 1.0) / (2.0 * sqrt(gm)) * sqrt(rho / pow(p, 1.0 / gm));
    } else {
        const double Jplus = v + 2.0 * a / (gm - 1.0);
        tmp = (Jplus - vstar) * (gm - 1.0) / (2.0 * sqrt(gm)) * sqrt(rho / pow(p, 1.0 / gm));
    }
    const double ptiny = 0.1;            // flowstate_limits.min_pressure, globalconfig.d:85
    double pstar = (tmp > 0.0) ? pow(tmp, 2.0 * gm / (gm - 1.0)) : ptiny;
    if (pstar > 1.1 * p) {
        int count = 0;
        double incr_pstar;
        do {
            double fv[2];
            const double dp = 0.001 * pstar;
            for (int n = 0; n < 2; ++n) {
                const double ps = (n == 0) ? pstar : pstar + dp;
                const double xi = ps / p;
                const double M1sq = 1.0 + (gm + 1.0) / 2.0 / gm * (xi - 1.0);
                const double v1 = sqrt(M1sq) * a;
                const double v2 = v1 * ((gm - 1.0) * M1sq + 2.0) / ((gm + 1.0) * M1sq);
                fv[n] = (side == 0) ? (vstar - v1 + v2 - v) : (vstar + v1 - v2 - v);
            }
            incr_pstar = -fv[0] * dp / (fv[1] - fv[0]);
            pstar += incr_pstar;
            count += 1;
        } while (fabs(incr_pstar) / pstar > 0.01 && count < 10);
    }
    pstar = fmin(pstar, p * 10.0);
#pragma unroll
    for (int q = 0; q < Lay::NCQ; ++q) F[q] = 0.0;
    F[Lay::iXMom] = pstar;
    F[Lay::iEnergy] = pstar * vstar;
}

// thermodynamic closure of a reconstructed state by the pair thermo_interpolator names; one copy per gas model
template <int GASM, int NSP>
__device__ __noinline__ bool thermo_by_interpolator(const EbGas* __restrict__ gas, int ti, Prim<NSP>* Q)
{
    if (ti == EB200_INTERP_PT) return thermo_from_pT<GASM, NSP>(gas, *Q);
    if (ti == EB200_INTERP_RHOP) return thermo_from_rhop<GASM, NSP>(gas, *Q);
    if (ti == EB200_INTERP_RHOT) return thermo_from_rhoT<GASM, NSP>(gas, *Q);
    return thermo_from_rhou<GASM, NSP>(gas, *Q);
}

// (general-metric blocks only: a job with such a wall is set up without the uniform-Cartesian fast path, eb200.h)
template <int DIM, int FLUX, int GASM, int NSP>
__device__ __noinline__ bool face_flux_one_sided(const EbParams& P, const EbGas* __restrict__ gas, const EbBlockDesc& D,
                                                 const EbArena& A, const double* __restrict__ prim,
                                                 long long c, long long st, int d, int nL, int nR, double* F)
{
    typedef Layout<DIM, NSP> Lay;
    const long long total = P.total;
    Frame fr;
    load_frame<DIM>(fr, A.face[d], total, c);
    auto loc = [&](double& x, double& y, double& z) { to_local<DIM>(fr, x, y, z); };
    auto glob = [&](double& x, double& y, double& z) { to_global<DIM>(fr, x, y, z); };
    int mode;
    if (nL == 0 && nR >= 2) mode = EB_ST_L0R2;
    else if (nL == 1 && nR >= 2) mode = EB_ST_L1R2;
    else if (nL >= 2 && nR == 1) mode = EB_ST_L2R1;
    else if (nL >= 2 && nR == 0) mode = EB_ST_L2R0;
    else if (nL >= 2 && nR >= 2) mode = EB_ST_L2R2;
    else return false;                      // "Stencils not suitable for standard interpolation."
    const bool has[4] = { nL >= 2, nL >= 1, nR >= 1, nR >= 2 };
    const long long cc[4] = { c - 2 * st, c - st, c, c + st };
    Prim<NSP> X[4];                         // L1, L0, R0, R1
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        if (has[m]) { load_prim<NSP>(X[m], prim, total, cc[m]); if (DIM == 2) X[m].vz = 0.0; }
    }
    if (!has[1]) X[1] = X[2];               // onedinterp.d:120-121
    if (!has[2]) X[2] = X[1];
    if (!has[0]) X[0] = X[1];               // never used by the stencil: defined values only
    if (!has[3]) X[3] = X[2];
    Prim<NSP> L = X[1], R = X[2];
    bool ok = true;
    const bool doL = (mode != EB_ST_L0R2), doR = (mode != EB_ST_L2R0);
    const bool extrap_copy = (mode == EB_ST_L2R0 || mode == EB_ST_L0R2) && P.extrema_clipping;
    if (P.interpolation_order > 1 && !extrap_copy) {
        const bool local_frame = (P.local_frame != 0);
        if (local_frame) {
#pragma unroll
            for (int m = 0; m < 4; ++m) loc(X[m].vx, X[m].vy, X[m].vz);
        }
        double len[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) len[m] = has[m] ? ldg(A.len[d] + cc[m]) : 0.0;
        if (mode == EB_ST_L2R0) { len[0] = ldg(A.len[0] + cc[0]); len[1] = ldg(A.len[0] + cc[1]); }   // :236 passes iLength
        EbOneSided I;
        I.mode = mode; I.limiter = P.apply_limiter != 0; I.clip = P.extrema_clipping != 0; I.eps = P.eps_va;
        I.w0 = 0.0; I.w1 = 0.0;
        const double lenL1 = len[0], lenL0 = len[1], lenR0 = len[2], lenR1 = len[3];
        if (mode == EB_ST_L2R2) l2r2_prepare(I.w, lenL1, lenL0, lenR0, lenR1);
        else if (mode == EB_ST_L2R1) {
            I.w.lenL0 = lenL0; I.w.lenR0 = lenR0;
            I.w.aL0 = eb_div(0.5 * lenL0, (lenL1 + 2.0 * lenL0 + lenR0));
            I.w.two_over_L0L1 = eb_div(2.0, (lenL0 + lenL1));
            I.w.two_over_R0L0 = eb_div(2.0, (lenR0 + lenL0));
            I.w.two_L0_plus_L1 = (2.0 * lenL0 + lenL1);
            I.w0 = eb_div(lenR0, (lenL0 + lenR0)); I.w1 = eb_div(lenL0, (lenL0 + lenR0));
        } else if (mode == EB_ST_L1R2) {
            I.w.lenL0 = lenL0; I.w.lenR0 = lenR0;
            I.w.aR0 = eb_div(0.5 * lenR0, (lenL0 + 2.0 * lenR0 + lenR1));
            I.w.two_over_R0L0 = eb_div(2.0, (lenR0 + lenL0));
            I.w.two_over_R1R0 = eb_div(2.0, (lenR1 + lenR0));
            I.w.two_R0_plus_R1 = (2.0 * lenR0 + lenR1);
            I.w0 = eb_div(lenR0, (lenL0 + lenR0)); I.w1 = eb_div(lenL0, (lenL0 + lenR0));
        } else if (mode == EB_ST_L2R0) {
            I.w0 = eb_div((2.0 * lenL0 + lenL1), (lenL0 + lenL1)); I.w1 = eb_div(-lenL0, (lenL0 + lenL1));
        } else {
            I.w0 = eb_div((2.0 * lenR0 + lenR1), (lenR0 + lenR1)); I.w1 = eb_div(-lenR0, (lenR0 + lenR1));
        }
        I.scalar(X[0].vx, X[1].vx, X[2].vx, X[3].vx, L.vx, R.vx);
        I.scalar(X[0].vy, X[1].vy, X[2].vy, X[3].vy, L.vy, R.vy);
        I.scalar(X[0].vz, X[1].vz, X[2].vz, X[3].vz, L.vz, R.vz);
        if (NSP > 1) {
            for (int i = 0; i < NSP; ++i) I.scalar(X[0].rho_s[i], X[1].rho_s[i], X[2].rho_s[i], X[3].rho_s[i], L.rho_s[i], R.rho_s[i]);
        }
        const int ti = P.thermo_interp;
        bool okL = true, okR = true;
        if (ti == EB200_INTERP_PT) {
            I.scalar(X[0].p, X[1].p, X[2].p, X[3].p, L.p, R.p);
            I.scalar(X[0].T, X[1].T, X[2].T, X[3].T, L.T, R.T);
            if (doL) okL = thermo_by_interpolator<GASM, NSP>(gas, ti, &L);
            if (doR) okR = thermo_by_interpolator<GASM, NSP>(gas, ti, &R);
        } else {
            if (NSP > 1) {
                double rho_L = 0.0, rho_R = 0.0;
                for (int i = 0; i < NSP; ++i) { rho_L += L.rho_s[i]; rho_R += R.rho_s[i]; }
                if (doL) { L.rho = rho_L; for (int i = 0; i < NSP; ++i) L.massf[i] = L.rho_s[i] / L.rho; ok &= scale_mass_fractions<NSP>(L.massf); }
                if (doR) { R.rho = rho_R; for (int i = 0; i < NSP; ++i) R.massf[i] = R.rho_s[i] / R.rho; ok &= scale_mass_fractions<NSP>(R.massf); }
            } else I.scalar(X[0].rho, X[1].rho, X[2].rho, X[3].rho, L.rho, R.rho);
            if (ti == EB200_INTERP_RHOP) I.scalar(X[0].p, X[1].p, X[2].p, X[3].p, L.p, R.p);
            else if (ti == EB200_INTERP_RHOT) I.scalar(X[0].T, X[1].T, X[2].T, X[3].T, L.T, R.T);
            else I.scalar(X[0].u, X[1].u, X[2].u, X[3].u, L.u, R.u);
            if (doL) okL = thermo_by_interpolator<GASM, NSP>(gas, ti, &L);
            if (doR) okR = thermo_by_interpolator<GASM, NSP>(gas, ti, &R);
        }
        if (!okL) L = X[1];                 // the cell's state, velocity in the frame the cells are in now
        if (!okR) R = X[2];
        if (ti == EB200_INTERP_PT && NSP > 1) {
            if (doL) { for (int i = 0; i < NSP; ++i) L.massf[i] = L.rho_s[i] / L.rho; ok &= scale_mass_fractions<NSP>(L.massf); }
            if (doR) { for (int i = 0; i < NSP; ++i) R.massf[i] = R.rho_s[i] / R.rho; ok &= scale_mass_fractions<NSP>(R.massf); }
        }
        if (local_frame) {                  // back to the global frame (the side that was not produced never left it)
            if (doL) glob(L.vx, L.vy, L.vz);
            if (doR) glob(R.vx, R.vy, R.vz);
        }
    }
    // into the face frame (fluxcalc.d:61-66, 192-193, 294-295)
    loc(L.vx, L.vy, L.vz); loc(R.vx, R.vy, R.vz);
    if (nL == 0) wall_flux_local<DIM, GASM, NSP>(gas, R, 0, F);
    else if (nR == 0) wall_flux_local<DIM, GASM, NSP>(gas, L, 1, F);
    else {
        const double alpha = (FluxPair<FLUX>::adaptive && A.Sf[d]) ? A.Sf[d][c] : 0.0;
        flux_in_face_frame<DIM, NSP, GASM, FLUX>(P, gas, L, R, alpha, F);
    }
    double fx = F[Lay::iXMom], fy = F[Lay::iYMom], fz = (DIM == 3) ? F[Lay::iZMom] : 0.0;
    glob(fx, fy, fz);
    F[Lay::iXMom] = fx; F[Lay::iYMom] = fy;
    if (DIM == 3) F[Lay::iZMom] = fz;
    return ok;
}

// One interface: reconstruction (onedinterp.d:751-988) + flux (fluxcalc.d:54-184).
// c = arena index of the cell on the plus side (right_cells[0]); st = stride along d.
// bcf = boundary face id if the interface lies on a block boundary, else -1.
// Returns false where the reference would throw.
template <int DIM, int FLUX, int GASM, int NSP, bool CART>
__device__ __forceinline__ bool face_flux(const EbParams& P, const EbGas* __restrict__ gas, const EbBlockDesc& D,
                                          const EbArena& A, const double* __restrict__ prim,
                                          long long c, long long st, int d, int bcf, double* F, int idx = 0, int nd = 0)
{
    typedef Layout<DIM, NSP> Lay;
    const long long total = P.total;
    if constexpr (!CART) {
        if (D.noghost_faces) {               // idx = index of the face along d, nd = cells of the block along d
            int nL = 2, nR = 2;
            if ((D.noghost_faces >> (2 * d)) & 1) nL = min(idx, 2);
            if ((D.noghost_faces >> (2 * d + 1)) & 1) nR = min(nd - idx, 2);
            if (nL < 2 || nR < 2) return face_flux_one_sided<DIM, FLUX, GASM, NSP>(P, gas, D, A, prim, c, st, d, nL, nR, F);
        }
    }
    Frame fr;
    if (!CART) load_frame<DIM>(fr, A.face[d], total, c);

    if (bcf >= 0 && D.bc_kind[bcf] == EB200_BC_OUTFLOW_SIMPLE_FLUX) {
        const int hi = bcf & 1;
        Prim<NSP> fs;
        load_prim<NSP>(fs, prim, total, hi ? c - st : c);
        if (DIM == 2) fs.vz = 0.0;
        double nx, ny, nz;
        if (CART) { nx = D.nvec[d][0]; ny = D.nvec[d][1]; nz = D.nvec[d][2]; }
        else { nx = fr.nx; ny = fr.ny; nz = fr.nz; }
        outflow_flux<DIM, NSP>(fs, hi ? 1 : -1, nx, ny, nz, F);
        return true;
    }

    const long long cL1 = c - 2 * st, cL0 = c - st, cR0 = c, cR1 = c + st;
    Prim<NSP> L, R;          // start as copies of cL0 / cR0 (onedinterp.d:120-123)
    bool ok = true;
    // cell-centre values that are never reconstructed
    L.a = ldg(prim + 4 * total + cL0); R.a = ldg(prim + 4 * total + cR0);
    if (GASM != EB200_GAS_IDEAL) { L.T = ldg(prim + 3 * total + cL0); R.T = ldg(prim + 3 * total + cR0); }
    else { L.T = 0.0; R.T = 0.0; }
    double vL0x = ldg(prim + 5 * total + cL0), vL0y = ldg(prim + 6 * total + cL0), vL0z = (DIM == 3) ? ldg(prim + 7 * total + cL0) : 0.0;
    double vR0x = ldg(prim + 5 * total + cR0), vR0y = ldg(prim + 6 * total + cR0), vR0z = (DIM == 3) ? ldg(prim + 7 * total + cR0) : 0.0;
    const double rhoL0 = ldg(prim + cL0), rhoR0 = ldg(prim + cR0);
    const double uL0 = ldg(prim + total + cL0), uR0 = ldg(prim + total + cR0);
    const bool local_frame = (P.local_frame != 0);

    if (P.interpolation_order > 1) {
        double vL1x = ldg(prim + 5 * total + cL1), vL1y = ldg(prim + 6 * total + cL1), vL1z = (DIM == 3) ? ldg(prim + 7 * total + cL1) : 0.0;
        double vR1x = ldg(prim + 5 * total + cR1), vR1y = ldg(prim + 6 * total + cR1), vR1z = (DIM == 3) ? ldg(prim + 7 * total + cR1) : 0.0;
        if (local_frame) {
            if (CART) {
                axis_to_local(D.fr[d], vL1x, vL1y, vL1z); axis_to_local(D.fr[d], vL0x, vL0y, vL0z);
                axis_to_local(D.fr[d], vR0x, vR0y, vR0z); axis_to_local(D.fr[d], vR1x, vR1y, vR1z);
            } else {
                to_local<DIM>(fr, vL1x, vL1y, vL1z); to_local<DIM>(fr, vL0x, vL0y, vL0z);
                to_local<DIM>(fr, vR0x, vR0y, vR0z); to_local<DIM>(fr, vR1x, vR1y, vR1z);
            }
        }
        EbWeights wl;
        if (!CART) {
            const double* ln = A.len[d];
            l2r2_prepare(wl, ldg(ln + cL1), ldg(ln + cL0), ldg(ln + cR0), ldg(ln + cR1));
        }
        const EbWeights& w = CART ? D.w[d] : wl;
        const bool lim = P.apply_limiter != 0, clip = P.extrema_clipping != 0, lmr = P.lmr != 0;
        const double eps = P.eps_va;
        interp_scalar(w, lim, clip, eps, vL1x, vL0x, vR0x, vR1x, L.vx, R.vx, lmr);
        interp_scalar(w, lim, clip, eps, vL1y, vL0y, vR0y, vR1y, L.vy, R.vy, lmr);
        if (DIM == 3) interp_scalar(w, lim, clip, eps, vL1z, vL0z, vR0z, vR1z, L.vz, R.vz, lmr);
        else { L.vz = 0.0; R.vz = 0.0; }
        const int ti = P.thermo_interp;
        bool okL, okR;
        if (NSP > 1) {
#pragma unroll
            for (int i = 0; i < NSP; ++i) {
                const double* ps = prim + (8 + NSP + i) * total;
                interp_scalar(w, lim, clip, eps, ldg(ps + cL1), ldg(ps + cL0), ldg(ps + cR0), ldg(ps + cR1), L.rho_s[i], R.rho_s[i], lmr);
            }
        }
        if (ti == EB200_INTERP_PT) {            // onedinterp.d:820-850: p and T, then the mass fractions
            const double* pp = prim + 2 * total;
            const double* pT = prim + 3 * total;
            interp_scalar(w, lim, clip, eps, ldg(pp + cL1), ldg(pp + cL0), ldg(pp + cR0), ldg(pp + cR1), L.p, R.p, lmr);
            interp_scalar(w, lim, clip, eps, ldg(pT + cL1), ldg(pT + cL0), ldg(pT + cR0), ldg(pT + cR1), L.T, R.T, lmr);
            if (NSP > 1) {                      // Lft/Rght start as copies of the cells: their mass fractions
#pragma unroll
                for (int i = 0; i < NSP; ++i) { L.massf[i] = ldg(prim + (8 + i) * total + cL0); R.massf[i] = ldg(prim + (8 + i) * total + cR0); }
            } else { L.massf[0] = 1.0; R.massf[0] = 1.0; }
            okL = thermo_from_pT<GASM, NSP>(gas, L);
            okR = thermo_from_pT<GASM, NSP>(gas, R);
        } else {
            if (NSP > 1) {
                double rho_L = 0.0, rho_R = 0.0;
#pragma unroll
                for (int i = 0; i < NSP; ++i) { rho_L += L.rho_s[i]; rho_R += R.rho_s[i]; }
                L.rho = rho_L; R.rho = rho_R;
#pragma unroll
                for (int i = 0; i < NSP; ++i) { L.massf[i] = L.rho_s[i] / L.rho; R.massf[i] = R.rho_s[i] / R.rho; }
                ok &= scale_mass_fractions<NSP>(L.massf);
                ok &= scale_mass_fractions<NSP>(R.massf);
            } else {
                interp_scalar(w, lim, clip, eps, ldg(prim + cL1), rhoL0, rhoR0, ldg(prim + cR1), L.rho, R.rho, lmr);
                L.massf[0] = 1.0; R.massf[0] = 1.0;
            }
            if (ti == EB200_INTERP_RHOP) {
                const double* pp = prim + 2 * total;
                interp_scalar(w, lim, clip, eps, ldg(pp + cL1), ldg(pp + cL0), ldg(pp + cR0), ldg(pp + cR1), L.p, R.p, lmr);
                okL = thermo_from_rhop<GASM, NSP>(gas, L);
                okR = thermo_from_rhop<GASM, NSP>(gas, R);
            } else if (ti == EB200_INTERP_RHOT) {
                const double* pT = prim + 3 * total;
                interp_scalar(w, lim, clip, eps, ldg(pT + cL1), ldg(pT + cL0), ldg(pT + cR0), ldg(pT + cR1), L.T, R.T, lmr);
                okL = thermo_from_rhoT<GASM, NSP>(gas, L);
                okR = thermo_from_rhoT<GASM, NSP>(gas, R);
            } else {
                interp_scalar(w, lim, clip, eps, ldg(prim + total + cL1), uL0, uR0, ldg(prim + total + cR1), L.u, R.u, lmr);
                okL = thermo_from_rhou<GASM, NSP>(gas, L);
                okR = thermo_from_rhou<GASM, NSP>(gas, R);
            }
        }
        // on a failed thermo update: fall back to the cell-centre state (onedinterp.d:45-74); lmr has no try/catch
        // there (lmr/onedinterp.d:296-360): the exception ends the step
        if (lmr && !(okL && okR)) ok = false;
        if (!okL) {
            load_prim<NSP>(L, prim, total, cL0);
            L.vx = vL0x; L.vy = vL0y; L.vz = vL0z;        // the cell's velocity, currently in the local frame
        }
        if (!okR) {
            load_prim<NSP>(R, prim, total, cR0);
            R.vx = vR0x; R.vy = vR0y; R.vz = vR0z;
        }
        if (ti == EB200_INTERP_PT && NSP > 1) {
#pragma unroll
            for (int i = 0; i < NSP; ++i) { L.massf[i] = L.rho_s[i] / L.rho; R.massf[i] = R.rho_s[i] / R.rho; }
            ok &= scale_mass_fractions<NSP>(L.massf);
            ok &= scale_mass_fractions<NSP>(R.massf);
        }
        if (!CART && local_frame) {
            // back to the global frame (onedinterp.d:979-987); the flux calculation rotates again
            to_global<DIM>(fr, L.vx, L.vy, L.vz); to_global<DIM>(fr, R.vx, R.vy, R.vz);
        }
        if (CART && !local_frame) { /* velocities are global: rotate below */ }
    } else {
        L.rho = rhoL0; R.rho = rhoR0; L.u = uL0; R.u = uR0;
        L.p = ldg(prim + 2 * total + cL0); R.p = ldg(prim + 2 * total + cR0);
        L.T = ldg(prim + 3 * total + cL0); R.T = ldg(prim + 3 * total + cR0);
        L.vx = vL0x; L.vy = vL0y; L.vz = vL0z; R.vx = vR0x; R.vy = vR0y; R.vz = vR0z;
        if (NSP > 1) {
#pragma unroll
            for (int i = 0; i < NSP; ++i) { L.massf[i] = ldg(prim + (8 + i) * total + cL0); R.massf[i] = ldg(prim + (8 + i) * total + cR0); }
        } else { L.massf[0] = 1.0; R.massf[0] = 1.0; }
    }
    // compute_interface_flux_interior: into the face frame
    const bool already_local = CART && local_frame && (P.interpolation_order > 1);
    if (!already_local) {
        if (CART) { axis_to_local(D.fr[d], L.vx, L.vy, L.vz); axis_to_local(D.fr[d], R.vx, R.vy, R.vz); }
        else { to_local<DIM>(fr, L.vx, L.vy, L.vz); to_local<DIM>(fr, R.vx, R.vy, R.vz); }
    }
    {
        const double alpha = (FluxPair<FLUX>::adaptive && A.Sf[d]) ? A.Sf[d][c] : 0.0;
        flux_in_face_frame<DIM, NSP, GASM, FLUX>(P, gas, L, R, alpha, F);
    }
    // momentum flux back to the global frame (fluxcalc.d:169-175)
    double fx = F[Lay::iXMom], fy = F[Lay::iYMom], fz = (DIM == 3) ? F[Lay::iZMom] : 0.0;
    if (CART) axis_to_global(D.fr[d], fx, fy, fz); else to_global<DIM>(fr, fx, fy, fz);
    F[Lay::iXMom] = fx; F[Lay::iYMom] = fy;
    if (DIM == 3) F[Lay::iZMom] = fz;
    return ok;
}

// Where else the FlowState of interior cell (i, j, k) at arena index c has to be stored: one ghost cell
// per direction at most (blocks that push have at least four cells in every direction).
template <int DIM>
__device__ __forceinline__ void push_targets(const EbBlockDesc& D, int i, int j, int k, long long c,
                                             long long& p0, long long& p1, long long& p2)
{
    p0 = p1 = p2 = -1;
    const int m = D.push_mask;
    if (!m) return;
    if ((m & 1) && i < 2) p0 = c + D.push_off[0]; else if ((m & 2) && i >= D.nic - 2) p0 = c + D.push_off[1];
    if ((m & 4) && j < 2) p1 = c + D.push_off[2]; else if ((m & 8) && j >= D.njc - 2) p1 = c + D.push_off[3];
    if (DIM == 3) { if ((m & 16) && k < 2) p2 = c + D.push_off[4]; else if ((m & 32) && k >= D.nkc - 2) p2 = c + D.push_off[5]; }
}

// Stage update (simcore_gasdynamic_step.d:1250-1357) + decode_conserved (fvcell.d:586-821) +
// check_data (fluidblock.d:607-675) for one cell whose residual dUdt[] is complete.
// U0pre / d0pre: U0 and dUdt_prev[0] of the cell if the caller loaded them ahead of time (else nullptr)
template <int DIM, int GASM, int NSP>
__device__ __forceinline__ void finish_cell(const EbParams& P, const EbGas* __restrict__ gas, const EbStageArgs& S,
                                            long long total, long long c, const double* dUdt, bool& fail, int& n_invalid,
                                            const double* U0pre = nullptr, const double* d0pre = nullptr,
                                            long long push0 = -1, long long push1 = -1, long long push2 = -1)
{
    constexpr int NCQ = Layout<DIM, NSP>::NCQ;
    double U[NCQ], U0[NCQ], d0[NCQ];
#pragma unroll
    for (int q = 0; q < NCQ; ++q) {
        U0[q] = U0pre ? U0pre[q] : ldg(S.U0 + q * total + c);
        d0[q] = (S.stage == 1) ? 0.0 : (d0pre ? d0pre[q] : ldg(S.dUdt_prev[0] + q * total + c));
    }
    if (S.stage == 1) {
#pragma unroll
        for (int q = 0; q < NCQ; ++q) U[q] = U0[q] + S.dt_g[0] * dUdt[q];
    } else if (S.stage == 2) {
#pragma unroll
        for (int q = 0; q < NCQ; ++q)
            U[q] = U0[q] + S.dt_g[4] * (S.dt_g[0] * d0[q] + S.dt_g[1] * dUdt[q]);
    } else if (S.stage == 3) {
#pragma unroll
        for (int q = 0; q < NCQ; ++q)
            U[q] = U0[q] + S.dt_g[4] * (S.dt_g[0] * d0[q] +
                                        S.dt_g[1] * ldg(S.dUdt_prev[1] + q * total + c) + S.dt_g[2] * dUdt[q]);
    } else {            // classic_rk4, simcore_gasdynamic_step.d:1376-1408
#pragma unroll
        for (int q = 0; q < NCQ; ++q)
            U[q] = U0[q] + S.dt_g[4] * (S.dt_g[0] * d0[q] + S.dt_g[1] * ldg(S.dUdt_prev[1] + q * total + c) +
                                        S.dt_g[2] * ldg(S.dUdt_prev[2] + q * total + c) + S.dt_g[3] * dUdt[q]);
    }
    if (S.dUdt_out) {
#pragma unroll
        for (int q = 0; q < NCQ; ++q) S.dUdt_out[q * total + c] = dUdt[q];
    }
    Prim<NSP> Q;
    Q.T = (GASM != EB200_GAS_IDEAL) ? ldg(S.prim_in + 3 * total + c) : 0.0;
    bool modified;
    int rc = decode_cell<DIM, GASM, NSP>(P, gas, U, Q, modified);
    if (rc) fail = true;
    else {
        store_prim<NSP>(Q, S.prim_out, total, c);
        // ghost cells of same-GPU neighbours that mirror this cell (EbBlockDesc::push_off)
        if (push0 >= 0) { store_prim<NSP>(Q, S.prim_out, total, push0); if (S.cellS) S.cellS[push0] = S.cellS[c]; }
        if (push1 >= 0) { store_prim<NSP>(Q, S.prim_out, total, push1); if (S.cellS) S.cellS[push1] = S.cellS[c]; }
        if (push2 >= 0) { store_prim<NSP>(Q, S.prim_out, total, push2); if (S.cellS) S.cellS[push2] = S.cellS[c]; }
        if (!check_data<NSP>(P, Q)) n_invalid++;
    }
    if (S.U_out) {
#pragma unroll
        for (int q = 0; q < NCQ; ++q) S.U_out[q * total + c] = U[q];
    }
}

template <int DIM, int FLUX, int GASM, int NSP, bool CART, int TY>
__global__ void __launch_bounds__(32 * TY, (GASM == EB200_GAS_IDEAL) ? EB_MIN_CTAS : EB_MIN_CTAS_TPG)
flux_update_kernel(const EbParams P, const EbGas* __restrict__ gas, const EbBlockDesc* __restrict__ descs, int nblocks,
                   const EbArena A, const EbStageArgs S)
{
    typedef Layout<DIM, NSP> Lay;
    constexpr int NCQ = Lay::NCQ;
    __shared__ EbBlockDesc D;
    __shared__ double sFN[2][NCQ][TY + 1][32];    // south-face fluxes of rows 0..TY (row TY = extra)
    __shared__ double sFE[2][NCQ][TY];            // flux of the face east of lane 31
    __shared__ int s_blk;

    const int lane = threadIdx.x, wy = threadIdx.y;
    const int tid = wy * 32 + lane;
    const long long cta = S.tile_list ? (long long)S.tile_list[blockIdx.x] : (long long)blockIdx.x;
    if (tid == 0) {
        int lo = 0, hi = nblocks - 1;
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (descs[mid].tile0 <= cta) lo = mid; else hi = mid - 1; }
        s_blk = lo;
    }
    __syncthreads();
    {
        const int* src = reinterpret_cast<const int*>(&descs[s_blk]);
        int* dst = reinterpret_cast<int*>(&D);
        for (int n = tid; n < (int)(sizeof(EbBlockDesc) / sizeof(int)); n += 32 * TY) dst[n] = src[n];
    }
    __syncthreads();
    // the kernel is compiled for one kind of block (CART or not); other kinds are skipped
    if ((D.cartesian != 0) != CART) return;

    const long long t = cta - D.tile0;
    const int ti = (int)(t % D.tiles_i);
    const int tj = (int)((t / D.tiles_i) % D.tiles_j);
    const int tm = (int)(t / ((long long)D.tiles_i * D.tiles_j));
    const int i0 = ti * 32, j0 = tj * TY;
    const int i = i0 + lane, j = j0 + wy;                 // interior (0-based) cell indices
    const int nic = D.nic, njc = D.njc, nkc = D.nkc;
    const long long sj = D.stride[1], sk = D.stride[2];
    const long long total = P.total;
    const int k0 = (DIM == 3) ? tm * D.chunk_m : 0;
    const int k1 = (DIM == 3) ? min(nkc, k0 + D.chunk_m) : 1;

    const bool cell_ok = (i < nic) && (j < njc);
    const bool faceW_ok = (i <= nic) && (j < njc);
    const bool faceS_ok = (i < nic) && (j <= njc);
    // extra pass: warp 0 -> faces east of lane 31 (one per row), warp 1 -> south faces of row j0+TY
    const bool extraE_ok = (wy == 0) && (lane < TY) && (i0 + 32 <= nic) && (j0 + lane < njc);
    const bool extraN_ok = (wy == 1) && (j0 + TY <= njc) && (i < nic);

    double acc[NCQ];          // partial surface integral of the cell of the previous plane
#pragma unroll
    for (int q = 0; q < NCQ; ++q) acc[q] = 0.0;
    bool fail = false;
    int n_invalid = 0;

    const int kend = (DIM == 3) ? k1 : 0;     // last plane index visited (3D: k1 = top face of the chunk)
    for (int k = k0; k <= kend; ++k) {
        const int buf = (k - k0) & 1;
        const bool plane_has_cells = (DIM == 3) ? (k < k1) : true;
        const long long c = D.cell0 + ((long long)(k + D.kg) * D.NJ + (j + EB_NG)) * D.NI + (i + EB_NG);
        double FW[NCQ], FB[NCQ];
#pragma unroll
        for (int q = 0; q < NCQ; ++q) { FW[q] = 0.0; FB[q] = 0.0; }

#pragma unroll 1
        for (int job = 0; job < 4; ++job) {
            bool active; long long cf, st; int d, bcf = -1; int row = 0, col = 0;
            int fidx, fn;               // index of the face along its direction, cells of the block along it
            if (job == 0) {             // west face of my cell
                active = plane_has_cells && faceW_ok; cf = c; st = 1; d = 0; fidx = i; fn = nic;
                if (i == 0) bcf = EB200_WEST; else if (i == nic) bcf = EB200_EAST;
            } else if (job == 1) {      // south face
                active = plane_has_cells && faceS_ok; cf = c; st = sj; d = 1; fidx = j; fn = njc;
                if (j == 0) bcf = EB200_SOUTH; else if (j == njc) bcf = EB200_NORTH;
            } else if (job == 2) {      // bottom face (3D)
                if (DIM != 3) continue;
                active = cell_ok; cf = c; st = sk; d = 2; fidx = k; fn = nkc;
                if (k == 0) bcf = EB200_BOTTOM; else if (k == nkc) bcf = EB200_TOP;
            } else {                    // tile-edge faces
                if (wy > 1) continue;
                if (wy == 0) {
                    active = plane_has_cells && extraE_ok; row = lane; d = 0; st = 1; fidx = i0 + 32; fn = nic;
                    cf = D.cell0 + ((long long)(k + D.kg) * D.NJ + (j0 + lane + EB_NG)) * D.NI + (i0 + 32 + EB_NG);
                    if (i0 + 32 == nic) bcf = EB200_EAST;
                } else {
                    active = plane_has_cells && extraN_ok; col = lane; d = 1; st = sj; fidx = j0 + TY; fn = njc;
                    cf = D.cell0 + ((long long)(k + D.kg) * D.NJ + (j0 + TY + EB_NG)) * D.NI + (i + EB_NG);
                    if (j0 + TY == njc) bcf = EB200_NORTH;
                }
            }
            double F[NCQ];
#pragma unroll
            for (int q = 0; q < NCQ; ++q) F[q] = 0.0;
            if (active) {
                bool ok = face_flux<DIM, FLUX, GASM, NSP, CART>(P, gas, D, A, S.prim_in, cf, st, d, bcf, F, fidx, fn);
                fail |= !ok;
            }
            if (job == 0) {
#pragma unroll
                for (int q = 0; q < NCQ; ++q) FW[q] = F[q];
            } else if (job == 1) {
#pragma unroll
                for (int q = 0; q < NCQ; ++q) sFN[buf][q][wy][lane] = F[q];
            } else if (job == 2) {
#pragma unroll
                for (int q = 0; q < NCQ; ++q) FB[q] = F[q];
            } else if (wy == 0) {
                if (lane < TY) {
#pragma unroll
                    for (int q = 0; q < NCQ; ++q) sFE[buf][q][row] = F[q];
                }
            } else {
#pragma unroll
                for (int q = 0; q < NCQ; ++q) sFN[buf][q][TY][col] = F[q];
            }
        }
        __syncthreads();

        // finish the cell of the previous plane: its top face is this plane's bottom face
        if (DIM == 3 && k > k0 && cell_ok) {
            const long long cp = c - sk;
            const double areaT = CART ? D.area[2] : ldg(A.face[2] + 9 * total + c);
            const double vol_inv = CART ? D.vol_inv : eb_div(1.0, ldg(A.vol + cp));
            double dUdt[NCQ];
#pragma unroll
            for (int q = 0; q < NCQ; ++q) { double si = acc[q] - FB[q] * areaT; dUdt[q] = vol_inv * si + 0.0; }
            long long p0, p1, p2;
            push_targets<DIM>(D, i, j, k - 1, cp, p0, p1, p2);
            finish_cell<DIM, GASM, NSP>(P, gas, S, total, cp, dUdt, fail, n_invalid, nullptr, nullptr, p0, p1, p2);
        }

        if (plane_has_cells) {
            // east flux from the next lane (lane 31: from the extra pass), north flux from the next row
            double FE[NCQ], FN[NCQ], FS[NCQ];
#pragma unroll
            for (int q = 0; q < NCQ; ++q) {
                double nb = __shfl_down_sync(0xffffffffu, FW[q], 1);
                FE[q] = (lane == 31) ? sFE[buf][q][wy] : nb;
                FS[q] = sFN[buf][q][wy][lane];
                FN[q] = sFN[buf][q][wy + 1][lane];
            }
            if (cell_ok) {
                double aW, aE, aS, aN, aB = 0.0;
                if (CART) { aW = aE = D.area[0]; aS = aN = D.area[1]; if (DIM == 3) aB = D.area[2]; }
                else {
                    aW = ldg(A.face[0] + 9 * total + c); aE = ldg(A.face[0] + 9 * total + c + 1);
                    aS = ldg(A.face[1] + 9 * total + c); aN = ldg(A.face[1] + 9 * total + c + sj);
                    if (DIM == 3) aB = ldg(A.face[2] + 9 * total + c);
                }
#pragma unroll
                for (int q = 0; q < NCQ; ++q) {
                    double si = FW[q] * aW;          // 0 - F*(-A)
                    si = si - FE[q] * aE;
                    si = si + FS[q] * aS;
                    si = si - FN[q] * aN;
                    if (DIM == 3) si = si + FB[q] * aB;
                    acc[q] = si;
                }
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
            finish_cell<DIM, GASM, NSP>(P, gas, S, total, c, dUdt, fail, n_invalid, nullptr, nullptr, p0, p1, p2);
        }
    }

    // status: one atomic per warp
    unsigned any_fail = __ballot_sync(0xffffffffu, fail);
    int inv = n_invalid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) inv += __shfl_down_sync(0xffffffffu, inv, o);
    if (lane == 0) {
        if (any_fail) atomicOr(&S.status[0], 1);
        if (inv) atomicAdd(&S.status[S.stage], inv);
    }
}

template <int DIM, int FLUX, int GASM, int NSP, bool CART>
void launch_one(const EbParams& P, const EbGas* gas, const EbBlockDesc* desc, int nblocks, long long ncta,
                const EbArena& A, const EbStageArgs& S, int tile_y, cudaStream_t st)
{
    (void)tile_y;
    flux_update_kernel<DIM, FLUX, GASM, NSP, CART, EB_TILE_Y><<<(unsigned)ncta, dim3(32, EB_TILE_Y), 0, st>>>(P, gas, desc, nblocks, A, S);
}

// Test hook: evaluate face_flux (general-metric path) for a batch of independent faces.
// Face n uses cells 4n..4n+3 of a small arena (L1, L0, R0, R1), stride 1.
template <int DIM, int FLUX, int GASM, int NSP>
__global__ void face_debug_kernel(const EbParams P, const EbGas* __restrict__ gas, const EbArena A, const double* __restrict__ prim,
                                  int nfaces, double* __restrict__ Fout, int* __restrict__ ok_out)
{
    constexpr int NCQ = Layout<DIM, NSP>::NCQ;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nfaces) return;
    __shared__ EbBlockDesc D;
    if (threadIdx.x == 0) { for (int f = 0; f < 6; ++f) D.bc_kind[f] = 0; }
    __syncthreads();
    double F[NCQ];
#pragma unroll
    for (int q = 0; q < NCQ; ++q) F[q] = 0.0;
    bool ok = face_flux<DIM, FLUX, GASM, NSP, false>(P, gas, D, A, prim, 4LL * n + 2, 1, 0, -1, F);
#pragma unroll
    for (int q = 0; q < NCQ; ++q) Fout[(long long)n * NCQ + q] = F[q];
    ok_out[n] = ok ? 1 : 0;
}

template <int FLUX>
void launch_face_debug_impl(const EbParams& P, int gas_model, const EbGas* gas, const EbArena& A, const double* prim,
                            int nfaces, double* Fout, int* ok_out, cudaStream_t st)
{
    const int threads = 128, blocks = (nfaces + threads - 1) / threads;
#define EB_DBG(DIM, GASM, NSP) face_debug_kernel<DIM, FLUX, GASM, NSP><<<blocks, threads, 0, st>>>(P, gas, A, prim, nfaces, Fout, ok_out)
    if (gas_model == EB200_GAS_IDEAL) { if (P.dims == 3) EB_DBG(3, EB200_GAS_IDEAL, 1); else EB_DBG(2, EB200_GAS_IDEAL, 1); }
    else {
#if EB_FLUX_HAS_TPG
#define EB_DBG_NSP(N) if (P.nsp == N) { if (P.dims == 3) EB_DBG(3, EB200_GAS_THERMALLY_PERFECT, N); else EB_DBG(2, EB200_GAS_THERMALLY_PERFECT, N); }
        EB_TPG_NSP_LIST(EB_DBG_NSP)
#undef EB_DBG_NSP
#endif
    }
#undef EB_DBG
}

// which: bit 0 = launch for Cartesian blocks, bit 1 = for general-metric blocks
template <int FLUX>
void launch_flux_update_impl(const EbParams& P, int gas_model, const EbGas* gas, const EbBlockDesc* desc, int nblocks,
                             long long ncta, const EbArena& A, const EbStageArgs& S, int tile_y, int which,
                             cudaStream_t st)
{
#define EB_LAUNCH(DIM, GASM, NSP)                                                                             \
    do {                                                                                                      \
        if (which & 1) launch_one<DIM, FLUX, GASM, NSP, true>(P, gas, desc, nblocks, ncta, A, S, tile_y, st);   \
        if (which & 2) launch_one<DIM, FLUX, GASM, NSP, false>(P, gas, desc, nblocks, ncta, A, S, tile_y, st);  \
    } while (0)
    if (gas_model == EB200_GAS_IDEAL) {
        if (P.dims == 3) EB_LAUNCH(3, EB200_GAS_IDEAL, 1); else EB_LAUNCH(2, EB200_GAS_IDEAL, 1);
    } else {
#if EB_FLUX_HAS_TPG
#ifndef EB_TP_DEV3D
#define EB_LAUNCH_NSP(N) if (P.nsp == N) { if (P.dims == 3) EB_LAUNCH(3, EB200_GAS_THERMALLY_PERFECT, N); else EB_LAUNCH(2, EB200_GAS_THERMALLY_PERFECT, N); }
        EB_TPG_NSP_LIST(EB_LAUNCH_NSP)
#undef EB_LAUNCH_NSP
#endif
#endif
    }
#undef EB_LAUNCH
}

}  // namespace EB_NS
