// flux_kernel_v3.cuh -- fused per-stage kernel for uniform-Cartesian blocks (ideal gas, l2r2 + van Albada):
// the configuration BASELINE.json's metric is quoted on.
//
// What is new against flux_kernel_v2.cuh:
//   * CELL-CENTRED reconstruction.  The van Albada limiter of a cell in a direction is ONE number
//     (onedinterp.d:357-384: sL of the face on the plus side and sR of the face on the minus side are built from
//     the same two slopes), so a thread evaluates it once per cell, direction and variable and produces both the
//     state at its minus face (the face's R side) and at its plus face (the L side of the next face).  The plus
//     state travels to the neighbour: by warp shuffle along i, through shared memory along j, in registers
//     along k.  A face needs 10 limiter evaluations less than in the face-centred form, and one reciprocal serves
//     two reconstructed values.
//   * the three directions are compile-time copies: no address selects, no job loop.
//   * a HELPER WARP (warp TY of the CTA) owns everything that belongs to the tile's halo: the plus states of the
//     cells west / south of the tile, the faces on the tile's east and north edge, and the TMA issue.  The TY
//     main warps therefore run divergence-free, identical work.
//   * planes are staged by TMA two planes ahead into a ring of four tile buffers; the k-stencil of a thread is
//     its own column in three of them (no per-thread ring, no global loads in the loop except U0 / dUdt).
//   * throughput build: p/rho = (gamma-1) u (no reciprocal of rho per face side), extrema clipping applied to the
//     increment (one integer sign test and one FP64 compare per reconstructed value).
// The FMA-free build uses the reference's expressions in the reference's order and is bit-identical to the oracle.
#pragma once
#include "flux_kernel_v2.cuh"

#ifndef EB_V3_VOTEFB
#define EB_V3_VOTEFB 1        // first-order fall-back behind a warp vote (a branch instead of selects)
#endif
#ifndef EB_V3_TRYSPLIT
#define EB_V3_TRYSPLIT 1      // poll the next plane's barrier early, consume the answer after independent loads
#endif
#ifndef EB_V3_CLIPVOTE
#define EB_V3_CLIPVOTE 0      // throughput build: extrema clipping behind a warp vote (measured: slower, 3.9 vs 5.1 G on the noisy box)
#endif
#ifndef EB_V3_NH
#define EB_V3_NH 2            // helper warps: 1 = one warp owns the whole tile halo, 2 = one for the halo along i (and the TMA issue), one along j
#endif
#ifndef EB_V3_SPIN
#define EB_V3_SPIN 1          // waiting on a CTA barrier: 0 = poll, 1 = try_wait with a suspend-time hint, 2 = poll with nanosleep back-off
#endif
#ifndef EB_V3_MIN_CTAS
#define EB_V3_MIN_CTAS ((EB_V3_TY >= 16) ? 1 : 2)     // 17 warps x 112 registers fill an SM; 9-warp CTAs come in pairs
#endif
#ifndef EB_V3_MIN_CTAS_2D
// 2D: a CTA lives for one plane, so its start-up (descriptor, barriers, the TMA load of its only tile) is not hidden by
// marching; two resident CTAs per SM (56 registers, one tile buffer each) hide it for each other.  Measured on the
// 4096 x 1024 forward-facing step: 4.43 G cell-updates/s with one CTA per SM, 6.83 G with two (8-row tiles with three
// or four CTAs: 6.5, 6.4 G)
#define EB_V3_MIN_CTAS_2D 2
#endif

namespace EB_NS {

// One cell, one direction, one variable: states at the minus face (qM, the R side of that face) and at the plus
// face (qP, the L side of that face).  qm, q0, qp: the cell below, the cell, the cell above along the direction.
template <bool CLIP>
__device__ __forceinline__ void recon_cell_scalar(const EbBlockDesc& D, int d, double eps, double qm, double q0, double qp,
                                                  double& qM, double& qP)
{
#ifdef EB_FAST_MATH
    // raw differences, constants of EbBlockDesc::uq (eps = epsilon_van_albada * uq[4])
    const double* __restrict__ K = D.uq[d];
    const double a = q0 - qm, b = qp - q0;
    const double ab = a * b;
    const double n = (ab + fabs(ab)) + eps;
    const double dn = fma(a, a, fma(b, b, eps));
    const double s = n * eb_rcp(dn);
    double iP = s * fma(a, K[1], b * K[0]);
    double iM = s * fma(a, K[3], b * K[2]);
    if (CLIP) {
        // limiters.d:43-51 on the increment: between 0 and the difference to the neighbour.  An increment that
        // points away from the neighbour (sign bits differ: integer test) becomes 0, one that overshoots becomes
        // the difference (only possible where epsilon dominates the limiter: |a| > 5 |b|, slopes ~ sqrt(eps))
        iP = ((__double2hiint(iP) ^ __double2hiint(b)) < 0) ? 0.0 : ((fabs(iP) > fabs(b)) ? b : iP);
        iM = ((__double2hiint(iM) ^ __double2hiint(a)) < 0) ? 0.0 : ((fabs(iM) > fabs(a)) ? a : iM);
    }
    qP = q0 + iP;
    qM = q0 - iM;
#else
    const EbWeights& w = D.w[d];
    const double delm = (q0 - qm) * w.two_over_L0L1;
    const double delp = (qp - q0) * w.two_over_R0L0;
    const double s = (delm * delp + fabs(delm * delp) + eps) / (delm * delm + delp * delp + eps);
    qP = q0 + s * w.aL0 * (delp * w.two_L0_plus_L1 + delm * w.lenR0);
    qM = q0 - s * w.aR0 * (delp * w.lenL0 + delm * w.two_R0_plus_R1);
    if (CLIP) {
        qP = clip_to_limits(qP, q0, qp);
        qM = clip_to_limits(qM, qm, q0);
    }
#endif
}

// variables: 0 rho, 1 u, 2..4 velocity components (x, y, z)
template <int DIM, int TY>
__device__ __forceinline__ void load_cell5(const double* __restrict__ t, int o, double* q)
{
    typedef Tile<DIM, TY> T;
    q[0] = t[T::F_RHO * T::FSZ + o]; q[1] = t[T::F_U * T::FSZ + o];
    q[2] = t[(T::F_V + 0) * T::FSZ + o]; q[3] = t[(T::F_V + 1) * T::FSZ + o];
    q[4] = (DIM == 3) ? t[(T::F_V + 2) * T::FSZ + o] : 0.0;
}

#if defined(EB_FAST_MATH) && EB_V3_CLIPVOTE
// The same reconstruction with the extrema clipping behind a warp vote: in smooth flow no lane of a warp needs it
// (an increment points away from its neighbour only at a local extremum, and overshoots only where epsilon rules
// the limiter), so the selects leave the main path.  Same values as recon_cell_scalar<true>.
__device__ __forceinline__ bool clip_needed(double inc, double d)
{
    return ((__double2hiint(inc) ^ __double2hiint(d)) < 0) || (fabs(inc) > fabs(d));
}
template <int NV>
__device__ __forceinline__ void recon_cell_voted(const EbBlockDesc& D, int d, double eps, const double* qm, const double* q0,
                                                 const double* qp, double* qM, double* qP)
{
    const double* __restrict__ K = D.uq[d];
    double iP[NV], iM[NV];
    bool need = false;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const double a = q0[v] - qm[v], b = qp[v] - q0[v];
        const double ab = a * b;
        const double n = (ab + fabs(ab)) + eps;
        const double dn = fma(a, a, fma(b, b, eps));
        const double s = n * eb_rcp(dn);
        iP[v] = s * fma(a, K[1], b * K[0]);
        iM[v] = s * fma(a, K[3], b * K[2]);
        need = need || clip_needed(iP[v], b) || clip_needed(iM[v], a);
    }
    if (__any_sync(__activemask(), need)) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const double a = q0[v] - qm[v], b = qp[v] - q0[v];
            iP[v] = ((__double2hiint(iP[v]) ^ __double2hiint(b)) < 0) ? 0.0 : ((fabs(iP[v]) > fabs(b)) ? b : iP[v]);
            iM[v] = ((__double2hiint(iM[v]) ^ __double2hiint(a)) < 0) ? 0.0 : ((fabs(iM[v]) > fabs(a)) ? a : iM[v]);
        }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) { qP[v] = q0[v] + iP[v]; qM[v] = q0[v] - iM[v]; }
}
#endif

// positive?  (sign bit clear and not zero; the integer compare keeps the FP64 pipe out of it in the throughput build)
__device__ __forceinline__ bool not_positive(double x)
{
#ifdef EB_FAST_MATH
    return __double2hiint(x) <= 0;
#else
    return x <= 0.0;
#endif
}

// Reconstruction of one cell along direction d (all variables) with the first-order fall-back of a side whose
// reconstructed state is not physical (onedinterp.d:45-74: update_thermo_from_rhou throws -> cell values).
// WANT: bit 0 = minus side, bit 1 = plus side (a side that is not wanted is still computed, the compiler drops it).
template <int DIM, bool CLIP>
__device__ __forceinline__ void recon_cell(const EbBlockDesc& D, int d, double eps, const double* qm, const double* q0,
                                           const double* qp, double* qM, double* qP, bool& fbM, bool& fbP)
{
    constexpr int NV = (DIM == 3) ? 5 : 4;
#if defined(EB_FAST_MATH) && EB_V3_CLIPVOTE
    if (CLIP) recon_cell_voted<NV>(D, d, eps, qm, q0, qp, qM, qP);
    else
#endif
    {
#pragma unroll
        for (int v = 0; v < NV; ++v) recon_cell_scalar<CLIP>(D, d, eps, qm[v], q0[v], qp[v], qM[v], qP[v]);
    }
    if (DIM == 2) { qM[4] = 0.0; qP[4] = 0.0; }
    fbM = not_positive(qM[1]) || not_positive(qM[0]);
    fbP = not_positive(qP[1]) || not_positive(qP[0]);
    // rare: a vote makes the condition warp-uniform, so that this stays a branch around the copies instead of
    // twenty selects on the main path
#if EB_V3_VOTEFB
    if (__any_sync(__activemask(), fbM || fbP))
#endif
    {
        if (fbM) {
#pragma unroll
            for (int v = 0; v < 5; ++v) qM[v] = q0[v];
        }
        if (fbP) {
#pragma unroll
            for (int v = 0; v < 5; ++v) qP[v] = q0[v];
        }
    }
}

// Flux of one face along direction DIR from its two reconstructed sides; F in conserved-quantity order with the
// momentum components in (x, y, z).  fb*: the side fell back to its cell (cL / cR = arena index of that cell).
template <int DIM, int FLUX, int DIR>
__device__ __forceinline__ void face_flux_v3(const EbParams& P, const EbGas* __restrict__ gas, const double* Ls, double aL, bool fbL,
                                             long long cL, const double* Rs, double aR, bool fbR, long long cR, double alpha,
                                             const double* __restrict__ prim, double* F)
{
    typedef Layout<DIM, 1> Lay;
    Prim<1> L, R;
    L.rho = Ls[0]; L.u = Ls[1]; L.vx = Ls[2]; L.vy = Ls[3]; L.vz = (DIM == 3) ? Ls[4] : 0.0; L.a = aL; L.massf[0] = 1.0;
    R.rho = Rs[0]; R.u = Rs[1]; R.vx = Rs[2]; R.vy = Rs[3]; R.vz = (DIM == 3) ? Rs[4] : 0.0; R.a = aR; R.massf[0] = 1.0;
    constexpr int SHOCK = FluxPair<FLUX>::shock, SMOOTH = FluxPair<FLUX>::smooth;
#ifdef EB_FAST_MATH
    (void)fbL; (void)fbR; (void)cL; (void)cR; (void)prim;
    const double gm1 = gas->Rgas * gas->Cvinv;
    const double pLrL = gm1 * L.u, pRrR = gm1 * R.u;
    L.p = L.rho * pLrL; R.p = R.rho * pRrR;
    constexpr bool has_component_form = (SHOCK < EB200_FLUX_ROE) && (SMOOTH < EB200_FLUX_ROE);
    if constexpr (has_component_form) {
        if (FluxPair<FLUX>::adaptive && alpha > 0.0) flux_components_pr<DIM, SHOCK>(L, R, DIR, false, P.M_inf, pLrL, pRrR, F);
        else flux_components_pr<DIM, SMOOTH>(L, R, DIR, P.entropy_fix != 0, P.M_inf, pLrL, pRrR, F);
        return;
    }
    L.T = L.u * gas->Cvinv; R.T = R.u * gas->Cvinv;
#else
    const long long total = P.total;
    if (fbL) { L.p = ldg(prim + 2 * total + cL); L.T = ldg(prim + 3 * total + cL); }
    else { L.T = L.u * gas->Cvinv; L.p = L.rho * gas->Rgas * L.T; }
    if (fbR) { R.p = ldg(prim + 2 * total + cR); R.T = ldg(prim + 3 * total + cR); }
    else { R.T = R.u * gas->Cvinv; R.p = R.rho * gas->Rgas * R.T; }
#endif
    // into the face frame: a renaming of the components (CartFrame conventions of flux_kernel_v2.cuh)
    if (DIM == 3) {
        const double lv[3] = { L.vx, L.vy, L.vz }, rv[3] = { R.vx, R.vy, R.vz };
        L.vx = lv[DIR]; L.vy = lv[(DIR + 1) % 3]; L.vz = lv[(DIR + 2) % 3];
        R.vx = rv[DIR]; R.vy = rv[(DIR + 1) % 3]; R.vz = rv[(DIR + 2) % 3];
    } else {
        const double lx = L.vx, ly = L.vy, rx = R.vx, ry = R.vy;
        L.vx = (DIR == 0) ? lx : ly; L.vy = (DIR == 0) ? -ly : lx;
        R.vx = (DIR == 0) ? rx : ry; R.vy = (DIR == 0) ? -ry : rx;
    }
    double Ff[Lay::NCQ];
    flux_in_face_frame<DIM, 1, EB200_GAS_IDEAL, FLUX>(P, gas, L, R, alpha, Ff);
    F[Lay::iMass] = Ff[Lay::iMass]; F[Lay::iEnergy] = Ff[Lay::iEnergy];
    if (DIM == 3) {
        F[Lay::iXMom + DIR] = Ff[Lay::iXMom];
        F[Lay::iXMom + (DIR + 1) % 3] = Ff[Lay::iYMom];
        F[Lay::iXMom + (DIR + 2) % 3] = Ff[Lay::iZMom];
    } else {
        F[Lay::iXMom] = (DIR == 0) ? Ff[Lay::iXMom] : Ff[Lay::iYMom];
        F[Lay::iYMom] = (DIR == 0) ? -Ff[Lay::iYMom] : Ff[Lay::iXMom];
    }
}

// BFE_SimpleOutflowFlux on a block-boundary face (bc/boundary_flux_effect.d:573-643).  Rare (faces on an outflow
// boundary only): kept out of line, one copy for every call site, so that the hot loop stays small.
template <int DIM>
__device__ __noinline__ void outflow_flux_nl(const double* __restrict__ prim, long long total, long long cell, int outsign,
                                             double nx, double ny, double nz, double* F)
{
    Prim<1> fs;
    load_prim<1>(fs, prim, total, cell);
    if (DIM == 2) fs.vz = 0.0;
    outflow_flux<DIM, 1>(fs, outsign, nx, ny, nz, F);
}

// returns true and fills F when the face with plus-side cell cf (index idx of n along DIR, stride st) lies on a
// boundary that carries this boundary condition
template <int DIM, int DIR>
__device__ __forceinline__ bool outflow_override(const EbParams& P, const EbBlockDesc& D, const double* __restrict__ prim,
                                                 int idx, int n, long long cf, long long st, double* F)
{
    if (!D.outflow_flux_faces) return false;
    int bcf = -1;
    if (idx == 0) bcf = 2 * DIR; else if (idx == n) bcf = 2 * DIR + 1;
    if (bcf < 0 || D.bc_kind[bcf] != EB200_BC_OUTFLOW_SIMPLE_FLUX) return false;
    const int hi = bcf & 1;
    double Ft[Layout<DIM, 1>::NCQ];
    outflow_flux_nl<DIM>(prim, P.total, hi ? cf - st : cf, hi ? 1 : -1, D.nvec[DIR][0], D.nvec[DIR][1], D.nvec[DIR][2], Ft);
#pragma unroll
    for (int q = 0; q < Layout<DIM, 1>::NCQ; ++q) F[q] = Ft[q];
    return true;
}

// a plus state leaves its thread with the fall-back flag folded into the sign of u (u > 0 for every valid state);
// only the FMA-free build needs the flag (it reads p and T of the cell instead of recomputing them)
__device__ __forceinline__ double fold_flag(double u, bool fb)
{
#ifdef EB_FAST_MATH
    (void)fb; return u;
#else
    return fb ? -u : u;
#endif
}
__device__ __forceinline__ bool unfold_flag(double& u)
{
#ifdef EB_FAST_MATH
    (void)u; return false;
#else
    const bool fb = u < 0.0; u = fabs(u); return fb;
#endif
}

// finish_cell of flux_kernel.cuh in two halves, so that only the new U (not U0, the residuals and the fluxes)
// has to survive between them.  First half: stage update (simcore_gasdynamic_step.d:1250-1357).
template <int NCQ>
__device__ __forceinline__ void stage_update_v3(const EbStageArgs& S, long long total, long long c, const double* U0, const double* d0,
                                                const double* dUdt, double* U)
{
    if (S.stage == 1) {
#pragma unroll
        for (int q = 0; q < NCQ; ++q) U[q] = U0[q] + S.dt_g[0] * dUdt[q];
    } else if (S.stage == 2) {
#pragma unroll
        for (int q = 0; q < NCQ; ++q) U[q] = U0[q] + S.dt_g[4] * (S.dt_g[0] * d0[q] + S.dt_g[1] * dUdt[q]);
    } else if (S.stage == 3) {
#pragma unroll
        for (int q = 0; q < NCQ; ++q)
            U[q] = U0[q] + S.dt_g[4] * (S.dt_g[0] * d0[q] + S.dt_g[1] * ldg(S.dUdt_prev[1] + q * total + c) + S.dt_g[2] * dUdt[q]);
    } else {
#pragma unroll
        for (int q = 0; q < NCQ; ++q)
            U[q] = U0[q] + S.dt_g[4] * (S.dt_g[0] * d0[q] + S.dt_g[1] * ldg(S.dUdt_prev[1] + q * total + c) +
                                        S.dt_g[2] * ldg(S.dUdt_prev[2] + q * total + c) + S.dt_g[3] * dUdt[q]);
    }
    if (S.dUdt_out) {
#pragma unroll
        for (int q = 0; q < NCQ; ++q) S.dUdt_out[q * total + c] = dUdt[q];
    }
}

// Second half: decode_conserved (fvcell.d:586-821), check_data (fluidblock.d:607-675), stores, ghost-cell pushes.
template <int DIM>
__device__ __forceinline__ void decode_store_v3(const EbParams& P, const EbGas* __restrict__ gas, const EbStageArgs& S, long long total,
                                                long long c, double* U, long long push0, long long push1, long long push2)
{
    constexpr int NCQ = Layout<DIM, 1>::NCQ;
    Prim<1> Q;
    Q.T = 0.0;
    bool modified;
    const int rc = decode_cell<DIM, EB200_GAS_IDEAL, 1>(P, gas, U, Q, modified);
    if (rc) atomicOr(&S.status[0], 1);          // rare: report at once, no flag to carry through the plane loop
    else {
        store_prim<1>(Q, S.prim_out, total, c);
        if ((push0 & push1 & push2) >= 0 || push0 >= 0 || push1 >= 0 || push2 >= 0) {     // ghost cells of same-GPU neighbours
#pragma unroll 1
            for (int n = 0; n < 3; ++n) {
                const long long pt = (n == 0) ? push0 : ((n == 1) ? push1 : push2);
                if (pt >= 0) { store_prim<1>(Q, S.prim_out, total, pt); if (S.cellS) S.cellS[pt] = S.cellS[c]; }
            }
        }
        if (!check_data<1>(P, Q)) atomicAdd(&S.status[S.stage], 1);
    }
    if (S.U_out) {
#pragma unroll
        for (int q = 0; q < NCQ; ++q) S.U_out[q * total + c] = U[q];
    }
}

template <int DIM, int TY>
struct V3Smem {
    typedef Tile<DIM, TY> T;
    static constexpr int NCQ = Layout<DIM, 1>::NCQ;
    static constexpr int NBUF = (DIM == 3) ? 3 : 1;
    static constexpr int O_TILE = 0;
    static constexpr int O_QPJ = O_TILE + NBUF * T::SIZE;        // [TY+1][5][32]: slot r = plus state (along j) of row r-1
    static constexpr int O_QPIH = O_QPJ + (TY + 1) * 5 * 32;     // [3][5][TY]: plus state (along i) of the cell west of the tile, slot = plane % 3
    static constexpr int O_FS = O_QPIH + 3 * 5 * TY;             // [NCQ][TY+1][32]: south-face fluxes; row TY = north edge
    static constexpr int O_FX = O_FS + NCQ * (TY + 1) * 32;      // [2][NCQ][TY]: east-edge fluxes, by plane parity (written a phase early)
    static constexpr int O_US = O_FX + 2 * NCQ * TY;                 // [2 NCQ][TY*32] (3D): U0 and dUdt0 of the cell to finish, staged by cp.async
    static constexpr int O_DESC = O_US + ((DIM == 3) ? 2 * NCQ * TY * 32 : 0);
    static constexpr size_t BYTES = sizeof(double) * O_DESC + sizeof(EbBlockDesc);
};

// one try of the barrier's phase; the result is consumed later so that independent work hides its latency
__device__ __forceinline__ unsigned mbar_try(unsigned long long* bar, unsigned parity)
{
    unsigned done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done;
}

// wait with a suspend-time hint: the warp sleeps in the barrier unit until the phase completes (or the hint, in ns,
// runs out) instead of spinning on try_wait and taking issue slots from the warps it is waiting for
__device__ __forceinline__ void mbar_wait_sleep(unsigned long long* bar, unsigned parity)
{
    unsigned done;
    do {
#if EB_V3_SPIN == 1
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
#else
        done = mbar_try(bar, parity);
#if EB_V3_SPIN == 2
        if (!done) __nanosleep(40);
#endif
#endif
    } while (!done);
}

// CTA-wide split-phase barrier on an mbarrier with one arrival per warp: arrive() after the warp's own writes,
// wait() before reading what the other warps wrote; independent work goes in between.
__device__ __forceinline__ void cta_arrive(unsigned long long* bar, int lane)
{
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

template <int DIM, int FLUX, bool CLIP, int TY>
__global__ void
#ifdef EB_V3_MAXNREG
__maxnreg__(EB_V3_MAXNREG)
#else
__launch_bounds__(32 * (TY + EB_V3_NH), (DIM == 2) ? EB_V3_MIN_CTAS_2D : EB_V3_MIN_CTAS)
#endif
flux_update_kernel_v3(const EbParams P, const EbGas* __restrict__ gas, const EbBlockDesc* __restrict__ descs, int nblocks,
                      const EbArena A, const EbStageArgs S)
{
    typedef Layout<DIM, 1> Lay;
    typedef Tile<DIM, TY> T;
    typedef V3Smem<DIM, TY> SM;
    constexpr int NCQ = Lay::NCQ;
    constexpr int NBUF = SM::NBUF;
    constexpr int NH = EB_V3_NH;
    constexpr int NT = 32 * (TY + NH);
    extern __shared__ __align__(128) double smem[];
    double* const tile = smem + SM::O_TILE;
    double* const qPj = smem + SM::O_QPJ;
    double* const qPiH = smem + SM::O_QPIH;
    double* const fS = smem + SM::O_FS;
    double* const fX = smem + SM::O_FX;
    double* const uS = smem + SM::O_US;
    EbBlockDesc& D = *reinterpret_cast<EbBlockDesc*>(smem + SM::O_DESC);
    __shared__ int s_blk;
    __shared__ int s_ij[2];                                  // i0, j0 of the tile
    __shared__ __align__(8) unsigned long long s_bar[3];     // tile buffers (TMA transaction barriers)
    __shared__ __align__(8) unsigned long long s_sync[3];    // R: plus states along j published; F: fluxes published; K: k-stencil read

    const int lane = threadIdx.x, wy = threadIdx.y;
    const int tid = wy * 32 + lane;
    const bool helper = (wy >= TY);
    const long long cta = S.tile_list ? (long long)S.tile_list[blockIdx.x] : (long long)blockIdx.x;
    if (tid == 0) {
        int lo = 0, hi = nblocks - 1;
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (descs[mid].tile0 <= cta) lo = mid; else hi = mid - 1; }
        s_blk = lo;
#pragma unroll
        for (int b = 0; b < NBUF; ++b) mbar_init(&s_bar[b], 1);
        mbar_init(&s_sync[0], TY + 1); mbar_init(&s_sync[1], TY + NH); mbar_init(&s_sync[2], TY);
        mbar_fence_init();
    }
    __syncthreads();
    {
        const int* src = reinterpret_cast<const int*>(&descs[s_blk]);
        int* dst = reinterpret_cast<int*>(&D);
        for (int n = tid; n < (int)(sizeof(EbBlockDesc) / sizeof(int)); n += NT) dst[n] = src[n];
    }
    __syncthreads();
    if (!D.v3) return;
    unsigned long long* const barR = &s_sync[0];
    unsigned long long* const barF = &s_sync[1];
    unsigned long long* const barK = &s_sync[2];

    const long long t = cta - D.tile0;
    const int ti = (int)(t % D.tiles_i);
    const int tj = (int)((t / D.tiles_i) % D.tiles_j);
    const int tm = (int)(t / ((long long)D.tiles_i * D.tiles_j));
    const int i0 = ti * 32, j0 = tj * TY;
    if (tid == 0) { s_ij[0] = i0; s_ij[1] = j0; }           // read after the start-up barrier
    const int nic = D.nic, njc = D.njc, nkc = D.nkc;
    const int NI = D.NI, NJ = D.NJ;
    const long long sj = D.stride[1], sk = D.stride[2];
    const long long total = P.total;
    const int k0 = (DIM == 3) ? tm * D.chunk_m : 0;
    const int k1 = (DIM == 3) ? min(nkc, k0 + D.chunk_m) : 1;
    const double eps = P.eps_va;
#ifdef EB_FAST_MATH
    const double eps_i = eps * D.uq[0][4], eps_j = eps * D.uq[1][4], eps_k = eps * D.uq[2][4];
#else
    const double eps_i = eps, eps_j = eps, eps_k = eps;
#endif

    // planes are numbered m = k - (k0 - 2) (3D; the first plane a chunk touches is k0 - 2): plane m lives in
    // tile[m % 3], its barrier completes phase (m / 3) & 1.  2D: one plane, one buffer.
    auto issue_plane = [&](int m) {        // one thread
        const int b = (DIM == 3) ? m % 3 : 0;
        double* dst = tile + b * T::SIZE;
        const void* tmap = reinterpret_cast<const char*>(S.tmaps) + (size_t)s_blk * 128;
        const int kp = (DIM == 3) ? (k0 - 2 + m) + D.kg : 0;
        mbar_expect_tx(&s_bar[b], (unsigned)(T::SIZE * sizeof(double)));
        tma_load_4d(dst, tmap, &s_bar[b], i0, j0, kp, 0);
        tma_load_4d(dst + 2 * T::FSZ, tmap, &s_bar[b], i0, j0, kp, 4);
        tma_load_4d(dst + 4 * T::FSZ, tmap, &s_bar[b], i0, j0, kp, 6);
    };
    auto wait_plane = [&](int m) { mbar_wait(&s_bar[(DIM == 3) ? m % 3 : 0], (unsigned)((m / 3) & 1)); };
    auto tile_of = [&](int m) -> const double* { return tile + ((DIM == 3) ? m % 3 : 0) * T::SIZE; };

    if (wy == TY && lane == 0) {
        issue_plane(0);
        if (DIM == 3) { issue_plane(1); issue_plane(2); }
    }

    // =============================== helper warp ===============================================================
    // per plane k: the plus state (along i) of the cells west of the tile for plane k + 1 (published a plane ahead,
    // so that the main warps need no barrier for their west faces), the east-edge faces of plane k from its own
    // reconstruction of the columns i0 + 31 and i0 + 32, the plus state (along j) of the row south of the tile, the
    // north-edge faces, and the TMA refill of the buffer whose plane has left the k-stencil.
    if (helper) {
        const int hr = lane & (TY - 1);                 // row served in the i-halo jobs
        const int hg = (lane / TY) & 1;                 // job 1: 0 = column i0 + 31 (col 33), 1 = column i0 + 32 (col 34)
        const int oN = (TY + 2) * T::COLS + (lane + 2); // north halo row
        const int oE = (hr + 2) * T::COLS + 34;         // cell east of the tile in row hr
        const int i = i0 + lane;
        const bool eastE = (lane >= TY) && (lane < 2 * TY) && (i0 + 32 <= nic) && (j0 + hr < njc);
        const bool northE = (j0 + TY <= njc) && (i < nic);
        // One reconstruction, four jobs per plane (one copy of the arithmetic keeps the code small):
        //   0: along i, cell west of the tile (lanes < TY), plane `kw`: plus state -> qPiH[kw & 1]
        //   1: along i, columns i0 + 31 and i0 + 32 (lanes < 2 TY): plus state of the first moves to the lanes of
        //      the second, which keep their minus state: the east-edge face
        //   2: along j, row south of the tile: plus state -> qPj[0]
        //   3: along j, row north of the tile: minus state kept for the north-edge face
        double eM[5], eL[5], nM[5];
        bool efb = false, nfb = false;
#pragma unroll
        for (int v = 0; v < 5; ++v) { eM[v] = 0.0; eL[v] = 0.0; nM[v] = 0.0; }
        auto halo_jobs = [&](int job_lo, int job_hi, const double* t0, const double* tw, int kw) {
#pragma unroll 1
            for (int job = job_lo; job <= job_hi; ++job) {
                const int d = job >> 1;
                const double* tp = (job == 0) ? tw : t0;
                const int so = (d == 0) ? 1 : T::COLS;
                int oc;
                bool act;
                if (job == 0) { oc = (hr + 2) * T::COLS + 1; act = lane < TY; }
                else if (job == 1) { oc = (hr + 2) * T::COLS + 33 + hg; act = lane < 2 * TY; }
                else if (job == 2) { oc = T::COLS + (lane + 2); act = true; }
                else { oc = oN; act = true; }
                double qM[5], qP[5];
                bool fbM = false, fbP = false;
#pragma unroll
                for (int v = 0; v < 5; ++v) { qM[v] = 0.0; qP[v] = 0.0; }
                if (act) {
                    double qm[5], q0[5], qp[5];
                    load_cell5<DIM, TY>(tp, oc - so, qm); load_cell5<DIM, TY>(tp, oc, q0); load_cell5<DIM, TY>(tp, oc + so, qp);
                    recon_cell<DIM, CLIP>(D, d, (d == 0) ? eps_i : eps_j, qm, q0, qp, qM, qP, fbM, fbP);
                    qP[1] = fold_flag(qP[1], fbP);
                }
                if (job == 0) {
                    if (act) {
#pragma unroll
                        for (int v = 0; v < 5; ++v) qPiH[(((kw % 3) * 5) + v) * TY + hr] = qP[v];
                    }
                } else if (job == 1) {
#pragma unroll
                    for (int v = 0; v < 5; ++v) { eL[v] = __shfl_up_sync(0xffffffffu, qP[v], TY); eM[v] = qM[v]; }
                    efb = fbM;
                } else if (job == 2) {
#pragma unroll
                    for (int v = 0; v < 5; ++v) qPj[v * 32 + lane] = qP[v];
                } else {
#pragma unroll
                    for (int v = 0; v < 5; ++v) nM[v] = qM[v];
                    nfb = fbM;
                }
            }
        };
        const bool helpI = (NH == 1) || (wy == TY);       // owns the halo along i and the TMA refills
        const bool helpJ = (NH == 1) || (wy == TY + 1);   // owns the halo along j
        // west halo of the first plane, before the start-up barrier
        if (helpI) {
            wait_plane((DIM == 3) ? 2 : 0);
            halo_jobs(0, 0, nullptr, tile_of((DIM == 3) ? 2 : 0), k0);
            if (DIM == 3) { wait_plane(0); wait_plane(1); }
        }
        __syncthreads();                                  // start-up barrier
        if (DIM == 3 && helpI && lane == 0) issue_plane(3);        // plane k0 + 1 -> the buffer plane k0 - 2 has left
        for (int k = k0; k < k1; ++k) {
            const int it = k - k0;
            const int m = (DIM == 3) ? it + 2 : 0;
            const double* t0 = tile_of(m);
            const bool next_has_cells = (DIM == 3) && (k + 1 < k1);
            if (helpI) {
                if (next_has_cells) wait_plane(m + 1);
                // the jobs along i write buffers nobody reads any more (qPiH: three slots, fX: two; this warp has
                // passed K of the last plane), so they run ahead of the main warps' last steps of the plane before
                halo_jobs(next_has_cells ? 0 : 1, 1, t0, tile_of(m + 1), k + 1);
                if (eastE) {
                    const long long cf = D.cell0 + ((long long)(k + D.kg) * NJ + (j0 + hr + EB_NG)) * NI + (i0 + 32 + EB_NG);
                    double F[NCQ];
                    if (!outflow_override<DIM, 0>(P, D, S.prim_in, i0 + 32, nic, cf, 1, F)) {
                        const bool fbL = unfold_flag(eL[1]);
                        const double alpha = FluxPair<FLUX>::adaptive ? A.Sf[0][cf] : 0.0;
                        face_flux_v3<DIM, FLUX, 0>(P, gas, eL, t0[T::F_A * T::FSZ + oE - 1], fbL, cf - 1, eM, t0[T::F_A * T::FSZ + oE], efb, cf,
                                                   alpha, S.prim_in, F);
                    }
#pragma unroll
                    for (int q = 0; q < NCQ; ++q) fX[((it & 1) * NCQ + q) * TY + hr] = F[q];
                }
                if (NH == 2) {
                    if (DIM == 3) {
                        mbar_wait_sleep(barK, (unsigned)(it & 1));      // every main warp has read plane k - 1 (and, having passed F of
                        if (lane == 0) issue_plane(m + 2);              // the plane before, so has the other helper): refill its buffer
                    }
                    cta_arrive(barF, lane);
                }
            }
            if (helpJ) {
                if (NH == 2) wait_plane(m);
                if (it > 0) mbar_wait_sleep(barF, (unsigned)((it - 1) & 1));      // the main warps have read last plane's states along j
                halo_jobs(2, 3, t0, nullptr, 0);
                cta_arrive(barR, lane);
                if (NH == 1 && DIM == 3) {
                    mbar_wait_sleep(barK, (unsigned)(it & 1));          // every main warp has read plane k - 1
                    if (lane == 0) issue_plane(m + 2);            // plane k + 2 -> its buffer (the last one needed is k1 + 1)
                }
                mbar_wait_sleep(barR, (unsigned)(it & 1));
                if (northE) {
                    const long long cf = D.cell0 + ((long long)(k + D.kg) * NJ + (j0 + TY + EB_NG)) * NI + (i + EB_NG);
                    double F[NCQ];
                    if (!outflow_override<DIM, 1>(P, D, S.prim_in, j0 + TY, njc, cf, sj, F)) {
                        double Ls[5];
#pragma unroll
                        for (int v = 0; v < 5; ++v) Ls[v] = qPj[(TY * 5 + v) * 32 + lane];
                        const bool fbL = unfold_flag(Ls[1]);
                        const double alpha = FluxPair<FLUX>::adaptive ? A.Sf[1][cf] : 0.0;
                        face_flux_v3<DIM, FLUX, 1>(P, gas, Ls, t0[T::F_A * T::FSZ + oN - T::COLS], fbL, cf - sj, nM, t0[T::F_A * T::FSZ + oN], nfb, cf,
                                                   alpha, S.prim_in, F);
                    }
#pragma unroll
                    for (int q = 0; q < NCQ; ++q) fS[(q * (TY + 1) + TY) * 32 + lane] = F[q];
                }
                cta_arrive(barF, lane);
            }
        }
        return;
    }

    // =============================== main warps: one thread per cell of the tile ================================
    // (i, j and the predicates below are recomputed from shared memory where they are needed: the asm statements of
    //  the barriers are compiler memory fences, so none of them stays in a register -- or worse, in a spill slot --
    //  across the plane loop)
#define i (s_ij[0] + lane)
#define j (s_ij[1] + wy)
#define cell_ok ((i < D.nic) && (j < D.njc))
#define faceW_ok ((i <= D.nic) && (j < D.njc))
#define faceS_ok ((i < D.nic) && (j <= D.njc))
    const int o = (wy + 2) * T::COLS + (lane + 2);          // own position in a tile field
    double* const myU = uS + wy * 32 + lane;                // own staging slots, field q at stride TY*32
    double acc[NCQ], FB[NCQ];
    double kP[5];                                           // plus state along k of the cell one plane below
    bool kPfb = false;
#pragma unroll
    for (int q = 0; q < NCQ; ++q) { acc[q] = 0.0; FB[q] = 0.0; }
#pragma unroll
    for (int v = 0; v < 5; ++v) kP[v] = 0.0;
    bool fail = false;
    int n_invalid = 0;

    if (DIM == 3) {
        wait_plane(0); wait_plane(1); wait_plane(2);
        double qm[5], q0[5], qp[5], qM[5];
        bool fbM;
        load_cell5<DIM, TY>(tile_of(0), o, qm); load_cell5<DIM, TY>(tile_of(1), o, q0); load_cell5<DIM, TY>(tile_of(2), o, qp);
        recon_cell<DIM, CLIP>(D, 2, eps_k, qm, q0, qp, qM, kP, fbM, kPfb);
    } else {
        wait_plane(0);
    }
    __syncthreads();            // start-up barrier: plane k0 - 2 has been read, the first west halo is published

    const int kend = (DIM == 3) ? k1 : 0;
    long long c = D.cell0 + ((long long)(k0 + D.kg) * NJ + (j + EB_NG)) * NI + (i + EB_NG);      // own cell, plane k
    int it = 0;
    for (int k = k0; k <= kend; ++k, ++it, c += D.stride[2]) {
        const int m = (DIM == 3) ? it + 2 : 0;
        const unsigned par = (unsigned)(it & 1);
        const bool has_cells = (DIM == 3) ? (k < k1) : true;
        const double* t0 = tile_of(m);
        const bool finish_below = (DIM == 3) && (k > k0) && cell_ok;

        // ---- 1. along k: own column in three tile buffers; the bottom face of this plane, which is the top face of
        //         the cell one plane below: that cell's residual is complete, its new U goes to the staging slots
        if (DIM == 3) {
#if EB_V3_TRYSPLIT
            const unsigned landed = mbar_try(&s_bar[(m + 1) % 3], (unsigned)(((m + 1) / 3) & 1));   // plane k + 1
#else
            const unsigned landed = 0;
#endif
            const double* tb = tile_of(m - 1);
            const double* ta = tile_of(m + 1);
            double qm[5], q0[5], qp[5], kM[5], kPn[5];
            bool fbM, fbP;
            load_cell5<DIM, TY>(tb, o, qm); load_cell5<DIM, TY>(t0, o, q0);
            if (!landed) wait_plane(m + 1);
            load_cell5<DIM, TY>(ta, o, qp);
            recon_cell<DIM, CLIP>(D, 2, eps_k, qm, q0, qp, kM, kPn, fbM, fbP);
            const double aB = tb[T::F_A * T::FSZ + o];
            if (has_cells) cta_arrive(barK, lane);        // plane k - 1 has left this warp's k-stencil
            if (cell_ok) {
                if (!outflow_override<DIM, 2>(P, D, S.prim_in, k, D.nkc, c, D.stride[2], FB)) {
                    const double alpha = FluxPair<FLUX>::adaptive ? A.Sf[2][c] : 0.0;
                    face_flux_v3<DIM, FLUX, 2>(P, gas, kP, aB, kPfb, c - D.stride[2], kM, t0[T::F_A * T::FSZ + o], fbM, c, alpha, S.prim_in, FB);
                }
            }
#pragma unroll
            for (int v = 0; v < 5; ++v) kP[v] = kPn[v];
            kPfb = fbP;
            if (finish_below) {
                cp_async_wait_all();                      // U0 (and the first residual) staged during the last iteration
                double U0p[NCQ], d0p[NCQ], dUb[NCQ], Un[NCQ];
#pragma unroll
                for (int q = 0; q < NCQ; ++q) {
                    double si = acc[q] - FB[q] * D.area[2];
                    dUb[q] = D.vol_inv * si + 0.0;
                    U0p[q] = myU[q * TY * 32];
                    d0p[q] = (S.stage != 1) ? myU[(NCQ + q) * TY * 32] : 0.0;
                }
                stage_update_v3<NCQ>(S, total, c - D.stride[2], U0p, d0p, dUb, Un);
#pragma unroll
                for (int q = 0; q < NCQ; ++q) myU[q * TY * 32] = Un[q];
            }
        }
#ifdef EB_FAST_MATH
        // throughput build: the surface integral is summed as the fluxes become available (B, W, E, S, N)
        if (DIM == 3) {
#pragma unroll
            for (int q = 0; q < NCQ; ++q) acc[q] = FB[q] * D.area[2];
        }
#endif

        double FW[NCQ], FS_[NCQ], FE[NCQ];
#pragma unroll
        for (int q = 0; q < NCQ; ++q) { FW[q] = 0.0; FS_[q] = 0.0; FE[q] = 0.0; }
        if (has_cells) {
            // ---- 2. along j: own minus state; the plus state goes to the row above through shared memory
            double jM[5];
            bool jfbM;
            {
                double qm[5], q0[5], qp[5], jP[5];
                bool fbP;
                load_cell5<DIM, TY>(t0, o - T::COLS, qm); load_cell5<DIM, TY>(t0, o, q0); load_cell5<DIM, TY>(t0, o + T::COLS, qp);
                recon_cell<DIM, CLIP>(D, 1, eps_j, qm, q0, qp, jM, jP, jfbM, fbP);
                jP[1] = fold_flag(jP[1], fbP);
#pragma unroll
                for (int v = 0; v < 5; ++v) qPj[((wy + 1) * 5 + v) * 32 + lane] = jP[v];
                cta_arrive(barR, lane);
            }
            // ---- 3. along i: the west neighbour's plus state by shuffle (lane 0: from the helper warp, published a plane ago)
            {
                double qm[5], q0[5], qp[5], iM[5], iP[5], Lw[5];
                bool ifbM, fbP;
                load_cell5<DIM, TY>(t0, o - 1, qm); load_cell5<DIM, TY>(t0, o, q0); load_cell5<DIM, TY>(t0, o + 1, qp);
                recon_cell<DIM, CLIP>(D, 0, eps_i, qm, q0, qp, iM, iP, ifbM, fbP);
                iP[1] = fold_flag(iP[1], fbP);
#pragma unroll
                for (int v = 0; v < 5; ++v) Lw[v] = __shfl_up_sync(0xffffffffu, iP[v], 1);
                if (lane == 0) {
#pragma unroll
                    for (int v = 0; v < 5; ++v) Lw[v] = qPiH[(((DIM == 3 ? k : k0) % 3) * 5 + v) * TY + wy];
                }
                if (faceW_ok) {
                    if (!outflow_override<DIM, 0>(P, D, S.prim_in, i, D.nic, c, 1, FW)) {
                        const bool fbL = unfold_flag(Lw[1]);
                        const double alpha = FluxPair<FLUX>::adaptive ? A.Sf[0][c] : 0.0;
                        face_flux_v3<DIM, FLUX, 0>(P, gas, Lw, t0[T::F_A * T::FSZ + o - 1], fbL, c - 1, iM, t0[T::F_A * T::FSZ + o], ifbM, c,
                                                   alpha, S.prim_in, FW);
                    }
                }
#pragma unroll
                for (int q = 0; q < NCQ; ++q) FE[q] = __shfl_down_sync(0xffffffffu, FW[q], 1);
#ifdef EB_FAST_MATH
#pragma unroll
                for (int q = 0; q < NCQ; ++q) {
                    acc[q] = fma(FW[q], D.area[0], acc[q]);
                    if (lane != 31) acc[q] = fma(-FE[q], D.area[0], acc[q]);
                }
#endif
            }
            // ---- 4. the south face, once the row below has published its plus state
            mbar_wait_sleep(barR, par);
            if (faceS_ok) {
                if (!outflow_override<DIM, 1>(P, D, S.prim_in, j, D.njc, c, D.stride[1], FS_)) {
                    double Ls[5];
#pragma unroll
                    for (int v = 0; v < 5; ++v) Ls[v] = qPj[(wy * 5 + v) * 32 + lane];
                    const bool fbL = unfold_flag(Ls[1]);
                    const double alpha = FluxPair<FLUX>::adaptive ? A.Sf[1][c] : 0.0;
                    face_flux_v3<DIM, FLUX, 1>(P, gas, Ls, t0[T::F_A * T::FSZ + o - T::COLS], fbL, c - D.stride[1], jM, t0[T::F_A * T::FSZ + o], jfbM, c,
                                               alpha, S.prim_in, FS_);
                }
            }
#pragma unroll
            for (int q = 0; q < NCQ; ++q) fS[(q * (TY + 1) + wy) * 32 + lane] = FS_[q];
            cta_arrive(barF, lane);
#ifdef EB_FAST_MATH
#pragma unroll
            for (int q = 0; q < NCQ; ++q) acc[q] = fma(FS_[q], D.area[1], acc[q]);
#endif
        }

        // ---- 5. decode and store the cell one plane below while the fluxes of the other rows arrive; then stage U0
        //         (and the first residual) of this plane's cell for the next iteration
        if (DIM == 3) {
            if (finish_below) {
                double Un[NCQ];
#pragma unroll
                for (int q = 0; q < NCQ; ++q) Un[q] = myU[q * TY * 32];
                const long long cp = c - D.stride[2];
                long long p0, p1, p2;
                push_targets<DIM>(D, i, j, k - 1, cp, p0, p1, p2);
                decode_store_v3<DIM>(P, gas, S, total, cp, Un, p0, p1, p2);
            }
            if (has_cells && cell_ok) {
#pragma unroll
                for (int q = 0; q < NCQ; ++q) cp_async8(myU + q * TY * 32, S.U0 + q * total + c);
                if (S.stage != 1) {
#pragma unroll
                    for (int q = 0; q < NCQ; ++q) cp_async8(myU + (NCQ + q) * TY * 32, S.dUdt_prev[0] + q * total + c);
                }
                cp_async_commit();
            }
        }

        // ---- 6. the rest of the surface integral of this plane's cell (all but the top face)
        if (has_cells) {
            mbar_wait_sleep(barF, par);
            if (cell_ok) {
#pragma unroll
                for (int q = 0; q < NCQ; ++q) {
                    const double fe = (lane == 31) ? fX[((int)par * NCQ + q) * TY + wy] : FE[q];
                    const double fn = fS[(q * (TY + 1) + wy + 1) * 32 + lane];
#ifdef EB_FAST_MATH
                    double si = fma(-fn, D.area[1], acc[q]);
                    if (lane == 31) si = fma(-fe, D.area[0], si);
#else
                    double si = FW[q] * D.area[0];          // 0 - F*(-A), summation order W, E, S, N, B, T (fvcell.d:824-854)
                    si = si - fe * D.area[0];
                    si = si + FS_[q] * D.area[1];
                    si = si - fn * D.area[1];
                    if (DIM == 3) si = si + FB[q] * D.area[2];
#endif
                    acc[q] = si;
                }
                if (DIM == 2) {
                    double dUdt[NCQ];
#pragma unroll
                    for (int q = 0; q < NCQ; ++q) dUdt[q] = D.vol_inv * acc[q] + 0.0;
                    long long p0, p1, p2;
                    push_targets<DIM>(D, i, j, 0, c, p0, p1, p2);
                    finish_cell<DIM, EB200_GAS_IDEAL, 1>(P, gas, S, total, c, dUdt, fail, n_invalid, nullptr, nullptr, p0, p1, p2);
                }
            }
        }
    }

#undef i
#undef j
#undef cell_ok
#undef faceW_ok
#undef faceS_ok
    unsigned any_fail = __ballot_sync(0xffffffffu, fail);
    int inv = n_invalid;
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) inv += __shfl_down_sync(0xffffffffu, inv, o2);
    if (lane == 0) {
        if (any_fail) atomicOr(&S.status[0], 1);
        if (inv) atomicAdd(&S.status[S.stage], inv);
    }
}

template <int DIM, int FLUX, bool CLIP>
void launch_one_v3(const EbParams& P, const EbGas* gas, const EbBlockDesc* desc, int nblocks, long long ncta,
                   const EbArena& A, const EbStageArgs& S, cudaStream_t st)
{
    constexpr int TY = EB_V3_TY;
    const size_t smem = V3Smem<DIM, TY>::BYTES;
    auto kern = flux_update_kernel_v3<DIM, FLUX, CLIP, TY>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured = true;
    }
    kern<<<(unsigned)ncta, dim3(32, TY + EB_V3_NH), smem, st>>>(P, gas, desc, nblocks, A, S);
}

// uniform-Cartesian blocks, ideal gas, interpolation_order = 2, apply_limiter = true, TMA staging available
template <int FLUX>
void launch_flux_update_v3_impl(const EbParams& P, const EbGas* gas, const EbBlockDesc* desc, int nblocks,
                                long long ncta, const EbArena& A, const EbStageArgs& S, cudaStream_t st)
{
    if (P.dims == 3) {
        if (P.extrema_clipping) launch_one_v3<3, FLUX, true>(P, gas, desc, nblocks, ncta, A, S, st);
        else launch_one_v3<3, FLUX, false>(P, gas, desc, nblocks, ncta, A, S, st);
    } else {
        if (P.extrema_clipping) launch_one_v3<2, FLUX, true>(P, gas, desc, nblocks, ncta, A, S, st);
        else launch_one_v3<2, FLUX, false>(P, gas, desc, nblocks, ncta, A, S, st);
    }
}

}  // namespace EB_NS
