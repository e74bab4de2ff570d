// flux_kernel_tp.cuh -- fused per-stage kernel for uniform-Cartesian blocks of a THERMALLY PERFECT gas mixture
// (config 5 of BASELINE.json: 5-species air, frozen chemistry; any species count the library is built for),
// l2r2 + van Albada reconstruction of (rho_s[], u, velocity) with config.thermo_interpolator = "rhou".
//
// Against the generic kernel (flux_kernel.cuh), which reconstructs both sides of every face from four cells and
// solves for the temperature of each side inside the face loop:
//   * CELL-CENTRED: a thread reconstructs its own cell along a direction once (one limiter per variable) and turns
//     the two results into complete FlowStates (rho = sum rho_s, massf, Newton solve for T from u, p = rho R T:
//     onedinterp.d:870-925, therm_perf_gas_mix_eos.d:68-160).  The plus state travels to the neighbour that needs
//     it: warp shuffle along i, shared memory along j, a per-thread shared-memory slot along k.
//   * the six Newton solves of a cell all start from the cell's temperature (the reference's starting guess), so in
//     the throughput build the species energies and heat capacities at that temperature are evaluated ONCE per
//     cell (TpCellTable); the first Newton step of every reconstructed state then costs two dot products with the
//     mass fractions.  Later iterates use the mixture polynomial of the CEA segment, ln T by a short series about
//     the cell value, and the iteration stops as soon as the next correction is below 1e-3 K (the following one
//     would be ~1e-11 K); the reference's exit state (u of the last evaluation when its own 1e-6 K test ends the
//     iteration) is reproduced.  Anything unusual (blend zones of the curves, iterates leaving the segment, no
//     convergence) takes the reference's route (thermo_from_rhou of device_math.cuh).
//   * two helper warps own the tile halo (west / south plus states, east / north edge faces), so the TY main warps
//     run the same work on every lane; no registers are held across planes (k state and partial surface integral
//     live in per-thread shared-memory slots), which keeps the Newton code free of spills.
//   * stencil values come through L1/L2 with coalesced read-only loads (every value is read by seven threads of
//     neighbouring warps within one plane step); shared memory carries only what threads exchange.
// The FMA-free build uses the reference's expressions in the reference's order: bit-identical to the oracle.
#pragma once
#include "flux_kernel_v3.cuh"

#ifndef EB_TP_MIN_CTAS
#define EB_TP_MIN_CTAS 1           // measured on 256^3: one CTA per SM with 168 registers 0.55 G cell-updates/s, two with 96 registers 0.52 G
#endif

namespace EB_NS {

template <int DIM, int NSP>
struct TpCfg {
    static constexpr int NV = NSP + 1 + (DIM == 3 ? 3 : 2);      // rho_s[], u, velocity components
    static constexpr int NS = 8 + NSP;                            // rho, u, p, T, a, vx, vy, vz, massf[]
    static constexpr int NCQ = Layout<DIM, NSP>::NCQ;
};

template <int DIM, int NSP>
__device__ __forceinline__ void tp_load_vars(const double* __restrict__ prim, long long total, long long c, double* q)
{
#pragma unroll
    for (int s = 0; s < NSP; ++s) q[s] = ldg(prim + (8 + NSP + s) * total + c);
    q[NSP] = ldg(prim + total + c);
    q[NSP + 1] = ldg(prim + 5 * total + c); q[NSP + 2] = ldg(prim + 6 * total + c);
    if (DIM == 3) q[NSP + 3] = ldg(prim + 7 * total + c);
}

// a FlowState in shared memory: field f at base[f * stride]
template <int NSP>
__device__ __forceinline__ void tp_put(double* base, int stride, const Prim<NSP>& X)
{
    base[0] = X.rho; base[stride] = X.u; base[2 * stride] = X.p; base[3 * stride] = X.T; base[4 * stride] = X.a;
    base[5 * stride] = X.vx; base[6 * stride] = X.vy; base[7 * stride] = X.vz;
#pragma unroll
    for (int s = 0; s < NSP; ++s) base[(8 + s) * stride] = X.massf[s];
}
template <int NSP>
__device__ __forceinline__ void tp_get(const double* base, int stride, Prim<NSP>& X)
{
    X.rho = base[0]; X.u = base[stride]; X.p = base[2 * stride]; X.T = base[3 * stride]; X.a = base[4 * stride];
    X.vx = base[5 * stride]; X.vy = base[6 * stride]; X.vz = base[7 * stride];
#pragma unroll
    for (int s = 0; s < NSP; ++s) X.massf[s] = base[(8 + s) * stride];
}
template <int NSP>
__device__ __forceinline__ void tp_shfl_up(Prim<NSP>& X)
{
    X.rho = __shfl_up_sync(0xffffffffu, X.rho, 1); X.u = __shfl_up_sync(0xffffffffu, X.u, 1);
    X.p = __shfl_up_sync(0xffffffffu, X.p, 1); X.T = __shfl_up_sync(0xffffffffu, X.T, 1);
    X.a = __shfl_up_sync(0xffffffffu, X.a, 1); X.vx = __shfl_up_sync(0xffffffffu, X.vx, 1);
    X.vy = __shfl_up_sync(0xffffffffu, X.vy, 1); X.vz = __shfl_up_sync(0xffffffffu, X.vz, 1);
#pragma unroll
    for (int s = 0; s < NSP; ++s) X.massf[s] = __shfl_up_sync(0xffffffffu, X.massf[s], 1);
}

// What the Newton code reads of the gas model, copied to shared memory once per CTA (the same words for every
// lane: broadcast loads instead of global ones)
template <int NSP>
struct TpGasS {
    double RA[EB_MAXSEG][8][NSP];
    double Rsp[NSP];
    double lo[EB_MAXSEG], hi[EB_MAXSEG];      // open temperature range in which segment s applies unblended
    int nseg, usable;
};

template <int NSP>
__device__ __forceinline__ void tp_fill_gas(const EbGas* __restrict__ g, TpGasS<NSP>& G, int tid, int nt)
{
    for (int n = tid; n < EB_MAXSEG * 8 * NSP; n += nt) {
        const int seg = n / (8 * NSP), k = (n / NSP) % 8, i = n % NSP;
        G.RA[seg][k][i] = g->RA[seg][k][i];
    }
    if (tid < NSP) G.Rsp[tid] = g->Rsp[tid];
    if (tid == 0) {
        const EbCurve& c = g->curves[0];
        const int nseg = c.nseg;
        G.nseg = nseg; G.usable = g->uniform_curves;
        for (int s = 0; s < EB_MAXSEG; ++s) {
            G.lo[s] = 1.0; G.hi[s] = 0.0;
            if (s < nseg) {
                G.lo[s] = (s == 0) ? c.T_low : c.T_breaks[s] + 0.5 * c.T_blends[s - 1];
                G.hi[s] = (s == nseg - 1) ? c.T_high : c.T_breaks[s + 1] - 0.5 * c.T_blends[s];
            }
        }
    }
}

#ifdef EB_FAST_MATH
// Species energies and heat capacities at the cell temperature (cea_thermo_curves.d:56-181 with the coefficients
// premultiplied by R_i: EbGas::RA).  seg < 0: the cell temperature lies in a blend zone or outside the curves, or
// the species do not share break points: the table is not used.
template <int NSP>
struct TpCellTable {
    double es[NSP], cvs[NSP];
    double T0, lnT0, lo, hi;      // (lo, hi): the open temperature range of segment seg
    int seg;
};

template <int NSP>
__device__ __forceinline__ void tp_table(const TpGasS<NSP>& G, double T0, TpCellTable<NSP>& tb)
{
    tb.T0 = T0; tb.seg = -1; tb.lnT0 = 0.0; tb.lo = 0.0; tb.hi = 0.0;
    if (!G.usable) return;
    int seg = -1;
#pragma unroll
    for (int s = 0; s < EB_MAXSEG; ++s) if (T0 > G.lo[s] && T0 < G.hi[s]) seg = s;
    if (seg < 0) return;
    tb.seg = seg; tb.lo = G.lo[seg]; tb.hi = G.hi[seg];
    const double rT = eb_rcp(T0), lnT = log(T0);
    tb.lnT0 = lnT;
#pragma unroll
    for (int i = 0; i < NSP; ++i) {
        const double a0 = G.RA[seg][0][i], a1 = G.RA[seg][1][i], a2 = G.RA[seg][2][i] - G.Rsp[i], a3 = G.RA[seg][3][i];
        const double a4 = G.RA[seg][4][i], a5 = G.RA[seg][5][i], a6 = G.RA[seg][6][i], a7 = G.RA[seg][7][i];
        tb.es[i] = -a0 * rT + a1 * lnT + a7 + T0 * (a2 + T0 * (0.5 * a3 + T0 * ((1.0 / 3.0) * a4 + T0 * (0.25 * a5 + T0 * (0.2 * a6)))));
        tb.cvs[i] = rT * (a0 * rT + a1) + a2 + T0 * (a3 + T0 * (a4 + T0 * (a5 + T0 * a6)));
    }
}

// ln(1 + r) for small r: 2 atanh(r / (2 + r))
__device__ __forceinline__ double tp_log1p_small(double r)
{
    const double s = r * eb_rcp(2.0 + r), s2 = s * s;
    return 2.0 * s * (1.0 + s2 * ((1.0 / 3.0) + s2 * (0.2 + s2 * ((1.0 / 7.0) + s2 * ((1.0 / 9.0) + s2 * (1.0 / 11.0))))));
}

// update_thermo_from_rhou by Newton from the cell temperature.  Returns false when the state has to take the
// reference's route (nothing written then).
template <int NSP>
__device__ __forceinline__ bool tp_thermo_fast(const TpGasS<NSP>& g, const TpCellTable<NSP>& tb, Prim<NSP>& Q)
{
    if (tb.seg < 0) return false;
    const double e = Q.u;
    double Rmix = 0.0, u0 = 0.0, cv0 = 0.0;
#pragma unroll
    for (int i = 0; i < NSP; ++i) { Rmix = fma(Q.massf[i], g.Rsp[i], Rmix); u0 = fma(Q.massf[i], tb.es[i], u0); cv0 = fma(Q.massf[i], tb.cvs[i], cv0); }
    double dx = (e - u0) * eb_rcp(cv0);
    double x = tb.T0 + dx;
    double Qu = u0;
    double adx = fabs(dx);
    if (!(adx < 1.0e-3)) {
        if (!(adx < 400.0) || !(x > tb.lo && x < tb.hi)) return false;
        double A[8];
        const int seg = tb.seg;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double a = 0.0;
#pragma unroll
            for (int i = 0; i < NSP; ++i) a = fma(Q.massf[i], g.RA[seg][k][i], a);
            A[k] = a;
        }
        const double a2 = A[2] - Rmix;
        const double r = dx * eb_rcp(tb.T0);
        double lnx = (fabs(r) < 0.06) ? tb.lnT0 + tp_log1p_small(r) : log(x);
        bool done = false;
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
            const double rT = eb_rcp(x);
            const double u = -A[0] * rT + A[1] * lnx + A[7] + x * (a2 + x * (0.5 * A[3] + x * ((1.0 / 3.0) * A[4] + x * (0.25 * A[5] + x * (0.2 * A[6])))));
            const double cv = rT * (A[0] * rT + A[1]) + a2 + x * (A[3] + x * (A[4] + x * (A[5] + x * A[6])));
            dx = (e - u) * eb_rcp(cv);
            Qu = u;
            x += dx;
            adx = fabs(dx);
            if (adx < 1.0e-3) { done = true; break; }
            if (!(x > tb.lo && x < tb.hi) || !(fabs(x - tb.T0) < 450.0)) return false;
            lnx = log(x);
        }
        if (!done) return false;
    }
    if (!(x > 0.0)) return false;
    Q.T = x;
    Q.u = (adx < 1.0e-6) ? Qu : e;       // newton.d:104-124: the reference leaves with u of its last evaluation
    Q.p = Q.rho * Rmix * x;
    return true;
}

template <int NSP>
__device__ __forceinline__ bool tp_scale_mass_fractions(double* massf)
{
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < NSP; ++i) { massf[i] = massf[i] >= 0.0 ? massf[i] : 0.0; sum += massf[i]; }
    if (fabs(sum - 1.0) > 0.1) return false;
    if (fabs(sum - 1.0) > 0.0) {
        const double r = eb_rcp(sum);
#pragma unroll
        for (int i = 0; i < NSP; ++i) massf[i] *= r;
    }
    return true;
}
#else
template <int NSP>
struct TpCellTable { double T0; };
template <int NSP>
__device__ __forceinline__ void tp_table(const TpGasS<NSP>&, double T0, TpCellTable<NSP>& tb) { tb.T0 = T0; }
#endif

// the reference's route (therm_perf_gas_mix_eos.d:68-160), one copy for every call site
template <int NSP>
__device__ __noinline__ bool tp_thermo_reference(const EbGas* __restrict__ g, Prim<NSP>* Q)
{
    return thermo_from_rhou<EB200_GAS_THERMALLY_PERFECT, NSP>(g, *Q);
}

// thermo of a state whose rho, massf, u are set and whose T holds the starting guess
template <int NSP>
__device__ __forceinline__ bool tp_thermo(const TpGasS<NSP>& G, const EbGas* __restrict__ g, const TpCellTable<NSP>& tb, Prim<NSP>& Q,
                                          Prim<NSP>& rare)
{
#ifdef EB_FAST_MATH
    if (tp_thermo_fast<NSP>(G, tb, Q)) return true;
    rare = Q;                     // rare: the one scratch state of the kernel (local memory) keeps Q itself in registers
    const bool ok = tp_thermo_reference<NSP>(g, &rare);
    Q = rare;
    return ok;
#else
    (void)tb; (void)G; (void)rare;
    return thermo_from_rhou<EB200_GAS_THERMALLY_PERFECT, NSP>(g, Q);
#endif
}

// One reconstructed side of cell c: q[] = (rho_s[], u, velocity) -> FlowState (onedinterp.d:870-925; on a failed
// thermo update the cell's own state, onedinterp.d:45-74).  fail: the reference throws (mass fractions off by > 0.1).
template <int DIM, int NSP>
__device__ __forceinline__ void tp_make_state(const TpGasS<NSP>& G, const EbGas* __restrict__ gas, const double* __restrict__ prim,
                                              long long total, long long c, const double* q, double a0, const TpCellTable<NSP>& tb,
                                              Prim<NSP>& X, bool& fail, Prim<NSP>& rare)
{
    double rho = 0.0;
#pragma unroll
    for (int s = 0; s < NSP; ++s) rho += q[s];
    X.rho = rho;
#ifdef EB_FAST_MATH
    const double rr = eb_rcp(rho);
#pragma unroll
    for (int s = 0; s < NSP; ++s) X.massf[s] = q[s] * rr;
    if (!tp_scale_mass_fractions<NSP>(X.massf)) fail = true;
#else
#pragma unroll
    for (int s = 0; s < NSP; ++s) X.massf[s] = q[s] / rho;
    if (!scale_mass_fractions<NSP>(X.massf)) fail = true;
#endif
    X.u = q[NSP]; X.T = tb.T0; X.p = 0.0;
    X.vx = q[NSP + 1]; X.vy = q[NSP + 2]; X.vz = (DIM == 3) ? q[NSP + 3] : 0.0;
    if (!tp_thermo<NSP>(G, gas, tb, X, rare)) {
        load_prim<NSP>(X, prim, total, c);
        if (DIM == 2) X.vz = 0.0;
    }
    X.a = a0;
}

// both sides of cell c along direction d (stride st): the state at its plus face goes straight to where its reader
// will look for it (dstP, field stride strideP: shared memory), the state at its minus face comes back in M -- one
// state in registers at a time
template <int DIM, int NSP>
__device__ __forceinline__ void tp_cell_states(const EbBlockDesc& D, const TpGasS<NSP>& G, const EbGas* __restrict__ gas, int d, double eps,
                                               bool clip, const double* __restrict__ prim, long long total, long long c, long long st,
                                               bool wantM, bool wantP, double* dstP, int strideP, Prim<NSP>& M, bool& fail, Prim<NSP>& rare,
                                               const TpCellTable<NSP>* cell_table = nullptr)
{
    constexpr int NV = TpCfg<DIM, NSP>::NV;
    double qM[NV], qP[NV];
    {
        double qm[NV], q0[NV], qp[NV];
        tp_load_vars<DIM, NSP>(prim, total, c - st, qm);
        tp_load_vars<DIM, NSP>(prim, total, c, q0);
        tp_load_vars<DIM, NSP>(prim, total, c + st, qp);
        if (clip) {
#pragma unroll
            for (int v = 0; v < NV; ++v) recon_cell_scalar<true>(D, d, eps, qm[v], q0[v], qp[v], qM[v], qP[v]);
        } else {
#pragma unroll
            for (int v = 0; v < NV; ++v) recon_cell_scalar<false>(D, d, eps, qm[v], q0[v], qp[v], qM[v], qP[v]);
        }
    }
    const double a0 = ldg(prim + 4 * total + c);
    TpCellTable<NSP> tb;
    if (cell_table) tb = *cell_table;           // the caller made the cell's table once for all directions
    else tp_table<NSP>(G, ldg(prim + 3 * total + c), tb);
    if (wantP) {
        Prim<NSP> X;
        tp_make_state<DIM, NSP>(G, gas, prim, total, c, qP, a0, tb, X, fail, rare);
        tp_put<NSP>(dstP, strideP, X);
    }
    if (wantM) tp_make_state<DIM, NSP>(G, gas, prim, total, c, qM, a0, tb, M, fail, rare);
}

// flux of a face along direction dir from complete states; F in conserved-quantity order, momentum in (x, y, z)
template <int DIM, int NSP, int FLUX>
__device__ __forceinline__ void tp_face_flux(const EbParams& P, const EbGas* __restrict__ gas, int dir, Prim<NSP> L, Prim<NSP> R, double alpha, double* F)
{
    typedef Layout<DIM, NSP> Lay;
    if (DIM == 3) {
        const double lx = L.vx, ly = L.vy, lz = L.vz, rx = R.vx, ry = R.vy, rz = R.vz;
        L.vx = (dir == 0) ? lx : ((dir == 1) ? ly : lz); L.vy = (dir == 0) ? ly : ((dir == 1) ? lz : lx); L.vz = (dir == 0) ? lz : ((dir == 1) ? lx : ly);
        R.vx = (dir == 0) ? rx : ((dir == 1) ? ry : rz); R.vy = (dir == 0) ? ry : ((dir == 1) ? rz : rx); R.vz = (dir == 0) ? rz : ((dir == 1) ? rx : ry);
    } else {
        const double lx = L.vx, ly = L.vy, rx = R.vx, ry = R.vy;
        L.vx = (dir == 0) ? lx : ly; L.vy = (dir == 0) ? -ly : lx;
        R.vx = (dir == 0) ? rx : ry; R.vy = (dir == 0) ? -ry : rx;
    }
    double Ff[Lay::NCQ];
    flux_in_face_frame<DIM, NSP, EB200_GAS_THERMALLY_PERFECT, FLUX>(P, gas, L, R, alpha, Ff);
    F[Lay::iMass] = Ff[Lay::iMass]; F[Lay::iEnergy] = Ff[Lay::iEnergy];
    if (DIM == 3) {
        const double fn = Ff[Lay::iXMom], f1 = Ff[Lay::iYMom], f2 = Ff[Lay::iZMom];
        F[Lay::iXMom] = (dir == 0) ? fn : ((dir == 1) ? f2 : f1);
        F[Lay::iYMom] = (dir == 0) ? f1 : ((dir == 1) ? fn : f2);
        F[Lay::iZMom] = (dir == 0) ? f2 : ((dir == 1) ? f1 : fn);
    } else {
        F[Lay::iXMom] = (dir == 0) ? Ff[Lay::iXMom] : Ff[Lay::iYMom];
        F[Lay::iYMom] = (dir == 0) ? -Ff[Lay::iYMom] : Ff[Lay::iXMom];
    }
#pragma unroll
    for (int s = 0; s < NSP; ++s) F[Lay::iSpecies + s] = Ff[Lay::iSpecies + s];
}

// BFE_SimpleOutflowFlux on a block-boundary face (bc/boundary_flux_effect.d:573-643): true and F filled when the
// face with plus-side cell cf (index idx of n along direction dir, stride st) carries it
template <int DIM, int NSP>
__device__ __forceinline__ bool tp_outflow_override(const EbParams& P, const EbBlockDesc& D, const double* __restrict__ prim,
                                                    int dir, int idx, int n, long long cf, long long st, double* F)
{
    if (!D.outflow_flux_faces) return false;
    int bcf = -1;
    if (idx == 0) bcf = 2 * dir; else if (idx == n) bcf = 2 * dir + 1;
    if (bcf < 0 || D.bc_kind[bcf] != EB200_BC_OUTFLOW_SIMPLE_FLUX) return false;
    const int hi = bcf & 1;
    Prim<NSP> fs;
    load_prim<NSP>(fs, prim, P.total, hi ? cf - st : cf);
    if (DIM == 2) fs.vz = 0.0;
    outflow_flux<DIM, NSP>(fs, hi ? 1 : -1, D.nvec[dir][0], D.nvec[dir][1], D.nvec[dir][2], F);
    return true;
}

// decode_conserved (fvcell.d:586-821) as decode_cell of device_math.cuh, with the Newton solve of this file
template <int DIM, int NSP>
__device__ __forceinline__ int tp_decode_cell(const EbParams& P, const TpGasS<NSP>& G, const EbGas* __restrict__ g, double* U, Prim<NSP>& Q, bool& U_modified,
                                              Prim<NSP>& rare)
{
#ifndef EB_FAST_MATH
    (void)G; (void)rare;
    return decode_cell<DIM, EB200_GAS_THERMALLY_PERFECT, NSP>(P, g, U, Q, U_modified);
#else
    typedef Layout<DIM, NSP> Lay;
    U_modified = false;
    const double rho = U[Lay::iMass];
    if (!(rho > 0.0)) return 1;
    Q.rho = rho;
    const double dinv = eb_rcp(rho);
    Q.vx = U[Lay::iXMom] * dinv; Q.vy = U[Lay::iYMom] * dinv;
    Q.vz = (DIM == 3) ? U[Lay::iZMom] * dinv : 0.0;
    const double ke = 0.5 * (Q.vx * Q.vx + Q.vy * Q.vy + Q.vz * Q.vz);
    Q.u = U[Lay::iEnergy] * dinv - ke;
    double rhos_sum = 0.0;
#pragma unroll
    for (int i = 0; i < NSP; ++i) {
        if (U[Lay::iSpecies + i] < 0.0) { U[Lay::iSpecies + i] = 0.0; U_modified = true; }
        rhos_sum += U[Lay::iSpecies + i];
    }
    if (fabs(rhos_sum - rho) > 0.1) return 1;
    if (fabs(rhos_sum - rho) > 0.0) {
        const double scale = rho * eb_rcp(rhos_sum);
#pragma unroll
        for (int i = 0; i < NSP; ++i) U[Lay::iSpecies + i] *= scale;
        U_modified = true;
    }
#pragma unroll
    for (int i = 0; i < NSP; ++i) { Q.massf[i] = U[Lay::iSpecies + i] * dinv; Q.rho_s[i] = U[Lay::iSpecies + i]; }
    TpCellTable<NSP> tb;
    tp_table<NSP>(G, Q.T, tb);
    bool fastpath = tp_thermo_fast<NSP>(G, tb, Q);
    if (fastpath) {
        // sound speed (therm_perf_gas.d:394-430): Cv of the mixture at the new temperature
        const double x = Q.T;
        if (x > tb.lo && x < tb.hi) {
            const int seg = tb.seg;
            double Rmix = 0.0, A0 = 0.0, A1 = 0.0, A2 = 0.0, A3 = 0.0, A4 = 0.0, A5 = 0.0, A6 = 0.0;
#pragma unroll
            for (int i = 0; i < NSP; ++i) {
                const double m = Q.massf[i];
                Rmix = fma(m, G.Rsp[i], Rmix);
                A0 = fma(m, G.RA[seg][0][i], A0); A1 = fma(m, G.RA[seg][1][i], A1); A2 = fma(m, G.RA[seg][2][i], A2);
                A3 = fma(m, G.RA[seg][3][i], A3); A4 = fma(m, G.RA[seg][4][i], A4); A5 = fma(m, G.RA[seg][5][i], A5);
                A6 = fma(m, G.RA[seg][6][i], A6);
            }
            const double rT = eb_rcp(x);
            const double Cv = rT * (A0 * rT + A1) + (A2 - Rmix) + x * (A3 + x * (A4 + x * (A5 + x * A6)));
            const double Cp = Cv + Rmix;
            Q.a = eb_sqrt(Cp * eb_rcp(Cv) * (Rmix * x));
            return 0;
        }
        return sound_speed<EB200_GAS_THERMALLY_PERFECT, NSP>(g, Q) ? 0 : 1;
    }
    rare = Q;
    const bool ok_ref = tp_thermo_reference<NSP>(g, &rare);
    Q = rare;
    if (!ok_ref) {
        if (P.ignore_low_T && (rho > 0.0)) {
            Q.T = P.low_T;
            if (!thermo_from_rhoT<EB200_GAS_THERMALLY_PERFECT, NSP>(g, Q)) return 1;
            U[Lay::iMass] = Q.rho;
            U[Lay::iXMom] = Q.rho * Q.vx; U[Lay::iYMom] = Q.rho * Q.vy;
            if (DIM == 3) U[Lay::iZMom] = Q.rho * Q.vz;
            const double ke2 = 0.5 * (Q.vx * Q.vx + Q.vy * Q.vy + Q.vz * Q.vz);
            U[Lay::iEnergy] = Q.rho * (Q.u + ke2);
#pragma unroll
            for (int i = 0; i < NSP; ++i) U[Lay::iSpecies + i] = Q.rho * Q.massf[i];
            U_modified = true;
        } else return 1;
    }
    if (Q.T <= 0.0) return 1;
    if (!sound_speed<EB200_GAS_THERMALLY_PERFECT, NSP>(g, Q)) return 1;
    return 0;
#endif
}

// stage update + decode + stores for a cell whose residual is complete (finish_cell of flux_kernel.cuh)
template <int DIM, int NSP>
__device__ __forceinline__ void tp_finish_cell(const EbParams& P, const TpGasS<NSP>& G, const EbGas* __restrict__ gas, const EbStageArgs& S, long long total,
                                               long long c, const double* dUdt, bool& fail, int& n_invalid,
                                               long long push0, long long push1, long long push2, Prim<NSP>& rare)
{
    constexpr int NCQ = Layout<DIM, NSP>::NCQ;
    double U[NCQ];
    if (S.stage == 1) {
#pragma unroll
        for (int q = 0; q < NCQ; ++q) U[q] = ldg(S.U0 + q * total + c) + S.dt_g[0] * dUdt[q];
    } else if (S.stage == 2) {
#pragma unroll
        for (int q = 0; q < NCQ; ++q)
            U[q] = ldg(S.U0 + q * total + c) + S.dt_g[4] * (S.dt_g[0] * ldg(S.dUdt_prev[0] + q * total + c) + S.dt_g[1] * dUdt[q]);
    } else if (S.stage == 3) {
#pragma unroll
        for (int q = 0; q < NCQ; ++q)
            U[q] = ldg(S.U0 + q * total + c) + S.dt_g[4] * (S.dt_g[0] * ldg(S.dUdt_prev[0] + q * total + c) +
                                                            S.dt_g[1] * ldg(S.dUdt_prev[1] + q * total + c) + S.dt_g[2] * dUdt[q]);
    } else {
#pragma unroll
        for (int q = 0; q < NCQ; ++q)
            U[q] = ldg(S.U0 + q * total + c) + S.dt_g[4] * (S.dt_g[0] * ldg(S.dUdt_prev[0] + q * total + c) +
                                                            S.dt_g[1] * ldg(S.dUdt_prev[1] + q * total + c) +
                                                            S.dt_g[2] * ldg(S.dUdt_prev[2] + q * total + c) + S.dt_g[3] * dUdt[q]);
    }
    if (S.dUdt_out) {
#pragma unroll
        for (int q = 0; q < NCQ; ++q) S.dUdt_out[q * total + c] = dUdt[q];
    }
    Prim<NSP> Q;
    Q.T = ldg(S.prim_in + 3 * total + c);
    bool modified;
    const int rc = tp_decode_cell<DIM, NSP>(P, G, gas, U, Q, modified, rare);
    if (rc) fail = true;
    else {
        store_prim<NSP>(Q, S.prim_out, total, c);
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

template <int DIM, int NSP, int TY>
struct TpSmem {
    typedef TpCfg<DIM, NSP> C;
    static constexpr int NS = C::NS, NCQ = C::NCQ;
    static constexpr int NTM = TY * 32;                               // main threads
#ifdef EB_FAST_MATH
    static constexpr int NACC = NCQ;
#else
    static constexpr int NACC = 2 * NCQ;                              // the bottom flux keeps its own slot: summation order W, E, S, N, B
#endif
    static constexpr int O_PJ = 0;                                    // [TY + 1][NS][32]: slot r = plus state (along j) of row r - 1
    static constexpr int O_FS = O_PJ + (TY + 1) * NS * 32;            // [TY + 1][NCQ][32]: south-face fluxes, row TY = north edge
    static constexpr int O_PIW = O_FS + (TY + 1) * NCQ * 32;          // [2][NS][TY]: plus state (along i) of the cell west of the tile, by plane parity
    static constexpr int O_PI = O_PIW + 2 * NS * TY;                  // [TY][NS][32]: plus state (along i) of every cell of the tile
    static constexpr int O_MJ = O_PI + TY * NS * 32;                  // [TY][NS][32]: minus state (along j), parked across the barrier
    static constexpr int O_FE = O_MJ + TY * NS * 32;                  // [NCQ][TY]: east-edge fluxes
    static constexpr int O_K = O_FE + NCQ * TY;                       // [2][NS][NTM]: plus state along k of the cell one plane below, by plane parity
    static constexpr int O_ACC = O_K + ((DIM == 3) ? 2 * NS * NTM : 0);   // [NACC][NTM]: partial surface integral
    static constexpr int O_DESC = O_ACC + NACC * NTM;
    static constexpr size_t BYTES = sizeof(double) * O_DESC + sizeof(EbBlockDesc);
};

template <int DIM, int FLUX, int NSP, int TY>
__global__ void
#ifdef EB_TP_MAXNREG
__maxnreg__(EB_TP_MAXNREG)
#else
__launch_bounds__(32 * (TY + 2), EB_TP_MIN_CTAS)
#endif
flux_update_kernel_tp(const EbParams P, const EbGas* __restrict__ gas, const EbBlockDesc* __restrict__ descs, int nblocks,
                      const EbArena A, const EbStageArgs S)
{
    typedef Layout<DIM, NSP> Lay;
    typedef TpSmem<DIM, NSP, TY> SM;
    constexpr int NCQ = Lay::NCQ, NS = SM::NS, NTM = SM::NTM;
    constexpr int NT = 32 * (TY + 2);
    extern __shared__ __align__(128) double smem[];
    double* const sPj = smem + SM::O_PJ;
    double* const sFS = smem + SM::O_FS;
    double* const sPiW = smem + SM::O_PIW;
    double* const sPi = smem + SM::O_PI;
    double* const sMj = smem + SM::O_MJ;
    double* const sFE = smem + SM::O_FE;
    double* const sK = smem + SM::O_K;
    double* const sAcc = smem + SM::O_ACC;
    EbBlockDesc& D = *reinterpret_cast<EbBlockDesc*>(smem + SM::O_DESC);
    __shared__ int s_blk;
    __shared__ TpGasS<NSP> G;

    const int lane = threadIdx.x, wy = threadIdx.y;
    const int tid = wy * 32 + lane;
    const long long cta = S.tile_list ? (long long)S.tile_list[blockIdx.x] : (long long)blockIdx.x;
    if (tid == 0) {
        int lo = 0, hi = nblocks - 1;
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (descs[mid].tile0 <= cta) lo = mid; else hi = mid - 1; }
        s_blk = lo;
    }
    tp_fill_gas<NSP>(gas, G, tid, NT);
    __syncthreads();
    {
        const int* src = reinterpret_cast<const int*>(&descs[s_blk]);
        int* dst = reinterpret_cast<int*>(&D);
        for (int n = tid; n < (int)(sizeof(EbBlockDesc) / sizeof(int)); n += NT) dst[n] = src[n];
    }
    __syncthreads();
    if (!D.cartesian) return;

    const long long t = cta - D.tile0;
    const int ti = (int)(t % D.tiles_i);
    const int tj = (int)((t / D.tiles_i) % D.tiles_j);
    const int tm = (int)(t / ((long long)D.tiles_i * D.tiles_j));
    const int i0 = ti * 32, j0 = tj * TY;
    const int nic = D.nic, njc = D.njc, nkc = D.nkc;
    const long long sj = D.stride[1], sk = D.stride[2];
    const long long total = P.total;
    const int k0 = (DIM == 3) ? tm * D.chunk_m : 0;
    const int k1 = (DIM == 3) ? min(nkc, k0 + D.chunk_m) : 1;
    const double* __restrict__ prim = S.prim_in;
    const bool clip = P.extrema_clipping != 0;
    auto eps_of = [&](int d) -> double {
#ifdef EB_FAST_MATH
        return P.eps_va * D.uq[d][4];
#else
        (void)d; return P.eps_va;
#endif
    };
    auto cell_at = [&](int ii, int jj, int kk) -> long long {
        return D.cell0 + ((long long)(kk + D.kg) * D.NJ + (jj + EB_NG)) * D.NI + (ii + EB_NG);
    };
    auto alpha_at = [&](int d, long long cf) -> double {
        if (!FluxPair<FLUX>::adaptive) return 0.0;
        const double* sf = (d == 0) ? A.Sf[0] : ((d == 1) ? A.Sf[1] : A.Sf[2]);
        return sf[cf];
    };
    bool fail = false;
    int n_invalid = 0;
    Prim<NSP> rare;                                   // scratch of the out-of-line reference route (see tp_thermo)
    const int kfirst = k0 - 1;       // a lead-in pass: the first states along k (3D) and the west halo of the first plane
    const int kend = (DIM == 3) ? k1 : 0;

    // =============================== helper warps ==============================================================
    // warp TY    : plus states of the row south of the tile (all lanes) and, a plane ahead, of the column west of it (lanes < TY)
    // warp TY + 1: minus states of the row north / the column east of the tile, then the faces on those two edges
    if (wy >= TY) {
        const bool h0 = (wy == TY);
        const int i = i0 + lane;
        const bool rowOk = h0 ? (i < nic) : ((j0 + TY <= njc) && (i < nic));
        const bool colOk = (lane < TY) && (j0 + lane < njc) && (h0 || (i0 + 32 <= nic));
        Prim<NSP> Y, Z;
        for (int k = kfirst; k <= kend; ++k) {
            const bool pre = (k < k0);
            const bool has_cells = !pre && (k < k1);
            if (!pre && !has_cells) break;
            const int par = (k - k0) & 1;
#pragma unroll 1
            for (int jb = 0; jb < 2; ++jb) {
                // jb 0: the row (along j);  jb 1: the column (along i; the west column belongs to the next plane)
                const int d = (jb == 0) ? 1 : 0;
                const int kk = (h0 && jb == 1) ? k + 1 : k;
                bool act = (jb == 0) ? (rowOk && !pre) : colOk;
                if (jb == 1 && h0 && kk >= k1) act = false;
                if (jb == 1 && !h0 && pre) act = false;
                const long long cc = (jb == 0) ? cell_at(i, h0 ? j0 - 1 : j0 + TY, kk) : cell_at(h0 ? i0 - 1 : i0 + 32, j0 + lane, kk);
                Prim<NSP> M;
                if (act) {
                    double* dstP = (jb == 0) ? sPj + lane : sPiW + (((par ^ 1) * NS) * TY) + lane;
                    tp_cell_states<DIM, NSP>(D, G, gas, d, eps_of(d), clip, prim, total, cc, (d == 0) ? 1 : sj, !h0, h0,
                                             dstP, (jb == 0) ? 32 : TY, M, fail, rare);
                    if (!h0) { if (jb == 0) Y = M; else Z = M; }
                }
            }
            __syncthreads();
            if (pre) continue;
            if (!h0) {
#pragma unroll 1
                for (int jb = 0; jb < 2; ++jb) {
                    const int d = (jb == 0) ? 1 : 0;
                    const bool act = (jb == 0) ? rowOk : colOk;
                    if (!act) continue;
                    const long long cf = (jb == 0) ? cell_at(i, j0 + TY, k) : cell_at(i0 + 32, j0 + lane, k);
                    double F[NCQ];
                    if (!tp_outflow_override<DIM, NSP>(P, D, prim, d, (jb == 0) ? j0 + TY : i0 + 32, (jb == 0) ? njc : nic, cf, (jb == 0) ? sj : 1, F)) {
                        Prim<NSP> L;
                        if (jb == 0) tp_get<NSP>(sPj + (TY * NS) * 32 + lane, 32, L);
                        else tp_get<NSP>(sPi + (lane * NS) * 32 + 31, 32, L);          // the plus state of lane 31 in row `lane`
                        tp_face_flux<DIM, NSP, FLUX>(P, gas, d, L, (jb == 0) ? Y : Z, alpha_at(d, cf), F);
                    }
                    if (jb == 0) {
#pragma unroll
                        for (int q = 0; q < NCQ; ++q) sFS[(TY * NCQ + q) * 32 + lane] = F[q];
                    } else {
#pragma unroll
                        for (int q = 0; q < NCQ; ++q) sFE[q * TY + lane] = F[q];
                    }
                }
            }
            __syncthreads();
        }
        if (__ballot_sync(0xffffffffu, fail) && lane == 0) atomicOr(&S.status[0], 1);
        return;
    }

    // =============================== main warps: one thread per cell of the tile ================================
    const int i = i0 + lane, j = j0 + wy;
    const bool cell_ok = (i < nic) && (j < njc);
    const bool doI = (i <= nic) && (j < njc);
    const bool doJ = (i < nic) && (j <= njc);
    double* const myK = sK + tid;             // field f at stride NTM, buffer b at b * NS * NTM
    double* const myAcc = sAcc + tid;
    double* const myPi = sPi + (wy * NS) * 32 + lane;
    double* const myMj = sMj + (wy * NS) * 32 + lane;
    const double aI = D.area[0], aJ = D.area[1], aK = D.area[2];

    for (int k = kfirst; k <= kend; ++k) {
        const bool pre = (k < k0);
        const bool has_cells = !pre && (k < k1);
        const int par = (k - k0) & 1;
        const long long c = cell_at(i, j, (DIM == 3) ? k : 0);
        const int nd = (DIM == 3) ? (has_cells ? 3 : 1) : (pre ? 0 : 2);
        // species energies and heat capacities at the cell temperature: once per cell, for all its directions
        TpCellTable<NSP> tbc;
        tp_table<NSP>(G, ldg(prim + 3 * total + c), tbc);
#pragma unroll 1
        for (int dd = 0; dd < nd; ++dd) {
            // 3D: along k first (the bottom face of this plane is the top face of the cell one plane below, which is then
            // complete), then along i (the west face; the east flux comes back from the next lane), then along j
            const int d = (DIM == 3) ? ((dd == 0) ? 2 : dd - 1) : dd;
            const bool act = (d == 2) ? cell_ok : ((d == 0) ? doI : doJ);
            const bool wM = !pre;
            const bool wP = (d == 2) ? (pre || has_cells) : cell_ok;
            const long long st = (d == 0) ? 1 : ((d == 1) ? sj : sk);
            // where the plus state goes: along k into the other buffer of the thread's slot (this plane still needs the
            // one below), along i into the row buffer (read by the next lane), along j into the slot of the row above
            double* dstP = (d == 2) ? myK + ((par ^ 1) * NS) * NTM : ((d == 0) ? myPi : sPj + ((wy + 1) * NS) * 32 + lane);
            const int strideP = (d == 2) ? NTM : 32;
            Prim<NSP> M;
            if (act) tp_cell_states<DIM, NSP>(D, G, gas, d, eps_of(d), clip, prim, total, c, st, wM, wP, dstP, strideP, M, fail, rare, &tbc);
            if (d == 1) {
                if (doJ) tp_put<NSP>(myMj, 32, M);
                continue;
            }
            if (d == 0) __syncwarp();
            double F[NCQ];
#pragma unroll
            for (int q = 0; q < NCQ; ++q) F[q] = 0.0;
            const bool dofl = (d == 2) ? (cell_ok && !pre) : doI;
            if (dofl) {
                const int idx = (d == 0) ? i : k, n = (d == 0) ? nic : nkc;
                if (!tp_outflow_override<DIM, NSP>(P, D, prim, d, idx, n, c, st, F)) {
                    Prim<NSP> L;
                    if (d == 2) tp_get<NSP>(myK + (par * NS) * NTM, NTM, L);
                    else if (lane == 0) tp_get<NSP>(sPiW + ((par * NS) * TY) + wy, TY, L);
                    else tp_get<NSP>(myPi - 1, 32, L);
                    tp_face_flux<DIM, NSP, FLUX>(P, gas, d, L, M, alpha_at(d, c), F);
                }
            }
            if (d == 2) {
                if (cell_ok && !pre) {
                    if (k > k0) {
                        double dUdt[NCQ];
#pragma unroll
                        for (int q = 0; q < NCQ; ++q) { const double si = myAcc[q * NTM] - F[q] * aK; dUdt[q] = D.vol_inv * si + 0.0; }
                        long long p0, p1, p2;
                        push_targets<DIM>(D, i, j, k - 1, c - sk, p0, p1, p2);
                        tp_finish_cell<DIM, NSP>(P, G, gas, S, total, c - sk, dUdt, fail, n_invalid, p0, p1, p2, rare);
                    }
                    if (has_cells) {
#ifdef EB_FAST_MATH
#pragma unroll
                        for (int q = 0; q < NCQ; ++q) myAcc[q * NTM] = F[q] * aK;
#else
#pragma unroll
                        for (int q = 0; q < NCQ; ++q) myAcc[(NCQ + q) * NTM] = F[q];
#endif
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < NCQ; ++q) {
                    const double fe = __shfl_down_sync(0xffffffffu, F[q], 1);
                    if (cell_ok) {
#ifdef EB_FAST_MATH
                        double si = (DIM == 3) ? myAcc[q * NTM] : 0.0;
                        si = fma(F[q], aI, si);
                        if (lane != 31) si = fma(-fe, aI, si);
#else
                        double si = F[q] * aI;          // 0 - F*(-A): summation order W, E, S, N, B, T (fvcell.d:824-854)
                        if (lane != 31) si = si - fe * aI;
#endif
                        myAcc[q * NTM] = si;
                    }
                }
            }
        }
        if (pre) { __syncthreads(); continue; }
        if (!has_cells) break;
        __syncthreads();
        {
            double FS_[NCQ];
#pragma unroll
            for (int q = 0; q < NCQ; ++q) FS_[q] = 0.0;
            if (doJ) {
                if (!tp_outflow_override<DIM, NSP>(P, D, prim, 1, j, njc, c, sj, FS_)) {
                    Prim<NSP> L, Mj;
                    tp_get<NSP>(sPj + (wy * NS) * 32 + lane, 32, L);
                    tp_get<NSP>(myMj, 32, Mj);
                    tp_face_flux<DIM, NSP, FLUX>(P, gas, 1, L, Mj, alpha_at(1, c), FS_);
                }
            }
#pragma unroll
            for (int q = 0; q < NCQ; ++q) sFS[(wy * NCQ + q) * 32 + lane] = FS_[q];
        }
        __syncthreads();

        // ---- the rest of the surface integral of this plane's cell (all but the top face)
        if (cell_ok) {
            double acc[NCQ];
#pragma unroll
            for (int q = 0; q < NCQ; ++q) {
                const double fs = sFS[(wy * NCQ + q) * 32 + lane];
                const double fn = sFS[((wy + 1) * NCQ + q) * 32 + lane];
                double si = myAcc[q * NTM];
#ifdef EB_FAST_MATH
                if (lane == 31) si = fma(-sFE[q * TY + wy], aI, si);
                si = fma(fs, aJ, si);
                si = fma(-fn, aJ, si);
#else
                if (lane == 31) si = si - sFE[q * TY + wy] * aI;
                si = si + fs * aJ;
                si = si - fn * aJ;
                if (DIM == 3) si = si + myAcc[(NCQ + q) * NTM] * aK;
#endif
                acc[q] = si;
            }
            if (DIM == 3) {
#pragma unroll
                for (int q = 0; q < NCQ; ++q) myAcc[q * NTM] = acc[q];
            } else {
                double Qy = 0.0;
                if (P.axisymmetric) Qy = ldg(prim + 2 * total + c) * D.areaxy / D.vol;      // fvcell.d:1161-1165
                double dUdt[NCQ];
#pragma unroll
                for (int q = 0; q < NCQ; ++q) dUdt[q] = D.vol_inv * acc[q] + ((q == Lay::iYMom) ? Qy : 0.0);
                long long p0, p1, p2;
                push_targets<DIM>(D, i, j, 0, c, p0, p1, p2);
                tp_finish_cell<DIM, NSP>(P, G, gas, S, total, c, dUdt, fail, n_invalid, p0, p1, p2, rare);
            }
        }
    }

    const unsigned any_fail = __ballot_sync(0xffffffffu, fail);
    int inv = n_invalid;
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) inv += __shfl_down_sync(0xffffffffu, inv, o2);
    if (lane == 0) {
        if (any_fail) atomicOr(&S.status[0], 1);
        if (inv) atomicAdd(&S.status[S.stage], inv);
    }
}

template <int DIM, int FLUX, int NSP>
void launch_one_tp(const EbParams& P, const EbGas* gas, const EbBlockDesc* desc, int nblocks, long long ncta,
                   const EbArena& A, const EbStageArgs& S, cudaStream_t st)
{
    constexpr int TY = EB_TILE_Y;
    const size_t smem = TpSmem<DIM, NSP, TY>::BYTES;
    auto kern = flux_update_kernel_tp<DIM, FLUX, NSP, TY>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    kern<<<(unsigned)ncta, dim3(32, TY + 2), smem, st>>>(P, gas, desc, nblocks, A, S);
}

// uniform-Cartesian blocks, thermally perfect gas, interpolation_order = 2, apply_limiter = true, thermo_interpolator = rhou.
// Returns false when no kernel is built for this species count.
template <int FLUX>
bool launch_flux_update_tp_impl(const EbParams& P, const EbGas* gas, const EbBlockDesc* desc, int nblocks,
                                long long ncta, const EbArena& A, const EbStageArgs& S, cudaStream_t st)
{
#define EB_TP(DIM, NSP) launch_one_tp<DIM, FLUX, NSP>(P, gas, desc, nblocks, ncta, A, S, st)
#define EB_TP_NSP(N) if (P.nsp == N) { if (P.dims == 3) EB_TP(3, N); else EB_TP(2, N); return true; }
    EB_TPG_NSP_LIST(EB_TP_NSP)
#undef EB_TP_NSP
#undef EB_TP
    return false;
}

}  // namespace EB_NS
