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

namespace EB_NS {

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
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

// Reconstruction + thermo + flux in the face frame.  Returns F in the face frame (momentum
// components along n, t1, t2).  ROT: general-metric path (velocities rotate with `fr`).
template <int DIM, int FLUX, bool CLIP, bool ROT>
__device__ __forceinline__ void face_core(const EbParams& P, const EbGas* __restrict__ gas, const EbWeights& w,
                                          Stencil& s, const Frame& fr, const double* __restrict__ prim_fallback,
                                          long long cL0, long long cR0, double* F)
{
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
        if (P.local_frame) {   // reference: back to global (onedinterp.d:979-987), then into the face frame again
            to_global<DIM>(fr, L.vx, L.vy, L.vz); to_global<DIM>(fr, R.vx, R.vy, R.vz);
        }
        to_local<DIM>(fr, L.vx, L.vy, L.vz); to_local<DIM>(fr, R.vx, R.vy, R.vz);
    }
    if (FLUX == EB200_FLUX_AUSMDV) flux_ausmdv<DIM, 1>(L, R, P.entropy_fix != 0, F);
    else if (FLUX == EB200_FLUX_HANEL) flux_hanel<DIM, 1>(L, R, F);
    else if (FLUX == EB200_FLUX_LDFSS0) flux_ldfss<DIM, 1, 0>(L, R, F);
    else if (FLUX == EB200_FLUX_LDFSS2) flux_ldfss<DIM, 1, 2>(L, R, F);
    else if (FLUX == EB200_FLUX_AUSM_PLUS_UP) flux_ausm_plus_up<DIM, 1>(L, R, P.M_inf, F);
    else flux_roe<DIM, 1>(L, R, gas->gamma_CpCv, gas->gamma_CpCv, F);
}

// Shared-memory tile of one k-plane: NF fields x ROWS x COLS doubles (halo of 2 on each side).
template <int DIM, int TY>
struct Tile {
    static constexpr int NF = (DIM == 3) ? 6 : 5;          // rho, u, vx, vy, [vz], a
    static constexpr int ROWS = TY + 4, COLS = 36;
    static constexpr int FSZ = ROWS * COLS;                // doubles per field
    static constexpr int SIZE = NF * FSZ;
    static constexpr int F_RHO = 0, F_U = 1, F_V = 2, F_A = (DIM == 3) ? 5 : 4;
    __device__ static constexpr int prim_index(int f)
    {
        return (f == 0) ? 0 : (f == 1) ? 1 : (f == F_A) ? 4 : 5 + (f - 2);
    }
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
__global__ void __launch_bounds__(32 * TY, 2)
flux_update_kernel_v2(const EbParams P, const EbGas* __restrict__ gas, const EbBlockDesc* __restrict__ descs, int nblocks,
                      const EbArena A, const EbStageArgs S)
{
    typedef Layout<DIM, 1> Lay;
    typedef Tile<DIM, TY> T;
    constexpr int NCQ = Lay::NCQ;
    constexpr int NT = 32 * TY;
    extern __shared__ double smem[];
    // layout: tile[2][SIZE] | fW[2][NCQ][TY][33] | fS[2][NCQ][TY+1][32] | desc
    double* tile = smem;
    double* fWs = tile + 2 * T::SIZE;
    double* fSs = fWs + 2 * NCQ * TY * 33;
    EbBlockDesc& D = *reinterpret_cast<EbBlockDesc*>(fSs + 2 * NCQ * (TY + 1) * 32);
    __shared__ int s_blk;

    const int lane = threadIdx.x, wy = threadIdx.y;
    const int tid = wy * 32 + lane;
    const long long cta = blockIdx.x;
    if (tid == 0) {
        int lo = 0, hi = nblocks - 1;
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (descs[mid].tile0 <= cta) lo = mid; else hi = mid - 1; }
        s_blk = lo;
    }
    __syncthreads();
    {
        const int* src = reinterpret_cast<const int*>(&descs[s_blk]);
        int* dst = reinterpret_cast<int*>(&D);
        for (int n = tid; n < (int)(sizeof(EbBlockDesc) / sizeof(int)); n += NT) dst[n] = src[n];
    }
    __syncthreads();
    if ((D.cartesian != 0) != CART) return;

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
    // second trip: warp 0 -> faces east of lane 31 (row = lane), warp 1 -> south faces of row TY
    const bool extraE_ok = (wy == 0) && (lane < TY) && (i0 + 32 <= nic) && (j0 + lane < njc);
    const bool extraN_ok = (wy == 1) && (j0 + TY <= njc) && (i < nic);

    // cp.async work list of this thread: tile positions tid and tid + NT (of ROWS*COLS)
    int pos_s[2], pos_g[2];
    bool pos_ok[2];
#pragma unroll
    for (int n = 0; n < 2; ++n) {
        const int p = tid + n * NT;
        const int r = p / T::COLS, c = p % T::COLS;
        pos_ok[n] = (p < T::ROWS * T::COLS) && (j0 + r < NJ) && (i0 + c < NI);
        pos_s[n] = r * T::COLS + c;
        pos_g[n] = (j0 + r) * NI + (i0 + c);             // offset within a padded k-plane
    }
    auto stage_plane = [&](int k, int buf) {
        // plane k (interior index; padded k + kg) -> tile[buf]
        const long long plane = D.cell0 + (long long)(k + D.kg) * NJ * NI;
        double* dst = tile + buf * T::SIZE;
#pragma unroll
        for (int f = 0; f < T::NF; ++f) {
            const double* src = S.prim_in + (long long)T::prim_index(f) * total + plane;
#pragma unroll
            for (int n = 0; n < 2; ++n)
                if (pos_ok[n]) cp_async8(dst + f * T::FSZ + pos_s[n], src + pos_g[n]);
        }
        cp_async_commit();
    };

    double acc[NCQ];
#pragma unroll
    for (int q = 0; q < NCQ; ++q) acc[q] = 0.0;
    bool fail = false;
    int n_invalid = 0;

    const int kend = (DIM == 3) ? k1 : 0;
    stage_plane(k0, 0);
    cp_async_wait_all();
    __syncthreads();

    for (int k = k0; k <= kend; ++k) {
        const int buf = (k - k0) & 1;
        const bool plane_has_cells = (DIM == 3) ? (k < k1) : true;
        // prefetch the next plane (its own cells are needed for the top face of the chunk too)
        if (DIM == 3 && k < kend) stage_plane(k + 1, buf ^ 1);
        const double* tl = tile + buf * T::SIZE;
        double* fW = fWs + buf * NCQ * TY * 33;
        double* fS = fSs + buf * NCQ * (TY + 1) * 32;
        const long long c = D.cell0 + ((long long)(k + D.kg) * NJ + (j + EB_NG)) * NI + (i + EB_NG);
        double FB[NCQ];
#pragma unroll
        for (int q = 0; q < NCQ; ++q) FB[q] = 0.0;

        // ---------------- the faces of this plane: one loop, ONE inlined copy of the face arithmetic -------
        // job 0: west face (d=0), job 1: south face (d=1), job 2: bottom face (d=2, 3D),
        // job 3: second trip for warp 0 (faces east of lane 31) and warp 1 (south faces of row TY)
#pragma unroll 1
        for (int job = 0; job < 4; ++job) {
            if (job == 2 && DIM != 3) continue;
            if (job == 3 && (wy > 1 || !plane_has_cells)) break;
            const int d = (job < 3) ? job : wy;
            int row = wy + 2, col = lane + 2;            // tile coordinates of the plus-side cell of the face
            int fi = i, fj = j;                          // its interior indices
            bool active;
            if (job == 0) active = plane_has_cells && (i <= nic) && (j < njc);
            else if (job == 1) active = plane_has_cells && (i < nic) && (j <= njc);
            else if (job == 2) active = cell_ok;
            else if (d == 0) { row = lane + 2; col = 34; fi = i0 + 32; fj = j0 + lane; active = extraE_ok; }
            else { row = TY + 2; fj = j0 + TY; active = extraN_ok; }
            double F[NCQ];
#pragma unroll
            for (int q = 0; q < NCQ; ++q) F[q] = 0.0;
            if (active) {
                const long long cf = D.cell0 + ((long long)(k + D.kg) * NJ + (fj + EB_NG)) * NI + (fi + EB_NG);
                const long long st = (d == 0) ? 1 : ((d == 1) ? sj : sk);
                int bcf = -1;
                if (d == 0) { if (fi == 0) bcf = EB200_WEST; else if (fi == nic) bcf = EB200_EAST; }
                else if (d == 1) { if (fj == 0) bcf = EB200_SOUTH; else if (fj == njc) bcf = EB200_NORTH; }
                else { if (k == 0) bcf = EB200_BOTTOM; else if (k == nkc) bcf = EB200_TOP; }
                if (bcf >= 0 && D.bc_kind[bcf] == EB200_BC_OUTFLOW_SIMPLE_FLUX) {
                    const int hi = bcf & 1;
                    Prim<1> fs;
                    load_prim<1>(fs, S.prim_in, total, hi ? cf - st : cf);
                    if (DIM == 2) fs.vz = 0.0;
                    double nx, ny, nz;
                    if (CART) { nx = D.nvec[d][0]; ny = D.nvec[d][1]; nz = D.nvec[d][2]; }
                    else { nx = ldg(A.face[d] + cf); ny = ldg(A.face[d] + total + cf); nz = (DIM == 3) ? ldg(A.face[d] + 2 * total + cf) : 0.0; }
                    outflow_flux<DIM, 1>(fs, hi ? 1 : -1, nx, ny, nz, F);
                } else {
                    const double* q0 = tl + row * T::COLS + col;     // own (R0) cell, field 0
                    Stencil s;
                    // ---- gather the stencil; velocities in (x, y, z) order first
                    double vx[4], vy[4], vz[4];
                    if (d < 2) {
                        const int so = (d == 0) ? 1 : T::COLS;       // tile stride along d
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            const double* qm = q0 + (m - 2) * so;
                            s.rho[m] = qm[T::F_RHO * T::FSZ]; s.u[m] = qm[T::F_U * T::FSZ];
                            vx[m] = qm[(T::F_V + 0) * T::FSZ]; vy[m] = qm[(T::F_V + 1) * T::FSZ];
                            vz[m] = (DIM == 3) ? qm[(T::F_V + 2) * T::FSZ] : 0.0;
                        }
                        s.aL = (q0 - so)[T::F_A * T::FSZ];
                    } else {
                        // own column: plane k from the tile, planes k-2, k-1, k+1 from global memory (L2)
                        const double* pin = S.prim_in;
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            if (m == 2) {
                                s.rho[m] = q0[T::F_RHO * T::FSZ]; s.u[m] = q0[T::F_U * T::FSZ];
                                vx[m] = q0[(T::F_V + 0) * T::FSZ]; vy[m] = q0[(T::F_V + 1) * T::FSZ];
                                vz[m] = (DIM == 3) ? q0[(T::F_V + 2) * T::FSZ] : 0.0;
                            } else {
                                const long long cm = cf + (m - 2) * sk;
                                s.rho[m] = ldg(pin + cm); s.u[m] = ldg(pin + total + cm);
                                vx[m] = ldg(pin + 5 * total + cm); vy[m] = ldg(pin + 6 * total + cm); vz[m] = ldg(pin + 7 * total + cm);
                            }
                        }
                        s.aL = ldg(pin + 4 * total + cf - sk);
                    }
                    s.aR = q0[T::F_A * T::FSZ];
                    // ---- into the face frame: a renaming on the Cartesian path (CartFrame conventions)
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        if (!CART) { s.v0[m] = vx[m]; s.v1[m] = vy[m]; s.v2[m] = vz[m]; }
                        else if (DIM == 3) {
                            s.v0[m] = (d == 0) ? vx[m] : ((d == 1) ? vy[m] : vz[m]);
                            s.v1[m] = (d == 0) ? vy[m] : ((d == 1) ? vz[m] : vx[m]);
                            s.v2[m] = (d == 0) ? vz[m] : ((d == 1) ? vx[m] : vy[m]);
                        } else {
                            s.v0[m] = (d == 0) ? vx[m] : vy[m];
                            s.v1[m] = (d == 0) ? -vy[m] : vx[m];
                            s.v2[m] = 0.0;
                        }
                    }
                    Frame fr;
                    EbWeights wl;
                    if (!CART) {
                        load_frame<DIM>(fr, A.face[d], total, cf);
                        const double* ln = A.len[d];
                        l2r2_prepare(wl, ldg(ln + cf - 2 * st), ldg(ln + cf - st), ldg(ln + cf), ldg(ln + cf + st));
                    }
                    const EbWeights& w = CART ? D.w[d] : wl;
                    double Fl[NCQ];
                    face_core<DIM, FLUX, CLIP, !CART>(P, gas, w, s, fr, S.prim_in, cf - st, cf, Fl);
                    // ---- momentum flux back to the global frame
                    F[Lay::iMass] = Fl[Lay::iMass]; F[Lay::iEnergy] = Fl[Lay::iEnergy];
                    if (!CART) {
                        double fx = Fl[Lay::iXMom], fy = Fl[Lay::iYMom], fz = (DIM == 3) ? Fl[Lay::iZMom] : 0.0;
                        to_global<DIM>(fr, fx, fy, fz);
                        F[Lay::iXMom] = fx; F[Lay::iYMom] = fy;
                        if (DIM == 3) F[Lay::iZMom] = fz;
                    } else if (DIM == 3) {
                        const double f0 = Fl[Lay::iXMom], f1 = Fl[Lay::iYMom], f2 = Fl[Lay::iZMom];
                        F[Lay::iXMom] = (d == 0) ? f0 : ((d == 1) ? f2 : f1);
                        F[Lay::iYMom] = (d == 0) ? f1 : ((d == 1) ? f0 : f2);
                        F[Lay::iZMom] = (d == 0) ? f2 : ((d == 1) ? f1 : f0);
                    } else {
                        const double f0 = Fl[Lay::iXMom], f1 = Fl[Lay::iYMom];
                        F[Lay::iXMom] = (d == 0) ? f0 : f1;
                        F[Lay::iYMom] = (d == 0) ? -f1 : f0;
                    }
                }
            }
            // ---- publish
            if (job == 2) {
#pragma unroll
                for (int q = 0; q < NCQ; ++q) FB[q] = F[q];
            } else if (d == 0) {
                if (job == 0 || lane < TY) {
#pragma unroll
                    for (int q = 0; q < NCQ; ++q) fW[(q * TY + (row - 2)) * 33 + (col - 2)] = F[q];
                }
            } else {
#pragma unroll
                for (int q = 0; q < NCQ; ++q) fS[(q * (TY + 1) + (row - 2)) * 32 + (col - 2)] = F[q];
            }
        }

        cp_async_wait_all();
        __syncthreads();

        // finish the cell of the previous plane: its top face is this plane's bottom face
        if (DIM == 3 && k > k0 && cell_ok) {
            const long long cp = c - sk;
            const double areaT = CART ? D.area[2] : ldg(A.face[2] + 9 * total + c);
            const double vol_inv = CART ? D.vol_inv : 1.0 / ldg(A.vol + cp);
            double dUdt[NCQ];
#pragma unroll
            for (int q = 0; q < NCQ; ++q) { double si = acc[q] - FB[q] * areaT; dUdt[q] = vol_inv * si + 0.0; }
            finish_cell<DIM, EB200_GAS_IDEAL, 1>(P, gas, S, total, cp, dUdt, fail, n_invalid);
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
                const double FW = fW[(q * TY + wy) * 33 + lane], FE = fW[(q * TY + wy) * 33 + lane + 1];
                const double FS = fS[(q * (TY + 1) + wy) * 32 + lane], FN = fS[(q * (TY + 1) + wy + 1) * 32 + lane];
                double si = FW * aW;          // 0 - F*(-A)
                si = si - FE * aE;
                si = si + FS * aS;
                si = si - FN * aN;
                if (DIM == 3) si = si + FB[q] * aB;
                acc[q] = si;
            }
        }

        if (DIM == 2 && cell_ok) {
            const double vol = CART ? D.vol : ldg(A.vol + c);
            const double vol_inv = CART ? D.vol_inv : 1.0 / vol;
            double Qy = 0.0;
            if (P.axisymmetric) {      // fvcell.d:1161-1165
                const double axy = CART ? D.areaxy : ldg(A.areaxy + c);
                Qy = ldg(S.prim_in + 2 * total + c) * axy / vol;
            }
            double dUdt[NCQ];
#pragma unroll
            for (int q = 0; q < NCQ; ++q) dUdt[q] = vol_inv * acc[q] + ((q == Lay::iYMom) ? Qy : 0.0);
            finish_cell<DIM, EB200_GAS_IDEAL, 1>(P, gas, S, total, c, dUdt, fail, n_invalid);
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
    return sizeof(double) * (2 * T::SIZE + 2 * NCQ * TY * 33 + 2 * NCQ * (TY + 1) * 32) + sizeof(EbBlockDesc);
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
