// eb200_internal.h -- structures shared by the host runtime (eb200.cu) and the kernels.
//
// Device data layout (all FP64, structure of arrays):
//   every per-cell field of every local block lives in one arena array of `total`
//   doubles; block b occupies [cell0, cell0 + NI*NJ*NK) in the padded block layout of
//   include/eb200.h (i fastest).  Field f of a group of nf fields is base + f*total.
//   prim[3]   : FlowState variables, nprim fields each: the start-of-step state is never
//               overwritten during a step (exact restore after a failed step), the two other
//               buffers ping-pong between stages
//   U[ns+1]   : conserved quantities per time level, ncq fields each
//   dUdt[ns]  : residuals per stage
//   geometry  : vol, areaxy, len[3], face[d][10]  (only allocated when some block is
//               not uniform-Cartesian; Cartesian blocks carry constants in BlockDesc)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/eb200.h"

#define EB_NG EB200_NGHOST
#define EB_MAXSP EB200_MAX_SPECIES
#define EB_MAXSEG EB200_MAX_SEGMENTS

// l2r2_prepare() results (reference src/eilmer/onedinterp.d:338-354)
#ifndef EB_V2_TY
#define EB_V2_TY 8            // rows of cells per CTA tile of the tuned kernel (CTA = 32 x EB_V2_TY threads)
#endif
#define EB_V2_COLS 36         // tile width with its two-cell halo
#ifndef EB_V3_TY
#define EB_V3_TY 16           // rows of cells per CTA tile of the cell-centred kernel (CTA = 32 x (EB_V3_TY + 1) threads, one per SM)
#endif

struct EbWeights {
    double aL0, aR0, lenL0, lenR0;
    double two_over_L0L1, two_over_R0L0, two_over_R1R0;
    double two_L0_plus_L1, two_R0_plus_R1;
};

// Per-direction frame of an axis-aligned face: local (x,y,z) = sgn[m]*v[perm[m]]
struct EbAxisFrame {
    int perm[3];
    int neg[3];   // 1: negate
};

struct EbBlockDesc {
    int nic, njc, nkc;
    int NI, NJ, NK, kg;
    int cartesian;
    int v3;                   // 1: the block is tiled for (and run by) the cell-centred kernel flux_update_kernel_v3
    int bc_kind[6];
    long long cell0;          // offset into the arena
    long long stride[3];
    // tiling of the fused flux kernel
    int tiles_i, tiles_j, tiles_m, chunk_m;
    long long tile0;          // first CTA index of this block
    // uniform-Cartesian constants
    double vol, vol_inv, areaxy;
    double len[3], area[3];
    EbWeights w[3];
    EbAxisFrame fr[3];
    double nvec[3][3];        // face normal per direction (exact +-1/0), used by flux BCs and CFL
    // throughput build, uniform spacing: reconstruction on raw differences.  With c = 2/(2*len):
    // uq[d] = { aL0*two_L0_plus_L1*c, aL0*lenR0*c, aR0*lenL0*c, aR0*two_R0_plus_R1*c, 1/c^2 }
    // (van Albada's epsilon is rescaled by 1/c^2 so that the limiter value is the same number)
    double uq[3][5];
    int outflow_flux_faces;   // bit f set: face f has EB200_BC_OUTFLOW_SIMPLE_FLUX
    int noghost_faces;        // bit f set: face f has no ghost-cell data (EB200_BC_WALL_WITH_SLIP1): one-sided stencils, wall flux
    // Same-GPU full-face neighbours behind opposite, equally sized faces get their ghost cells written by
    // the kernel that produces the FlowStates ("push"): a cell within two layers of face f is also stored
    // at arena index c + push_off[f], the ghost cell of the neighbour that mirrors it (full_face_copy.d
    // semantics for aligned blocks).  bit f of push_mask: face f pushes.
    int push_mask;
    long long push_off[6];
};

struct EbCurve {             // reference src/gas/thermo/cea_thermo_curves.d
    double R;
    int nseg, nbreaks;
    double T_breaks[EB_MAXSEG + 1], T_blends[EB_MAXSEG];
    double coeffs[EB_MAXSEG][9];
    double T_low, T_high, Cp_low, Cp_high, h_low, h_high;
};

struct EbGas {
    int model, nsp;
    // ideal gas (reference src/gas/ideal_gas.d:64-68)
    double Rgas, gamma, Cv, Cvinv, Cp, gamma_CpCv;
    double Rsp[EB_MAXSP];
    EbCurve curves[EB_MAXSP];
    // throughput build: when all species share break points and blend ranges, the mixture's energy and
    // Cv are ONE polynomial in T whose coefficients are sum_i massf_i * RA[seg][k][i], RA = R_i * a_i[seg][k]
    int uniform_curves, pad1;
    double RA[EB_MAXSEG][8][EB_MAXSP];
};

struct EbParams {            // passed by value to kernels (kept small)
    int dims, axisymmetric, nsp, ncq, nprim;
    int interpolation_order, apply_limiter, extrema_clipping, local_frame, entropy_fix;
    int ignore_low_T;
    int iZMom, iEnergy, iSpecies;
    int shock_detect, strict_shock;       // adaptive flux calculators: PJ shock detector on
    int thermo_interp, lmr;               // eb200_thermo_interpolator; lmr != 0: eb200_config.solver_variant = 1 (src/lmr formulas)
    double eps_va, M_inf, max_velocity, max_temp, min_temp, low_T;
    double comp_tol, shear_tol;
    long long total;          // arena length (field stride)
};

struct EbArena {             // device pointers
    double* prim[3];          // [cur] = state at the start of the step (kept intact), the other two ping-pong
    double* U[5];
    double* dUdt[4];
    double* vol; double* areaxy; double* len[3]; double* face[3];
    double* S;                 // FlowState.S of cells incl. ghost cells (shock detector), one copy
    double* Sf[3];             // IFace.fs.S per direction
};

struct EbStageArgs {
    const double* prim_in; double* prim_out;
    const double* U0; double* U_out;       // U_out: only written in the final stage (else nullptr)
    const double* dUdt_prev[3];            // residuals of earlier stages (nullptr when unused)
    double* dUdt_out;                      // nullptr when no later stage needs it
    double dt_g[5];                        // stage 1: dt*g0 in [0]; later stages: g[0..3]; dt in [4]
    int stage, n_stages;
    int* status;                           // [0] step-failed flag, [1..4] invalid-cell count per stage
    const int* tile_list;                  // CTA -> tile id (nullptr: identity); used to run interior tiles first
    double* cellS;                         // FlowState.S per cell when the shock detector is on (travels with pushed ghost cells), else nullptr
    const void* tmaps;                     // CUtensorMap[local block] over prim_in (i, j, k, field), or nullptr: stage with cp.async
};

// ghost-cell work lists (indices into the arena)
struct EbCopyItem { int dst, src; };
struct EbReflectItem { int dst, src, fidx, meta; };   // meta = blk*4 + dir
// Species counts for which the thermally-perfect-gas kernels are instantiated: a build-time list
// (make TPG_NSP="2 3 5 7", then for every flux calculator); eb200_init refuses other counts with a message that says so.
// Default: five species for every flux calculator, three as well for ausmdv (each count costs minutes of compile time
// per flux calculator and arithmetic mode).
#ifndef EB_TPG_NSP_LIST
#define EB_TPG_NSP_DEFAULT 1
#if !defined(EB_FLUX) || EB_FLUX == 0
#define EB_TPG_NSP_LIST(X) X(3) X(5)
#else
#define EB_TPG_NSP_LIST(X) X(5)
#endif
#endif
#define EB_P2P_MAXPEERS 64                    /* flags per slot in the region the halo peers of one rank share */
struct EbFillItem { int dst, param; };

// launchers implemented once per arithmetic mode (namespaces eb_strict / eb_fast)
// `which`: bit 0 = launch the uniform-Cartesian kernel, bit 1 = the general-metric kernel
#define EB_DECLARE_LAUNCHERS(ns)                                                                         \
    namespace ns {                                                                                       \
    void launch_flux_update(int flux_calc, int gas_model, const EbParams& P, const EbGas* gas,           \
                            const EbBlockDesc* desc, int nblocks, long long ncta, const EbArena& A,      \
                            const EbStageArgs& S, int which, cudaStream_t st);                           \
    void launch_face_debug(int flux_calc, int gas_model, const EbParams& P, const EbGas* gas,            \
                           const EbArena& A, const double* prim, int nfaces, double* Fout, int* ok_out,  \
                           cudaStream_t st);                                                             \
    void launch_decode(const EbParams& P, int gas_model, const EbGas* gas, const EbBlockDesc& hdesc,     \
                       const double* prim_in, double* prim_out, double* U, int do_encode, int* status,   \
                       cudaStream_t st);                                                                 \
    void launch_signal(const EbParams& P, const EbBlockDesc* descs, int nblocks, long long max_cells,   \
                       const EbArena& A, const double* prim, double dt_current, double cfl_value,        \
                       unsigned long long* red, double* last_signal, cudaStream_t st);                   \
    void launch_ghosts(const EbParams& P, const EbGas* gas, const EbBlockDesc* desc, const EbArena& A, double* prim, \
                       const EbCopyItem* copy, long long ncopy, const EbReflectItem* refl,               \
                       long long nrefl, const EbFillItem* fill, long long nfill, const double* params,   \
                       cudaStream_t st);                                                                 \
    void launch_pack(const EbParams& P, const double* prim, const double* S, const int* idx, long long n, \
                     double* buf, cudaStream_t st);                                                      \
    void launch_unpack(const EbParams& P, double* prim, double* S, const int* idx, long long n,          \
                       const double* buf, cudaStream_t st);                                              \
    void launch_put(const EbParams& P, long long total_dst, const double* prim_src, const double* S_src, double* prim_dst,  \
                    double* S_dst, const int* src_idx, const int* dst_idx, long long n, cudaStream_t st);  \
    void launch_halo_signal(unsigned long long* const* remote_flags, int npeers, unsigned long long seq, int slot, cudaStream_t st); \
    void launch_halo_wait(const unsigned long long* flags, int npeers, unsigned long long seq, int slot, int* status, cudaStream_t st); \
    void launch_detect_shocks_all(const EbParams& P, const EbBlockDesc* d_desc, int nblocks, long long max_positions,  \
                                  const EbArena& A, const double* prim, cudaStream_t st);                \
    }

EB_DECLARE_LAUNCHERS(eb_strict)
EB_DECLARE_LAUNCHERS(eb_fast)
