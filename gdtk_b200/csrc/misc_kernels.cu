// misc_kernels.cu -- the small kernels around the fused flux kernel: encode/decode at
// upload, CFL signal reduction, ghost-cell fill (boundary conditions and same-GPU
// full-face copies) and halo pack/unpack.  Compiled once per arithmetic mode
// (-DEB_NS=eb_strict -fmad=false, -DEB_NS=eb_fast).
#include "device_math.cuh"

namespace EB_NS {

// per-flux launchers live in flux_inst.cu (one object per flux calculator)
#define EB_DECL_FLUX(k)                                                                                        \
    void launch_flux_update_k##k(const EbParams& P, int gas_model, const EbGas* gas, const EbBlockDesc* desc,  \
                                 int nblocks, long long ncta, const EbArena& A, const EbStageArgs& S,          \
                                 int tile_y, int which, cudaStream_t st);
EB_DECL_FLUX(0)
#ifndef EB_DEV_FLUX0_ONLY
EB_DECL_FLUX(1) EB_DECL_FLUX(2) EB_DECL_FLUX(3) EB_DECL_FLUX(4) EB_DECL_FLUX(5)
EB_DECL_FLUX(6) EB_DECL_FLUX(7) EB_DECL_FLUX(8) EB_DECL_FLUX(9) EB_DECL_FLUX(10) EB_DECL_FLUX(11) EB_DECL_FLUX(12)
#endif
#define EB_DECL_DBG(k)                                                                                         \
    void launch_face_debug_k##k(const EbParams& P, int gas_model, const EbGas* gas, const EbArena& A,           \
                                const double* prim, int nfaces, double* Fout, int* ok_out, cudaStream_t st);
EB_DECL_DBG(0)
#ifndef EB_DEV_FLUX0_ONLY
EB_DECL_DBG(1) EB_DECL_DBG(2) EB_DECL_DBG(3) EB_DECL_DBG(4) EB_DECL_DBG(5)
EB_DECL_DBG(6) EB_DECL_DBG(7) EB_DECL_DBG(8) EB_DECL_DBG(9) EB_DECL_DBG(10) EB_DECL_DBG(11) EB_DECL_DBG(12)
#endif

void launch_face_debug(int flux_calc, int gm, const EbParams& P, const EbGas* gas, const EbArena& A, const double* prim,
                       int nfaces, double* Fout, int* ok_out, cudaStream_t st)
{
    switch (flux_calc) {
    case 0: launch_face_debug_k0(P, gm, gas, A, prim, nfaces, Fout, ok_out, st); break;
#ifndef EB_DEV_FLUX0_ONLY
    case 1: launch_face_debug_k1(P, gm, gas, A, prim, nfaces, Fout, ok_out, st); break;
    case 2: launch_face_debug_k2(P, gm, gas, A, prim, nfaces, Fout, ok_out, st); break;
    case 3: launch_face_debug_k3(P, gm, gas, A, prim, nfaces, Fout, ok_out, st); break;
    case 4: launch_face_debug_k4(P, gm, gas, A, prim, nfaces, Fout, ok_out, st); break;
    case 5: launch_face_debug_k5(P, gm, gas, A, prim, nfaces, Fout, ok_out, st); break;
    case 6: launch_face_debug_k6(P, gm, gas, A, prim, nfaces, Fout, ok_out, st); break;
    case 7: launch_face_debug_k7(P, gm, gas, A, prim, nfaces, Fout, ok_out, st); break;
    case 8: launch_face_debug_k8(P, gm, gas, A, prim, nfaces, Fout, ok_out, st); break;
    case 9: launch_face_debug_k9(P, gm, gas, A, prim, nfaces, Fout, ok_out, st); break;
    case 10: launch_face_debug_k10(P, gm, gas, A, prim, nfaces, Fout, ok_out, st); break;
    case 11: launch_face_debug_k11(P, gm, gas, A, prim, nfaces, Fout, ok_out, st); break;
    case 12: launch_face_debug_k12(P, gm, gas, A, prim, nfaces, Fout, ok_out, st); break;
#endif
    }
}

// which: bit 0 Cartesian blocks, bit 1 general-metric blocks, bit 2 force the generic kernel
void launch_flux_update(int flux_calc, int gm, const EbParams& P, const EbGas* gas, const EbBlockDesc* desc, int nblocks,
                        long long ncta, const EbArena& A, const EbStageArgs& S, int which, cudaStream_t st)
{
    const int ty = (which & 4) ? -1 : ((which & 8) ? 1 : 0);      // -1: generic kernel, 1: v2 instead of v3
    which &= 3;
    switch (flux_calc) {
    case 0: launch_flux_update_k0(P, gm, gas, desc, nblocks, ncta, A, S, ty, which, st); break;
#ifndef EB_DEV_FLUX0_ONLY
    case 1: launch_flux_update_k1(P, gm, gas, desc, nblocks, ncta, A, S, ty, which, st); break;
    case 2: launch_flux_update_k2(P, gm, gas, desc, nblocks, ncta, A, S, ty, which, st); break;
    case 3: launch_flux_update_k3(P, gm, gas, desc, nblocks, ncta, A, S, ty, which, st); break;
    case 4: launch_flux_update_k4(P, gm, gas, desc, nblocks, ncta, A, S, ty, which, st); break;
    case 5: launch_flux_update_k5(P, gm, gas, desc, nblocks, ncta, A, S, ty, which, st); break;
    case 6: launch_flux_update_k6(P, gm, gas, desc, nblocks, ncta, A, S, ty, which, st); break;
    case 7: launch_flux_update_k7(P, gm, gas, desc, nblocks, ncta, A, S, ty, which, st); break;
    case 8: launch_flux_update_k8(P, gm, gas, desc, nblocks, ncta, A, S, ty, which, st); break;
    case 9: launch_flux_update_k9(P, gm, gas, desc, nblocks, ncta, A, S, ty, which, st); break;
    case 10: launch_flux_update_k10(P, gm, gas, desc, nblocks, ncta, A, S, ty, which, st); break;
    case 11: launch_flux_update_k11(P, gm, gas, desc, nblocks, ncta, A, S, ty, which, st); break;
    case 12: launch_flux_update_k12(P, gm, gas, desc, nblocks, ncta, A, S, ty, which, st); break;
#endif
    }
}

// ---------------------------------------------------------------------------------------
// encode_conserved (fvcell.d:511-583) + decode_conserved for every interior cell of one
// block: what init_simulation does after reading the flow (simcore.d:325-334), and the
// restore path after a failed step.

template <int DIM, int GASM, int NSP>
__global__ void decode_kernel(const EbParams P, const EbGas* __restrict__ gas, const EbBlockDesc D,
                              const double* __restrict__ prim_in, double* __restrict__ prim_out,
                              double* __restrict__ U, int do_encode, int* status)
{
    typedef Layout<DIM, NSP> Lay;
    constexpr int NCQ = Lay::NCQ;
    const long long n = (long long)D.nic * D.njc * D.nkc;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = (int)(t % D.nic), j = (int)((t / D.nic) % D.njc), k = (int)(t / ((long long)D.nic * D.njc));
    const long long c = D.cell0 + ((long long)(k + D.kg) * D.NJ + (j + EB_NG)) * D.NI + (i + EB_NG);
    const long long total = P.total;
    Prim<NSP> Q;
    load_prim<NSP>(Q, prim_in, total, c);
    if (DIM == 2) Q.vz = 0.0;
    double Uc[NCQ];
    if (do_encode) {
        Uc[Lay::iMass] = Q.rho;
        Uc[Lay::iXMom] = Q.rho * Q.vx; Uc[Lay::iYMom] = Q.rho * Q.vy;
        if (DIM == 3) Uc[Lay::iZMom] = Q.rho * Q.vz;
        double ke = 0.5 * (Q.vx * Q.vx + Q.vy * Q.vy + Q.vz * Q.vz);
        Uc[Lay::iEnergy] = Q.rho * (Q.u + ke);
        if (NSP > 1) {
#pragma unroll
            for (int s = 0; s < NSP; ++s) Uc[Lay::iSpecies + s] = Q.rho * Q.massf[s];
        }
    } else {
#pragma unroll
        for (int q = 0; q < NCQ; ++q) Uc[q] = U[q * total + c];
    }
    bool modified;
    int rc = decode_cell<DIM, GASM, NSP>(P, gas, Uc, Q, modified);
    if (rc) { atomicOr(&status[0], 1); return; }
    store_prim<NSP>(Q, prim_out, total, c);
    if (do_encode || modified) {
#pragma unroll
        for (int q = 0; q < NCQ; ++q) U[q * total + c] = Uc[q];
    }
}

void launch_decode(const EbParams& P, int gas_model, const EbGas* gas, const EbBlockDesc& hdesc,
                   const double* prim_in, double* prim_out, double* U, int do_encode, int* status, cudaStream_t st)
{
    const long long n = (long long)hdesc.nic * hdesc.njc * hdesc.nkc;
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
#define EB_DEC(DIM, GASM, NSP) decode_kernel<DIM, GASM, NSP><<<blocks, threads, 0, st>>>(P, gas, hdesc, prim_in, prim_out, U, do_encode, status)
    if (gas_model == EB200_GAS_IDEAL) { if (P.dims == 3) EB_DEC(3, EB200_GAS_IDEAL, 1); else EB_DEC(2, EB200_GAS_IDEAL, 1); }
#ifndef EB_NO_TPG
#define EB_DEC_NSP(N) else if (P.nsp == N) { if (P.dims == 3) EB_DEC(3, EB200_GAS_THERMALLY_PERFECT, N); else EB_DEC(2, EB200_GAS_THERMALLY_PERFECT, N); }
    EB_TPG_NSP_LIST(EB_DEC_NSP)
#undef EB_DEC_NSP
#endif
#undef EB_DEC
}

// ---------------------------------------------------------------------------------------
// FVCell.signal_frequency (fvcell.d:975-1058, structured, non-stringent, inviscid) and the
// per-block min/max of FluidBlock.determine_time_step_size (fluidblock.d:1031-1069).
// red[0] = min dt_local, red[1] = max cfl_local as bit patterns of non-negative doubles.

template <int DIM>
__global__ void signal_kernel(const EbParams P, const EbBlockDesc* __restrict__ descs, const EbArena A, const double* __restrict__ prim,
                              double dt_current, double cfl_value, unsigned long long* red_all, double* last_all)
{
    // blockIdx.y = local block; red_all[2*b], red_all[2*b+1], last_all[b] belong to block b
    __shared__ EbBlockDesc D;
    {
        const int* src = reinterpret_cast<const int*>(&descs[blockIdx.y]);
        int* dst = reinterpret_cast<int*>(&D);
        for (int m = threadIdx.x; m < (int)(sizeof(EbBlockDesc) / sizeof(int)); m += blockDim.x) dst[m] = src[m];
    }
    __syncthreads();
    unsigned long long* red = red_all + 2 * blockIdx.y;
    double* last_signal = last_all + blockIdx.y;
    const long long n = (long long)D.nic * D.njc * D.nkc;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = P.total;
    double dt_local = __longlong_as_double(0x7ff0000000000000LL), cfl_local = 0.0;
    if (t < n) {
        const int i = (int)(t % D.nic), j = (int)((t / D.nic) % D.njc), k = (int)(t / ((long long)D.nic * D.njc));
        const long long c = D.cell0 + ((long long)(k + D.kg) * D.NJ + (j + EB_NG)) * D.NI + (i + EB_NG);
        const double vx = ldg(prim + 5 * total + c), vy = ldg(prim + 6 * total + c);
        const double vz = (DIM == 3) ? ldg(prim + 7 * total + c) : 0.0;
        const double a = ldg(prim + 4 * total + c);
        double nN[3], nE[3], nT[3] = {0.0, 0.0, 0.0}, lenI, lenJ, lenK = 1.0;
        if (D.cartesian) {
            for (int m = 0; m < 3; ++m) { nE[m] = D.nvec[0][m]; nN[m] = D.nvec[1][m]; nT[m] = D.nvec[2][m]; }
            lenI = D.len[0]; lenJ = D.len[1]; lenK = D.len[2];
        } else {
            const long long cE = c + 1, cN = c + D.stride[1], cT = c + D.stride[2];
            for (int m = 0; m < 3; ++m) {
                nE[m] = ldg(A.face[0] + m * total + cE); nN[m] = ldg(A.face[1] + m * total + cN);
                if (DIM == 3) nT[m] = ldg(A.face[2] + m * total + cT);
            }
            lenI = ldg(A.len[0] + c); lenJ = ldg(A.len[1] + c);
            if (DIM == 3) lenK = ldg(A.len[2] + c);
        }
        double un_N = fabs(vx * nN[0] + vy * nN[1] + vz * nN[2]);
        double un_E = fabs(vx * nE[0] + vy * nE[1] + vz * nE[2]);
        double signal = 0.0;
        signal = fmax(signal, (un_N + a) / lenJ);
        signal = fmax(signal, (un_E + a) / lenI);
        if (DIM == 3) {
            double un_T = fabs(vx * nT[0] + vy * nT[1] + vz * nT[2]);
            signal = fmax(signal, (un_T + a) / lenK);
        }
        cfl_local = dt_current * signal;
        dt_local = cfl_value / signal;
        if (t == n - 1) *last_signal = signal;
        if (!(dt_local == dt_local)) dt_local = __longlong_as_double(0x7ff0000000000000LL);
        if (!(cfl_local == cfl_local) || cfl_local < 0.0) cfl_local = 0.0;
        if (dt_local < 0.0) dt_local = 0.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dt_local = fmin(dt_local, __shfl_down_sync(0xffffffffu, dt_local, o));
        cfl_local = fmax(cfl_local, __shfl_down_sync(0xffffffffu, cfl_local, o));
    }
    __shared__ double s_dt[32], s_cfl[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { s_dt[w] = dt_local; s_cfl[w] = cfl_local; }
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        dt_local = (lane < nw) ? s_dt[lane] : __longlong_as_double(0x7ff0000000000000LL);
        cfl_local = (lane < nw) ? s_cfl[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dt_local = fmin(dt_local, __shfl_down_sync(0xffffffffu, dt_local, o));
            cfl_local = fmax(cfl_local, __shfl_down_sync(0xffffffffu, cfl_local, o));
        }
        if (lane == 0) {
            atomicMin(&red[0], (unsigned long long)__double_as_longlong(dt_local));
            atomicMax(&red[1], (unsigned long long)__double_as_longlong(cfl_local));
        }
    }
}

// all local blocks in one launch; max_cells = cells of the largest block
void launch_signal(const EbParams& P, const EbBlockDesc* descs, int nblocks, long long max_cells, const EbArena& A, const double* prim,
                   double dt_current, double cfl_value, unsigned long long* red, double* last_signal, cudaStream_t st)
{
    const int threads = 256;
    const dim3 grid((unsigned)((max_cells + threads - 1) / threads), (unsigned)nblocks);
    if (P.dims == 3) signal_kernel<3><<<grid, threads, 0, st>>>(P, descs, A, prim, dt_current, cfl_value, red, last_signal);
    else signal_kernel<2><<<grid, threads, 0, st>>>(P, descs, A, prim, dt_current, cfl_value, red, last_signal);
}

// ---------------------------------------------------------------------------------------
// Ghost cells.  copy: same-GPU full-face copies (full_face_copy.d:1891-1899) and zero-order
// extrapolation (extrapolate_copy.d:129-145); reflect: internal_copy_then_reflect.d:111-134
// with ghost_cell.d:33-43; fill: flow_state_copy.d:92-107.

__global__ void ghost_kernel(const EbParams P, const EbGas* __restrict__ gas, const EbBlockDesc* __restrict__ descs, const EbArena A, double* __restrict__ prim,
                             const EbCopyItem* __restrict__ copy, long long ncopy,
                             const EbReflectItem* __restrict__ refl, long long nrefl,
                             const EbFillItem* __restrict__ fill, long long nfill, const double* __restrict__ params)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = P.total;
    const int nprim = P.nprim;
    if (t < ncopy) {
        // sources are interior cells, destinations ghost cells: disjoint, so all loads of a batch can be
        // in flight before the first store (the compiler cannot know that, prim is read and written)
        const EbCopyItem it = copy[t];
        const double* __restrict__ src = prim + it.src;
        for (int v0 = 0; v0 < nprim; v0 += 8) {
            double tmp[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) if (v0 + m < nprim) tmp[m] = __ldg(src + (long long)(v0 + m) * total);
#pragma unroll
            for (int m = 0; m < 8; ++m) if (v0 + m < nprim) prim[(long long)(v0 + m) * total + it.dst] = tmp[m];
        }
        if (P.shock_detect) A.S[it.dst] = A.S[it.src];          // FlowState.S travels with the FlowState
    } else if (t < ncopy + nrefl) {
        const EbReflectItem it = refl[t - ncopy];
        const double* __restrict__ src = prim + it.src;
        if ((it.meta & 3) == 3) {
            // OutFlowBC_FixedP / FixedPT (fixed_p.d, fixed_pt.d): ghost layer n = interior layer n, then p (and T)
            // from the boundary condition and update_thermo_from_pT; fidx = index of { p_outside, T_outside | -1 }
            for (int v = 0; v < nprim; ++v) prim[(long long)v * total + it.dst] = __ldg(src + (long long)v * total);
            const double p_out = params[(long long)it.fidx * nprim], T_out = params[(long long)it.fidx * nprim + 1];
            const double T = (T_out > 0.0) ? T_out : __ldg(src + 3 * total);
            double rho, u;
            if (gas->model == EB200_GAS_IDEAL) {              // ideal_gas.d:89-97
                rho = p_out / (T * gas->Rgas);
                u = gas->Cv * T;
            } else {                                          // perf_gas_mix_eos.d:55-62, therm_perf_gas_mix_eos.d:61-66
                double Rmix = 0.0;
                for (int i = 0; i < gas->nsp; ++i) Rmix += __ldg(src + (long long)(8 + i) * total) * gas->Rsp[i];
                const double denom = Rmix * T;
                rho = p_out / denom;
                const double logT = log(T);
                u = 0.0;
                for (int i = 0; i < gas->nsp; ++i) {
                    double h;
                    cea_h(gas->curves[i], T, logT, h);
                    u += __ldg(src + (long long)(8 + i) * total) * (h - gas->Rsp[i] * T);
                }
            }
            prim[it.dst] = rho; prim[total + it.dst] = u; prim[2 * total + it.dst] = p_out; prim[3 * total + it.dst] = T;
            if (P.shock_detect) A.S[it.dst] = A.S[it.src];
            return;
        }
        double x = __ldg(src + 5 * total), y = __ldg(src + 6 * total), z = __ldg(src + 7 * total);
        {
            double tmp[5];
#pragma unroll
            for (int m = 0; m < 5; ++m) tmp[m] = __ldg(src + (long long)m * total);
#pragma unroll
            for (int m = 0; m < 5; ++m) prim[(long long)m * total + it.dst] = tmp[m];
        }
        for (int v = 8; v < nprim; ++v) prim[v * total + it.dst] = __ldg(src + (long long)v * total);
        const int blk = it.meta >> 2, d = it.meta & 3;
        const EbBlockDesc& D = descs[blk];
        if (D.cartesian) {
            EbAxisFrame f = D.fr[d];
            axis_to_local(f, x, y, z); x = -x; axis_to_global(f, x, y, z);
        } else {
            Frame f;
            if (P.dims == 3) { load_frame<3>(f, A.face[d], total, it.fidx); to_local<3>(f, x, y, z); x = -x; to_global<3>(f, x, y, z); }
            else { load_frame<2>(f, A.face[d], total, it.fidx); to_local<2>(f, x, y, z); x = -x; to_global<2>(f, x, y, z); }
        }
        prim[5 * total + it.dst] = x; prim[6 * total + it.dst] = y; prim[7 * total + it.dst] = z;
        if (P.shock_detect) A.S[it.dst] = A.S[it.src];
    } else if (t < ncopy + nrefl + nfill) {
        const EbFillItem it = fill[t - ncopy - nrefl];
        for (int v = 0; v < nprim; ++v) prim[v * total + it.dst] = params[(long long)it.param * nprim + v];
        if (P.shock_detect) A.S[it.dst] = 0.0;                  // S of a FlowState made in the job script
    }
}

void launch_ghosts(const EbParams& P, const EbGas* gas, const EbBlockDesc* desc, const EbArena& A, double* prim,
                   const EbCopyItem* copy, long long ncopy, const EbReflectItem* refl, long long nrefl,
                   const EbFillItem* fill, long long nfill, const double* params, cudaStream_t st)
{
    const long long n = ncopy + nrefl + nfill;
    if (n == 0) return;
    const int threads = 256;
    ghost_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(P, gas, desc, A, prim, copy, ncopy, refl, nrefl, fill, nfill, params);
}

// halo pack / unpack for blocks owned by other processes: buf[v*n + t]; with the shock detector on,
// FlowState.S is variable number nprim
__global__ void pack_kernel(long long total, int nprim, const double* __restrict__ prim, const double* __restrict__ S,
                            const int* __restrict__ idx, long long n, double* __restrict__ buf)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long c = idx[t];
    for (int v = 0; v < nprim; ++v) buf[v * n + t] = prim[v * total + c];
    if (S) buf[(long long)nprim * n + t] = S[c];
}
__global__ void unpack_kernel(long long total, int nprim, double* __restrict__ prim, double* __restrict__ S,
                              const int* __restrict__ idx, long long n, const double* __restrict__ buf)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long c = idx[t];
    for (int v = 0; v < nprim; ++v) prim[v * total + c] = buf[v * n + t];
    if (S) S[c] = buf[(long long)nprim * n + t];
}
// Halo over NVLink without staging buffers: the interior cells a neighbouring rank needs go straight into that
// rank's ghost cells (its arena, mapped through CUDA IPC); total_dst is the field stride of the remote arena.
__global__ void put_kernel(long long total_src, long long total_dst, int nprim, const double* __restrict__ prim_src,
                           const double* __restrict__ S_src, double* __restrict__ prim_dst, double* __restrict__ S_dst,
                           const int* __restrict__ src_idx, const int* __restrict__ dst_idx, long long n)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long cs = src_idx[t], cd = dst_idx[t];
    for (int v = 0; v < nprim; ++v) prim_dst[v * total_dst + cd] = prim_src[v * total_src + cs];
    if (S_src) S_dst[cd] = S_src[cs];
}
// After the puts of one exchange (stream order): tell every peer that exchange `seq` has landed.
// slot 0 = "my puts of exchange seq are in your memory", slot 1 = "I am done reading what exchange seq - 1 brought"
// (raised before the puts: FlowState.S has one buffer only, so a peer must not overwrite it while the previous
// stage still reads it; the three FlowState buffers need no such handshake).
__global__ void halo_signal_kernel(unsigned long long* const* __restrict__ remote_flags, int npeers, unsigned long long seq, int slot)
{
    const int p = threadIdx.x;
    if (p >= npeers) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote_flags[p] + slot * EB_P2P_MAXPEERS), "l"(seq) : "memory");
}
// Before the tiles that read those ghost cells: wait until every peer has signalled exchange `seq`.
// (gives up after ~10 s of GPU clock and raises bit 1 of the step status: a peer that never signals -- a rank that
//  died or took another number of steps -- must not hang the device)
__global__ void halo_wait_kernel(const unsigned long long* __restrict__ flags, int npeers, unsigned long long seq, int slot, int* status)
{
    const int p = threadIdx.x;
    if (p < npeers) {
        unsigned long long v;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + slot * EB_P2P_MAXPEERS + p) : "memory");
            if (v < seq && clock64() - t0 > 20000000000LL) { atomicOr(status, 3); break; }
        } while (v < seq);
    }
    __threadfence_system();
}
void launch_put(const EbParams& P, long long total_dst, const double* prim_src, const double* S_src, double* prim_dst, double* S_dst,
                const int* src_idx, const int* dst_idx, long long n, cudaStream_t st)
{
    if (n > 0) put_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P.total, total_dst, P.nprim, prim_src, P.shock_detect ? S_src : nullptr,
                                                                      prim_dst, S_dst, src_idx, dst_idx, n);
}
void launch_halo_signal(unsigned long long* const* remote_flags, int npeers, unsigned long long seq, int slot, cudaStream_t st)
{
    halo_signal_kernel<<<1, 64, 0, st>>>(remote_flags, npeers, seq, slot);
}
void launch_halo_wait(const unsigned long long* flags, int npeers, unsigned long long seq, int slot, int* status, cudaStream_t st)
{
    halo_wait_kernel<<<1, 64, 0, st>>>(flags, npeers, seq, slot, status);
}

void launch_pack(const EbParams& P, const double* prim, const double* S, const int* idx, long long n, double* buf, cudaStream_t st)
{
    if (n > 0) pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P.total, P.nprim, prim, P.shock_detect ? S : nullptr, idx, n, buf);
}
void launch_unpack(const EbParams& P, double* prim, double* S, const int* idx, long long n, const double* buf, cudaStream_t st)
{
    if (n > 0) unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P.total, P.nprim, prim, P.shock_detect ? S : nullptr, idx, n, buf);
}

// ---------------------------------------------------------------------------------------
// detect_shocks (simcore_gasdynamic_step.d:3197-3224) with shock_detector_smoothing = 0, one block:
//   detect_shock_points  (fluidblock.d:479-520): PJ_ShockDetector (shockdetectors.d:22-93) on every face
//   shock_faces_to_cells (:573-583)
//   enforce_strict_shock_detector (:585-605); ghost-cell S is what the last ghost fill copied.
// One thread per position of the block extended by one cell on the plus side of every direction.

// PJ_ShockDetector (shockdetectors.d:22-93) on values in registers.
// side: 0 = the face has a cell on both sides; 1 = only the left cell exists (a wall without ghost-cell data on the
// right), 2 = only the right cell: shockdetectors.d:48-84, the gas velocity relative to the wall (gvel = 0)
struct PjCell { double vx, vy, vz, a; };
struct PjFrame { double n[3], t1[3], t2[3]; };

template <int DIM>
__device__ __forceinline__ PjCell pj_load_cell(const double* __restrict__ prim, long long total, long long c)
{
    PjCell q;
    q.vx = ldg(prim + 5 * total + c); q.vy = ldg(prim + 6 * total + c);
    q.vz = (DIM == 3) ? ldg(prim + 7 * total + c) : 0.0;
    q.a = ldg(prim + 4 * total + c);
    return q;
}

// the frame of the face on the minus-d side of cell c: the exact +-1/0 vectors of a uniform-Cartesian block from its
// descriptor (compares against the loop index: no dynamically indexed local arrays), else the stored (n, t1, t2)
__device__ __forceinline__ void pj_load_frame(const EbBlockDesc& D, const EbArena& A, long long total, int d, long long c, PjFrame& f)
{
    if (D.cartesian) {
        const int p0 = D.fr[d].perm[0], p1 = D.fr[d].perm[1], p2 = D.fr[d].perm[2];
        const double s0 = D.fr[d].neg[0] ? -1.0 : 1.0, s1 = D.fr[d].neg[1] ? -1.0 : 1.0, s2 = D.fr[d].neg[2] ? -1.0 : 1.0;
#pragma unroll
        for (int m = 0; m < 3; ++m) { f.n[m] = (p0 == m) ? s0 : 0.0; f.t1[m] = (p1 == m) ? s1 : 0.0; f.t2[m] = (p2 == m) ? s2 : 0.0; }
    } else {
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            f.n[m] = ldg(A.face[d] + m * total + c); f.t1[m] = ldg(A.face[d] + (3 + m) * total + c);
            f.t2[m] = ldg(A.face[d] + (6 + m) * total + c);
        }
    }
}

__device__ __forceinline__ double pj_detector(const EbParams& P, const PjCell& cL, const PjCell& cR, const PjFrame& f, int side)
{
    if (side != 0) {
        const PjCell& q = (side == 1) ? cL : cR;
        const double u = q.vx * f.n[0] + q.vy * f.n[1] + q.vz * f.n[2];
        const double v = q.vx * f.t1[0] + q.vy * f.t1[1] + q.vz * f.t1[2];
        const double w = q.vx * f.t2[0] + q.vy * f.t2[1] + q.vz * f.t2[2];
#ifdef EB_FAST_MATH
        // sound speeds are positive: compare the numerators with tolerance x denominator, no division
        const double un = (side == 1) ? -u : u;
        return ((fmax(fabs(v), fabs(w)) < P.shear_tol * q.a) && (un < P.comp_tol * q.a)) ? 1.0 : 0.0;
#else
        const double comp = (side == 1) ? ((-u) / q.a) : (u / q.a);
        const double shear = fmax(fabs(v) / q.a, fabs(w) / q.a);
        return ((shear < P.shear_tol) && (comp < P.comp_tol)) ? 1.0 : 0.0;
#endif
    }
    const double uL = cL.vx * f.n[0] + cL.vy * f.n[1] + cL.vz * f.n[2];
    const double uR = cR.vx * f.n[0] + cR.vy * f.n[1] + cR.vz * f.n[2];
    const double a_min = (cL.a < cR.a) ? cL.a : cR.a;
#ifdef EB_FAST_MATH
    const double comp = 0.0;
#else
    const double comp = ((uR - uL) / a_min);
#endif
    const double vL = cL.vx * f.t1[0] + cL.vy * f.t1[1] + cL.vz * f.t1[2], vR = cR.vx * f.t1[0] + cR.vy * f.t1[1] + cR.vz * f.t1[2];
    const double wL = cL.vx * f.t2[0] + cL.vy * f.t2[1] + cL.vz * f.t2[2], wR = cR.vx * f.t2[0] + cR.vy * f.t2[1] + cR.vz * f.t2[2];
    const double sound_speed = 0.5 * (cL.a + cR.a);
#ifdef EB_FAST_MATH
    (void)comp;
    return ((fmax(fabs(vL - vR), fabs(wL - wR)) < P.shear_tol * sound_speed) && ((uR - uL) < P.comp_tol * a_min)) ? 1.0 : 0.0;
#else
    const double shear_y = fabs(vL - vR) / sound_speed;
    const double shear_z = fabs(wL - wR) / sound_speed;
    const double shear = fmax(shear_y, shear_z);
    return ((shear < P.shear_tol) && (comp < P.comp_tol)) ? 1.0 : 0.0;
#endif
}

// pass 0: the detector on the minus-side faces of every position AND, in the same launch, shock_faces_to_cells for
// interior cells: the detector of a cell's plus-side faces is evaluated a second time by the cell itself (a pure
// function of the same FlowStates: the same value the owner of that face stores), which saves a pass over the block.
// A thread loads its own cell once and one neighbour per side and direction.
// pass 2: enforce_strict_shock_detector.
template <int DIM>
__global__ void shock_kernel(const EbParams P, const EbBlockDesc* __restrict__ descs, const EbArena A, const double* __restrict__ prim, int pass)
{
    // blockIdx.y = local block: all blocks of the process in one launch
    __shared__ EbBlockDesc D;
    {
        const int* src = reinterpret_cast<const int*>(&descs[blockIdx.y]);
        int* dst = reinterpret_cast<int*>(&D);
        for (int m = threadIdx.x; m < (int)(sizeof(EbBlockDesc) / sizeof(int)); m += blockDim.x) dst[m] = src[m];
    }
    __syncthreads();
    const int ei = D.nic + 1, ej = D.njc + 1, ek = (DIM == 3) ? D.nkc + 1 : 1;
    const long long n = (long long)ei * ej * ek;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = (int)(t % ei), j = (int)((t / ei) % ej), k = (int)(t / ((long long)ei * ej));
    const long long c = D.cell0 + ((long long)(k + D.kg) * D.NJ + (j + EB_NG)) * D.NI + (i + EB_NG);
    const long long total = P.total;
    const bool in_i = i < D.nic, in_j = j < D.njc, in_k = (DIM == 3) ? (k < D.nkc) : true;
    const bool interior = in_i && in_j && in_k;
    double Smax = 0.0;                                     // iface order W,E,S,N,B,T; max is order-independent
    PjCell own;
    if (pass == 0) own = pj_load_cell<DIM>(prim, total, c);
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        // the face on the minus-d side of this position exists if the other two indices are interior
        const bool face_ok = (d == 0) ? (in_j && in_k) : ((d == 1) ? (in_i && in_k) : (in_i && in_j));
        if (!face_ok) continue;
        const long long st = D.stride[d];
        // faces on a boundary without ghost-cell data have a cell on one side only
        const int idx = (d == 0) ? i : ((d == 1) ? j : k), nd = (d == 0) ? D.nic : ((d == 1) ? D.njc : D.nkc);
        const bool wallL = (D.noghost_faces >> (2 * d)) & 1, wallR = (D.noghost_faces >> (2 * d + 1)) & 1;
        const int side = (wallL && idx == 0) ? 2 : ((wallR && idx == nd) ? 1 : 0);
        if (pass == 0) {
            PjFrame f;
            pj_load_frame(D, A, total, d, c, f);
            PjCell below = own;
            if (side != 2) below = pj_load_cell<DIM>(prim, total, c - st);
            const double Sf = pj_detector(P, below, own, f, side);
            A.Sf[d][c] = Sf;
            if (interior) {
                const int side1 = (wallR && idx + 1 == nd) ? 1 : 0;
                PjCell above = own;
                if (side1 != 1) above = pj_load_cell<DIM>(prim, total, c + st);
                if (!D.cartesian) pj_load_frame(D, A, total, d, c + st, f);
                Smax = fmax(Smax, Sf);
                Smax = fmax(Smax, pj_detector(P, own, above, f, side1));
            }
        } else {
            double Sf = A.Sf[d][c];
            if (Sf > 0.0 || (side != 2 && A.S[c - st] > 0.0) || (side != 1 && A.S[c] > 0.0)) Sf = 1.0;     // fluidblock.d:585-605: cells that exist
            A.Sf[d][c] = Sf;
        }
    }
    if (pass == 0 && interior) A.S[c] = Smax;
}

void launch_detect_shocks_all(const EbParams& P, const EbBlockDesc* d_desc, int nblocks, long long max_positions, const EbArena& A,
                              const double* prim, cudaStream_t st)
{
    const int threads = 256;
    const dim3 grid((unsigned)((max_positions + threads - 1) / threads), (unsigned)nblocks);
    for (int pass = 0; pass <= (P.strict_shock ? 2 : 0); pass += 2) {
        if (P.dims == 3) shock_kernel<3><<<grid, threads, 0, st>>>(P, d_desc, A, prim, pass);
        else shock_kernel<2><<<grid, threads, 0, st>>>(P, d_desc, A, prim, pass);
    }
}

}  // namespace EB_NS
