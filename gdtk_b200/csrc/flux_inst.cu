// flux_inst.cu -- instantiates the fused flux+update kernel for ONE flux calculator
// (-DEB_FLUX=0..10, enum eb200_flux_calculator) in ONE arithmetic mode (-DEB_NS=...).
#ifndef EB_FLUX
#error "EB_FLUX must be defined"
#endif
#ifdef EB_NO_TPG
#define EB_FLUX_HAS_TPG 0
#else
#define EB_FLUX_HAS_TPG (EB_FLUX != 7 && EB_FLUX != 8 && EB_FLUX != 10 && EB_FLUX != 11 && EB_FLUX != 12)   /* hllc, hlle2 with several species are not on this path;
                                                                          of the adaptive ones only the default is built for TPG */
#endif
#include "flux_kernel.cuh"
#include "flux_kernel_v2.cuh"
#include "flux_kernel_v3.cuh"
#if EB_FLUX_HAS_TPG
#include "flux_kernel_tp.cuh"
#endif

#define EB_CAT2(a, b) a##b
#define EB_CAT(a, b) EB_CAT2(a, b)

namespace EB_NS {
void EB_CAT(launch_flux_update_k, EB_FLUX)(const EbParams& P, int gas_model, const EbGas* gas, const EbBlockDesc* desc,
                                          int nblocks, long long ncta, const EbArena& A, const EbStageArgs& S,
                                          int tile_y, int which, cudaStream_t st)
{
    // tile_y < 0: force the generic kernel (A/B testing); otherwise the tuned kernel handles the reference's
    // default configuration (ideal gas, second-order reconstruction, limiter on)
    const bool tuned = (tile_y >= 0) && gas_model == EB200_GAS_IDEAL && P.interpolation_order == 2 && P.apply_limiter != 0 &&
                       P.thermo_interp == EB200_INTERP_RHOU;
#if EB_FLUX_HAS_TPG
    // thermally perfect mixtures on uniform-Cartesian blocks: the cell-centred kernel of flux_kernel_tp.cuh
    if (tile_y >= 0 && (which & 1) && gas_model == EB200_GAS_THERMALLY_PERFECT && P.interpolation_order == 2 && P.apply_limiter != 0 &&
        P.thermo_interp == EB200_INTERP_RHOU) {
        if (launch_flux_update_tp_impl<EB_FLUX>(P, gas, desc, nblocks, ncta, A, S, st)) which &= ~1;
        if (!which) return;
    }
#endif
    if (tuned) {
        // uniform-Cartesian blocks whose tiles can be staged by TMA: the cell-centred kernel (tile_y == 1 keeps v2, A/B testing)
        if ((which & 1) && tile_y == 0 && S.tmaps != nullptr) {
            launch_flux_update_v3_impl<EB_FLUX>(P, gas, desc, nblocks, ncta, A, S, st);
            which &= ~1;
        }
        if (which) launch_flux_update_v2_impl<EB_FLUX>(P, gas, desc, nblocks, ncta, A, S, which, st);
    } else launch_flux_update_impl<EB_FLUX>(P, gas_model, gas, desc, nblocks, ncta, A, S, tile_y, which, st);
}
void EB_CAT(launch_face_debug_k, EB_FLUX)(const EbParams& P, int gas_model, const EbGas* gas, const EbArena& A,
                                         const double* prim, int nfaces, double* Fout, int* ok_out, cudaStream_t st)
{
    launch_face_debug_impl<EB_FLUX>(P, gas_model, gas, A, prim, nfaces, Fout, ok_out, st);
}
}  // namespace EB_NS
