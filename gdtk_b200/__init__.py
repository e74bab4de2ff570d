"""gdtk_b200: B200-native explicit structured-block update for Eilmer (gdtk-uq/gdtk).

Only what the hot path needs: ``csrc/`` (CUDA kernels + the C ABI of include/eb200.h)
and the host-side mirror of the reference's job-script / time-marching interface.
"""
from . import _abi  # noqa: F401
from .gas import IdealGas, ThermallyPerfectGas, FlowState, set_gas_model  # noqa: F401
from .sim import (Config, FluidBlock, Simulation, WallBC_WithSlip, WallBC_WithSlip1, UserDefinedBC, InFlowBC_Supersonic,  # noqa: F401
                  OutFlowBC_Simple, OutFlowBC_SimpleFlux, OutFlowBC_SimpleExtrapolate, OutFlowBC_FixedP, OutFlowBC_FixedPT,
                  ExchangeBC_FullFace, identify_block_connections)
