/*
 * eb200.h -- C ABI of the B200-native explicit structured-block update for Eilmer.
 *
 * This is the drop-in boundary for ONE hot path of gdtk-uq/gdtk (Eilmer 4):
 *
 *   gasdynamic_explicit_increment_with_fixed_grid()   src/eilmer/simcore_gasdynamic_step.d:906-1575
 *   determine_time_step_size() (per-block part)       src/eilmer/simcore_gasdynamic_step.d:79-111,
 *                                                     src/eilmer/fluidblock.d:987-1084
 *   exchange_ghost_cell_boundary_data()               src/eilmer/simcore_exchange.d:96-135
 *
 * The reference has no plugin/FFI interface for this path; the seam is the
 * `gasdynamic_step` function pointer selected in integrate_in_time()
 * (src/eilmer/simcore.d:968-985, called at :1069).  The calling conventions
 * below follow the reference's own two C-boundary precedents:
 *   - extern(C) kernel_launcher(double*, size_t ...)   src/eilmer/cuda_gpu_chem.d:22-26
 *   - src/gas/gas_cwrap.d:35-93: integer handles, return 0 / negative,
 *     message kept for the caller, no exception crosses the ABI, caller owns
 *     every host buffer.
 *
 * All arithmetic is IEEE double (`number = double`, src/nm/number.d:11-15).
 * No torch types appear here: plain pointers, ints and doubles only.
 *
 * Array layout ("padded block layout")
 * ------------------------------------
 * A block has nic x njc x nkc interior cells (nkc = 1 in 2D) and
 * EB200_NGHOST = 2 ghost layers (n_ghost_cell_layers, src/eilmer/globalconfig.d:975)
 * in every index direction that exists:
 *     NI = nic + 4, NJ = njc + 4, NK = (3D) nkc + 4 : (2D) 1
 *     padded cell index  c = (k*NJ + j)*NI + i,   interior i in [2, 2+nic) ...
 * Every per-cell array passed across this ABI has NI*NJ*NK doubles in this
 * order (the interior sub-box is the reference's cell order
 * (k*njc + j)*nic + i of src/eilmer/sfluidblock.d:299-312).
 * Per-face arrays of index direction d (d = 0,1,2 for i,j,k faces) also have
 * NI*NJ*NK entries: entry c holds the face on the MINUS-d side of padded cell
 * c, i.e. the face whose right_cells[0] is cell c (src/eilmer/sfluidblock.d:544-612).
 */
#ifndef EB200_H
#define EB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define EB200_NGHOST 2
#define EB200_MAX_SPECIES 8
#define EB200_MAX_SEGMENTS 4

/* Face names and order: src/geom/elements/nomenclature.d:12-19 */
enum eb200_face {
    EB200_WEST = 0, EB200_EAST = 1, EB200_SOUTH = 2, EB200_NORTH = 3,
    EB200_BOTTOM = 4, EB200_TOP = 5
};

/* Flux calculators on this path; names are Eilmer's config.flux_calculator
 * strings (src/eilmer/globalconfig.d:293-345). */
enum eb200_flux_calculator {
    EB200_FLUX_AUSMDV = 0,        /* fluxcalc.d:474-647   */
    EB200_FLUX_HANEL = 1,         /* fluxcalc.d:1028-1128 */
    EB200_FLUX_LDFSS0 = 2,        /* fluxcalc.d:819-914   */
    EB200_FLUX_LDFSS2 = 3,        /* fluxcalc.d:917-1025  */
    EB200_FLUX_AUSM_PLUS_UP = 4,  /* fluxcalc.d:1415-1602 */
    EB200_FLUX_ROE = 5,           /* fluxcalc.d:1929-2120 */
    /* adaptive calculators (fluxcalc.d:1315-1412): the first scheme where the shock detector marks
     * the face (IFace.fs.S = 1), the second elsewhere; they switch the shock detector on
     * (configCheckPoint1, globalconfig.d:2676-2690; detect_shocks, simcore_gasdynamic_step.d:3197-3224,
     * PJ_ShockDetector shockdetectors.d:22-93, fluidblock.d:479-605) */
    EB200_FLUX_ADAPTIVE_HANEL_AUSMDV = 6,        /* the reference's default (globalconfig.d:1031) */
    EB200_FLUX_ADAPTIVE_HANEL_AUSM_PLUS_UP = 7,
    EB200_FLUX_ADAPTIVE_LDFSS0_LDFSS2 = 8,
    EB200_FLUX_EFM = 9,                          /* fluxcalc.d:1131-1312 (equilibrium flux method) */
    EB200_FLUX_ADAPTIVE_EFM_AUSMDV = 10,         /* config.flux_calculator = "adaptive" (globalconfig.d:333) */
    EB200_FLUX_HLLC = 11,                        /* fluxcalc.d:650-816 (single-species gas on this path, like roe) */
    EB200_FLUX_HLLE2 = 12                        /* fluxcalc.d:1779-1926, with the reference's y/z momentum slip at :1901 */
};

/* config.thermo_interpolator (globalconfig.d:1073, onedinterp.d:771-978) */
enum eb200_thermo_interpolator {
    EB200_INTERP_RHOU = 0, EB200_INTERP_PT = 1, EB200_INTERP_RHOP = 2, EB200_INTERP_RHOT = 3
};

/* config.gasdynamic_update_scheme (src/eilmer/globalconfig.d:126-200);
 * gamma tables at simcore_gasdynamic_step.d:1235-1395. */
enum eb200_update_scheme {
    EB200_UPDATE_EULER = 0,
    EB200_UPDATE_PC = 1,          /* predictor-corrector, the default */
    EB200_UPDATE_MIDPOINT = 2,
    EB200_UPDATE_CLASSIC_RK3 = 3,
    EB200_UPDATE_TVD_RK3 = 4,
    EB200_UPDATE_DENMAN_RK3 = 5,  /* every stage continues from the U of the stage before (:1303, :1352) */
    EB200_UPDATE_CLASSIC_RK4 = 6  /* four stages */
};

enum eb200_gas_model {
    EB200_GAS_IDEAL = 0,              /* src/gas/ideal_gas.d */
    EB200_GAS_THERMALLY_PERFECT = 1   /* src/gas/therm_perf_gas.d */
};

/* Boundary conditions (user-level names in src/eilmer/bc.lua). */
enum eb200_bc_kind {
    /* WallBC_WithSlip (bc.lua:757-781): GhostCellInternalCopyThenReflect
     * bc/ghost_cell_effect/internal_copy_then_reflect.d:111-134 */
    EB200_BC_WALL_WITH_SLIP = 0,
    /* InFlowBC_Supersonic (bc.lua:1320): GhostCellFlowStateCopy
     * bc/ghost_cell_effect/flow_state_copy.d:92-107; params = FlowState */
    EB200_BC_INFLOW_SUPERSONIC = 1,
    /* OutFlowBC_SimpleExtrapolate (bc.lua:1580), xOrder = 0:
     * bc/ghost_cell_effect/extrapolate_copy.d:124-145 */
    EB200_BC_OUTFLOW_SIMPLE_EXTRAPOLATE = 2,
    /* OutFlowBC_Simple = OutFlowBC_SimpleFlux (bc.lua:1601-1625): ExtrapolateCopy
     * ghost fill + BFE_SimpleOutflowFlux bc/boundary_flux_effect.d:573-643,
     * convective_flux_computed_in_bc = true (sfluidblock.d:2240) */
    EB200_BC_OUTFLOW_SIMPLE_FLUX = 3,
    /* ExchangeBC_FullFace (bc.lua:1672): GhostCellFullFaceCopy
     * bc/ghost_cell_effect/full_face_copy.d:1657-1903 */
    EB200_BC_EXCHANGE_FULL_FACE = 4,
    /* OutFlowBC_FixedP (bc.lua:1627-1647): ExtrapolateCopy, then GhostCellFixedP
     * (bc/ghost_cell_effect/fixed_p.d): ghost cell n = interior cell n with p = p_outside and
     * update_thermo_from_pT; params = { p_outside } */
    EB200_BC_OUTFLOW_FIXED_P = 5,
    /* OutFlowBC_FixedPT (bc.lua:1649-1670), GhostCellFixedPT (fixed_pt.d): also T = T_outside;
     * params = { p_outside, T_outside } */
    EB200_BC_OUTFLOW_FIXED_PT = 6,
    /* WallBC_WithSlip1 (bc.lua:783-806): ghost_cell_data_available = false, no ghost-cell effect.  The wall face has
     * no left (or right) cells and takes compute_flux_at_left_wall / _right_wall (fluxcalc.d:41-51, 187-385); it and
     * the next face in are reconstructed from the one-sided stencils l0r2 / l2r0 and l1r2 / l2r1
     * (onedinterp.d:117-273, 386-485, 991-1838).  A job with such a wall on any block runs the generic kernel on
     * every block and must be initialised, on every process of the job, with eb200_config.reserved_i[0] = 1 (no
     * uniform-Cartesian fast path: the one-sided code is built for the general-metric kernel only; eb200_commit
     * says so otherwise) and reserved_i[1] = 1 (a process that owns no such wall must pick the same kernel as the
     * others for N processes to reproduce one). */
    EB200_BC_WALL_WITH_SLIP1 = 7,
    /* UserDefinedBC (bc.lua) whose ghostCells() function does not depend on time or on the flow
     * (bc/user_defined_effects.d:237-310 evaluates it at the ghost-cell centres every stage and gets the same
     * FlowStates every time): the caller evaluates it once.  params = one FlowState (nprim doubles, the order of
     * upload_flow) per ghost cell, nparams = 2 * n1 * n2 * nprim with the ghost cell of layer l (0 = next to the
     * face) behind the face cell (a1, a2) at ((a2 * n1 + a1) * 2 + l) * nprim; a1 runs along direction (d + 1) % 3,
     * a2 along (d + 2) % 3, d = direction of the face normal (0 = i). */
    EB200_BC_GHOST_PROFILE = 8
};

/* Order of the primitive (FlowState) variables in upload/download and in the
 * FlowState parameter of EB200_BC_INFLOW_SUPERSONIC:
 *   rho, u (internal energy), p, T, a, velx, vely, velz,
 *   then, only if n_species > 1:  massf[0..nsp), rho_s[0..nsp)
 * (src/eilmer/flowstate.d:49-63, src/gas/gas_state.d:18-49). */
#define EB200_PRIM_RHO 0
#define EB200_PRIM_U 1
#define EB200_PRIM_P 2
#define EB200_PRIM_T 3
#define EB200_PRIM_A 4
#define EB200_PRIM_VELX 5
#define EB200_PRIM_VELY 6
#define EB200_PRIM_VELZ 7
#define EB200_NPRIM_BASE 8
#define EB200_NPRIM_SHORT 5   /* eb200_upload_flow, single-species gas: rho, u, velx, vely, velz only */

/* NASA/CEA thermo curve of one species
 * (src/gas/thermo/cea_thermo_curves.d:24-55, data e.g.
 * src/gas/sample-data/therm-perf-5-species-air.lua). */
typedef struct eb200_species {
    double mol_mass;                         /* kg/mol; R_s = 8.31451/M (physical_constants.d:13) */
    int nsegments;
    double T_break_points[EB200_MAX_SEGMENTS + 1];
    double T_blend_ranges[EB200_MAX_SEGMENTS];
    double coeffs[EB200_MAX_SEGMENTS][9];
} eb200_species;

/* Everything the path reads from GlobalConfig/LocalConfig
 * (src/eilmer/globalconfig.d:923-1130,1262-1295; defaults in comments). */
typedef struct eb200_config {
    int dimensions;                 /* 2 or 3 */
    int axisymmetric;               /* 0; 2D only */
    int gas_model;                  /* eb200_gas_model */
    int n_species;                  /* 1 for ideal gas */
    int flux_calculator;            /* eb200_flux_calculator */
    int interpolation_order;        /* 2 (1 = copy cell values) */
    int apply_limiter;              /* 1 */
    int extrema_clipping;           /* 1 */
    int interpolate_in_local_frame; /* 1 */
    int apply_entropy_fix;          /* 1 (AUSMDV) */
    int update_scheme;              /* EB200_UPDATE_PC */
    int max_invalid_cells;          /* 0; compared with the invalid cells of ALL local blocks of a stage together (the
                                       reference counts per block, simcore_gasdynamic_step.d:1442: stricter when > 0) */
    int strict_fp;                  /* 1: kernels built with FMA contraction off, bit-comparable
                                       with the reference's generic x86-64 (no FMA) arithmetic;
                                       0: fused multiply-add allowed (throughput build) */
    int rank;                       /* this process's rank (0 when single process) */
    int device;                     /* CUDA device ordinal used by this process */
    int reserved_i[4];              /* testing knobs: [0] != 0 never use the uniform-Cartesian fast path;
                                       [1] == 1 always use the generic flux kernel, == 2 the face-centred tuned
                                       kernel where the cell-centred one would run;
                                       [2] != 0 stage tiles with cp.async even where TMA could be used;
                                       [3] != 0 fill every ghost cell with the ghost-cell kernel (no stores into
                                       neighbouring blocks from the flux kernel) */
    int thermo_interpolator;        /* eb200_thermo_interpolator: which pair of thermodynamic variables is
                                       reconstructed (config.thermo_interpolator, onedinterp.d:771-978); 0 = rhou */
    double epsilon_van_albada;      /* 1e-12 */
    double M_inf;                   /* 0.01 (ausm_plus_up) */
    double max_velocity;            /* flowstate_limits: 30000 */
    double max_temp;                /* 50000 */
    double min_temp;                /* 0 */
    double suggested_low_T_value;   /* 200; used when ignore_low_T_thermo_update_failure */
    int ignore_low_T_thermo_update_failure; /* 1 */
    int strict_shock_detector;      /* 1 (globalconfig.d:1064); only read by the adaptive flux calculators */
    /* Ideal gas (src/gas/ideal_gas.d:43-69) */
    double ideal_mol_mass;          /* kg/mol */
    double ideal_gamma;
    double compression_tolerance;   /* -0.30 (globalconfig.d:1123), PJ shock detector */
    double shear_tolerance;         /* 0.20 (globalconfig.d:1108) */
    double solver_variant;          /* 0: Eilmer 4 (src/eilmer), every citation in this header; 1: the formulas of Eilmer 5
                                       (src/lmr) where they change numbers on this path: van Albada's epsilon scaled by the
                                       local magnitude and cell spacing (lmr/onedinterp.d:147-151), the smooth maximum for
                                       AUSMDV's common sound speed (lmr/fluxcalc.d:553-561), no fall-back to the cell state
                                       when a reconstructed state has no thermodynamic closure (lmr/onedinterp.d:296-360:
                                       the step fails instead).  Single-species gas, generic kernel. */
    double reserved_d[3];
    /* Thermally perfect gas mixture */
    eb200_species species[EB200_MAX_SPECIES];
} eb200_config;

/* Callback used when blocks owned by other processes are connected to local
 * blocks (one process per GPU).  The library packs every outgoing face into
 * one device buffer per peer rank, calls this function once per exchange, and
 * unpacks after it returns.  `send`/`recv` are DEVICE pointers; counts are in
 * doubles.  The callee moves send[p] to rank peers[p] and fills recv[p] from
 * rank peers[p] (e.g. NCCL send/recv on the stream `cuda_stream`).
 * Replaces MPI_Irecv/MPI_Send/MPI_Wait of full_face_copy.d:1681,1803-1842. */
typedef int (*eb200_exchange_fn)(void* user, int npeers, const int* peers,
                                 double* const* send, const long long* send_count,
                                 double* const* recv, const long long* recv_count,
                                 void* cuda_stream);

/* ---- lifetime ---------------------------------------------------------- */

/* Create a simulation; returns a handle >= 0, or < 0 on error.
 * Called at the end of init_simulation() (src/eilmer/simcore.d:794-798). */
int eb200_init(const eb200_config* cfg);

/* Free everything owned by the handle (finalize_simulation, simcore.d:1677). */
int eb200_finalize(int sim);

/* Copy the last error message (NUL-terminated, truncated to n) into dest;
 * returns its full length.  Convention of src/gas/gas_cwrap.d:35-93. */
int eb200_last_error(char* dest, int n);

/* ---- block set-up (after compute_primary_cell_geometric_data +
 *      exchange_ghost_cell_geometry_data, simcore.d:245-263,448) ----------- */

/* Declare a block of this simulation.  blk_id is the global (universe) block
 * id; owner_rank says which process holds it.  Every process declares every
 * block it is connected to; arrays are only uploaded for local blocks.
 * Returns 0 or < 0. */
int eb200_block_create(int sim, int blk_id, int nic, int njc, int nkc, int owner_rank);

/* Static geometry of a local block, padded block layout.
 *   vol, areaxy      : cell volume[gtl=0] and xy-plane area (fvcell.d:396-438; areaxy may be
 *                      NULL unless axisymmetric)
 *   len_i,len_j,len_k: iLength, jLength, kLength INCLUDING ghost cells
 *                      (ghost values as set by sfluidblock.d:897-1086 or received by
 *                      full_face_copy.d:1576-1626); len_k may be NULL in 2D
 *   face[d]          : for d = 0,1,2 a pointer to 10 consecutive per-face arrays
 *                      n.x n.y n.z t1.x t1.y t1.z t2.x t2.y t2.z area  (fvinterface.d:301-355),
 *                      each NI*NJ*NK long; face[2] is NULL in 2D. */
int eb200_block_set_geometry(int sim, int blk_id,
                             const double* vol, const double* areaxy,
                             const double* len_i, const double* len_j, const double* len_k,
                             const double* const face[3]);

/* Boundary condition of one block face.
 *   kind = EB200_BC_INFLOW_SUPERSONIC: params = FlowState, nparams = 8 (+ 2*nsp if nsp > 1)
 *   kind = EB200_BC_OUTFLOW_FIXED_P: params = { p_outside }; EB200_BC_OUTFLOW_FIXED_PT: { p_outside, T_outside }
 *   kind = EB200_BC_EXCHANGE_FULL_FACE: other_blk/other_face/orientation as in
 *          full_face_copy.d:141-1380 (orientation 0 only in 3D; 2D all face pairs). */
int eb200_block_set_bc(int sim, int blk_id, int face, int kind,
                       const double* params, int nparams,
                       int other_blk, int other_face, int orientation);

/* Explicit cell mapping for an EB200_BC_EXCHANGE_FULL_FACE face (call after eb200_block_set_bc): the
 * faces of two 3D blocks can meet in any of 6 x 6 x 4 ways (full_face_copy.d:141-1380 spells out every
 * case; ExchangeBC_MappedCell locates source cells by position, mapped_cell_copy.d).  Here the caller
 * names the source cell of every ghost cell: ghost cell m behind `face`, m = (a2 * n1 + a1) * 2 + layer
 * with a1, a2 the cell indices along block directions (d+1)%3 and (d+2)%3 (d = face / 2, n1 cells along
 * the first) and layer 0 next to the face, takes the FlowState of interior cell
 * (src_ijk[3m], src_ijk[3m+1], src_ijk[3m+2]) of block other_blk.  n = number of ghost cells.
 * Every process declares the maps of all blocks, its own and the others'. */
int eb200_block_set_face_map(int sim, int blk_id, int face, const int* src_ijk, long long n);

/* Finish set-up: builds device tables (ghost maps, exchange lists).  Must be
 * called once after all blocks, geometry and BCs are in. */
int eb200_commit(int sim);

/* Install the inter-process exchange callback (NULL: all blocks local). */
int eb200_set_exchange(int sim, eb200_exchange_fn fn, void* user);

/* ---- flow data --------------------------------------------------------- */

/* Upload primitive variables of a local block: prims[v] is a padded array for
 * variable v in the EB200_PRIM_* order (ghost values ignored).  Then, like
 * init_simulation (simcore.d:325-334), encode_conserved (fvcell.d:511-583)
 * and decode_conserved (fvcell.d:586-821) are applied to every cell.
 * For a single-species gas nprims may also be EB200_NPRIM_SHORT: prims = { rho, u, velx, vely, velz }, the
 * variables decode_conserved starts from; p, T and a are then computed on the device (5/8 of the bytes cross PCIe).
 * The call only enqueues the copies and the kernel (page-locked host arrays must stay untouched until the next
 * call that returns results; pageable ones are staged before the call returns); a FlowState that cannot be decoded is reported by the next
 * eb200_compute_dt, eb200_step or eb200_run_steps (return < 0, eb200_last_error names the upload). */
int eb200_upload_flow(int sim, int blk_id, const double* const* prims, int nprims);

/* Download primitive variables (interior + current ghost values). */
int eb200_download_flow(int sim, int blk_id, double* const* prims, int nprims);

/* Download conserved quantities U[0] (ncq padded arrays; order of
 * ConservedQuantitiesIndices, conservedquantities.d:67-197:
 * mass, xMom, yMom, [zMom], totEnergy, [species x nsp]). */
int eb200_download_conserved(int sim, int blk_id, double* const* U, int ncq);

/* The same download without waiting for it: the copies are queued behind everything enqueued so far, on their own
 * stream, and the call returns.  U must stay valid (and, to overlap, be page-locked) until eb200_wait_downloads
 * returns.  The next eb200_upload_flow of this block and every step wait for the copies on the device, so a loop
 * "upload all blocks, step, download all blocks" keeps host -> device and device -> host copies of consecutive
 * steps in flight together.  (No counterpart in the reference: its FlowStates never leave host memory; this is the
 * transfer pattern of a D shim that refreshes FVCell.fs every step, simcore.d:1029-1039.) */
int eb200_download_conserved_async(int sim, int blk_id, double* const* U, int ncq);
int eb200_wait_downloads(int sim);

/* FlowStates of single cells of local blocks, e.g. the history points of a job
 * (setHistoryPoint; history.d:80-95 writes one line per history cell).
 * Probe m is interior cell (ijk[3m], ijk[3m+1], ijk[3m+2]) of block blk_ids[m];
 * out[m * nprims + v] is its variable v in the EB200_PRIM_* order. */
int eb200_probe_cells(int sim, int n, const int* blk_ids, const int* ijk, double* out, int nprims);

/* ---- the hot path ------------------------------------------------------ */

/* Per-block time-step limits, min-reduced over all local blocks
 * (FluidBlock.determine_time_step_size, fluidblock.d:987-1084, and the
 * serial reduction of simcore_gasdynamic_step.d:94-100).
 * out[0] = dt_allow, out[1] = cfl_max, out[2] = dt_allow_parab (0: inviscid).
 * The growth/shrink policy (:135-156) stays with the caller. */
int eb200_compute_dt(int sim, double dt_current, double cfl_value, int check_cfl, double out[3]);

/* One whole time step of size dt from U[0] (all stages, ghost-cell exchange,
 * boundary conditions, flux, update, decode, bad-cell count).
 * Returns 0: success, U[0] and the FlowStates hold the new solution.
 *         1: step failed (a cell could not be decoded); U[0] and FlowStates are
 *            restored to the start of the step; the caller reduces dt (x0.2)
 *            and retries (simcore_gasdynamic_step.d:995-999,1545).
 *       < 0: fatal (too many bad cells :1440-1453, CUDA error ...).
 * n_bad_cells (may be NULL) receives the invalid-cell count of the last stage. */
int eb200_step(int sim, double t0, double dt, int* n_bad_cells);

/* ---- introspection used by benchmarks and tests ------------------------ */

/* Number of CUDA kernels this library has launched since eb200_init. */
long long eb200_kernel_launches(int sim);

/* Device time (ms, CUDA events on the library's stream) spent in the fused
 * flux+update kernel since the last call with reset != 0, and its launch
 * count. */
int eb200_flux_kernel_time(int sim, int reset, double* ms, long long* launches);

/* Run `nsteps` steps of fixed dt back to back without host synchronisation
 * between them (bench inner loop; same work as nsteps calls of eb200_step).
 * Returns like eb200_step. */
int eb200_run_steps(int sim, double t0, double dt, int nsteps, int* n_bad_cells);

/* Take back the last eb200_step, which must have returned 0 and must be the last call that changed the state:
 * FlowStates and U[0] are again those of the start of that step (nothing is copied: the start-of-step buffers are
 * intact until the next step).  For multi-process runs, where the reference decides collectively
 * (MPI_Allreduce of step_failed, simcore_gasdynamic_step.d:1545-1554): a rank whose own step succeeded while
 * another rank's failed takes its step back and retries with the smaller dt like everybody else.  0 or < 0. */
int eb200_undo_step(int sim);

/* ---- direct halo exchange between the processes of one node ---------------------------
 * Replaces MPI_Irecv/MPI_Send/MPI_Wait of full_face_copy.d:1681,1803-1842 (and the exchange callback above) by
 * stores over NVLink: after eb200_commit every process exports one opaque blob per halo peer
 * (eb200_p2p_export), the host layer moves the blobs (MPI_Alltoall / torch.distributed in the Python host), and
 * every process imports the blobs its peers made for it (eb200_p2p_import).  From then on eb200_step writes the
 * cells a peer needs straight into that peer's ghost cells (CUDA IPC mapping of its arena) and raises a flag in
 * the peer's memory; no host code runs per stage.  Until every peer has been imported the callback is used.
 *
 * eb200_p2p_export: writes the blob this process makes for `peer_rank` into `blob` (nbytes of room) and returns
 * its size; with blob == NULL or too little room it only returns the size needed.  < 0 on error.
 * eb200_p2p_import: returns 1 when all peers are imported (direct exchange is on), 0 when some are missing,
 * < 0 on error (e.g. the processes do not share a node, or the peer runs another configuration). */
int eb200_p2p_export(int sim, int peer_rank, void* blob, int nbytes);
int eb200_p2p_import(int sim, int peer_rank, const void* blob, int nbytes);

/* One line of text about how the simulation was set up (blocks per kernel, tiles, TMA, halo transport):
 * copied into dest (NUL-terminated, truncated to n); returns its full length.  Diagnostics only. */
int eb200_describe(int sim, char* dest, int n);

/* The cudaStream_t (as void*) on which the library enqueues all its work; lets the caller
 * record CUDA events around calls and order its own transfers.  NULL on error. */
void* eb200_cuda_stream(int sim);

/* 1 if the block uses the uniform-Cartesian fast path (metrics folded into
 * per-block constants), 0 for the general-metric path, < 0 on error. */
int eb200_block_is_cartesian(int sim, int blk_id);

/* Test hook: reconstruction (onedinterp.d:751-988) + interface flux (fluxcalc.d:54-184) of
 * `nfaces` independent faces through the general-metric device code.
 *   cells[nfaces][4][nprim] : FlowStates of L1, L0, R0, R1 (EB200_PRIM order)
 *   len[nfaces][4]          : cell lengths along the stencil
 *   geo[nfaces][10]         : n t1 t2 area
 *   F[nfaces][ncq]          : flux per unit area, global frame;  ok[nfaces] (may be NULL): 0 where
 *                             the reference would throw. */
int eb200_debug_face_flux(int sim, int nfaces, const double* cells, const double* len, const double* geo,
                          double* F, int* ok);

#ifdef __cplusplus
}
#endif
#endif /* EB200_H */
