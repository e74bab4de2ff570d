// eb200.cu -- host runtime behind the C ABI of include/eb200.h.
//
// Owns the device arena (SoA fields of all local blocks), the block table, the ghost-cell
// work lists and the stage loop of gasdynamic_explicit_increment_with_fixed_grid
// (reference src/eilmer/simcore_gasdynamic_step.d:906-1575).  Every numerical operation
// of a time step runs in the CUDA kernels of this directory; there is no CPU fallback.
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "eb200_internal.h"


#ifndef EB_TILE_Y
#define EB_TILE_Y 8
#endif

namespace {

thread_local char g_err[1024] = "";
void set_err(const char* fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}

#define CUDA_OK(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            set_err("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__,   \
                    cudaGetErrorString(e_));                                                   \
            return -100;                                                                       \
        }                                                                                      \
    } while (0)

struct BC {
    int kind = EB200_BC_WALL_WITH_SLIP, other_blk = -1, other_face = -1, orientation = 0;
    std::vector<int> map;          // eb200_block_set_face_map: (i, j, k) of the source cell of every ghost cell
    std::vector<double> params;
    int param_index = -1;
};

struct Block {
    int id = 0, nic = 0, njc = 0, nkc = 0, owner = 0;
    bool local = false;
    int NI = 0, NJ = 0, NK = 0, kg = 0;
    long long ncp = 0, cell0 = -1, stride[3] = {0, 0, 0};
    int local_index = -1;
    cudaEvent_t ev_d2h = nullptr;             // last asynchronous download of this block (eb200_download_conserved_async)
    bool d2h_pending = false;
    bool has_geometry = false, cartesian = false;
    EbBlockDesc cartD;                        // constants of the uniform-Cartesian fast path
    std::vector<double> vol, areaxy, len[3], face[3];
    BC bc[6];
    long long cidx(int i, int j, int k) const { return ((long long)(k + kg) * NJ + (j + EB_NG)) * NI + (i + EB_NG); }
};

struct Peer {
    int rank = -1;
    std::vector<int> send_idx, recv_idx;      // arena indices
    int *d_send_idx = nullptr, *d_recv_idx = nullptr;
    double *d_send = nullptr, *d_recv = nullptr;
    // direct halo over NVLink (eb200_p2p_export / eb200_p2p_import): the peer's arena mapped through CUDA IPC
    bool p2p = false;
    long long recv_off = 0;                   // byte offset of my recv index list for this peer inside the shared region
    double* r_prim[3] = { nullptr, nullptr, nullptr };
    double* r_S = nullptr;
    unsigned long long* r_flag = nullptr;     // the peer's flag that I raise
    long long r_total = 0;
    int* d_dst_idx = nullptr;                 // the peer's ghost cells (its arena indices) in wire order
    std::vector<void*> opened;
};

#define EB_P2P_MAGIC 0x65623270u

struct P2PBlob {                              // what a rank tells one peer about itself; plain data, moved by the host layer
    unsigned magic;
    int exporter_rank, nprim, has_S;
    long long total;                          // field stride of the exporter's arena
    cudaIpcMemHandle_t prim[3]; long long prim_off[3];
    cudaIpcMemHandle_t S; long long S_off;
    cudaIpcMemHandle_t shared; long long shared_off;   // the exporter's shared region: flags, then its recv index lists
    long long flag_off;                       // inside the region: the flag the importer raises
    long long recv_idx_off, recv_count;       // inside the region: the exporter's ghost cells the importer fills, wire order
};

struct Sim {
    eb200_config cfg;
    EbParams P;
    EbGas hgas;
    EbGas* d_gas = nullptr;
    int n_stages = 0, threeD = 0, nfaces = 4;
    std::vector<std::unique_ptr<Block>> blocks;
    std::vector<Block*> local;
    bool committed = false;
    EbArena A;
    std::vector<void*> allocs;
    std::vector<EbBlockDesc> hdesc;
    EbBlockDesc* d_desc = nullptr;
    CUtensorMap* d_tmaps = nullptr;        // [3 prim buffers][local blocks], nullptr when TMA staging is not used
    long long ncta = 0;
    int which = 0;
    EbCopyItem* d_copy = nullptr; long long ncopy = 0;   // ncopy: items every stage needs
    long long ncopy_full = 0;              // + the copies the flux kernel does itself (push); needed after an upload
    bool ghosts_stale = true;              // FlowStates came from the host: the pushed ghost cells are not filled
    EbReflectItem* d_refl = nullptr; long long nrefl = 0;
    EbFillItem* d_fill = nullptr; long long nfill = 0;
    double* d_params = nullptr;
    std::vector<Peer> peers;
    char* d_shared = nullptr; size_t shared_bytes = 0;       // flags[EB_P2P_MAXPEERS] + recv index lists, one allocation = one IPC handle
    unsigned long long** d_remote_flags = nullptr;           // device array: r_flag of every peer
    unsigned long long halo_seq = 0;
    bool p2p_ready = false;
    eb200_exchange_fn exchange = nullptr; void* exchange_user = nullptr;
    int* d_status = nullptr; int* h_status = nullptr;
    unsigned long long* d_red = nullptr; double* d_last = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t comm_stream = nullptr;       // halo traffic of other ranks, overlapped with interior tiles
    cudaStream_t d2h_stream = nullptr;        // asynchronous downloads: device -> host copies run beside the uploads of the next step
    cudaEvent_t ev_state = nullptr, ev_d2h_all = nullptr;
    bool d2h_inflight = false;
    cudaStream_t bnd_stream = nullptr;        // boundary tiles: start when the halo is in, run under the interior launch's tail
    cudaEvent_t ev_pack = nullptr, ev_comm = nullptr, ev_ghost = nullptr, ev_bnd = nullptr;
    int* d_tiles_int = nullptr; long long n_tiles_int = 0;   // tiles that read no ghost cell of another rank
    int* d_tiles_bnd = nullptr; long long n_tiles_bnd = 0;   // tiles that do
    bool upload_unchecked = false;            // eb200_upload_flow was called since the status was last read
    int undo_cur = -1;                        // >= 0: the last step succeeded and can be taken back (eb200_undo_step)
    int cur = 0;                              // index of the primitive buffer holding the current state
    int U0 = 0;                               // index of the U level that currently plays U[0]
    std::vector<int> Ulev;                    // permutation of U levels (swap at the end of a step)
    long long launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    size_t ev_used = 0;
    double flux_ms_acc = 0.0; long long flux_launches = 0;
};

std::vector<std::unique_ptr<Sim>> g_sims;

Sim* get_sim(int h)
{
    if (h < 0 || h >= (int)g_sims.size() || !g_sims[h]) { set_err("invalid sim handle %d", h); return nullptr; }
    return g_sims[h].get();
}
Block* get_blk(Sim* s, int id)
{
    for (auto& b : s->blocks) if (b->id == id) return b.get();
    set_err("unknown block id %d", id);
    return nullptr;
}

int n_stages_for(int scheme)
{
    switch (scheme) {
    case EB200_UPDATE_EULER: return 1;
    case EB200_UPDATE_PC: case EB200_UPDATE_MIDPOINT: return 2;
    case EB200_UPDATE_CLASSIC_RK3: case EB200_UPDATE_TVD_RK3: case EB200_UPDATE_DENMAN_RK3: return 3;
    case EB200_UPDATE_CLASSIC_RK4: return 4;
    }
    return 0;
}
// gamma tables of simcore_gasdynamic_step.d:1235-1395
void stage_gammas(int scheme, int stage, double g[4])
{
    g[0] = g[1] = g[2] = g[3] = 0.0;
    if (stage == 1) {
        switch (scheme) {
        case EB200_UPDATE_EULER: case EB200_UPDATE_PC: case EB200_UPDATE_TVD_RK3: g[0] = 1.0; break;
        case EB200_UPDATE_MIDPOINT: case EB200_UPDATE_CLASSIC_RK3: g[0] = 0.5; break;
        case EB200_UPDATE_DENMAN_RK3: g[0] = 8.0 / 15.0; break;
        case EB200_UPDATE_CLASSIC_RK4: g[0] = 1.0 / 2.0; break;
        }
    } else if (stage == 2) {
        switch (scheme) {
        case EB200_UPDATE_PC: g[0] = 0.5; g[1] = 0.5; break;
        case EB200_UPDATE_MIDPOINT: g[0] = 0.0; g[1] = 1.0; break;
        case EB200_UPDATE_CLASSIC_RK3: g[0] = -1.0; g[1] = 2.0; break;
        case EB200_UPDATE_TVD_RK3: g[0] = 0.25; g[1] = 0.25; break;
        case EB200_UPDATE_DENMAN_RK3: g[0] = -17.0 / 60.0; g[1] = 5.0 / 12.0; break;
        case EB200_UPDATE_CLASSIC_RK4: g[0] = 0.0; g[1] = 1.0 / 2.0; break;
        }
    } else if (stage == 3) {
        switch (scheme) {
        case EB200_UPDATE_CLASSIC_RK3: g[0] = 1.0 / 6.0; g[1] = 4.0 / 6.0; g[2] = 1.0 / 6.0; break;
        case EB200_UPDATE_TVD_RK3: g[0] = 1.0 / 6.0; g[1] = 1.0 / 6.0; g[2] = 4.0 / 6.0; break;
        case EB200_UPDATE_DENMAN_RK3: g[0] = 0.0; g[1] = -5.0 / 12.0; g[2] = 3.0 / 4.0; break;
        case EB200_UPDATE_CLASSIC_RK4: g[0] = 0.0; g[1] = 0.0; g[2] = 1.0; break;
        }
    } else {
        g[0] = 1.0 / 6.0; g[1] = 1.0 / 3.0; g[2] = 1.0 / 3.0; g[3] = 1.0 / 6.0;       // classic_rk4
    }
}

// --- host copies of the curve evaluation, used once to fill the constants of EbCurve --------
bool h_coeffs(const EbCurve& c, double T, double a[9])
{
    int nb = c.nbreaks;
    if (c.nseg == 1 || T < (c.T_breaks[1] - 0.5 * c.T_blends[0])) { memcpy(a, c.coeffs[0], 72); return true; }
    if (T > (c.T_breaks[nb - 2] + 0.5 * c.T_blends[c.nseg - 2])) { memcpy(a, c.coeffs[c.nseg - 1], 72); return true; }
    for (int i = 1; i < nb - 1; ++i) {
        double lo = c.T_breaks[i] - 0.5 * c.T_blends[i - 1], hi = c.T_breaks[i] + 0.5 * c.T_blends[i - 1];
        if (T >= lo && T <= hi) {
            double wB = (1. / c.T_blends[i - 1]) * (T - lo), wA = 1.0 - wB;
            for (int j = 0; j < 9; ++j) a[j] = wA * c.coeffs[i - 1][j] + wB * c.coeffs[i][j];
            return true;
        }
        if (T > hi && T < (c.T_breaks[i + 1] - 0.5 * c.T_blends[i])) { memcpy(a, c.coeffs[i], 72); return true; }
    }
    return false;
}
double h_Cp(const EbCurve& c, double T)
{
    double a[9]; h_coeffs(c, T, a);
    double Cp_on_R = a[0] / (T * T) + a[1] / T + a[2] + a[3] * T;
    Cp_on_R += a[4] * T * T + a[5] * T * T * T + a[6] * T * T * T * T;
    return c.R * Cp_on_R;
}
double h_h(const EbCurve& c, double T)
{
    double a[9]; h_coeffs(c, T, a);
    double logT = std::log(T);
    double h_on_RT = -a[0] / T + a[1] * logT + a[2] * T + a[3] * T * T / 2.0;
    h_on_RT += a[4] * T * T * T / 3.0 + a[5] * T * T * T * T / 4.0 + a[6] * T * T * T * T * T / 5.0 + a[7];
    return c.R * h_on_RT;
}

template <class T>
int dev_alloc(Sim* s, T** p, size_t count, bool zero = true)
{
    void* q = nullptr;
    CUDA_OK(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
    if (zero) CUDA_OK(cudaMemsetAsync(q, 0, std::max<size_t>(count, 1) * sizeof(T), s->stream));
    s->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}

template <class T>
int dev_upload(Sim* s, T** p, const std::vector<T>& v)
{
    if (dev_alloc(s, p, v.size(), false)) return -100;
    if (!v.empty()) CUDA_OK(cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s->stream));
    return 0;
}

// Is the block uniform-Cartesian?  All face frames of a direction identical and axis-aligned
// (signed permutation), all lengths / areas / volumes identical bit for bit.  If so, fill the
// constants of the descriptor.  Works on the caller's arrays (padded block layout).
bool detect_cartesian(const Sim* s, const Block* b, const double* vol, const double* areaxy,
                      const double* const len[3], const double* const face[3], EbBlockDesc& D)
{
    if (s->cfg.axisymmetric) return false;
    if (s->cfg.reserved_i[0]) return false;        // testing knob: force the general-metric path
    const int dims = s->cfg.dimensions;
    const int n[3] = { b->nic, b->njc, b->nkc };
    const double v0 = vol[b->cidx(0, 0, 0)];
    for (int k = 0; k < n[2]; ++k) for (int j = 0; j < n[1]; ++j) {
        const double* row = vol + b->cidx(0, j, k);
        for (int i = 0; i < n[0]; ++i) if (row[i] != v0) return false;
    }
    D.vol = v0; D.vol_inv = 1.0 / v0; D.areaxy = (dims == 2 && areaxy) ? areaxy[b->cidx(0, 0, 0)] : 0.0;
    for (int d = 0; d < 3; ++d) { D.len[d] = 1.0; D.area[d] = 0.0; }
    for (int d = 0; d < dims; ++d) {
        // lengths: interior cells plus the two ghost layers either side along d
        int lo[3] = { 0, 0, 0 }, hi[3] = { n[0], n[1], n[2] };
        lo[d] = -EB_NG; hi[d] = n[d] + EB_NG;
        const double l0 = len[d][b->cidx(0, 0, 0)];
        for (int k = lo[2]; k < hi[2]; ++k) for (int j = lo[1]; j < hi[1]; ++j) {
            const double* row = len[d] + b->cidx(0, j, k);
            for (int i = lo[0]; i < hi[0]; ++i) if (row[i] != l0) return false;
        }
        D.len[d] = l0;
        // faces: index range [0, n[d]] along d
        int fhi[3] = { n[0], n[1], n[2] }; fhi[d] = n[d] + 1;
        double f0[10];
        const long long c0 = b->cidx(0, 0, 0);
        for (int m = 0; m < 10; ++m) f0[m] = face[d][(long long)m * b->ncp + c0];
        for (int m = 0; m < 10; ++m) {
            const double* fm = face[d] + (long long)m * b->ncp;
            for (int k = 0; k < fhi[2]; ++k) for (int j = 0; j < fhi[1]; ++j) {
                const double* row = fm + b->cidx(0, j, k);
                for (int i = 0; i < fhi[0]; ++i) if (row[i] != f0[m]) return false;
            }
        }
        // axis aligned?
        for (int v = 0; v < 3; ++v) {
            int nz = 0, which = -1;
            for (int m = 0; m < 3; ++m) {
                double x = f0[3 * v + m];
                if (x != 0.0) { ++nz; which = m; if (std::fabs(x) != 1.0) return false; }
            }
            if (nz != 1) return false;
            D.fr[d].perm[v] = which;
            D.fr[d].neg[v] = f0[3 * v + which] < 0.0 ? 1 : 0;
        }
        if (D.fr[d].perm[0] == D.fr[d].perm[1] || D.fr[d].perm[0] == D.fr[d].perm[2] || D.fr[d].perm[1] == D.fr[d].perm[2]) return false;
        // the fast path is compiled for the frames Eilmer builds on a box grid (CartFrame in
        // flux_kernel_v2.cuh): 3D (e_d, e_d+1, e_d+2); 2D i-face (x, -y, z), j-face (y, x, z)
        {
            int ep[3], en[3] = { 0, 0, 0 };
            if (dims == 3) { ep[0] = d; ep[1] = (d + 1) % 3; ep[2] = (d + 2) % 3; }
            else { ep[0] = d; ep[1] = 1 - d; ep[2] = 2; en[1] = (d == 0) ? 1 : 0; }
            for (int v = 0; v < 3; ++v) if (D.fr[d].perm[v] != ep[v] || D.fr[d].neg[v] != en[v]) return false;
        }
        for (int m = 0; m < 3; ++m) D.nvec[d][m] = f0[m];
        D.area[d] = f0[9];
        // l2r2_prepare on four equal lengths (onedinterp.d:338-354)
        EbWeights& w = D.w[d];
        const double l = l0;
        w.lenL0 = l; w.lenR0 = l;
        w.aL0 = 0.5 * l / (l + 2.0 * l + l);
        w.aR0 = 0.5 * l / (l + 2.0 * l + l);
        w.two_over_L0L1 = 2.0 / (l + l); w.two_over_R0L0 = 2.0 / (l + l); w.two_over_R1R0 = 2.0 / (l + l);
        w.two_L0_plus_L1 = (2.0 * l + l); w.two_R0_plus_R1 = (2.0 * l + l);
        const double c = w.two_over_R0L0;
        D.uq[d][0] = w.aL0 * w.two_L0_plus_L1 * c; D.uq[d][1] = w.aL0 * w.lenR0 * c;
        D.uq[d][2] = w.aR0 * w.lenL0 * c; D.uq[d][3] = w.aR0 * w.two_R0_plus_R1 * c;
        D.uq[d][4] = 1.0 / (c * c);
    }
    if (dims == 2) {
        for (int m = 0; m < 5; ++m) D.uq[2][m] = D.uq[0][m];
        D.fr[2].perm[0] = 2; D.fr[2].perm[1] = 0; D.fr[2].perm[2] = 1; D.fr[2].neg[0] = D.fr[2].neg[1] = D.fr[2].neg[2] = 0;
        D.nvec[2][0] = 0; D.nvec[2][1] = 0; D.nvec[2][2] = 1; D.w[2] = D.w[0];
    }
    return true;
}

// full_face_copy.d:704-870 (2D) and :958-1014 (3D, orientation 0): interior cell (i,j,k) of `ot`
// that feeds ghost layer `layer` behind the boundary-face cell (t1,t2) of face `face` of a block.
int full_face_source(const Sim* s, int face, const Block* ot, int oface, int t1, int t2, int layer, int ijk[3])
{
    if (!s->threeD) {
        const int t = t1;
        const bool grpA = (face == EB200_NORTH || face == EB200_WEST);
        const bool grpB = (oface == EB200_NORTH || oface == EB200_WEST);
        const bool rev = (grpA == grpB);
        switch (oface) {
        case EB200_NORTH: ijk[0] = rev ? ot->nic - t - 1 : t; ijk[1] = ot->njc - 1 - layer; break;
        case EB200_EAST: ijk[0] = ot->nic - 1 - layer; ijk[1] = rev ? ot->njc - t - 1 : t; break;
        case EB200_SOUTH: ijk[0] = rev ? ot->nic - t - 1 : t; ijk[1] = layer; break;
        case EB200_WEST: ijk[0] = layer; ijk[1] = rev ? ot->njc - t - 1 : t; break;
        default: set_err("bad face"); return -1;
        }
        ijk[2] = 0;
        return 0;
    }
    if ((face ^ 1) != oface) { set_err("3D full-face copy: only aligned opposite faces (orientation 0) are supported"); return -1; }
    const int d = face / 2, d1 = (d + 1) % 3, d2 = (d + 2) % 3;
    const int on[3] = { ot->nic, ot->njc, ot->nkc };
    ijk[d1] = t1; ijk[d2] = t2;
    ijk[d] = (oface & 1) ? on[d] - 1 - layer : layer;
    return 0;
}

// Enumerate the ghost cells behind one block face in a fixed order: (t2, t1, layer).
template <class Fn>
void for_face_ghosts(const Sim* s, const Block* b, int face, Fn fn)
{
    const int d = face / 2, hi = face & 1;
    const int n[3] = { b->nic, b->njc, b->nkc };
    const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
    for (int a2 = 0; a2 < n[d2]; ++a2) for (int a1 = 0; a1 < n[d1]; ++a1) {
        int idx[3]; idx[d1] = a1; idx[d2] = a2; idx[d] = hi ? n[d] : 0;
        const long long cf = b->cidx(idx[0], idx[1], idx[2]);     // plus-side cell of the boundary face
        const int t1 = s->threeD ? a1 : (d == 0 ? idx[1] : idx[0]);
        for (int layer = 0; layer < EB_NG; ++layer) {
            const long long ghost = hi ? cf + layer * b->stride[d] : cf - (1 + layer) * b->stride[d];
            const long long mirror = hi ? cf - (1 + layer) * b->stride[d] : cf + layer * b->stride[d];
            const long long first = hi ? cf - b->stride[d] : cf;
            fn(t1, a2, layer, cf, ghost, mirror, first);
        }
    }
}

int record_flux_events(Sim* s, cudaEvent_t* e0, cudaEvent_t* e1)
{
    if (s->ev_used >= s->ev_pool.size()) {
        if (s->ev_pool.size() >= 4096) { *e0 = *e1 = nullptr; return 0; }
        cudaEvent_t a, b;
        CUDA_OK(cudaEventCreate(&a)); CUDA_OK(cudaEventCreate(&b));
        s->ev_pool.push_back({ a, b });
    }
    *e0 = s->ev_pool[s->ev_used].first; *e1 = s->ev_pool[s->ev_used].second;
    s->ev_used++;
    return 0;
}

int drain_flux_events(Sim* s)
{
    for (size_t i = 0; i < s->ev_used; ++i) {
        float ms = 0.f;
        CUDA_OK(cudaEventSynchronize(s->ev_pool[i].second));
        CUDA_OK(cudaEventElapsedTime(&ms, s->ev_pool[i].first, s->ev_pool[i].second));
        s->flux_ms_acc += ms; s->flux_launches++;
    }
    s->ev_used = 0;
    return 0;
}

#define MODE_CALL(s, fn, ...)                                         \
    do {                                                              \
        if ((s)->cfg.strict_fp) eb_strict::fn(__VA_ARGS__);           \
        else eb_fast::fn(__VA_ARGS__);                                \
        (s)->launches++;                                              \
    } while (0)

// Halo traffic with other processes (phase 02 of the reference step): pack on the main stream,
// then callback + unpack on the communication stream so that the main stream can meanwhile work on
// tiles that do not touch those ghost cells.  ev_comm is recorded when the ghost cells are in place.
int exchange_remote(Sim* s, double* prim, int buf)
{
    if (s->peers.empty()) return 0;
    if (s->p2p_ready) {
        // direct stores into the neighbours' ghost cells, a flag per peer; no host code, no staging buffers
        const int np = (int)s->peers.size();
        const unsigned long long seq = ++s->halo_seq;
        CUDA_OK(cudaEventRecord(s->ev_pack, s->stream));
        CUDA_OK(cudaStreamWaitEvent(s->comm_stream, s->ev_pack, 0));
        if (s->P.shock_detect) {
            // FlowState.S has a single buffer: the peers may overwrite my S ghost cells only after my previous stage
            MODE_CALL(s, launch_halo_signal, s->d_remote_flags, np, seq, 1, s->comm_stream);
            MODE_CALL(s, launch_halo_wait, (const unsigned long long*)s->d_shared, np, seq, 1, s->d_status, s->comm_stream);
        }
        for (int p = 0; p < np; ++p) {
            Peer& pr = s->peers[p];
            MODE_CALL(s, launch_put, s->P, pr.r_total, prim, s->A.S, pr.r_prim[buf], pr.r_S, pr.d_send_idx, pr.d_dst_idx,
                      (long long)pr.send_idx.size(), s->comm_stream);
        }
        MODE_CALL(s, launch_halo_signal, s->d_remote_flags, np, seq, 0, s->comm_stream);
        MODE_CALL(s, launch_halo_wait, (const unsigned long long*)s->d_shared, np, seq, 0, s->d_status, s->comm_stream);
        CUDA_OK(cudaEventRecord(s->ev_comm, s->comm_stream));
        return 0;
    }
    if (!s->exchange) { set_err("blocks on other ranks are connected but no exchange callback is installed"); return -5; }
    const int np = (int)s->peers.size();
    std::vector<int> ranks(np);
    std::vector<double*> sp(np), rp(np);
    std::vector<long long> sc(np), rc(np);
    for (int p = 0; p < np; ++p) {
        Peer& pr = s->peers[p];
        MODE_CALL(s, launch_pack, s->P, prim, s->A.S, pr.d_send_idx, (long long)pr.send_idx.size(), pr.d_send, s->stream);
        ranks[p] = pr.rank; sp[p] = pr.d_send; rp[p] = pr.d_recv;
        const int nv = s->P.nprim + (s->P.shock_detect ? 1 : 0);
        sc[p] = (long long)pr.send_idx.size() * nv; rc[p] = (long long)pr.recv_idx.size() * nv;
    }
    CUDA_OK(cudaEventRecord(s->ev_pack, s->stream));
    CUDA_OK(cudaStreamWaitEvent(s->comm_stream, s->ev_pack, 0));
    int rc_cb = s->exchange(s->exchange_user, np, ranks.data(), sp.data(), sc.data(), rp.data(), rc.data(), (void*)s->comm_stream);
    if (rc_cb != 0) { set_err("exchange callback failed (%d)", rc_cb); return -6; }
    for (int p = 0; p < np; ++p) {
        Peer& pr = s->peers[p];
        MODE_CALL(s, launch_unpack, s->P, prim, s->A.S, pr.d_recv_idx, (long long)pr.recv_idx.size(), pr.d_recv, s->comm_stream);
    }
    CUDA_OK(cudaEventRecord(s->ev_comm, s->comm_stream));
    return 0;
}

// TMA descriptors for the tile staging of the tuned flux kernel: per prim buffer and local block a
// 4D tensor (i, j, k, field) in the padded block layout; one box = (36, TY+4, 1, 2 fields) with the TY of the
// kernel that runs the block.  TMA wants 16-byte multiples for the strides, i.e. an even NI; otherwise the
// face-centred kernel stages with cp.async.
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiled tensor_map_encoder()
{
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return (EncodeTiled)fn;
}

// Can the tiles be staged by TMA?  (16-byte strides: even padded widths; the driver entry point exists)
bool tma_available(const Sim* s)
{
    if (s->cfg.reserved_i[2]) return false;             // testing knob: never use TMA
    for (const Block* b : s->local) if (b->NI % 2) return false;
    return tensor_map_encoder() != nullptr;
}

int build_tensor_maps(Sim* s)
{
    s->d_tmaps = nullptr;
    if (!tma_available(s)) return 0;
    EncodeTiled encode = tensor_map_encoder();
    const size_t nl = s->local.size();
    std::vector<CUtensorMap> maps(3 * nl);
    for (int p = 0; p < 3; ++p) {
        for (size_t n = 0; n < nl; ++n) {
            const Block* b = s->local[n];
            cuuint64_t gdim[4] = { (cuuint64_t)b->NI, (cuuint64_t)b->NJ, (cuuint64_t)b->NK, (cuuint64_t)s->P.nprim };
            cuuint64_t gstride[3] = { (cuuint64_t)b->NI * 8, (cuuint64_t)b->NI * b->NJ * 8, (cuuint64_t)s->P.total * 8 };
            cuuint32_t box[4] = { EB_V2_COLS, (cuuint32_t)((s->hdesc[n].v3 ? EB_V3_TY : EB_V2_TY) + 4), 1, 2 };
            cuuint32_t estr[4] = { 1, 1, 1, 1 };
            CUresult r = encode(&maps[p * nl + n], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, s->A.prim[p] + b->cell0, gdim, gstride, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_err("cuTensorMapEncodeTiled failed (%d)", (int)r); return -7; }
        }
    }
    return dev_upload(s, &s->d_tmaps, maps);
}

// Same-GPU full-face copies and boundary conditions (phases 02/03 of the reference step).
int fill_local_ghost_cells(Sim* s, double* prim, bool all_copies)
{
    const long long ncopy = all_copies ? s->ncopy_full : s->ncopy;
    if (ncopy + s->nrefl + s->nfill > 0)
        MODE_CALL(s, launch_ghosts, s->P, s->d_gas, s->d_desc, s->A, prim, s->d_copy, ncopy, s->d_refl, s->nrefl, s->d_fill, s->nfill, s->d_params, s->stream);
    return 0;
}

// Does face f of local block b push its two cell layers into the ghost cells of its neighbour?
// Same GPU, opposite faces, equal block sizes (then one constant index offset maps cell to ghost cell),
// at least four cells per direction (a cell has at most one target per direction).
bool face_pushes(const Sim* s, const Block* b, int f, const Block** other)
{
    if (s->cfg.reserved_i[3]) return false;
    const BC& bc = b->bc[f];
    if (bc.kind != EB200_BC_EXCHANGE_FULL_FACE) return false;
    const Block* ot = nullptr;
    for (auto& q : s->blocks) if (q->id == bc.other_blk) ot = q.get();
    if (!ot || !ot->local || !b->local || ot == b) return false;
    if (bc.other_face != (f ^ 1) || !bc.map.empty()) return false;
    const BC& back = ot->bc[f ^ 1];
    if (!back.map.empty()) return false;
    if (back.kind != EB200_BC_EXCHANGE_FULL_FACE || back.other_blk != b->id || back.other_face != f) return false;
    if (ot->nic != b->nic || ot->njc != b->njc || ot->nkc != b->nkc) return false;
    if (b->nic < 4 || b->njc < 4 || (s->threeD && b->nkc < 4)) return false;
    if (other) *other = ot;
    return true;
}

// All stages of one step, enqueued on the stream (no host synchronisation).
int enqueue_step(Sim* s, double dt)
{
    const int ns = s->n_stages;
    // the start-of-step buffer stays intact; the two others ping-pong between stages
    const int work[2] = { (s->cur + 1) % 3, (s->cur + 2) % 3 };
    int in_buf = s->cur;
    for (int stage = 1; stage <= ns; ++stage) {
        const int out_buf = work[(stage - 1) & 1];
        double* prim_in = s->A.prim[in_buf];
        double* prim_out = s->A.prim[out_buf];
        int rc = exchange_remote(s, prim_in, in_buf);
        if (rc) return rc;
        rc = fill_local_ghost_cells(s, prim_in, stage == 1 && s->ghosts_stale);
        if (rc) return rc;
        if (stage == 1) s->ghosts_stale = false;
        if (s->P.shock_detect && stage == 1) {
            // detect_shocks (phase 04) needs every ghost cell: wait for the halo of other ranks first
            if (!s->peers.empty()) CUDA_OK(cudaStreamWaitEvent(s->stream, s->ev_comm, 0));
            long long max_pos = 0;
            for (const EbBlockDesc& D : s->hdesc)
                max_pos = std::max(max_pos, (long long)(D.nic + 1) * (D.njc + 1) * (s->threeD ? D.nkc + 1 : 1));
            MODE_CALL(s, launch_detect_shocks_all, s->P, s->d_desc, (int)s->hdesc.size(), max_pos, s->A, prim_in, s->stream);
            s->launches += s->P.strict_shock ? 1 : 0;
        }
        EbStageArgs S;
        memset(&S, 0, sizeof S);
        S.prim_in = prim_in; S.prim_out = prim_out;
        S.cellS = s->P.shock_detect ? s->A.S : nullptr;      // FlowState.S goes with the FlowState (the ghost-cell kernel copies it too)
        S.tmaps = s->d_tmaps ? (const void*)(s->d_tmaps + (size_t)in_buf * s->local.size()) : nullptr;
        // Denman's scheme builds every stage on the U of the stage before (U_old = cell.U[1], cell.U[2],
        // simcore_gasdynamic_step.d:1303,1352), so every stage stores its U; the others start from U[0]
        const bool denman = s->cfg.update_scheme == EB200_UPDATE_DENMAN_RK3;
        S.U0 = s->A.U[s->Ulev[denman ? stage - 1 : 0]];
        S.U_out = (stage == ns || denman) ? s->A.U[s->Ulev[stage]] : nullptr;
        for (int m = 0; m < 3; ++m) S.dUdt_prev[m] = (m < stage - 1) ? s->A.dUdt[m] : nullptr;
        S.dUdt_out = (stage < ns) ? s->A.dUdt[stage - 1] : nullptr;
        double g[4]; stage_gammas(s->cfg.update_scheme, stage, g);
        if (stage == 1) { S.dt_g[0] = dt * g[0]; S.dt_g[4] = dt; }
        else { S.dt_g[0] = g[0]; S.dt_g[1] = g[1]; S.dt_g[2] = g[2]; S.dt_g[3] = g[3]; S.dt_g[4] = dt; }
        S.stage = stage; S.n_stages = ns; S.status = s->d_status;
        cudaEvent_t e0, e1;
        if (record_flux_events(s, &e0, &e1)) return -100;
        if (e0) CUDA_OK(cudaEventRecord(e0, s->stream));
        auto launch = [&](const int* tiles, long long n, cudaStream_t st) {
            if (n <= 0) return;
            S.tile_list = tiles;
            if (s->cfg.strict_fp) eb_strict::launch_flux_update(s->cfg.flux_calculator, s->cfg.gas_model, s->P, s->d_gas, s->d_desc, (int)s->hdesc.size(), n, s->A, S, s->which, st);
            else eb_fast::launch_flux_update(s->cfg.flux_calculator, s->cfg.gas_model, s->P, s->d_gas, s->d_desc, (int)s->hdesc.size(), n, s->A, S, s->which, st);
            s->launches += ((s->which & 1) ? 1 : 0) + ((s->which & 2) ? 1 : 0);
        };
        if (s->peers.empty()) {
            launch(nullptr, s->ncta, s->stream);
        } else {
            // interior tiles now (they overlap with the halo traffic); boundary tiles on their own stream as
            // soon as the halo and the local ghost cells are in, so that they fill in under the interior tail
            CUDA_OK(cudaEventRecord(s->ev_ghost, s->stream));
            launch(s->d_tiles_int, s->n_tiles_int, s->stream);
            CUDA_OK(cudaStreamWaitEvent(s->bnd_stream, s->ev_ghost, 0));
            CUDA_OK(cudaStreamWaitEvent(s->bnd_stream, s->ev_comm, 0));
            launch(s->d_tiles_bnd, s->n_tiles_bnd, s->bnd_stream);
            CUDA_OK(cudaEventRecord(s->ev_bnd, s->bnd_stream));
            CUDA_OK(cudaStreamWaitEvent(s->stream, s->ev_bnd, 0));
        }
        if (e1) CUDA_OK(cudaEventRecord(e1, s->stream));
        CUDA_OK(cudaGetLastError());
        in_buf = out_buf;
    }
    s->cur = in_buf;
    return 0;
}

int read_status(Sim* s)
{
    CUDA_OK(cudaMemcpyAsync(s->h_status, s->d_status, 8 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CUDA_OK(cudaStreamSynchronize(s->stream));
    return 0;
}

}  // namespace

// ================================================================================================
extern "C" {

int eb200_last_error(char* dest, int n)
{
    int len = (int)strlen(g_err);
    if (dest && n > 0) { strncpy(dest, g_err, n - 1); dest[n - 1] = 0; }
    return len;
}

int eb200_init(const eb200_config* cfg)
{
    if (!cfg) { set_err("null config"); return -1; }
    if (cfg->dimensions != 2 && cfg->dimensions != 3) { set_err("dimensions must be 2 or 3"); return -1; }
    if (cfg->n_species < 1 || cfg->n_species > EB_MAXSP) { set_err("bad n_species"); return -1; }
    if (cfg->gas_model == EB200_GAS_IDEAL && cfg->n_species != 1) { set_err("ideal gas has one species"); return -1; }
    if (cfg->gas_model == EB200_GAS_THERMALLY_PERFECT) {
        bool built = false;
        std::string list;
#define EB_CHK_NSP(N) { if (cfg->n_species == N) built = true; list += (list.empty() ? "" : ", "); list += #N; }
        EB_TPG_NSP_LIST(EB_CHK_NSP)
#undef EB_CHK_NSP
        if (!built) {
            set_err("thermally perfect gas: this library is built for %s species (got %d); rebuild with make TPG_NSP=\"... %d\"",
                    list.c_str(), cfg->n_species, cfg->n_species);
            return -1;
        }
#ifdef EB_TPG_NSP_DEFAULT
        if (cfg->n_species != 5 && cfg->flux_calculator != EB200_FLUX_AUSMDV) {
            set_err("thermally perfect gas with %d species: the default build has these kernels for ausmdv only "
                    "(five species for every flux calculator); rebuild with make TPG_NSP=\"%d 5\"", cfg->n_species, cfg->n_species);
            return -1;
        }
#endif
    }
    if (cfg->flux_calculator < 0 || cfg->flux_calculator > EB200_FLUX_HLLE2) { set_err("unknown flux calculator %d", cfg->flux_calculator); return -1; }
    if (cfg->solver_variant != 0.0 && (cfg->n_species != 1 || !cfg->interpolate_in_local_frame)) {
        set_err("solver_variant = lmr: single-species gas with interpolate_in_local_frame only (lmr/onedinterp.d:218-227, 270-292)"); return -1;
    }
    if ((cfg->flux_calculator == EB200_FLUX_HLLC || cfg->flux_calculator == EB200_FLUX_HLLE2) && cfg->n_species > 1) {
        set_err("hllc and hlle2 with multiple species are not on this path yet"); return -1;
    }
    if (!n_stages_for(cfg->update_scheme)) { set_err("unsupported update scheme %d", cfg->update_scheme); return -1; }
    const bool adaptive = (cfg->flux_calculator >= EB200_FLUX_ADAPTIVE_HANEL_AUSMDV && cfg->flux_calculator <= EB200_FLUX_ADAPTIVE_LDFSS0_LDFSS2) ||
                          cfg->flux_calculator == EB200_FLUX_ADAPTIVE_EFM_AUSMDV;
    if (adaptive && cfg->compression_tolerance > 0.0) { set_err("compression_tolerance should be negative!"); return -1; }
    if (adaptive && cfg->n_species > 1 && cfg->flux_calculator != EB200_FLUX_ADAPTIVE_HANEL_AUSMDV) {
        set_err("thermally perfect gas: of the adaptive flux calculators only adaptive_hanel_ausmdv is built"); return -1;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_err("no CUDA device available: eb200 has no CPU fallback");
        return -2;
    }
    CUDA_OK(cudaSetDevice(cfg->device));
    auto s = std::make_unique<Sim>();
    s->cfg = *cfg;
    s->threeD = cfg->dimensions == 3; s->nfaces = s->threeD ? 6 : 4;
    s->n_stages = n_stages_for(cfg->update_scheme);
    memset(&s->A, 0, sizeof s->A);
    EbParams& P = s->P; memset(&P, 0, sizeof P);
    P.dims = cfg->dimensions; P.axisymmetric = cfg->axisymmetric; P.nsp = cfg->n_species;
    P.iZMom = s->threeD ? 3 : -1; P.iEnergy = s->threeD ? 4 : 3;
    P.ncq = P.iEnergy + 1; P.iSpecies = -1;
    if (P.nsp > 1) { P.iSpecies = P.ncq; P.ncq += P.nsp; }
    P.nprim = EB200_NPRIM_BASE + (P.nsp > 1 ? 2 * P.nsp : 0);
    P.interpolation_order = cfg->interpolation_order; P.apply_limiter = cfg->apply_limiter;
    P.extrema_clipping = cfg->extrema_clipping; P.local_frame = cfg->interpolate_in_local_frame;
    P.entropy_fix = cfg->apply_entropy_fix; P.ignore_low_T = cfg->ignore_low_T_thermo_update_failure;
    P.eps_va = cfg->epsilon_van_albada; P.M_inf = cfg->M_inf; P.max_velocity = cfg->max_velocity;
    P.max_temp = cfg->max_temp; P.min_temp = cfg->min_temp; P.low_T = cfg->suggested_low_T_value;
    P.shock_detect = adaptive ? 1 : 0; P.strict_shock = cfg->strict_shock_detector;
    if (cfg->thermo_interpolator < EB200_INTERP_RHOU || cfg->thermo_interpolator > EB200_INTERP_RHOT) {
        set_err("unknown thermo_interpolator %d", cfg->thermo_interpolator); return -1;
    }
    P.thermo_interp = cfg->thermo_interpolator; P.lmr = (cfg->solver_variant != 0.0) ? 1 : 0;
    P.comp_tol = cfg->compression_tolerance; P.shear_tol = cfg->shear_tolerance;
    EbGas& g = s->hgas; memset(&g, 0, sizeof g);
    g.model = cfg->gas_model; g.nsp = cfg->n_species;
    if (cfg->gas_model == EB200_GAS_IDEAL) {
        g.Rgas = 8.31451 / cfg->ideal_mol_mass;              // ideal_gas.d:64-68
        g.gamma = cfg->ideal_gamma;
        g.Cv = g.Rgas / (g.gamma - 1.0);
        g.Cvinv = 1.0 / g.Cv;
        g.Cp = g.Rgas * g.gamma / (g.gamma - 1.0);
        g.gamma_CpCv = g.Cp / g.Cv;                           // gas_model.d:205
    } else {
        for (int i = 0; i < g.nsp; ++i) {
            const eb200_species& sp = cfg->species[i];
            EbCurve& c = g.curves[i];
            g.Rsp[i] = 8.31451 / sp.mol_mass;
            c.R = g.Rsp[i]; c.nseg = sp.nsegments; c.nbreaks = sp.nsegments + 1;
            if (c.nseg < 1 || c.nseg > EB_MAXSEG) { set_err("bad nsegments for species %d", i); return -1; }
            for (int k = 0; k <= c.nseg; ++k) c.T_breaks[k] = sp.T_break_points[k];
            for (int k = 0; k < c.nseg; ++k) c.T_blends[k] = sp.T_blend_ranges[k];
            memcpy(c.coeffs, sp.coeffs, sizeof c.coeffs);
            c.T_low = c.T_breaks[0]; c.T_high = c.T_breaks[c.nbreaks - 1];
            c.Cp_low = h_Cp(c, c.T_low); c.Cp_high = h_Cp(c, c.T_high);
            c.h_low = h_h(c, c.T_low); c.h_high = h_h(c, c.T_high);
        }
        g.uniform_curves = (g.curves[0].nseg >= 2) ? 1 : 0;
        for (int i = 1; i < g.nsp; ++i) {
            const EbCurve& a = g.curves[0]; const EbCurve& b = g.curves[i];
            if (a.nseg != b.nseg) { g.uniform_curves = 0; break; }
            for (int k = 0; k <= a.nseg; ++k) if (a.T_breaks[k] != b.T_breaks[k]) g.uniform_curves = 0;
            for (int k = 0; k < a.nseg; ++k) if (a.T_blends[k] != b.T_blends[k]) g.uniform_curves = 0;
        }
        for (int i = 0; i < g.nsp; ++i)
            for (int sg = 0; sg < g.curves[i].nseg; ++sg)
                for (int k = 0; k < 8; ++k) g.RA[sg][k][i] = g.curves[i].R * g.curves[i].coeffs[sg][k];
    }
    CUDA_OK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;                       // halo traffic gets the highest stream priority
        CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_OK(cudaStreamCreateWithPriority(&s->comm_stream, cudaStreamNonBlocking, hi));
    }
    CUDA_OK(cudaStreamCreateWithFlags(&s->bnd_stream, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&s->d2h_stream, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&s->ev_state, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&s->ev_d2h_all, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&s->ev_ghost, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&s->ev_bnd, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&s->ev_pack, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&s->ev_comm, cudaEventDisableTiming));
    for (int l = 0; l <= s->n_stages; ++l) s->Ulev.push_back(l);
    int h = -1;
    for (size_t i = 0; i < g_sims.size(); ++i) if (!g_sims[i]) { h = (int)i; break; }
    if (h < 0) { g_sims.emplace_back(); h = (int)g_sims.size() - 1; }
    g_sims[h] = std::move(s);
    return h;
}

int eb200_finalize(int sim)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    cudaSetDevice(s->cfg.device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (Peer& pr : s->peers) for (void* q : pr.opened) cudaIpcCloseMemHandle(q);
    for (void* p : s->allocs) cudaFree(p);
    for (auto& e : s->ev_pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    if (s->h_status) cudaFreeHost(s->h_status);
    if (s->comm_stream) { cudaStreamSynchronize(s->comm_stream); cudaStreamDestroy(s->comm_stream); }
    if (s->bnd_stream) { cudaStreamSynchronize(s->bnd_stream); cudaStreamDestroy(s->bnd_stream); }
    if (s->d2h_stream) { cudaStreamSynchronize(s->d2h_stream); cudaStreamDestroy(s->d2h_stream); }
    for (auto& b : s->blocks) if (b->ev_d2h) cudaEventDestroy(b->ev_d2h);
    if (s->ev_ghost) cudaEventDestroy(s->ev_ghost);
    if (s->ev_bnd) cudaEventDestroy(s->ev_bnd);
    if (s->ev_pack) cudaEventDestroy(s->ev_pack);
    if (s->ev_comm) cudaEventDestroy(s->ev_comm);
    if (s->stream) cudaStreamDestroy(s->stream);
    g_sims[sim].reset();
    return 0;
}

int eb200_block_create(int sim, int blk_id, int nic, int njc, int nkc, int owner_rank)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    if (s->committed) { set_err("block_create after commit"); return -1; }
    if (!s->threeD && nkc != 1) { set_err("nkc must be 1 in 2D"); return -1; }
    if (nic < EB_NG || njc < EB_NG || (s->threeD && nkc < EB_NG)) { set_err("Too few cells for ghost cell copies."); return -1; }
    for (auto& b : s->blocks) if (b->id == blk_id) { set_err("block %d declared twice", blk_id); return -1; }
    auto b = std::make_unique<Block>();
    b->id = blk_id; b->nic = nic; b->njc = njc; b->nkc = nkc; b->owner = owner_rank;
    b->local = (owner_rank == s->cfg.rank);
    b->NI = nic + 2 * EB_NG; b->NJ = njc + 2 * EB_NG; b->NK = s->threeD ? nkc + 2 * EB_NG : 1;
    b->kg = s->threeD ? EB_NG : 0;
    b->ncp = (long long)b->NI * b->NJ * b->NK;
    b->stride[0] = 1; b->stride[1] = b->NI; b->stride[2] = (long long)b->NI * b->NJ;
    s->blocks.push_back(std::move(b));
    return 0;
}

int eb200_block_set_geometry(int sim, int blk_id, const double* vol, const double* areaxy,
                             const double* len_i, const double* len_j, const double* len_k,
                             const double* const face[3])
{
    Sim* s = get_sim(sim); if (!s) return -1;
    Block* b = get_blk(s, blk_id); if (!b) return -1;
    if (!b->local) { set_err("geometry given for non-local block %d", blk_id); return -1; }
    if (s->committed) { set_err("set_geometry after commit"); return -1; }
    const long long n = b->ncp;
    if (!vol || !len_i || !len_j || !face) { set_err("null geometry array"); return -1; }
    if (!areaxy && s->cfg.axisymmetric) { set_err("areaxy required for axisymmetric"); return -1; }
    if (s->threeD && !len_k) { set_err("len_k required in 3D"); return -1; }
    for (int d = 0; d < s->cfg.dimensions; ++d)
        if (!face[d]) { set_err("face geometry missing for direction %d", d); return -1; }
    const double* lens[3] = { len_i, len_j, len_k };
    memset(&b->cartD, 0, sizeof b->cartD);
    b->cartesian = detect_cartesian(s, b, vol, areaxy, lens, face, b->cartD);
    if (!b->cartesian) {          // the general-metric path needs the arrays on the device: keep a copy until commit
        b->vol.assign(vol, vol + n);
        if (areaxy) b->areaxy.assign(areaxy, areaxy + n); else b->areaxy.assign(n, 0.0);
        b->len[0].assign(len_i, len_i + n); b->len[1].assign(len_j, len_j + n);
        if (s->threeD) b->len[2].assign(len_k, len_k + n);
        for (int d = 0; d < s->cfg.dimensions; ++d) b->face[d].assign(face[d], face[d] + 10 * n);
    }
    b->has_geometry = true;
    return 0;
}

int eb200_block_set_bc(int sim, int blk_id, int face, int kind, const double* params, int nparams,
                       int other_blk, int other_face, int orientation)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    Block* b = get_blk(s, blk_id); if (!b) return -1;
    if (s->committed) { set_err("set_bc after commit"); return -1; }
    if (face < 0 || face >= s->nfaces) { set_err("bad face %d", face); return -1; }
    if (kind < 0 || kind > EB200_BC_GHOST_PROFILE) { set_err("unknown bc kind %d", kind); return -1; }
    if (kind == EB200_BC_WALL_WITH_SLIP1) {
        const int nd = (face / 2 == 0) ? b->nic : ((face / 2 == 1) ? b->njc : b->nkc);
        if (nd < 4) { set_err("block %d: a boundary without ghost-cell data needs at least 4 cells along its normal (got %d)", blk_id, nd); return -1; }
    }
    BC& bc = b->bc[face];
    bc.kind = kind; bc.other_blk = other_blk; bc.other_face = other_face; bc.orientation = orientation;
    if (kind == EB200_BC_INFLOW_SUPERSONIC) {
        if (nparams != s->P.nprim || !params) { set_err("inflow FlowState needs %d values", s->P.nprim); return -1; }
        bc.params.assign(params, params + nparams);
    }
    if (kind == EB200_BC_GHOST_PROFILE) {
        const int d = face / 2;
        const int nn[3] = { b->nic, b->njc, b->nkc };
        const long long need = (long long)EB_NG * nn[(d + 1) % 3] * nn[(d + 2) % 3] * s->P.nprim;
        if (nparams != need || !params) { set_err("ghost profile of block %d face %d needs %lld values (got %d)", blk_id, face, need, nparams); return -1; }
        bc.params.assign(params, params + nparams);
    }
    if (kind == EB200_BC_OUTFLOW_FIXED_P || kind == EB200_BC_OUTFLOW_FIXED_PT) {
        const int need = (kind == EB200_BC_OUTFLOW_FIXED_P) ? 1 : 2;
        if (nparams != need || !params) { set_err("bc kind %d needs %d parameter(s)", kind, need); return -1; }
        bc.params.assign((size_t)s->P.nprim, 0.0);          // one slot of the parameter table: { p_outside, T_outside | -1 }
        bc.params[0] = params[0]; bc.params[1] = (need == 2) ? params[1] : -1.0;
    }
    if (kind == EB200_BC_EXCHANGE_FULL_FACE) {
        if (s->threeD && orientation != 0) { set_err("only orientation 0 is supported in 3D"); return -1; }
        if (other_face < 0 || other_face >= s->nfaces) { set_err("bad other_face"); return -1; }
    }
    return 0;
}

int eb200_block_set_face_map(int sim, int blk_id, int face, const int* src_ijk, long long n)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    Block* b = get_blk(s, blk_id); if (!b) return -1;
    if (s->committed) { set_err("set_face_map after commit"); return -1; }
    if (face < 0 || face >= s->nfaces) { set_err("bad face %d", face); return -1; }
    BC& bc = b->bc[face];
    if (bc.kind != EB200_BC_EXCHANGE_FULL_FACE) { set_err("block %d face %d: a cell map needs a full-face exchange", blk_id, face); return -1; }
    Block* ot = get_blk(s, bc.other_blk); if (!ot) return -1;
    const int d = face / 2;
    const int nn[3] = { b->nic, b->njc, b->nkc };
    const long long want = (long long)EB_NG * nn[(d + 1) % 3] * nn[(d + 2) % 3];
    if (n != want || !src_ijk) { set_err("block %d face %d: the cell map needs %lld entries", blk_id, face, want); return -1; }
    for (long long m = 0; m < n; ++m) {
        const int i = src_ijk[3 * m], j = src_ijk[3 * m + 1], k = src_ijk[3 * m + 2];
        if (i < 0 || i >= ot->nic || j < 0 || j >= ot->njc || k < 0 || k >= ot->nkc) {
            set_err("block %d face %d: map entry %lld (%d,%d,%d) is outside block %d", blk_id, face, m, i, j, k, ot->id); return -1;
        }
    }
    bc.map.assign(src_ijk, src_ijk + 3 * n);
    return 0;
}

int eb200_set_exchange(int sim, eb200_exchange_fn fn, void* user)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    s->exchange = fn; s->exchange_user = user;
    return 0;
}

int eb200_commit(int sim)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    if (s->committed) { set_err("commit called twice"); return -1; }
    CUDA_OK(cudaSetDevice(s->cfg.device));
    // 1. arena offsets
    long long total = 0;
    for (auto& b : s->blocks) {
        if (!b->local) continue;
        if (!b->has_geometry) { set_err("local block %d has no geometry", b->id); return -1; }
        b->cell0 = total; b->local_index = (int)s->local.size();
        total += (b->ncp + 31) / 32 * 32;
        s->local.push_back(b.get());
    }
    if (s->local.empty()) { set_err("no local blocks on rank %d", s->cfg.rank); return -1; }
    if (total >= (1LL << 31)) { set_err("too many cells on one device (%lld): indices are 32-bit", total); return -1; }
    s->P.total = total;
    const int nprim = s->P.nprim, ncq = s->P.ncq, ns = s->n_stages, dims = s->cfg.dimensions;
    // 2. descriptors + Cartesian detection + tiling
    s->hdesc.resize(s->local.size());
    bool any_general = false, any_cart = false;
    long long cells_total = 0;
    for (size_t n = 0; n < s->local.size(); ++n) {
        Block* b = s->local[n];
        EbBlockDesc& D = s->hdesc[n];
        if (b->cartesian) D = b->cartD; else memset(&D, 0, sizeof D);
        D.nic = b->nic; D.njc = b->njc; D.nkc = b->nkc; D.NI = b->NI; D.NJ = b->NJ; D.NK = b->NK; D.kg = b->kg;
        D.cell0 = b->cell0; for (int d = 0; d < 3; ++d) D.stride[d] = b->stride[d];
        D.outflow_flux_faces = 0;
        D.noghost_faces = 0;
        for (int f = 0; f < 6; ++f) {
            D.bc_kind[f] = b->bc[f].kind;
            if (b->bc[f].kind == EB200_BC_OUTFLOW_SIMPLE_FLUX) D.outflow_flux_faces |= (1 << f);
            if (f < s->nfaces && b->bc[f].kind == EB200_BC_WALL_WITH_SLIP1) D.noghost_faces |= (1 << f);
        }
        D.push_mask = 0;
        for (int f = 0; f < 6; ++f) D.push_off[f] = 0;
        for (int f = 0; f < s->nfaces; ++f) {
            const Block* ot = nullptr;
            if (!face_pushes(s, b, f, &ot)) continue;
            const int d = f / 2;
            const long long nd = (d == 0) ? b->nic : (d == 1) ? b->njc : b->nkc;
            D.push_mask |= (1 << f);
            D.push_off[f] = (ot->cell0 - b->cell0) + ((f & 1) ? -nd : nd) * b->stride[d];
        }
        D.cartesian = b->cartesian ? 1 : 0;
        (b->cartesian ? any_cart : any_general) = true;
        cells_total += (long long)b->nic * b->njc * b->nkc;
    }
    // a boundary without ghost-cell data anywhere in the job (any rank: every rank must pick the same kernels for the
    // multi-GPU result to equal the single-GPU one): the generic kernel, which knows the one-sided stencils
    bool one_sided = false;
    for (auto& b : s->blocks) for (int f = 0; f < s->nfaces; ++f) if (b->bc[f].kind == EB200_BC_WALL_WITH_SLIP1) one_sided = true;
    if (one_sided && any_cart) {
        set_err("a job with EB200_BC_WALL_WITH_SLIP1 walls must be initialised with eb200_config.reserved_i[0] = 1: "
                "the one-sided stencils are served by the general-metric kernel");
        return -1;
    }
    if (one_sided && s->P.lmr) { set_err("solver_variant = lmr has no one-sided stencils on this path (lmr/onedinterp.d has l2r2 and l3r3 only)"); return -1; }
    const bool force_generic = s->cfg.reserved_i[1] == 1 || one_sided || s->P.lmr;     // the variant formulas live in the generic kernel
    s->which = (any_cart ? 1 : 0) | (any_general ? 2 : 0) | (force_generic ? 4 : 0) | (s->cfg.reserved_i[1] == 2 ? 8 : 0);
    // Which fused kernel runs a block decides its tiling.  The cell-centred kernel (flux_kernel_v3.cuh) takes the
    // uniform-Cartesian blocks of the reference's default configuration (same conditions as in flux_inst.cu) when
    // their tiles can be staged by TMA.
    const bool tma_ok = tma_available(s);
    const bool v3_config = tma_ok && s->cfg.reserved_i[1] == 0 && !one_sided && !s->P.lmr && s->cfg.gas_model == EB200_GAS_IDEAL &&
                           s->P.interpolation_order == 2 && s->P.apply_limiter != 0 && s->P.thermo_interp == EB200_INTERP_RHOU;
    for (size_t n = 0; n < s->local.size(); ++n) s->hdesc[n].v3 = (v3_config && s->hdesc[n].cartesian) ? 1 : 0;
    {
        // k-chunking (3D): a CTA marches over `chunk` planes of its tile.  Aim at >= 8 waves of CTAs
        // (148 SMs x 2 resident CTAs) so that the last, partly filled wave costs little; every chunk
        // costs a start-up (three planes to stage before the first face) and four extra planes of tile
        // traffic, so chunks are kept >= 32 planes.
        long long tiles_plane = 0;
        for (size_t n = 0; n < s->local.size(); ++n) {
            EbBlockDesc& D = s->hdesc[n];
            const int ty = D.v3 ? EB_V3_TY : EB_TILE_Y;
            D.tiles_i = (D.nic + 31) / 32; D.tiles_j = (D.njc + ty - 1) / ty;
            tiles_plane += (long long)D.tiles_i * D.tiles_j;
        }
        // (EB200_CHUNK_WAVES / EB200_CHUNK_MIN: development knobs)
        const long long waves = getenv("EB200_CHUNK_WAVES") ? atoll(getenv("EB200_CHUNK_WAVES")) : 8;
        const long long chunk_min = getenv("EB200_CHUNK_MIN") ? atoll(getenv("EB200_CHUNK_MIN")) : 32;
        const long long want = 148LL * 2 * waves;
        long long tile0 = 0;
        for (size_t n = 0; n < s->local.size(); ++n) {
            EbBlockDesc& D = s->hdesc[n];
            int chunk = D.nkc;
            if (s->threeD && tiles_plane < want) {
                long long nch = (want + tiles_plane - 1) / tiles_plane;
                chunk = (int)std::max<long long>(chunk_min, (D.nkc + nch - 1) / nch);
                chunk = std::min(chunk, D.nkc);
            }
            D.chunk_m = s->threeD ? chunk : 1;
            D.tiles_m = s->threeD ? (D.nkc + chunk - 1) / chunk : 1;
            D.tile0 = tile0;
            tile0 += (long long)D.tiles_i * D.tiles_j * D.tiles_m;
        }
        s->ncta = tile0;
    }
    // 3. device memory
    for (int p = 0; p < 3; ++p) if (dev_alloc(s, &s->A.prim[p], (size_t)nprim * total)) return -100;
    for (int l = 0; l <= ns; ++l) if (dev_alloc(s, &s->A.U[l], (size_t)ncq * total)) return -100;
    for (int l = 0; l < ns - 1; ++l) if (dev_alloc(s, &s->A.dUdt[l], (size_t)ncq * total)) return -100;   // residuals needed by later stages
    if (s->P.shock_detect) {
        if (dev_alloc(s, &s->A.S, (size_t)total)) return -100;
        for (int d = 0; d < dims; ++d) if (dev_alloc(s, &s->A.Sf[d], (size_t)total)) return -100;
    }
    if (any_general) {
        if (dev_alloc(s, &s->A.vol, (size_t)total)) return -100;
        if (s->cfg.axisymmetric) if (dev_alloc(s, &s->A.areaxy, (size_t)total)) return -100;
        for (int d = 0; d < dims; ++d) {
            if (dev_alloc(s, &s->A.len[d], (size_t)total)) return -100;
            if (dev_alloc(s, &s->A.face[d], (size_t)10 * total)) return -100;
        }
        for (Block* b : s->local) {
            if (b->cartesian) continue;
            const size_t bytes = (size_t)b->ncp * sizeof(double);
            CUDA_OK(cudaMemcpyAsync(s->A.vol + b->cell0, b->vol.data(), bytes, cudaMemcpyHostToDevice, s->stream));
            if (s->cfg.axisymmetric) CUDA_OK(cudaMemcpyAsync(s->A.areaxy + b->cell0, b->areaxy.data(), bytes, cudaMemcpyHostToDevice, s->stream));
            for (int d = 0; d < dims; ++d) {
                CUDA_OK(cudaMemcpyAsync(s->A.len[d] + b->cell0, b->len[d].data(), bytes, cudaMemcpyHostToDevice, s->stream));
                for (int m = 0; m < 10; ++m)
                    CUDA_OK(cudaMemcpyAsync(s->A.face[d] + (long long)m * total + b->cell0, b->face[d].data() + (long long)m * b->ncp,
                                            bytes, cudaMemcpyHostToDevice, s->stream));
            }
        }
    }
    if (dev_upload(s, &s->d_desc, s->hdesc)) return -100;
    if (build_tensor_maps(s)) return -100;
    {
        std::vector<EbGas> gv(1, s->hgas);
        if (dev_upload(s, &s->d_gas, gv)) return -100;
    }
    if (dev_alloc(s, &s->d_status, 8)) return -100;
    if (dev_alloc(s, &s->d_red, 2 * s->local.size())) return -100;
    if (dev_alloc(s, &s->d_last, s->local.size())) return -100;
    CUDA_OK(cudaMallocHost((void**)&s->h_status, 8 * sizeof(int)));
    // 4. ghost-cell work lists
    std::vector<EbCopyItem> copy, copy_pushed; std::vector<EbReflectItem> refl; std::vector<EbFillItem> fill;
    std::vector<double> params;
    std::map<int, Peer> peers;
    struct Key { int blk, face; };
    // outgoing halo: for every face of a REMOTE block that is connected to one of my blocks,
    // gather my interior cells in that block's ghost enumeration order.
    std::map<int, std::vector<std::pair<std::pair<int, int>, std::vector<int>>>> send_sets, recv_sets;
    for (Block* b : s->local) {
        for (int f = 0; f < s->nfaces; ++f) {
            BC& bc = b->bc[f];
            const int d = f / 2;
            if (bc.kind == EB200_BC_EXCHANGE_FULL_FACE) {
                Block* ot = get_blk(s, bc.other_blk);
                if (!ot) { set_err("block %d face %d: neighbour block %d was not declared", b->id, f, bc.other_blk); return -1; }
                std::vector<int> recv_list, send_list;
                std::vector<std::pair<long long, int>> recv_keyed, send_keyed;   // (receiver's cell index in its block, arena index)
                int err = 0;
                // the neighbour's flux kernel writes these ghost cells itself when it pushes through its face
                std::vector<EbCopyItem>& copies = (ot->local && face_pushes(s, ot, bc.other_face, nullptr)) ? copy_pushed : copy;
                long long m = 0;
                for_face_ghosts(s, b, f, [&](int t1, int t2, int layer, long long, long long ghost, long long, long long) {
                    int ijk[3];
                    if (!bc.map.empty()) { ijk[0] = bc.map[3 * m]; ijk[1] = bc.map[3 * m + 1]; ijk[2] = bc.map[3 * m + 2]; }
                    else if (full_face_source(s, f, ot, bc.other_face, t1, t2, layer, ijk)) { err = 1; return; }
                    ++m;
                    if (ot->local) copies.push_back({ (int)(b->cell0 + ghost), (int)(ot->cell0 + ot->cidx(ijk[0], ijk[1], ijk[2])) });
                    else recv_keyed.push_back({ ghost, (int)(b->cell0 + ghost) });
                });
                if (err) return -1;
                if (!ot->local) {
                    // what the neighbour needs from me: its ghost cells behind (ot, other_face), in ITS order
                    // (the neighbour's own boundary condition is only declared here when it carries a cell map)
                    const BC& obc = ot->bc[bc.other_face];
                    const bool omap = obc.kind == EB200_BC_EXCHANGE_FULL_FACE && obc.other_blk == b->id && !obc.map.empty();
                    if (!bc.map.empty() && !omap) {
                        set_err("block %d face %d has a cell map but its neighbour, block %d face %d, was declared without one", b->id, f, ot->id, bc.other_face);
                        return -1;
                    }
                    long long mo = 0;
                    for_face_ghosts(s, ot, bc.other_face, [&](int t1, int t2, int layer, long long, long long ghost, long long, long long) {
                        int ijk[3];
                        if (omap) { ijk[0] = obc.map[3 * mo]; ijk[1] = obc.map[3 * mo + 1]; ijk[2] = obc.map[3 * mo + 2]; }
                        else if (full_face_source(s, bc.other_face, b, f, t1, t2, layer, ijk)) { err = 1; return; }
                        ++mo;
                        send_keyed.push_back({ ghost, (int)(b->cell0 + b->cidx(ijk[0], ijk[1], ijk[2])) });
                    });
                    if (err) return -1;
                    // wire order of a face set: ascending cell index of the RECEIVING block (both sides can compute it)
                    std::sort(recv_keyed.begin(), recv_keyed.end());
                    std::sort(send_keyed.begin(), send_keyed.end());
                    for (auto& e : recv_keyed) recv_list.push_back(e.second);
                    for (auto& e : send_keyed) send_list.push_back(e.second);
                    recv_sets[ot->owner].push_back({ { b->id, f }, recv_list });
                    send_sets[ot->owner].push_back({ { ot->id, bc.other_face }, send_list });
                }
            } else if (bc.kind == EB200_BC_WALL_WITH_SLIP1) {
                // no ghost-cell data: nothing fills these cells and nothing reads them
            } else if (bc.kind == EB200_BC_WALL_WITH_SLIP) {
                for_face_ghosts(s, b, f, [&](int, int, int, long long cf, long long ghost, long long mirror, long long) {
                    refl.push_back({ (int)(b->cell0 + ghost), (int)(b->cell0 + mirror), (int)(b->cell0 + cf), b->local_index * 4 + d });
                });
            } else if (bc.kind == EB200_BC_INFLOW_SUPERSONIC) {
                bc.param_index = (int)(params.size() / nprim);
                params.insert(params.end(), bc.params.begin(), bc.params.end());
                for_face_ghosts(s, b, f, [&](int, int, int, long long, long long ghost, long long, long long) {
                    fill.push_back({ (int)(b->cell0 + ghost), bc.param_index });
                });
            } else if (bc.kind == EB200_BC_GHOST_PROFILE) {
                // one row of the parameter table per ghost cell, in the enumeration order of for_face_ghosts
                bc.param_index = (int)(params.size() / nprim);
                params.insert(params.end(), bc.params.begin(), bc.params.end());
                int row = bc.param_index;
                for_face_ghosts(s, b, f, [&](int, int, int, long long, long long ghost, long long, long long) {
                    fill.push_back({ (int)(b->cell0 + ghost), row++ });
                });
            } else if (bc.kind == EB200_BC_OUTFLOW_FIXED_P || bc.kind == EB200_BC_OUTFLOW_FIXED_PT) {
                bc.param_index = (int)(params.size() / nprim);
                params.insert(params.end(), bc.params.begin(), bc.params.end());
                for_face_ghosts(s, b, f, [&](int, int, int, long long, long long ghost, long long mirror, long long) {
                    refl.push_back({ (int)(b->cell0 + ghost), (int)(b->cell0 + mirror), bc.param_index, b->local_index * 4 + 3 });
                });
            } else {   // zero-order extrapolation (both outflow kinds)
                for_face_ghosts(s, b, f, [&](int, int, int, long long, long long ghost, long long, long long first) {
                    copy.push_back({ (int)(b->cell0 + ghost), (int)(b->cell0 + first) });
                });
            }
        }
    }
    // order the work by destination address: neighbouring threads then touch neighbouring ghost cells
    std::sort(copy.begin(), copy.end(), [](const EbCopyItem& a, const EbCopyItem& b) { return a.dst < b.dst; });
    std::sort(copy_pushed.begin(), copy_pushed.end(), [](const EbCopyItem& a, const EbCopyItem& b) { return a.dst < b.dst; });
    std::sort(refl.begin(), refl.end(), [](const EbReflectItem& a, const EbReflectItem& b) { return a.dst < b.dst; });
    std::sort(fill.begin(), fill.end(), [](const EbFillItem& a, const EbFillItem& b) { return a.dst < b.dst; });
    s->ncopy = (long long)copy.size(); s->nrefl = (long long)refl.size(); s->nfill = (long long)fill.size();
    copy.insert(copy.end(), copy_pushed.begin(), copy_pushed.end());
    s->ncopy_full = (long long)copy.size();
    if (getenv("EB200_DEBUG"))
        fprintf(stderr, "[eb200] rank %d: ghost work per stage: %lld copies (+%lld pushed by the flux kernel), %lld reflections, %lld fills; tma=%d\n",
                s->cfg.rank, s->ncopy, s->ncopy_full - s->ncopy, s->nrefl, s->nfill, s->d_tmaps ? 1 : 0);
    if (dev_upload(s, &s->d_copy, copy)) return -100;
    if (dev_upload(s, &s->d_refl, refl)) return -100;
    if (dev_upload(s, &s->d_fill, fill)) return -100;
    if (dev_upload(s, &s->d_params, params)) return -100;
    // peers: both sides order their sets by (receiving block id, receiving face)
    for (auto& kv : recv_sets) {
        Peer p; p.rank = kv.first;
        auto rs = kv.second; auto ss = send_sets[kv.first];
        std::sort(rs.begin(), rs.end(), [](auto& a, auto& b) { return a.first < b.first; });
        std::sort(ss.begin(), ss.end(), [](auto& a, auto& b) { return a.first < b.first; });
        for (auto& e : rs) p.recv_idx.insert(p.recv_idx.end(), e.second.begin(), e.second.end());
        for (auto& e : ss) p.send_idx.insert(p.send_idx.end(), e.second.begin(), e.second.end());
        if (dev_upload(s, &p.d_send_idx, p.send_idx)) return -100;
        if (dev_upload(s, &p.d_recv_idx, p.recv_idx)) return -100;
        if (dev_alloc(s, &p.d_send, p.send_idx.size() * (size_t)(nprim + 1))) return -100;
        if (dev_alloc(s, &p.d_recv, p.recv_idx.size() * (size_t)(nprim + 1))) return -100;
        s->peers.push_back(std::move(p));
    }
    if (!s->peers.empty()) {
        // one allocation (one IPC handle) with what the peers need to see: the flags they raise and, per peer, the
        // arena indices of the ghost cells they fill (wire order)
        if (s->peers.size() > EB_P2P_MAXPEERS) { set_err("more than %d peer ranks", EB_P2P_MAXPEERS); return -1; }
        size_t bytes = 2 * EB_P2P_MAXPEERS * sizeof(unsigned long long);     // "landed" and "ready" flags
        for (Peer& p : s->peers) { p.recv_off = (long long)bytes; bytes += (p.recv_idx.size() * sizeof(int) + 255) / 256 * 256; }
        bytes = std::max<size_t>(bytes, (size_t)2 << 20);     // a whole allocation granule: nothing else shares the mapping
        if (dev_alloc(s, &s->d_shared, bytes)) return -100;
        s->shared_bytes = bytes;
        for (Peer& p : s->peers)
            if (!p.recv_idx.empty())
                CUDA_OK(cudaMemcpyAsync(s->d_shared + p.recv_off, p.recv_idx.data(), p.recv_idx.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    }
    if (!s->peers.empty()) {
        // tiles whose stencils reach ghost cells filled by another rank go last (after the halo arrived)
        std::vector<int> t_int, t_bnd;
        for (size_t n = 0; n < s->local.size(); ++n) {
            const Block* b = s->local[n];
            const EbBlockDesc& D = s->hdesc[n];
            bool remote[6] = { false, false, false, false, false, false };
            for (int f = 0; f < s->nfaces; ++f)
                if (b->bc[f].kind == EB200_BC_EXCHANGE_FULL_FACE) { Block* ot = get_blk(s, b->bc[f].other_blk); remote[f] = ot && !ot->local; }
            // a tile is "boundary" when its stencils (two cells beyond its far-edge faces) reach ghost cells that another
            // rank fills: also the last-but-one tile when the last one is narrower than two cells
            const int ty = D.v3 ? EB_V3_TY : EB_TILE_Y;
            for (int tm = 0; tm < D.tiles_m; ++tm) for (int tj = 0; tj < D.tiles_j; ++tj) for (int ti = 0; ti < D.tiles_i; ++ti) {
                const bool bnd = (remote[EB200_WEST] && ti == 0) || (remote[EB200_EAST] && (ti + 1) * 32 + 2 > D.nic) ||
                                 (remote[EB200_SOUTH] && tj == 0) || (remote[EB200_NORTH] && (tj + 1) * ty + 2 > D.njc) ||
                                 (remote[EB200_BOTTOM] && tm == 0) || (remote[EB200_TOP] && (tm + 1) * D.chunk_m + 2 > D.nkc);
                const int id = (int)(D.tile0 + ((long long)tm * D.tiles_j + tj) * D.tiles_i + ti);
                (bnd ? t_bnd : t_int).push_back(id);
            }
        }
        s->n_tiles_int = (long long)t_int.size(); s->n_tiles_bnd = (long long)t_bnd.size();
        if (dev_upload(s, &s->d_tiles_int, t_int)) return -100;
        if (dev_upload(s, &s->d_tiles_bnd, t_bnd)) return -100;
    }
    CUDA_OK(cudaStreamSynchronize(s->stream));
    // host geometry copies are no longer needed
    for (Block* b : s->local) {
        std::vector<double>().swap(b->vol); std::vector<double>().swap(b->areaxy);
        for (int d = 0; d < 3; ++d) { std::vector<double>().swap(b->len[d]); std::vector<double>().swap(b->face[d]); }
    }
    s->committed = true;
    (void)cells_total;
    return 0;
}

// the deferred verdict on the uploads since the last check (status word 7)
static int check_uploads(Sim* s, bool status_is_fresh)
{
    if (!s->upload_unchecked) return 0;
    if (!status_is_fresh && read_status(s)) return -100;
    s->upload_unchecked = false;
    if (s->h_status[7]) {
        CUDA_OK(cudaMemsetAsync(s->d_status + 7, 0, sizeof(int), s->stream));
        set_err("decode_conserved failed for a FlowState given to eb200_upload_flow");
        return -1;
    }
    return 0;
}

int eb200_upload_flow(int sim, int blk_id, const double* const* prims, int nprims)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    Block* b = get_blk(s, blk_id); if (!b) return -1;
    if (!s->committed || !b->local) { set_err("upload_flow needs a committed local block"); return -1; }
    // short form (single-species gas): rho, u, velx, vely, velz -- the independent variables; p, T and a follow from
    // them on the device (the encode + decode pass below computes them anyway), 5/8 of the bytes over PCIe
    const bool short_form = (nprims == EB200_NPRIM_SHORT && s->P.nprim == EB200_NPRIM_BASE);
    if (nprims != s->P.nprim && !short_form) { set_err("expected %d primitive arrays", s->P.nprim); return -1; }
    CUDA_OK(cudaSetDevice(s->cfg.device));
    const size_t bytes = (size_t)b->ncp * sizeof(double);
    double* prim = s->A.prim[s->cur];
    s->ghosts_stale = true;
    s->undo_cur = -1;
    static const int short_field[EB200_NPRIM_SHORT] = { 0, 1, 5, 6, 7 };
    if (b->d2h_pending) {      // the block's last asynchronous download reads what this upload rewrites
        CUDA_OK(cudaStreamWaitEvent(s->stream, b->ev_d2h, 0));
        b->d2h_pending = false;
    }
    for (int v = 0; v < nprims; ++v) {
        const int f = short_form ? short_field[v] : v;
        CUDA_OK(cudaMemcpyAsync(prim + (long long)f * s->P.total + b->cell0, prims[v], bytes, cudaMemcpyHostToDevice, s->stream));
    }
    // no host synchronisation here: a cell that cannot be decoded raises status word 7, which the next eb200_step /
    // eb200_run_steps / eb200_compute_dt reports (uploading 64 blocks then costs 64 copies, not 64 round trips)
    MODE_CALL(s, launch_decode, s->P, s->cfg.gas_model, s->d_gas, s->hdesc[b->local_index], prim, prim, s->A.U[s->Ulev[0]], 1, s->d_status + 7, s->stream);
    CUDA_OK(cudaGetLastError());
    s->upload_unchecked = true;
    return 0;
}

int eb200_download_flow(int sim, int blk_id, double* const* prims, int nprims)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    Block* b = get_blk(s, blk_id); if (!b) return -1;
    if (!s->committed || !b->local) { set_err("download_flow needs a committed local block"); return -1; }
    if (nprims != s->P.nprim) { set_err("expected %d primitive arrays", s->P.nprim); return -1; }
    CUDA_OK(cudaSetDevice(s->cfg.device));
    const size_t bytes = (size_t)b->ncp * sizeof(double);
    const double* prim = s->A.prim[s->cur];
    for (int v = 0; v < nprims; ++v)
        CUDA_OK(cudaMemcpyAsync(prims[v], prim + (long long)v * s->P.total + b->cell0, bytes, cudaMemcpyDeviceToHost, s->stream));
    CUDA_OK(cudaStreamSynchronize(s->stream));
    return 0;
}

int eb200_probe_cells(int sim, int n, const int* blk_ids, const int* ijk, double* out, int nprims)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    if (!s->committed) { set_err("probe_cells needs a committed simulation"); return -1; }
    if (nprims != s->P.nprim) { set_err("expected %d primitive variables", s->P.nprim); return -1; }
    CUDA_OK(cudaSetDevice(s->cfg.device));
    const double* prim = s->A.prim[s->cur];
    for (int m = 0; m < n; ++m) {
        Block* b = get_blk(s, blk_ids[m]); if (!b) return -1;
        const int i = ijk[3 * m], j = ijk[3 * m + 1], k = ijk[3 * m + 2];
        if (!b->local || i < 0 || i >= b->nic || j < 0 || j >= b->njc || k < 0 || k >= b->nkc) {
            set_err("probe %d: cell (%d,%d,%d) is not an interior cell of a local block %d", m, i, j, k, blk_ids[m]);
            return -1;
        }
        const long long c = b->cell0 + b->cidx(i, j, k);
        // one strided copy: nprims rows of one double, row pitch = one field of the arena
        CUDA_OK(cudaMemcpy2DAsync(out + (size_t)m * nprims, sizeof(double), prim + c, (size_t)s->P.total * sizeof(double),
                                  sizeof(double), (size_t)nprims, cudaMemcpyDeviceToHost, s->stream));
    }
    CUDA_OK(cudaStreamSynchronize(s->stream));
    return 0;
}

int eb200_download_conserved(int sim, int blk_id, double* const* U, int ncq)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    Block* b = get_blk(s, blk_id); if (!b) return -1;
    if (!s->committed || !b->local) { set_err("download_conserved needs a committed local block"); return -1; }
    if (ncq != s->P.ncq) { set_err("expected %d conserved arrays", s->P.ncq); return -1; }
    CUDA_OK(cudaSetDevice(s->cfg.device));
    const size_t bytes = (size_t)b->ncp * sizeof(double);
    const double* Ud = s->A.U[s->Ulev[0]];
    for (int q = 0; q < ncq; ++q)
        CUDA_OK(cudaMemcpyAsync(U[q], Ud + (long long)q * s->P.total + b->cell0, bytes, cudaMemcpyDeviceToHost, s->stream));
    CUDA_OK(cudaStreamSynchronize(s->stream));
    return 0;
}

// Asynchronous form of eb200_download_conserved: the copies go to a second stream, behind everything enqueued so far,
// and the call returns at once.  The next eb200_upload_flow of the same block waits (on the device) for them, every
// other entry point that changes the state waits for all of them; eb200_wait_downloads is the host-side wait.  With
// page-locked host arrays the device -> host copies of one step then run beside the host -> device copies of the
// next one (PCIe is full duplex).
int eb200_download_conserved_async(int sim, int blk_id, double* const* U, int ncq)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    Block* b = get_blk(s, blk_id); if (!b) return -1;
    if (!s->committed || !b->local) { set_err("download_conserved_async needs a committed local block"); return -1; }
    if (ncq != s->P.ncq) { set_err("expected %d conserved arrays", s->P.ncq); return -1; }
    CUDA_OK(cudaSetDevice(s->cfg.device));
    if (!b->ev_d2h) CUDA_OK(cudaEventCreateWithFlags(&b->ev_d2h, cudaEventDisableTiming));
    CUDA_OK(cudaEventRecord(s->ev_state, s->stream));
    CUDA_OK(cudaStreamWaitEvent(s->d2h_stream, s->ev_state, 0));
    const size_t bytes = (size_t)b->ncp * sizeof(double);
    const double* Ud = s->A.U[s->Ulev[0]];
    for (int q = 0; q < ncq; ++q)
        CUDA_OK(cudaMemcpyAsync(U[q], Ud + (long long)q * s->P.total + b->cell0, bytes, cudaMemcpyDeviceToHost, s->d2h_stream));
    CUDA_OK(cudaEventRecord(b->ev_d2h, s->d2h_stream));
    CUDA_OK(cudaEventRecord(s->ev_d2h_all, s->d2h_stream));
    b->d2h_pending = true;
    s->d2h_inflight = true;
    return 0;
}

int eb200_wait_downloads(int sim)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    if (!s->committed) { set_err("not committed"); return -1; }
    CUDA_OK(cudaSetDevice(s->cfg.device));
    CUDA_OK(cudaStreamSynchronize(s->d2h_stream));
    s->d2h_inflight = false;
    for (Block* b : s->local) b->d2h_pending = false;
    return 0;
}

// the compute stream does not overwrite conserved quantities that an asynchronous download is still reading
static int order_after_downloads(Sim* s)
{
    if (s->d2h_inflight) {
        CUDA_OK(cudaStreamWaitEvent(s->stream, s->ev_d2h_all, 0));
        s->d2h_inflight = false;
    }
    return 0;
}

int eb200_compute_dt(int sim, double dt_current, double cfl_value, int check_cfl, double out[3])
{
    Sim* s = get_sim(sim); if (!s) return -1;
    if (!s->committed) { set_err("not committed"); return -1; }
    CUDA_OK(cudaSetDevice(s->cfg.device));
    if (check_uploads(s, false)) return -1;
    double cfl_allow;
    switch (s->n_stages) { case 1: cfl_allow = 0.9; break; case 2: cfl_allow = 1.2; break; case 3: cfl_allow = 1.6; break; default: cfl_allow = 0.9; }
    const double cfl_adjust = 0.5;
    double dt_allow_g = 1.7976931348623157e308, cfl_max_g = 0.0;
    // one launch for all local blocks; per-block results so that the reference's per-block clamp applies
    const size_t nl = s->local.size();
    std::vector<unsigned long long> red(2 * nl);
    std::vector<double> last(nl);
    long long max_cells = 0;
    for (size_t n = 0; n < nl; ++n) {
        red[2 * n] = 0x7ff0000000000000ULL; red[2 * n + 1] = 0ULL;
        max_cells = std::max(max_cells, (long long)s->local[n]->nic * s->local[n]->njc * s->local[n]->nkc);
    }
    CUDA_OK(cudaMemcpyAsync(s->d_red, red.data(), 2 * nl * sizeof(unsigned long long), cudaMemcpyHostToDevice, s->stream));
    MODE_CALL(s, launch_signal, s->P, s->d_desc, (int)nl, max_cells, s->A, s->A.prim[s->cur], dt_current, cfl_value, s->d_red, s->d_last, s->stream);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(red.data(), s->d_red, 2 * nl * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
    CUDA_OK(cudaMemcpyAsync(last.data(), s->d_last, nl * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CUDA_OK(cudaStreamSynchronize(s->stream));
    for (size_t n = 0; n < nl; ++n) {
        double dt_allow, cfl_max;
        memcpy(&dt_allow, &red[2 * n], 8); memcpy(&cfl_max, &red[2 * n + 1], 8);
        if (check_cfl && (cfl_max < 0.0 || cfl_max > cfl_allow)) {        // fluidblock.d:1072-1082
            cfl_max = cfl_adjust * cfl_allow;
            dt_allow = cfl_max / last[n];
        }
        dt_allow_g = std::min(dt_allow_g, dt_allow);
        cfl_max_g = std::max(cfl_max_g, cfl_max);
    }
    out[0] = dt_allow_g; out[1] = cfl_max_g; out[2] = 0.0;
    return 0;
}

static int finish_steps(Sim* s, int cur0, int* n_bad_cells, bool can_restore)
{
    if (read_status(s)) return -100;
    if (check_uploads(s, true)) return -1;
    const int ns = s->n_stages;
    int bad = s->h_status[ns];
    int worst = 0;
    for (int st = 1; st <= ns; ++st) worst = std::max(worst, s->h_status[st]);
    if (n_bad_cells) *n_bad_cells = bad;
    if (s->h_status[0] != 0) {
        if (!can_restore) { set_err("a cell could not be decoded during run_steps"); return -3; }
        // U[0] and the start-of-step FlowStates (buffer cur0) were never written: just point back
        s->cur = cur0;
        return 1;
    }
    if (worst > s->cfg.max_invalid_cells) {
        set_err("Too many bad cells during explicit gasdynamic update (%d).", worst);
        return -2;
    }
    return 0;
}

int eb200_step(int sim, double t0, double dt, int* n_bad_cells)
{
    (void)t0;
    Sim* s = get_sim(sim); if (!s) return -1;
    if (!s->committed) { set_err("not committed"); return -1; }
    CUDA_OK(cudaSetDevice(s->cfg.device));
    if (order_after_downloads(s)) return -100;
    CUDA_OK(cudaMemsetAsync(s->d_status, 0, 7 * sizeof(int), s->stream));
    const int cur0 = s->cur;
    int rc = enqueue_step(s, dt);
    if (rc) return rc;
    rc = finish_steps(s, cur0, n_bad_cells, true);
    s->undo_cur = -1;
    if (rc == 0) { std::swap(s->Ulev[0], s->Ulev[s->n_stages]); s->undo_cur = cur0; }     // :1557-1561 swap(U[0], U[end])
    if (s->ev_used > 2048) drain_flux_events(s);
    return rc;
}

int eb200_undo_step(int sim)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    if (s->undo_cur < 0) { set_err("undo_step: no successful step to take back"); return -1; }
    // the start-of-step FlowStates (buffer undo_cur) and the old U[0] were never written: point back at them
    std::swap(s->Ulev[0], s->Ulev[s->n_stages]);
    s->cur = s->undo_cur;
    s->undo_cur = -1;
    return 0;
}

int eb200_run_steps(int sim, double t0, double dt, int nsteps, int* n_bad_cells)
{
    (void)t0;
    Sim* s = get_sim(sim); if (!s) return -1;
    if (!s->committed) { set_err("not committed"); return -1; }
    s->undo_cur = -1;
    CUDA_OK(cudaSetDevice(s->cfg.device));
    if (order_after_downloads(s)) return -100;
    CUDA_OK(cudaMemsetAsync(s->d_status, 0, 7 * sizeof(int), s->stream));
    for (int n = 0; n < nsteps; ++n) {
        int rc = enqueue_step(s, dt);
        if (rc) return rc;
        std::swap(s->Ulev[0], s->Ulev[s->n_stages]);
        if (s->ev_used > 2048) drain_flux_events(s);
    }
    return finish_steps(s, s->cur, n_bad_cells, false);
}

long long eb200_kernel_launches(int sim)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    return s->launches;
}

int eb200_flux_kernel_time(int sim, int reset, double* ms, long long* launches)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    CUDA_OK(cudaSetDevice(s->cfg.device));
    CUDA_OK(cudaStreamSynchronize(s->stream));
    if (drain_flux_events(s)) return -100;
    if (ms) *ms = s->flux_ms_acc;
    if (launches) *launches = s->flux_launches;
    if (reset) { s->flux_ms_acc = 0.0; s->flux_launches = 0; }
    return 0;
}

// Test hook (declared in include/eb200.h): reconstruction + flux for a batch of independent faces.
int eb200_debug_face_flux(int sim, int nfaces, const double* cells, const double* len, const double* geo,
                          double* F, int* ok)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    if (nfaces <= 0) return 0;
    CUDA_OK(cudaSetDevice(s->cfg.device));
    const int nprim = s->P.nprim, ncq = s->P.ncq;
    const long long total = 4LL * nfaces;
    std::vector<double> hprim((size_t)nprim * total), hlen((size_t)total), hface((size_t)10 * total, 0.0);
    for (int n = 0; n < nfaces; ++n) {
        for (int m = 0; m < 4; ++m) {
            for (int v = 0; v < nprim; ++v) hprim[(size_t)v * total + 4 * n + m] = cells[((size_t)n * 4 + m) * nprim + v];
            hlen[4 * n + m] = len[(size_t)n * 4 + m];
        }
        for (int g = 0; g < 10; ++g) hface[(size_t)g * total + 4 * n + 2] = geo[(size_t)n * 10 + g];
    }
    double *dprim = nullptr, *dlen = nullptr, *dface = nullptr, *dF = nullptr; int* dok = nullptr;
    CUDA_OK(cudaMalloc(&dprim, hprim.size() * 8)); CUDA_OK(cudaMalloc(&dlen, hlen.size() * 8));
    CUDA_OK(cudaMalloc(&dface, hface.size() * 8)); CUDA_OK(cudaMalloc(&dF, (size_t)nfaces * ncq * 8));
    CUDA_OK(cudaMalloc(&dok, (size_t)nfaces * sizeof(int)));
    CUDA_OK(cudaMemcpy(dprim, hprim.data(), hprim.size() * 8, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(dlen, hlen.data(), hlen.size() * 8, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(dface, hface.data(), hface.size() * 8, cudaMemcpyHostToDevice));
    EbParams P = s->P; P.total = total;
    EbArena A; memset(&A, 0, sizeof A);
    A.len[0] = dlen; A.face[0] = dface;
    EbGas* dgas = s->d_gas;
    if (!dgas) {     // before commit: upload the gas parameters on the fly
        CUDA_OK(cudaMalloc((void**)&dgas, sizeof(EbGas)));
        CUDA_OK(cudaMemcpy(dgas, &s->hgas, sizeof(EbGas), cudaMemcpyHostToDevice));
    }
    if (s->cfg.strict_fp) eb_strict::launch_face_debug(s->cfg.flux_calculator, s->cfg.gas_model, P, dgas, A, dprim, nfaces, dF, dok, s->stream);
    else eb_fast::launch_face_debug(s->cfg.flux_calculator, s->cfg.gas_model, P, dgas, A, dprim, nfaces, dF, dok, s->stream);
    s->launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(s->stream));
    CUDA_OK(cudaMemcpy(F, dF, (size_t)nfaces * ncq * 8, cudaMemcpyDeviceToHost));
    if (ok) CUDA_OK(cudaMemcpy(ok, dok, (size_t)nfaces * sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(dprim); cudaFree(dlen); cudaFree(dface); cudaFree(dF); cudaFree(dok);
    if (!s->d_gas) cudaFree(dgas);
    return 0;
}

// ---- direct halo exchange between the processes of one node (CUDA IPC + NVLink peer stores) ------------------------
namespace {
// base address and size of the allocation a device pointer lies in (an IPC handle always names the whole allocation)
int allocation_base(const void* ptr, char** base)
{
    typedef CUresult (*GetRange)(CUdeviceptr*, size_t*, CUdeviceptr);
    static GetRange fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
            (void)cudaGetLastError(); set_err("cuMemGetAddressRange is not available"); return -1;
        }
        fn = (GetRange)f;
    }
    CUdeviceptr b = 0; size_t sz = 0;
    if (fn(&b, &sz, (CUdeviceptr)ptr) != CUDA_SUCCESS) { set_err("cuMemGetAddressRange failed"); return -1; }
    *base = (char*)b;
    return 0;
}
int export_handle(const void* ptr, cudaIpcMemHandle_t* h, long long* off)
{
    char* base = nullptr;
    if (allocation_base(ptr, &base)) return -1;
    CUDA_OK(cudaIpcGetMemHandle(h, base));
    *off = (const char*)ptr - base;
    return 0;
}
int open_handle(Peer& pr, const cudaIpcMemHandle_t& h, long long off, void** out)
{
    void* base = nullptr;
    CUDA_OK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    pr.opened.push_back(base);
    *out = (char*)base + off;
    return 0;
}
}  // namespace

int eb200_p2p_export(int sim, int peer_rank, void* blob, int nbytes)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    if (!s->committed) { set_err("p2p_export needs a committed simulation"); return -1; }
    if (!blob || nbytes < (int)sizeof(P2PBlob)) return (int)sizeof(P2PBlob);       // tells the caller how much room it takes
    CUDA_OK(cudaSetDevice(s->cfg.device));
    int idx = -1;
    for (size_t p = 0; p < s->peers.size(); ++p) if (s->peers[p].rank == peer_rank) idx = (int)p;
    if (idx < 0) { set_err("rank %d is not a halo peer of rank %d", peer_rank, s->cfg.rank); return -1; }
    P2PBlob B; memset(&B, 0, sizeof B);
    B.magic = EB_P2P_MAGIC; B.exporter_rank = s->cfg.rank; B.nprim = s->P.nprim; B.has_S = s->P.shock_detect ? 1 : 0;
    B.total = s->P.total;
    for (int b = 0; b < 3; ++b) if (export_handle(s->A.prim[b], &B.prim[b], &B.prim_off[b])) return -100;
    if (B.has_S && export_handle(s->A.S, &B.S, &B.S_off)) return -100;
    if (export_handle(s->d_shared, &B.shared, &B.shared_off)) return -100;
    B.flag_off = (long long)idx * (long long)sizeof(unsigned long long);
    B.recv_idx_off = s->peers[idx].recv_off; B.recv_count = (long long)s->peers[idx].recv_idx.size();
    CUDA_OK(cudaStreamSynchronize(s->stream));          // the index lists are in place before anybody reads them
    memcpy(blob, &B, sizeof B);
    return (int)sizeof(P2PBlob);
}

int eb200_p2p_import(int sim, int peer_rank, const void* blob, int nbytes)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    if (!s->committed) { set_err("p2p_import needs a committed simulation"); return -1; }
    if (!blob || nbytes < (int)sizeof(P2PBlob)) { set_err("p2p_import: blob too small"); return -1; }
    CUDA_OK(cudaSetDevice(s->cfg.device));
    P2PBlob B; memcpy(&B, blob, sizeof B);
    if (B.magic != EB_P2P_MAGIC || B.exporter_rank != peer_rank) { set_err("p2p_import: not a blob of rank %d", peer_rank); return -1; }
    Peer* pr = nullptr;
    for (Peer& p : s->peers) if (p.rank == peer_rank) pr = &p;
    if (!pr) { set_err("rank %d is not a halo peer of rank %d", peer_rank, s->cfg.rank); return -1; }
    if (pr->p2p) { set_err("rank %d was imported already", peer_rank); return -1; }
    if (B.nprim != s->P.nprim || B.has_S != (s->P.shock_detect ? 1 : 0)) { set_err("p2p_import: rank %d runs another configuration", peer_rank); return -1; }
    if (B.recv_count != (long long)pr->send_idx.size()) {
        set_err("p2p_import: rank %d expects %lld cells from rank %d, which sends %lld", peer_rank, B.recv_count, s->cfg.rank, (long long)pr->send_idx.size());
        return -1;
    }
    void* q = nullptr;
    for (int b = 0; b < 3; ++b) { if (open_handle(*pr, B.prim[b], B.prim_off[b], &q)) return -100; pr->r_prim[b] = (double*)q; }
    if (B.has_S) { if (open_handle(*pr, B.S, B.S_off, &q)) return -100; pr->r_S = (double*)q; }
    if (open_handle(*pr, B.shared, B.shared_off, &q)) return -100;
    char* region = (char*)q;
    pr->r_flag = (unsigned long long*)(region + B.flag_off);
    pr->r_total = B.total;
    if (dev_alloc(s, &pr->d_dst_idx, (size_t)B.recv_count, false)) return -100;
    if (B.recv_count > 0)
        CUDA_OK(cudaMemcpyAsync(pr->d_dst_idx, region + B.recv_idx_off, (size_t)B.recv_count * sizeof(int), cudaMemcpyDeviceToDevice, s->stream));
    CUDA_OK(cudaStreamSynchronize(s->stream));
    pr->p2p = true;
    bool all = true;
    for (Peer& p : s->peers) all = all && p.p2p;
    if (all) {
        std::vector<unsigned long long*> rf;
        for (Peer& p : s->peers) rf.push_back(p.r_flag);
        if (dev_upload(s, &s->d_remote_flags, rf)) return -100;
        CUDA_OK(cudaStreamSynchronize(s->stream));
        s->p2p_ready = true;
    }
    return s->p2p_ready ? 1 : 0;
}

int eb200_describe(int sim, char* dest, int n)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    int ncart = 0, nv3 = 0;
    for (size_t b = 0; b < s->hdesc.size(); ++b) { ncart += s->hdesc[b].cartesian ? 1 : 0; nv3 += s->hdesc[b].v3 ? 1 : 0; }
    char buf[512];
    snprintf(buf, sizeof buf,
             "rank %d: %zu local blocks (%d uniform-Cartesian, %d run by flux_update_kernel_v3), %lld tiles, tma=%d, "
             "halo peers=%zu (%s), arithmetic=%s",
             s->cfg.rank, s->local.size(), ncart, nv3, s->ncta, s->d_tmaps ? 1 : 0, s->peers.size(),
             s->peers.empty() ? "none" : (s->p2p_ready ? "direct NVLink stores" : "exchange callback"),
             s->cfg.strict_fp ? "strict (no FMA)" : "throughput");
    int len = (int)strlen(buf);
    if (dest && n > 0) { strncpy(dest, buf, n - 1); dest[n - 1] = 0; }
    return len;
}

void* eb200_cuda_stream(int sim)
{
    Sim* s = get_sim(sim); if (!s) return nullptr;
    return (void*)s->stream;
}

int eb200_block_is_cartesian(int sim, int blk_id)
{
    Sim* s = get_sim(sim); if (!s) return -1;
    Block* b = get_blk(s, blk_id); if (!b) return -1;
    if (!s->committed || !b->local) { set_err("block_is_cartesian needs a committed local block"); return -1; }
    return b->cartesian ? 1 : 0;
}

}  // extern "C"
