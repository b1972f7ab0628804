// mhhb200 -- host side of the C ABI, shared by the translation units (host_core.cu, host_tend.cu, host_pres.cu):
// the context, error / launch macros and the declarations of the per-stage drivers.  No CPU compute path exists.
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstdint>
#include <initializer_list>
#include <cstring>
#include <string>
#include <vector>
#include <new>

#include "../../include/mhhb200.h"
#include "common.cuh"
#include "stencil_kernels.cuh"
#include "poisson_kernels.cuh"
#include "fft_warp.cuh"
#include "poisson_fused.cuh"
#include "tile_kernels.cuh"
#include "tile2_kernels.cuh"
#include "tile3_kernels.cuh"
#include "evisc3_kernels.cuh"
#include "slab_kernels.cuh"
#include "surface_kernels.cuh"
#include "forcing_kernels.cuh"
#include "tke2_kernels.cuh"
#include "order2_kernels.cuh"
#include "order4_kernels.cuh"
#include "order2i_kernels.cuh"
#include "pres4_kernels.cuh"
#include "thermo_buoy_kernels.cuh"
#include "thermo_moist_kernels.cuh"
#include <cudaTypedefs.h>
#include <dlfcn.h>
#include <nccl.h>

using namespace mhh;

struct mhh_ctx
{
    int dtype = MHH_F64;
    int device = 0;
    std::string err;
    cudaStream_t own_stream = nullptr;
    // side stream for work that is independent of the pressure solve (the RK3 update of the scalars): forked / joined with the
    // two events inside one sub-step.  MHH_OVERLAP=0 keeps everything on the main stream.
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool overlap = true;
    cudaStream_t stream = nullptr;
    long long launches = 0;
    long long ws_bytes = 0;
    int num_sms = 148;
    // optional per-kernel timing: one event after every launch; a kernel's time is the gap to the
    // previous event on the (in-order) stream
    int tile_y = 8;             // MHH_TILE_Y=8|16: tile height of the z-marching kernels
    bool force_plain = false;   // MHH_FORCE_PLAIN=1: use the point-wise kernels everywhere (A/B comparisons)
    bool no_tma = false;        // MHH_NO_TMA=1: keep the cp.async tile kernels (A/B comparisons)
    int tile3_y = 0;            // MHH_TILE3_Y: rows per CTA of the warp-specialised kernel; 0 = 3 rows with the scalar group (13 warps), 4 without
    bool fuse_scalar = true;    // MHH_FUSE_SCALAR=0: keep scalar 0 out of the momentum kernel (A/B comparisons)
    bool scal_tma = true;       // MHH_SCAL_TMA=0: further scalars with the cp.async tile kernel instead of the TMA-staged scalar kernel (A/B switch)
    bool fuse_advec2 = true;    // MHH_FUSE_ADVEC2=0: Advec_2 + Diff_smag2 as two point-wise kernels instead of the fused TMA kernel (A/B switch)
    bool stream_vec2 = true;    // MHH_STREAM_VEC2=0: one cell per thread in rk3 / pres_out_rk3 (A/B switch)
    bool evisc_tma = true;      // MHH_EVISC_TMA=0: keep the cp.async eddy-viscosity kernel (A/B switch)
    int evisc3_npl = 0;         // MHH_EVISC3_NPL=1|2: points per lane of the TMA eddy-viscosity kernel (tile 32 or 64 wide; 0: 1 for fp64, 2 for fp32)
    int evisc3_mb = 0;          // MHH_EVISC3_MB=2|3|4: resident CTAs per SM the TMA eddy-viscosity kernel is compiled for (0: 3)
    int evisc_mb = 4;           // MHH_EVISC_MB=2|3|4 (measured 512^3 fp64: 2.72 | 2.51 | 2.11 ms): resident CTAs per SM the eddy-viscosity kernel is compiled for (register cap)
    int prefetch = 1;           // MHH_PREFETCH: L2 prefetch distance (levels) of the TMA tile kernels, 0 = off
    bool prof = false;
    std::vector<std::pair<const char*, cudaEvent_t>> prof_events;
    std::vector<cudaEvent_t> prof_pool;
    std::string prof_json;
    // y-slab decomposition (npx = 1, npy = nranks): NCCL communicator over the slab ranks
    int nranks = 1, rank = 0;
    ncclComm_t comm = nullptr;
    // Buffer / Force registered for the fused sub-steps (mhh_dycore_set_forcing)
    mhh_forcing forcing{}; bool forcing_set = false;
    // CUDA graph of one full RK3 step (mhh_dycore_step / _step_host on a single GPU): the ~80 launches per sub-step are
    // launch-bound on the small grids (drycblles 128^3, moser180), so the second step with identical arguments is captured on
    // the context's own stream and replayed from then on.  MHH_GRAPH=0 keeps every step eager; profiling, slabs (NCCL), a failed
    // capture or changing arguments (adaptive dt) run eagerly as well.
    // MHH_GRAPH unset: replay when a field is at most 512 MiB (measured: 128^3 fp64 1.18 -> 1.00 ms/step, moser180-shaped DNS
    // 4.59 -> 4.16, 512x512x256 fp32 18.6 -> 18.4; nothing to gain at 512^3 fp64 and beyond); 1: always; 0: never
    bool use_graph = true; int graph_mode = -1;
    cudaGraphExec_t graph_exec = nullptr;
    unsigned long long graph_key = 0, graph_seen = 0;
    long long graph_launches = 0, graph_replays = 0;
    bool graph_failed = false;
    cudaEvent_t ev_g0 = nullptr, ev_g1 = nullptr;
    void drop_graph() { if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; } graph_key = graph_seen = 0; }
    // restart IO staging (device + pinned host), grown on demand by mhh_field3d_save / _load
    void *io_dev = nullptr, *io_host = nullptr; size_t io_cap = 0;
    // Diff_tke2 registered for the fused sub-steps (mhh_dycore_set_tke2)
    mhh_tke2 tke2{}; bool tke2_set = false;
    // Thermo_buoy registered for the fused sub-steps (mhh_dycore_set_thermo_buoy; prm->swthermo = 2)
    mhh_thermo_buoy buoy{}; bool buoy_set = false;
    // Thermo_moist registered for the fused sub-steps (mhh_dycore_set_thermo_moist; prm->swthermo = 3)
    mhh_thermo_moist moist{}; bool moist_set = false;
    virtual ~mhh_ctx() {}
};

// ---- NCCL, bound at run time (dlopen) so that single-GPU users need no NCCL at all; inside a torch
// process this resolves to the libnccl.so.2 torch has already loaded.
struct NcclApi
{
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};


NcclApi* nccl_api(std::string& err);
void prof_mark(mhh_ctx* c, const char* name);

namespace mhhhost {
using namespace mhh;


#define CUDA_TRY(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_); return MHH_E_CUDA; } } while (0)

#define NCCL_TRY(ctx, api, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { \
    (ctx)->err = std::string(#call) + ": " + (api)->GetErrorString(r_); return MHH_E_CUDA; } } while (0)

#define KCHECKN(ctx, name) do { (ctx)->launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
    (ctx)->err = std::string("kernel launch ") + name + ": " + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + std::to_string(__LINE__); \
    return MHH_E_CUDA; } prof_mark(ctx, name); } while (0)


FftPlan make_plan(int n, bool& ok);

template <typename TF>
struct Ctx : mhh_ctx
{
    GridDev<TF> g{};
    mhh_grid_desc desc{};
    // device copies of the profiles
    TF *d_prof = nullptr;          // 12 profiles x kcells
    TF *d_mlen0 = nullptr;
    // Pres_2
    int nm = 0;
    FftPlan plan_x{}, plan_y{};
    cplx<TF> *tw_xh = nullptr, *tw_xf = nullptr, *tw_y = nullptr;
    TF *d_bmati = nullptr, *d_bmatj = nullptr, *d_a = nullptr, *d_c = nullptr, *d_dz2rho = nullptr, *d_dz2 = nullptr;
    TF *spec = nullptr;            // spectral workspace, x side: nm*jmax*ktot complex
    TF *specT = nullptr;           // y side: mcl*jtot*ktot complex (== spec on a single GPU)
    TF *fac = nullptr;             // tdma factors, mcl*jtot*ktot
    // Pres_4: band coefficients (7 x kmax), 4th-order modified wavenumbers, LU factors of every mode (7 x (kmax+4) x ncol)
    TF *d_m7 = nullptr, *d_bmati4 = nullptr, *d_bmatj4 = nullptr, *lu4 = nullptr;
    std::vector<TF> h_dzi4, h_dzhi4, h_z;
    SpecLayout lay{};
    PeerPtrs<TF> peers{};          // IPC-mapped workspaces of all slab ranks (peers.on: fused transposes)
    int *d_barrier = nullptr;      // dummy word for the all-reduce that closes a fused transpose
    // peer halos: two alternating sets of (from north, from south) receive buffers, exported over CUDA IPC; the
    // neighbours' sets are mapped in peer_halo_{south,north}
    TF *phalo = nullptr;           // own: [set][dir] x phalo_cap elements
    size_t phalo_cap = 0;
    TF *phalo_south = nullptr, *phalo_north = nullptr;     // base of the south / north neighbour's phalo
    unsigned phalo_count = 0;
    TF *halo = nullptr;            // 4 staging buffers (send south/north, recv north/south) of halo_cap elements
    size_t halo_cap = 0;
    bool basestate_set = false;
    // Boundary_surface lookup table (z/L nodes and evaluation function, float like the reference), built by mhh_boundary_surface_init
    float *d_zL_sl = nullptr, *d_f_sl = nullptr;
    // Buffer / Force: per-level damping factors (u, v, scalars | w) and the two device sums of the fixed-mass-flux forcing
    TF *d_sigmaz = nullptr; double *d_sums = nullptr;
    double buf_key[3] = {-1., -1., -1.}; int buf_k = 0, buf_kh = 0;
    int surf_mbcbot = -1, surf_thermobc = -1;
    std::vector<TF> h_thref, h_threfh;
    // Thermo_moist: pref, prefh, rhoref, rhorefh, exnref, exnrefh of its base state + the mean profiles of thl and qt (8 x kcells;
    // thvref / thvrefh live in g.thref / g.threfh), and the count of non-converged saturation adjustments
    TF *d_moist = nullptr; int *d_moist_flag = nullptr; bool moist_profiles_set = false;
    MoistProfiles<TF> moist_profiles() const
    {
        const int kc = g.kcells;
        return {d_moist, d_moist + kc, d_moist + 2 * kc, d_moist + 3 * kc, const_cast<TF*>(g.thref), const_cast<TF*>(g.threfh),
                d_moist + 4 * kc, d_moist + 5 * kc};
    }
    double *d_red = nullptr;       // reduction scalar
    double *h_red = nullptr;       // pinned
    std::vector<TF> h_rhoref, h_rhorefh, h_dz, h_dzhi;
    int rows_x = 1, mc_y = 4;
    bool wfft_x = false, wfft_y = false;   // warp-per-sequence FFT kernels (power-of-two lengths)
    // Pres_2 version 2 (poisson_fused.cuh): y transforms fused with the Thomas sweeps.  In this mode `spec` is the X side,
    // `specT` the Y side (the names the IPC export uses), `fac` holds the reciprocal pivots T[ml][k][pos].
    bool fused = false;
    Spec2 lay2{};
    int jlog2 = 0;
    cplx<TF>* stage2 = nullptr;    // NCCL transport only: send staging of the backward transpose
    size_t smem_x = 0, smem_y = 0;

    ~Ctx() override
    {
        cudaSetDevice(device);
        drop_graph();
        if (ev_g0) cudaEventDestroy(ev_g0);
        if (ev_g1) cudaEventDestroy(ev_g1);
        cudaFree(io_dev); if (io_host) cudaFreeHost(io_host);
        cudaFree(d_moist); cudaFree(d_moist_flag);
        cudaFree(d_zL_sl); cudaFree(d_f_sl); cudaFree(d_sigmaz); cudaFree(d_sums);
        cudaFree(d_prof); cudaFree(d_mlen0); cudaFree(tw_xh); cudaFree(tw_xf); cudaFree(tw_y);
        cudaFree(d_bmati); cudaFree(d_bmatj); cudaFree(d_a); cudaFree(d_c); cudaFree(d_dz2rho); cudaFree(d_dz2);
        for (int r = 0; r < MAX_SLAB_RANKS; ++r)
            if (peers.on && r != rank) { if (peers.x[r]) cudaIpcCloseMemHandle(peers.x[r]); if (peers.y[r]) cudaIpcCloseMemHandle(peers.y[r]); }
        if (phalo_south && phalo_south != phalo) cudaIpcCloseMemHandle(phalo_south);
        if (phalo_north && phalo_north != phalo && phalo_north != phalo_south) cudaIpcCloseMemHandle(phalo_north);
        cudaFree(phalo);
        cudaFree(d_barrier);
        cudaFree(d_m7); cudaFree(d_bmati4); cudaFree(d_bmatj4); cudaFree(lu4); cudaFree(stage2);
        if (specT != spec) cudaFree(specT);
        cudaFree(spec); cudaFree(fac); cudaFree(d_red); cudaFree(halo);
        if (comm) { std::string e; NcclApi* api = nccl_api(e); if (api) api->CommDestroy(comm); }
        if (h_red) cudaFreeHost(h_red);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (side_stream) cudaStreamDestroy(side_stream);
        if (own_stream) cudaStreamDestroy(own_stream);
    }

    TdmaCoef<TF> coef() const { return {d_a, d_c, d_dz2rho, d_dz2, d_bmati, d_bmatj}; }

    dim3 blk() const { return dim3(64, 4, 1); }
    dim3 grd_interior() const { return dim3((g.imax + 63) / 64, (g.jmax + 3) / 4, g.kmax); }
    dim3 grd_all() const { return dim3((g.icells + 63) / 64, (g.jcells + 3) / 4, g.kcells); }
};

template <typename TF> inline TF* P(void* p) { return static_cast<TF*>(p); }
template <typename TF> inline const TF* P(const void* p) { return static_cast<const TF*>(p); }

#define NEED_BASE(c) do { if (!(c)->basestate_set) { (c)->err = "mhh_set_basestate has not been called"; return MHH_E_INVALID; } } while (0)
#define NEED(c, ptr, what) do { if (!(ptr)) { (c)->err = std::string(what) + " is NULL"; return MHH_E_INVALID; } } while (0)



// ---- per-stage drivers (defined in one translation unit each, instantiated for double and float) ----
// host_core.cu
template <typename TF> int slab_barrier(Ctx<TF>* c, const char* name);
template <typename TF> int exchange_ns(Ctx<TF>* c, TF* const* flds, int nf, int w, int nk);
template <typename TF> int cyclic_impl(Ctx<TF>* c, TF* fld, int edge, bool two_d);
template <typename TF> int cyclic_fields(Ctx<TF>* c, TF* const* flds, int nf);
template <typename TF> int buffer_exec_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_forcing* fo);
template <typename TF> int force_exec_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_forcing* fo, double sub_dt);
// host_tend.cu
template <typename TF> int check_mom(Ctx<TF>* c, const mhh_fields* f, bool need_evisc, bool surface);
template <typename TF> int evisc_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, const TF* n2);
// adv2: Advec_2's fluxes inside the fused TMA kernel; returns MHH_NOT_FUSED (nothing launched) when that variant does not apply
#define MHH_NOT_FUSED 1
template <typename TF> int tend_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, bool adv, bool diff, bool buoy, const mhh_tke2* tke = nullptr, bool adv2 = false);
template <typename TF> int o2_impl(Ctx<TF>* c, const mhh_fields* f, bool adv, bool diff, bool buoy);
template <typename TF> int o4_impl(Ctx<TF>* c, const mhh_fields* f, int adv_sw, bool diff);
template <typename TF> int adv2i_impl(Ctx<TF>* c, const mhh_fields* f, int adv_sw);
template <typename TF> int o2_cfl_impl(Ctx<TF>* c, const mhh_fields* f, double* out, int order);
template <typename TF> int reduce_mode_impl(Ctx<TF>* c, int mode, const TF* u, const TF* v, const TF* w, TF p0, TF p1, TF p2, double* out);
// host_tke2.cu
template <typename TF> int tke2_check(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, const mhh_tke2* tke);
template <typename TF> int tke2_visc_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, const mhh_tke2* tke, const TF* n2);
template <typename TF> int limiter_impl(Ctx<TF>* c, TF* at, const TF* a, TF min_value, TF sub_dt);
// host_thermo.cu
template <typename TF> int thermo_buoy_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_thermo_buoy* tb);
template <typename TF> int thermo_moist_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_thermo_moist* tm);
// host_pres.cu
template <typename TF> int pres_create(Ctx<TF>* c);
template <typename TF> int pres_set_values(Ctx<TF>* c);
template <typename TF> int pres_solve_impl(Ctx<TF>* c, const mhh_fields* f, double sub_dt);
template <typename TF> int pres_exec_impl(Ctx<TF>* c, const mhh_fields* f, double sub_dt);
template <typename TF> int pres4_exec_impl(Ctx<TF>* c, const mhh_fields* f, double sub_dt);
template <typename TF> int pres4_div_impl(Ctx<TF>* c, const mhh_fields* f, double* out);
template <typename TF> int fft_roundtrip_impl(Ctx<TF>* c, const TF* in, TF* out, int solve);

} // namespace mhhhost
