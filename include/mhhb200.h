/*
 * mhhb200.h -- C ABI of libmhhb200.so, the B200-native (sm_100a) dynamical core for MicroHH.
 *
 * This is the drop-in boundary for the hot path of one RK3 sub-step: every entry point
 * replaces one member function of the reference's Advec / Diff / Pres / Boundary_cyclic /
 * Timeloop classes (cited per function, paths relative to the reference tree).  A thin C++
 * adapter (microhh_b200/host/mhh_adapters.hpp) forwards the class methods to
 * these functions so `microhh init/run`, the .ini/.nc case files and the Field3d layout stay
 * unchanged.
 *
 * Conventions
 *  - Plain C: pointers and sizes only, no C++/torch types.
 *  - Precision is a property of the context (MHH_F64 = default build, MHH_F32 = USESP build);
 *    field pointers are `void*` to arrays of that type.
 *  - Field arrays are DEVICE pointers in the reference's ghosted layout
 *    ijk = i + j*icells + k*icells*jcells (include/grid.h:49-134); the caller owns them.
 *    2-D companions (flux_bot, grad_bot, dudz_mo, ...) are DEVICE arrays of icells*jcells.
 *  - Metric/base-state profiles passed at create / set_basestate time are HOST arrays of kcells.
 *  - Every function returns 0 on success or a negative MHH_E_* code; mhh_last_error() gives the
 *    message.  No exception crosses the ABI.
 *  - Calls are asynchronous on the context's stream and ordered; mhh_sync() waits.
 *    A context is not re-entrant; use one per GPU.
 *  - There is NO CPU fallback: without a CUDA device every compute call fails with MHH_E_CUDA.
 */
#ifndef MHHB200_H
#define MHHB200_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#  define MHH_API __attribute__((visibility("default")))
#else
#  define MHH_API
#endif

#define MHH_F64 0
#define MHH_F32 1

#define MHH_OK          0
#define MHH_E_INVALID  -1   /* bad argument / unsupported configuration */
#define MHH_E_CUDA     -2   /* CUDA runtime error (message has the detail) */
#define MHH_E_NOMEM    -3
#define MHH_E_IO       -4   /* restart file cannot be created / opened / is too short (the reference's save/load return 1) */

#define MHH_MAX_SCALARS 8

/* Boundary_cyclic Edge (include/boundary_cyclic.h:33) */
#define MHH_EDGE_EAST_WEST   0
#define MHH_EDGE_NORTH_SOUTH 1
#define MHH_EDGE_BOTH        2

/* Boundary_type subset used for the vertical ghost cells (src/boundary.cxx:700-772) */
#define MHH_BC_NONE     -1
#define MHH_BC_DIRICHLET 0
#define MHH_BC_NEUMANN   1   /* also Flux_type */

typedef struct mhh_ctx mhh_ctx;

/* POD copy of Grid_data<TF> (include/grid.h:49-134).  z..dzhi: HOST arrays, kcells entries,
 * element type = the context's precision. */
typedef struct mhh_grid_desc
{
    int itot, jtot, ktot;        /* global size */
    int imax, jmax, kmax;        /* local block (== global for npx=npy=1) */
    int igc, jgc, kgc;           /* ghost cells */
    double xsize, ysize, zsize;
    const void* z;
    const void* zh;
    const void* dz;
    const void* dzh;
    const void* dzi;
    const void* dzhi;
    int npx, npy;                /* process grid (src/master_parallel.cxx:103-153); 1,1 = single GPU; npx must be 1 (y slabs) */
    int mpicoordx, mpicoordy;
    /* 4th-order grid only (swspatialorder = 4, src/grid.cxx:306-375): HOST arrays of kcells, else NULL */
    const void* dzi4;
    const void* dzhi4;
} mhh_grid_desc;

/* Device pointers to the 3-D fields and their 2-D companions (Fields maps mp/mt/sp/st/sd,
 * include/fields.h:134-143; Field3d companions include/field3d.h:49-78). */
typedef struct mhh_fields
{
    void *u, *v, *w;             /* fields.mp */
    void *ut, *vt, *wt;          /* fields.mt */
    void *evisc;                 /* fields.sd["evisc"] */
    void *p;                     /* fields.sd["p"] */
    int   ns;                    /* number of prognostic scalars */
    void *s[MHH_MAX_SCALARS];    /* fields.sp */
    void *st[MHH_MAX_SCALARS];   /* fields.st */
    double svisc[MHH_MAX_SCALARS];
    double visc;
    /* 2-D companions; may be NULL when the surface model is off */
    void *u_fluxbot, *u_fluxtop, *v_fluxbot, *v_fluxtop;
    void *s_fluxbot[MHH_MAX_SCALARS], *s_fluxtop[MHH_MAX_SCALARS];
    /* Boundary_surface outputs consumed by Diff_smag2 (Boundary::get_dudz/dvdz/dbdz/z0m) */
    void *dudz_mo, *dvdz_mo, *dbdz_mo, *z0m;
    /* vertical ghost-cell inputs: value (Dirichlet) or gradient (Neumann/flux) at bottom/top */
    void *u_bot, *u_gradbot, *u_top, *u_gradtop;
    void *v_bot, *v_gradbot, *v_top, *v_gradtop;
    void *s_bot[MHH_MAX_SCALARS], *s_gradbot[MHH_MAX_SCALARS], *s_top[MHH_MAX_SCALARS], *s_gradtop[MHH_MAX_SCALARS];
    /* != 0: the scalar is in [advec] fluxlimit_list -> Advec_2i5 advects it with the Koren (1993) flux limiter
     * (include/advec_monotonic.h:98-202, src/advec_2i5.cxx:1046-1056) */
    int   s_fluxlimit[MHH_MAX_SCALARS];
} mhh_fields;

/* Run-time switches of the path ([advec] [diff] [boundary] [thermo] of the .ini). */
typedef struct mhh_params
{
    int    swadvec;              /* 25 = 2i5, 2 = 2, 24 = 2i4, 262 = 2i62, 4 = 4, 41 = 4m */
    int    swdiff;               /* 1 = smag2, 2 = 2, 3 = tke2 (needs mhh_dycore_set_tke2), 4 = 4 (the 4th-order configuration:
                                  * 4 + 4 + pres_4 on a 4th-order grid) */
    int    swthermo;             /* 0 = off, 1 = dry (buoyancy from scalar 0 = th), 2 = buoy (scalar 0 IS the buoyancy b;
                                  * the fused sub-steps need mhh_dycore_set_thermo_buoy), 3 = moist (scalars thl and qt, thl first;
                                  * needs mhh_dycore_set_thermo_moist and a base state, see mhh_thermo_moist_*) */
    int    surface_model;        /* Boundary switch != "default"  (Surface_model::Enabled) */
    int    sw_mason;             /* [diff] swmason */
    double cs, tPr;              /* [diff] cs, tPr */
    int    mbcbot, mbctop;       /* MHH_BC_* for u,v */
    int    sbcbot[MHH_MAX_SCALARS], sbctop[MHH_MAX_SCALARS];
} mhh_params;

/* State of the Monin-Obukhov surface model (Boundary_surface<TF>, include/boundary_surface.h): 2-D device arrays of
 * ijcells entries, caller-owned (in MicroHH they are Boundary_surface's ustar_g, obuk_g, nobuk_g, z0m_g, z0h_g). */
typedef struct mhh_surface
{
    void *ustar, *obuk;          /* friction velocity and Obukhov length (start from Constants::dsmall like init_surface) */
    int  *nobuk;                 /* last lookup-table index per column (search hint), start from 0 */
    void *z0m, *z0h;             /* roughness lengths */
    void *dutot;                 /* scratch plane */
    int   sbcbot[MHH_MAX_SCALARS];   /* per scalar: MHH_SBC_DIRICHLET (surface value given -> flux computed), MHH_SBC_FLUX (flux given ->
                                      * surface value computed), anything else: left alone (src/boundary_surface.cxx:292-340) */
} mhh_surface;
#define MHH_SBC_DIRICHLET 0
#define MHH_SBC_FLUX      2

/* Damping layer (Buffer<TF>, src/buffer.cxx) and large-scale forcings (Force<TF>, src/force.cxx).  Profiles are DEVICE arrays
 * of kcells entries (the reference's bufferprofs / ug_g / vg_g / lsprofs_g / wls_g); NULL switches the term off. */
#define MHH_LSPRES_OFF   0
#define MHH_LSPRES_UFLUX 1    /* [force] swlspres=uflux: fixed mass flux (enforce_fixed_flux, src/force.cxx:65-75, 612-624) */
#define MHH_LSPRES_DPDX  2    /* swlspres=dpdx: constant pressure gradient (:626-635) */
#define MHH_LSPRES_GEO   3    /* swlspres=geo: Coriolis force with a geostrophic wind (:637-660; 2nd / 4th order by the grid) */
typedef struct mhh_forcing
{
    int    swbuffer;                         /* [buffer] swbuffer */
    double buffer_zstart, buffer_sigma, buffer_beta;
    const void *bufferprof_u, *bufferprof_v, *bufferprof_w;
    const void *bufferprof_s[MHH_MAX_SCALARS];
    int    swlspres;
    double uflux, dpdx, fc;                  /* [force] uflux, dpdx, fc */
    const void *ug, *vg;                     /* geostrophic wind profiles */
    double utrans, vtrans;                   /* [grid] utrans, vtrans (Galilean transformation) */
    const void *ls_s[MHH_MAX_SCALARS];       /* [force] swls: large-scale source profile per scalar */
    const void *wls;                         /* [force] swwls=local: subsidence velocity profile applied to every scalar */
} mhh_forcing;

/* Deardorff (1980) SGS-TKE closure (Diff_tke2<TF>, src/diff_tke2.cxx): the prognostic scalar `sgstke` is one of the
 * scalars of mhh_fields (fields.sp["sgstke"], src/diff_tke2.cxx:544), `eviscs` the eddy viscosity for heat / scalars
 * (fields.sd["eviscs"], only with buoyancy: prm->swthermo != 0).  Constants: [diff] ap, cf, ce1, ce2, cm, ch1, ch2, cn
 * (reference defaults 1.5, 2.5, 0.19, 0.51, 0.12, 1, 2, 0.76; src/diff_tke2.cxx:525-532). */
typedef struct mhh_tke2
{
    int    isgstke;              /* index of sgstke in mhh_fields.s / .st */
    void  *eviscs;               /* DEVICE field (ghosted layout); may be NULL when swthermo == 0 */
    double ap, cf, ce1, ce2, cm, ch1, ch2, cn;
} mhh_tke2;

/* Thermo_buoy<TF> (src/thermo_buoy.cxx:306-330): [thermo] alpha (slope angle, rad), N2 (background stratification),
 * swbaroclinic / dbdy_ls, and [grid] utrans.  Slope-enabled thermodynamics is on when |alpha| > 0 or |N2| > 0. */
typedef struct mhh_thermo_buoy
{
    double alpha, n2;
    double utrans;               /* Grid_data::utrans (calc_buoyancy_tend_b adds it to the interpolated u) */
    int    swbaroclinic;
    double dbdy_ls;
} mhh_thermo_buoy;

/* Thermo_moist<TF> (src/thermo_moist.cxx:1080-1133): the prognostic scalars thl and qt, [thermo] pbot and swupdatebasestate. */
typedef struct mhh_thermo_moist
{
    int    ithl, iqt;            /* indices of thl and qt in mhh_fields.s / .st */
    double pbot;                 /* surface pressure [Pa] */
    int    swupdatebasestate;    /* != 0: exec re-integrates the hydrostatic base state from the horizontal means of thl and qt */
} mhh_thermo_moist;
/* get_thermo_field names served on the device */
#define MHH_MOIST_B  0
#define MHH_MOIST_QL 1
#define MHH_MOIST_N2 2

/* ---- context ----------------------------------------------------------------------------- */
MHH_API int  mhh_ctx_create(const mhh_grid_desc* grid, int dtype, int device, mhh_ctx** out);
MHH_API void mhh_ctx_destroy(mhh_ctx* ctx);
MHH_API const char* mhh_last_error(const mhh_ctx* ctx);
MHH_API int  mhh_sync(mhh_ctx* ctx);
/* Run on an externally owned CUDA stream (cudaStream_t as void*); NULL = the CUDA legacy default
 * stream.  A new context runs on its own non-blocking stream until this is called. */
MHH_API int  mhh_set_stream(mhh_ctx* ctx, void* cuda_stream);
/* Fields::rhoref/rhorefh (include/fields.h:162-188) and Thermo_dry's thref/threfh; HOST arrays (kcells).
 * thref/threfh may be NULL when swthermo == 0.  Also (re)builds the Pres_2 coefficient tables. */
MHH_API int  mhh_set_basestate(mhh_ctx* ctx, const void* rhoref, const void* rhorefh, const void* thref, const void* threfh);
/* number of kernel launches issued by this context so far (for bench.py's gpu_launches) */
MHH_API long long mhh_launch_count(const mhh_ctx* ctx);
/* Per-kernel device timing (CUDA events on the context's stream): start, run any calls, stop.
 * `json` receives a context-owned string {"kernel": {"n": launches, "ms": total}, ...}. */
MHH_API int mhh_profile_start(mhh_ctx* ctx);
MHH_API int mhh_profile_stop(mhh_ctx* ctx, const char** json);
/* bytes of device memory owned by the context (workspace, tables) */
MHH_API long long mhh_workspace_bytes(const mhh_ctx* ctx);

/* ---- multi-GPU: y-slab decomposition (npx = 1, npy = P ranks, one context / process per GPU) -----
 * The reference cannot combine MPI and CUDA (CMakeLists.txt:50-52); its CPU-MPI code is the semantic
 * model: the process grid of Master (src/master_parallel.cxx:103-153; here mpicoordy = rank),
 * Boundary_cyclic's neighbour exchange (src/boundary_cyclic.cxx:115-176), Transpose::exec_xy/exec_yx
 * (src/transpose.cxx:117-271) inside FFT::exec_forward/backward (src/fft.cxx:455-587) and the
 * MPI_Allreduce(MAX) of cfl / dn / divergence (src/master_parallel.cxx:270-290).
 * A context created with npy > 1 owns the slab jmax = jtot/npy; once mhh_comm_init() has connected
 * the ranks, EVERY hot-path call below is collective over the slab ranks (NCCL semantics) and does
 * its own north/south ghost-row exchange, all-to-all transposes and reductions on the context's
 * stream.  NCCL is bound at run time (dlopen of libnccl.so.2); single-GPU use needs no NCCL.
 * Bootstrap: rank 0 calls mhh_comm_get_unique_id, the host broadcasts the bytes (MPI_Bcast in
 * MicroHH's Master, torch.distributed in the tests), every rank calls mhh_comm_init. */
#define MHH_COMM_ID_BYTES 128
MHH_API int mhh_comm_get_unique_id(void* id, int nbytes);
MHH_API int mhh_comm_init(mhh_ctx* ctx, const void* id, int nbytes);
/* Fused transposes over NVLink/NVSwitch peer memory (optional, same node): every rank exports CUDA IPC handles of its
 * two spectral workspaces, the host all-gathers them in rank order (MPI_Allgather / torch.distributed.all_gather),
 * every rank opens its peers'.  From then on the x transform stores its modes straight into the owning rank's y-side
 * workspace and the inverse y transform stores straight into the row owner's x-side workspace -- the all-to-all is the
 * store phase of the FFT kernels; one 4-byte all-reduce closes each transpose.  The ghost-row exchange likewise pushes
 * its strips straight into the neighbours' receive buffers.  Without these calls (or with MHH_NO_PEER=1) both use
 * grouped ncclSend/ncclRecv. */
#define MHH_IPC_BYTES 192
MHH_API int mhh_comm_get_ipc_handles(mhh_ctx* ctx, void* out, int nbytes);
MHH_API int mhh_comm_open_peers(mhh_ctx* ctx, const void* all_handles, int nbytes);
/* Drop the peer mappings again and use NCCL transport (e.g. when mapping failed on some rank: the host decides
 * collectively and every rank calls this). */
MHH_API int mhh_comm_disable_peers(mhh_ctx* ctx);
/* 0 = single GPU / not connected, 1 = NCCL send/recv transposes and halos, 2 = fused NVLink peer stores */
MHH_API int mhh_comm_transport(const mhh_ctx* ctx);

/* Spectral workspace layout of the slab decomposition (pure host functions, no GPU needed): which
 * x-modes a rank owns after the forward transpose and where element (row, mode) / (k, j, mode) lives
 * in the x-side / y-side buffers (complex-element index).  Exposed so that the host-side tests can
 * replay the all-to-all with any transport (world_size-2 gloo tests on CPU). */
typedef struct mhh_slab_info
{
    int nm;                  /* x-modes in total: itot/2 + 1 */
    int mcl, m_off;          /* modes owned by this rank: [m_off, m_off + mcl) */
    int jmax;                /* rows of the slab per level */
    long long rows;          /* jmax*ktot */
    long long xside_elems;   /* complex elements of the x-side buffer: nm*rows */
    long long yside_elems;   /* complex elements of the y-side buffer: mcl*jtot*ktot */
} mhh_slab_info;
MHH_API int mhh_slab_layout(int itot, int jtot, int ktot, int npy, int rank, mhh_slab_info* out);
MHH_API long long mhh_slab_xindex(int itot, int jtot, int ktot, int npy, int rank, long long row, int m);
/* x-side index in the 8-mode-panel layout used with the fused peer transposes (block d = [k][panel][jl][8]); *total
 * receives the padded size of the x-side buffer in complex elements */
MHH_API long long mhh_slab_xindex_tiled(int itot, int jtot, int ktot, int npy, int rank, long long row, int m, long long* total);
MHH_API long long mhh_slab_yindex(int itot, int jtot, int ktot, int npy, int rank, int k, int j, int ml);

/* Layout of the FUSED Pres_2 path (power-of-two itot/2 and jtot; y transforms fused with the two-sided Thomas sweeps):
 *   Y side of the rank that owns a mode: [source rank s][local mode ml][level k][local row jl]  -- a y sequence of one
 *     (mode, level) is npy contiguous pieces, block s IS the message from rank s;
 *   X side of the rank that owns a row:  [mode owner d][level k][row panel jl/8][local mode ml][jl%8] -- the 8 rows of a mode are
 *     one 128-byte chunk, block d IS the message from rank d.
 * ksplit: levels below it are eliminated upwards, the others downwards (two CTAs per mode). */
typedef struct mhh_slab2_info
{
    int nm, mcl, m_off, jmax, npan, ksplit;
    long long xside_elems, yside_elems;
} mhh_slab2_info;
MHH_API int mhh_slab2_layout(int itot, int jtot, int ktot, int npy, int rank, mhh_slab2_info* out);
MHH_API long long mhh_slab2_yindex(int itot, int jtot, int ktot, int npy, int owner, int src, int ml, int k, int jl);
MHH_API long long mhh_slab2_xindex(int itot, int jtot, int ktot, int npy, int mode_owner, int k, int jl, int ml);

/* ---- Boundary_cyclic<TF>::exec / exec_2d  (src/boundary_cyclic.cxx:369-507) ---------------- */
MHH_API int mhh_boundary_cyclic(mhh_ctx* ctx, void* fld, int edge);
MHH_API int mhh_boundary_cyclic_2d(mhh_ctx* ctx, void* fld2d);

/* ---- Boundary<TF>::set_ghost_cells, 2nd order, one field (src/boundary.cxx:700-772, 933-961) */
MHH_API int mhh_boundary_ghost_cells_2nd(mhh_ctx* ctx, void* fld, int bcbot, const void* bot, const void* gradbot,
                                 int bctop, const void* top, const void* gradtop);

/* ---- 4th-order grids: Boundary<TF>::set_ghost_cells, one field (src/boundary.cxx:776-848, 963-991), and
 *      Boundary<TF>::set_ghost_cells_w (src/boundary.cxx:850-922, 999-1021; conservation != 0: Conservation_type) ---- */
MHH_API int mhh_boundary_ghost_cells_4th(mhh_ctx* ctx, void* fld, int bcbot, const void* bot, const void* gradbot,
                                 int bctop, const void* top, const void* gradtop);
MHH_API int mhh_boundary_ghost_cells_w_4th(mhh_ctx* ctx, void* w, int conservation);

/* ---- Advec<TF>::exec / get_cfl  (swadvec = 25: Advec_2i5, src/advec_2i5.cxx:955-1063;
 *      swadvec = 2: Advec_2, src/advec_2.cxx:288-345; swadvec = 4: Advec_4, src/advec_4.cxx:573-684;
 *      swadvec = 41: Advec_4m, src/advec_4m.cxx:511-615 -- the fully conservative scheme cases/moser180 ships with) */
MHH_API int mhh_advec_exec(mhh_ctx* ctx, int swadvec, const mhh_fields* f);
MHH_API int mhh_advec_get_cfl(mhh_ctx* ctx, int swadvec, const mhh_fields* f, double dt, double* cfl);

/* ---- Diff_smag2<TF>::exec_viscosity / exec / get_dn  (src/diff_smag2.cxx:312-356, 381-607) --
 * n2: device N2 field from Thermo::get_thermo_field("N2"), or NULL to derive it from scalar 0
 * (th) as Thermo_dry does (src/thermo_dry.cxx:66-78). */
MHH_API int mhh_diff_smag2_exec_viscosity(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, const void* n2);
MHH_API int mhh_diff_smag2_exec(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm);
MHH_API int mhh_diff_smag2_get_dn(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, double dt, double* dn);

/* ---- Diff_tke2<TF> (src/diff_tke2.cxx; the surface model is mandatory, :557):
 *   create          cold start: sgstke = max(sgstke, Constants::sgstke_min) + cyclic fill (:641-660)
 *   exec_viscosity  strain^2, evisc (and eviscs) from sgstke, and the buoyancy / dissipation / shear sources added to the
 *                   tendency of sgstke, fused into one kernel (:799-983); n2 as in mhh_diff_smag2_exec_viscosity
 *   exec            diff_u / diff_v / diff_w with evisc; diff_c with tPr = 1: sgstke with evisc, the other scalars with
 *                   eviscs when there is buoyancy (:676-797)
 *   get_dn          calc_dnmul on eviscs (buoyancy) or evisc, tPr = 1 (:611-638)
 * and Limiter<TF>::exec's tendency_limiter on one (tendency, field) pair (src/limiter.cxx:35-59, 117-129; for sgstke:
 * min_value = Constants::sgstke_min = 1e-7).
 * mhh_dycore_set_tke2 registers the closure (a copy of the struct) for the fused sub-steps with prm->swdiff = 3, which then
 * run exec_viscosity, exec and the limiter on sgstke where Model::exec does (src/model.cxx:376, 414, 440); NULL unregisters. */
MHH_API int mhh_diff_tke2_create(mhh_ctx* ctx, void* sgstke);
MHH_API int mhh_diff_tke2_exec_viscosity(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, const mhh_tke2* tke, const void* n2);
MHH_API int mhh_diff_tke2_exec(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, const mhh_tke2* tke);
MHH_API int mhh_diff_tke2_get_dn(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, const mhh_tke2* tke, double dt, double* dn);
MHH_API int mhh_limiter_exec(mhh_ctx* ctx, void* at, const void* a, double min_value, double sub_dt);
MHH_API int mhh_dycore_set_tke2(mhh_ctx* ctx, const mhh_tke2* tke);

/* ---- Diff_2<TF>::exec / get_dn  (src/diff_2.cxx:133-190): nu * laplacian on u, v, w and every scalar ---- */
MHH_API int mhh_diff_2_exec(mhh_ctx* ctx, const mhh_fields* f);
MHH_API int mhh_diff_2_get_dn(mhh_ctx* ctx, const mhh_fields* f, double dt, double* dn);
/* ---- Diff_4<TF>::exec (src/diff_4.cxx:255-310; get_dn is Diff_2's formula, :226-245 -> mhh_diff_2_get_dn).  Needs a
 * 4th-order grid.  Advec_4 is mhh_advec_exec / mhh_advec_get_cfl with swadvec = 4 (src/advec_4.cxx:573-684). */
MHH_API int mhh_diff_4_exec(mhh_ctx* ctx, const mhh_fields* f);

/* ---- Thermo_dry<TF>::exec (buoyancy on wt) and get_thermo_field("N2")  (src/thermo_dry.cxx) -- */
MHH_API int mhh_thermo_dry_exec(mhh_ctx* ctx, void* wt, const void* th);
MHH_API int mhh_thermo_dry_n2(mhh_ctx* ctx, void* n2, const void* th);
/* ---- Thermo_buoy<TF>::exec (src/thermo_buoy.cxx:345-391; 2nd- or 4th-order by the context's grid): buoyancy on wt, with
 * slope-enabled thermodynamics also on ut and on the buoyancy tendency st[0], the baroclinic term on st[0]; one kernel.
 * Reads s[0] = b (cyclic + vertical ghost cells filled), u, v, w.  get_thermo_field("N2") (:410-413) = _n2 with bg_n2 = [thermo] N2.
 * mhh_dycore_set_thermo_buoy registers the parameters for the fused sub-steps with prm->swthermo = 2 (NULL: unregister);
 * thermo.exec then runs where Model::exec has it (src/model.cxx:388), before the advection. */
MHH_API int mhh_thermo_buoy_exec(mhh_ctx* ctx, const mhh_fields* f, const mhh_thermo_buoy* tb);
MHH_API int mhh_thermo_buoy_n2(mhh_ctx* ctx, void* n2, const void* b, double bg_n2);
MHH_API int mhh_dycore_set_thermo_buoy(mhh_ctx* ctx, const mhh_thermo_buoy* tb);

/* ---- Thermo_moist<TF> (src/thermo_moist.cxx, include/thermo_moist_functions.h) ---------------------------------------------
 * Base state `bs` (pref, prefh, rhoref, rhorefh, thvref, thvrefh, exnref, exnrefh; kcells entries) lives in the context;
 * thvref / thvrefh share storage with the thref / threfh of mhh_set_basestate (the closures' N2 and the surface model read them
 * there, as Thermo_moist::get_thermo_field("N2") uses thvref).  fields.rhoref / rhorefh of the dynamics stay what
 * mhh_set_basestate set (create_basestate step 6 copies the INITIAL rhoref there; they are not updated afterwards).
 *   _calc_base_state  Thermo_moist_functions::calc_base_state (functions.h:271-340) on the device from HOST mean profiles
 *                     thl0 / qt0 (kcells, ghost entries kstart-1 and kend set, e.g. by calc_top_and_bot) = create_basestate step 4
 *   _set_profiles     overwrite any of the eight profiles from HOST arrays (NULL = keep; pref with exnref and prefh with exnrefh,
 *                     both or neither: the kernels take the exner function from the profile): the Boussinesq override of
 *                     create_basestate step 5, or a base state loaded from a restart file (Thermo_moist::load)
 *   _get_profiles     copy them to HOST arrays (NULL = skip); synchronises
 *   _exec             Thermo_moist::exec (:1415-1447): with swupdatebasestate the mean profiles of thl and qt
 *                     (Field3d_operators::calc_mean_profile) and calc_base_state run on the device, stream-ordered, no host
 *                     round trip (y slabs: one ncclAllReduce(sum) of the two partial mean profiles, as master.sum does); then
 *                     calc_buoyancy_tend_2nd (saturation adjustment at the half levels) on wt
 *   _get_thermo_field get_thermo_field(name) for "b", "ql", "N2" (MHH_MOIST_*); out: DEVICE field
 *   _get_buoyancy_surf / _fluxbot   get_buoyancy_surf (b at kstart + b_bot from s_bot[ithl], s_bot[iqt]) and
 *                     get_buoyancy_fluxbot (from s_fluxbot[ithl], s_fluxbot[iqt]); bbot / bfluxbot: DEVICE planes (ijcells)
 *   _nonconverged     number of saturation adjustments that hit the iteration cap since the last call (the reference throws
 *                     "Non-converging saturation adjustment"); synchronises */
MHH_API int mhh_thermo_moist_calc_base_state(mhh_ctx* ctx, const void* thl0, const void* qt0, double pbot);
MHH_API int mhh_thermo_moist_set_profiles(mhh_ctx* ctx, const void* pref, const void* prefh, const void* rhoref, const void* rhorefh,
                                          const void* thvref, const void* thvrefh, const void* exnref, const void* exnrefh);
MHH_API int mhh_thermo_moist_get_profiles(mhh_ctx* ctx, void* pref, void* prefh, void* rhoref, void* rhorefh,
                                          void* thvref, void* thvrefh, void* exnref, void* exnrefh);
MHH_API int mhh_thermo_moist_exec(mhh_ctx* ctx, const mhh_fields* f, const mhh_thermo_moist* tm);
MHH_API int mhh_thermo_moist_get_thermo_field(mhh_ctx* ctx, int which, void* out, const mhh_fields* f, const mhh_thermo_moist* tm);
MHH_API int mhh_thermo_moist_get_buoyancy_surf(mhh_ctx* ctx, void* b, void* bbot, const mhh_fields* f, const mhh_thermo_moist* tm);
MHH_API int mhh_thermo_moist_get_buoyancy_fluxbot(mhh_ctx* ctx, void* bfluxbot, const mhh_fields* f, const mhh_thermo_moist* tm);
MHH_API int mhh_thermo_moist_nonconverged(mhh_ctx* ctx, long long* count);
/* diagnostics: fixed-point sweeps the last base-state integration on the device took (see thermo_moist_kernels.cuh); synchronises */
MHH_API int mhh_thermo_moist_base_state_sweeps(mhh_ctx* ctx, int* sweeps);
MHH_API int mhh_dycore_set_thermo_moist(mhh_ctx* ctx, const mhh_thermo_moist* tm);

/* ---- Pres<TF>::exec / check_divergence: swpres = 2 -> Pres_2 (src/pres_2.cxx:66-105);
 *      swpres = 4 -> Pres_4 (src/pres_4.cxx:76-156; 7-band solve per mode; needs a 4th-order grid; single GPU) -- */
MHH_API int mhh_pres_exec(mhh_ctx* ctx, int swpres, const mhh_fields* f, double sub_dt);
MHH_API int mhh_pres_check_divergence(mhh_ctx* ctx, int swpres, const mhh_fields* f, double* divmax);
/* Spectral pieces, exposed for tests / cuFFT comparison: forward x+y transform of a compact
 * (kmax,jmax,imax) real device array into the context workspace and back. */
MHH_API int mhh_pres_fft_roundtrip(mhh_ctx* ctx, const void* in_compact, void* out_compact, int solve);

/* ---- Timeloop<TF>::exec, rk3 of one (field, tendency) pair  (src/timeloop.cxx:250-286) ------- */
MHH_API int mhh_timeloop_rk3(mhh_ctx* ctx, void* a, void* at, int substep, double dt);

/* ---- One whole RK3 sub-step of the dynamical core, fused (Model::exec, src/model.cxx:356-504):
 * cyclic + vertical ghost cells -> eddy viscosity -> advection+diffusion(+buoyancy) tendencies ->
 * pressure solve -> pressure correction fused with the RK3 update of u,v,w and the scalars. */
MHH_API int mhh_dycore_substep(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, int substep, double dt);
/* The same sub-step in stages, for a host that keeps running its own surface model (Boundary_surface::exec) at the point
 * where Model::exec runs it (src/model.cxx:375-401: exec_viscosity -> thermo.exec -> boundary.exec -> set_ghost_cells ->
 * advec.exec ...):
 *   mhh_dycore_substep_pre        boundary.set_prognostic_cyclic_bcs + set_ghost_cells + diff.exec_viscosity
 *   (host: boundary->exec updates dudz_mo / fluxbot / gradbot ...)
 *   mhh_dycore_set_ghost_cells    boundary.set_ghost_cells of u, v and the scalars again (src/model.cxx:401)
 *   mhh_dycore_substep_post       thermo.exec + advec.exec + diff.exec (fused), pres.exec, timeloop.exec
 * mhh_dycore_substep == pre + post (the 2-D companions are then the caller's values for the whole sub-step).
 * mhh_dycore_tendencies is the fused tendency stage of `post` alone (tests, profiling). LES configurations only. */
MHH_API int mhh_dycore_substep_pre(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm);
MHH_API int mhh_dycore_set_ghost_cells(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm);
MHH_API int mhh_dycore_tendencies(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm);
MHH_API int mhh_dycore_substep_post(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, int substep, double dt);
/* ---- Boundary_surface<TF>: Monin-Obukhov surface model with constant z0 and the z/L lookup solver
 * (src/boundary_surface.cxx:818-990; include/boundary_surface_kernels.h).  init builds the lookup table for (z0m, z0h, z[kstart])
 * like Boundary_surface::init_solver; mbcbot must be MHH_BC_DIRICHLET (no-slip), thermobc MHH_SBC_DIRICHLET or MHH_SBC_FLUX
 * (the bottom BC of th).  exec updates ustar / obuk, the momentum fluxes and gradients, the scalars' surface value or flux and
 * gradient, and the MO gradients dudz_mo / dvdz_mo / dbdz_mo that exec_viscosity uses -- for Thermo_dry (prm->swthermo = 1)
 * or without thermo (stability_neutral). */
MHH_API int mhh_boundary_surface_init(mhh_ctx* ctx, double z0m, double z0h, int mbcbot, int thermobc);
MHH_API int mhh_boundary_surface_exec(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, const mhh_surface* s);
/* The self-driven LES sub-step in Model::exec's order (src/model.cxx:368-504): mhh_dycore_substep_pre, the surface model,
 * Boundary::set_ghost_cells again, mhh_dycore_substep_post. */
MHH_API int mhh_dycore_substep_surface(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, const mhh_surface* s, int substep, double dt);
/* ---- Buffer<TF>::exec (src/buffer.cxx:170-205) and Force<TF>::exec (src/force.cxx:608-700) on the tendencies in `f`.
 * mhh_dycore_set_forcing registers them (a copy of the struct) for the fused sub-steps, which then run buffer.exec and
 * force.exec between diff.exec and pres.exec as Model::exec does (src/model.cxx:416-430); NULL unregisters. */
MHH_API int mhh_buffer_exec(mhh_ctx* ctx, const mhh_fields* f, const mhh_forcing* forcing);
MHH_API int mhh_force_exec(mhh_ctx* ctx, const mhh_fields* f, const mhh_forcing* forcing, double sub_dt);
MHH_API int mhh_dycore_set_forcing(mhh_ctx* ctx, const mhh_forcing* forcing);
/* ---- Field3d_io<TF>::save_field3d / load_field3d (src/field3d_io.cxx:669-751 serial, :57-246 MPI-IO): restart IO of one
 * DEVICE field in the reference's unchanged file layout -- the interior levels [kstart, kend) x jtot x itot as raw TF, no
 * header (Fields::save / load call it for every prognostic field with offset 0, src/fields.cxx:1243-1320).  On y slabs every
 * rank reads / writes its rows of the ONE file (the MPI build's subarray view); single GPU: an existing file is an error as with
 * fopen(..., "wbx").  Synchronous: returns after the file operation. */
MHH_API int mhh_field3d_save(mhh_ctx* ctx, const void* fld, const char* filename, double offset, int kstart, int kend);
MHH_API int mhh_field3d_load(mhh_ctx* ctx, void* fld, const char* filename, double offset, int kstart, int kend);
/* Three sub-steps.  On a single GPU the second call with identical arguments (same structs, same dt) is captured into a CUDA
 * graph on the context's own stream and replayed from then on, ordered after / before the work on the stream set with
 * mhh_set_stream: the ~240 launches of a step are launch-bound on small grids.  Changing arguments (adaptive dt), profiling,
 * slabs, or MHH_GRAPH=0 in the environment run the step eagerly.  mhh_graph_replays: how many steps were replays so far. */
MHH_API int mhh_dycore_step(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, double dt);
MHH_API long long mhh_graph_replays(const mhh_ctx* ctx);
/* End-to-end variant with HOST buffers (ghosted layout): copies u,v,w and the scalars to the
 * device fields in `f`, zeroes nothing (tendencies in `f` are used as they are), runs `nsteps`
 * full RK3 steps and copies u,v,w,scalars back.  Host pointers should be pinned. */
MHH_API int mhh_dycore_step_host(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, double dt, int nsteps,
                         void* h_u, void* h_v, void* h_w, void* const* h_s);

#ifdef __cplusplus
}
#endif
#endif /* MHHB200_H */
