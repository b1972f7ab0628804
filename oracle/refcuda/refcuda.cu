// TEST / MEASUREMENT INFRASTRUCTURE ONLY -- the reference's own single-GPU CUDA kernels, compiled for sm_100a from the
// headers where they lie under /root/reference (nothing is copied), behind a tiny C ABI so that tools/ref_cuda_bench.py can
// time them on the B200 "for context" (BASELINE.json north_star).  The kernels are launched through the reference's own
// launch_grid_kernel fallback path (include/cuda_launcher.h: default block sizes, kernel_launcher auto-tuning disabled,
// which is also what a plain `cmake -DUSECUDA=TRUE` build does).  Never linked into or called by the product library.
#include "advec_2i5_kernels.cuh"
#include "diff_kl_kernels.cuh"
#include "diff_smag2_kl_kernels.cuh"
#include "pres_2_kernels.cuh"
#include "timeloop_kernels.cuh"
#include "cuda_launcher.h"

struct RefArgs
{
    int istart, iend, jstart, jend, kstart, kend, icells, ijcells, imax, jmax, kmax, igc, jgc, kgc;
    double dxi, dyi, dt, tPri, visc;
    void *u, *v, *w, *s, *ut, *vt, *wt, *st, *evisc, *p, *tmp1, *tmp2, *n2;
    void *fluxbotu, *fluxtopu, *fluxbotv, *fluxtopv, *fluxbots, *fluxtops, *dudz, *dvdz, *dbdz, *z0m;
    void *z, *dz, *dzi, *dzhi, *rhoref, *rhorefh, *rhorefi, *rhorefhi, *mlen, *a, *c, *bmati, *bmatj;
};

template <typename TF>
static int launch(int which, const RefArgs& r)
{
    auto P = [](void* p) { return static_cast<TF*>(p); };
    auto C = [](void* p) { return static_cast<const TF*>(p); };
    Grid_layout g = {r.istart, r.iend, r.jstart, r.jend, r.kstart, r.kend, 1, r.icells, r.ijcells};
    Grid_layout gn = {0, r.imax, 0, r.jmax, 0, r.kmax, 1, r.imax, r.imax * r.jmax};           // compact, no ghost cells
    Grid_layout g2 = {0, r.imax, 0, r.jmax, 0, 1, 1, r.imax, r.imax * r.jmax};
    const TF dxi = (TF)r.dxi, dyi = (TF)r.dyi;
    switch (which)
    {
        case 0: launch_grid_kernel<Advec_2i5_kernels::advec_u_g<TF>>(g, P(r.ut), C(r.u), C(r.v), C(r.w), C(r.rhorefi), C(r.rhorefh), C(r.dzi), dxi, dyi); break;
        case 1: launch_grid_kernel<Advec_2i5_kernels::advec_v_g<TF>>(g, P(r.vt), C(r.u), C(r.v), C(r.w), C(r.rhorefi), C(r.rhorefh), C(r.dzi), dxi, dyi); break;
        case 2: launch_grid_kernel<Advec_2i5_kernels::advec_w_g<TF>>(g, P(r.wt), C(r.u), C(r.v), C(r.w), C(r.rhoref), C(r.rhorefhi), C(r.dzhi), dxi, dyi); break;
        case 3: launch_grid_kernel<Advec_2i5_kernels::advec_s_g<TF>>(g, P(r.st), C(r.s), C(r.u), C(r.v), C(r.w), C(r.rhorefi), C(r.rhorefh), C(r.dzi), dxi, dyi); break;
        case 4: launch_grid_kernel<Diff_les_kernels::calc_strain2_g<TF, true>>(g, P(r.evisc), C(r.u), C(r.v), C(r.w), C(r.dudz), C(r.dvdz), C(r.dzi), C(r.dzhi), dxi, dyi); break;
        case 5: launch_grid_kernel<Diff_smag2_kernels::evisc_g<TF, true>>(g, P(r.evisc), C(r.n2), C(r.dbdz), C(r.mlen), C(r.z0m), C(r.z), (TF)r.tPri); break;
        case 6: launch_grid_kernel<Diff_les_kernels::diff_uvw_g<TF, true>>(g, P(r.ut), P(r.vt), P(r.wt), C(r.evisc), C(r.u), C(r.v), C(r.w),
                    C(r.fluxbotu), C(r.fluxtopu), C(r.fluxbotv), C(r.fluxtopv), C(r.dzi), C(r.dzhi), dxi, dyi,
                    C(r.rhoref), C(r.rhorefh), C(r.rhorefi), C(r.rhorefhi), (TF)r.visc); break;
        case 7: launch_grid_kernel<Diff_les_kernels::diff_c_g<TF, true>>(g, P(r.st), C(r.s), C(r.evisc), C(r.fluxbots), C(r.fluxtops), C(r.dzi), C(r.dzhi),
                    dxi * dxi, dyi * dyi, C(r.rhorefi), C(r.rhorefh), (TF)r.tPri, (TF)r.visc); break;
        case 8: launch_grid_kernel<Pres_2_kernels::pres_in_g<TF>>(gn, P(r.p), C(r.u), C(r.v), C(r.w), C(r.ut), C(r.vt), C(r.wt), C(r.dzi), C(r.rhoref), C(r.rhorefh),
                    dxi, dyi, (TF)(1. / r.dt), r.icells, r.ijcells, r.igc, r.jgc, r.kgc); break;
        case 9: launch_grid_kernel<Pres_2_kernels::solve_in_g<TF>>(gn, P(r.p), C(r.tmp1), P(r.tmp2), C(r.a), C(r.c), C(r.dz), C(r.rhoref), C(r.bmati), C(r.bmatj), r.kstart, r.kmax); break;
        case 10: launch_grid_kernel<Pres_2_kernels::tdma_g<TF>>(g2, C(r.a), C(r.tmp2), C(r.c), P(r.p), P(r.tmp1), r.kmax); break;
        case 11: launch_grid_kernel<Pres_2_kernels::solve_out_g<TF>>(gn, P(r.p), C(r.tmp1), r.istart, r.jstart, r.kstart, r.icells, r.ijcells); break;
        case 12: launch_grid_kernel<Pres_2_kernels::pres_out_g<TF>>(g, P(r.ut), P(r.vt), P(r.wt), C(r.p), C(r.dzhi), dxi, dyi); break;
        case 13: launch_grid_kernel<Timeloop_kernels::rk3_g<TF, 0>>(g, P(r.u), P(r.ut), (TF)r.dt); break;
        default: return -1;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

extern "C" __attribute__((visibility("default"))) int refcuda_launch(int which, int is_double, const RefArgs* r)
{
    try { return is_double ? launch<double>(which, *r) : launch<float>(which, *r); }
    catch (...) { return -3; }
}
extern "C" __attribute__((visibility("default"))) int refcuda_args_size() { return (int)sizeof(RefArgs); }
