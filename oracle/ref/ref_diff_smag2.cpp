// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's Smagorinsky CPU kernels: Diff_kernels::calc_strain2 / diff_u /
// diff_v / diff_w / diff_c / calc_dnmul (reference include/diff_kernels.h:34-511) and
// calc_evisc / calc_evisc_neutral (reference src/diff_smag2.cxx:47-269).
#include <src/diff_smag2.cxx>
#include "ref_common.h"

#define GEOM const Ref_geom& g = ref_geom
#define RANGE g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells
#define RANGE3 g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.jcells, g.icells*g.jcells

template<typename TF> static Boundary_cyclic<TF>& cyclic_stub()
{
    alignas(16) static char buf[sizeof(Boundary_cyclic<TF>)];
    return *reinterpret_cast<Boundary_cyclic<TF>*>(buf);
}

#define DEFINE(TF, SFX) \
MHH_EXPORT void ref_diff_strain2_##SFX(TF* strain2, const TF* u, const TF* v, const TF* w, const TF* ugradbot, const TF* vgradbot, \
        const TF* z, const TF* dzi, const TF* dzhi, TF dxi, TF dyi, int surface) \
{ GEOM; if (surface) dk::calc_strain2<TF, Surface_model::Enabled>(strain2, u, v, w, ugradbot, vgradbot, z, dzi, dzhi, dxi, dyi, RANGE); \
        else         dk::calc_strain2<TF, Surface_model::Disabled>(strain2, u, v, w, ugradbot, vgradbot, z, dzi, dzhi, dxi, dyi, RANGE); } \
MHH_EXPORT void ref_diff_evisc_##SFX(TF* evisc, const TF* u, const TF* v, const TF* w, const TF* N2, const TF* bgradbot, \
        const TF* z, const TF* dz, const TF* dzi, const TF* z0m, TF dx, TF dy, TF cs, TF tPr, int surface, int mason) \
{ GEOM; Boundary_cyclic<TF>& bc = cyclic_stub<TF>(); \
  if (surface && mason)  calc_evisc<TF, Surface_model::Enabled, true >(evisc, u, v, w, N2, bgradbot, z, dz, dzi, z0m, dx, dy, cs, tPr, RANGE3, bc); \
  else if (surface)      calc_evisc<TF, Surface_model::Enabled, false>(evisc, u, v, w, N2, bgradbot, z, dz, dzi, z0m, dx, dy, cs, tPr, RANGE3, bc); \
  else if (mason)        calc_evisc<TF, Surface_model::Disabled, true >(evisc, u, v, w, N2, bgradbot, z, dz, dzi, z0m, dx, dy, cs, tPr, RANGE3, bc); \
  else                   calc_evisc<TF, Surface_model::Disabled, false>(evisc, u, v, w, N2, bgradbot, z, dz, dzi, z0m, dx, dy, cs, tPr, RANGE3, bc); } \
MHH_EXPORT void ref_diff_evisc_neutral_##SFX(TF* evisc, const TF* u, const TF* v, const TF* w, const TF* ufluxbot, const TF* vfluxbot, \
        const TF* z, const TF* dz, const TF* dzhi, const TF* z0m, TF dx, TF dy, TF zsize, TF cs, TF visc, int surface, int mason) \
{ GEOM; Boundary_cyclic<TF>& bc = cyclic_stub<TF>(); \
  if (surface && mason)  calc_evisc_neutral<TF, Surface_model::Enabled, true >(evisc, u, v, w, ufluxbot, vfluxbot, z, dz, dzhi, z0m, dx, dy, zsize, cs, visc, RANGE3, bc); \
  else if (surface)      calc_evisc_neutral<TF, Surface_model::Enabled, false>(evisc, u, v, w, ufluxbot, vfluxbot, z, dz, dzhi, z0m, dx, dy, zsize, cs, visc, RANGE3, bc); \
  else if (mason)        calc_evisc_neutral<TF, Surface_model::Disabled, true >(evisc, u, v, w, ufluxbot, vfluxbot, z, dz, dzhi, z0m, dx, dy, zsize, cs, visc, RANGE3, bc); \
  else                   calc_evisc_neutral<TF, Surface_model::Disabled, false>(evisc, u, v, w, ufluxbot, vfluxbot, z, dz, dzhi, z0m, dx, dy, zsize, cs, visc, RANGE3, bc); } \
MHH_EXPORT void ref_diff_u_##SFX(TF* ut, const TF* u, const TF* v, const TF* w, const TF* dzi, const TF* dzhi, TF dxi, TF dyi, \
        const TF* evisc, const TF* fluxbot, const TF* fluxtop, const TF* rhoref, const TF* rhorefh, TF visc, int surface) \
{ GEOM; if (surface) dk::diff_u<TF, Surface_model::Enabled >(ut, u, v, w, dzi, dzhi, dxi, dyi, evisc, fluxbot, fluxtop, rhoref, rhorefh, visc, RANGE); \
        else         dk::diff_u<TF, Surface_model::Disabled>(ut, u, v, w, dzi, dzhi, dxi, dyi, evisc, fluxbot, fluxtop, rhoref, rhorefh, visc, RANGE); } \
MHH_EXPORT void ref_diff_v_##SFX(TF* vt, const TF* u, const TF* v, const TF* w, const TF* dzi, const TF* dzhi, TF dxi, TF dyi, \
        const TF* evisc, const TF* fluxbot, const TF* fluxtop, const TF* rhoref, const TF* rhorefh, TF visc, int surface) \
{ GEOM; if (surface) dk::diff_v<TF, Surface_model::Enabled >(vt, u, v, w, dzi, dzhi, dxi, dyi, evisc, fluxbot, fluxtop, rhoref, rhorefh, visc, RANGE); \
        else         dk::diff_v<TF, Surface_model::Disabled>(vt, u, v, w, dzi, dzhi, dxi, dyi, evisc, fluxbot, fluxtop, rhoref, rhorefh, visc, RANGE); } \
MHH_EXPORT void ref_diff_w_##SFX(TF* wt, const TF* u, const TF* v, const TF* w, const TF* dzi, const TF* dzhi, TF dxi, TF dyi, \
        const TF* evisc, const TF* rhoref, const TF* rhorefh, TF visc) \
{ GEOM; dk::diff_w<TF>(wt, u, v, w, dzi, dzhi, dxi, dyi, evisc, rhoref, rhorefh, visc, RANGE); } \
MHH_EXPORT void ref_diff_c_##SFX(TF* at, const TF* a, const TF* dzi, const TF* dzhi, TF dxidxi, TF dyidyi, \
        const TF* evisc, const TF* fluxbot, const TF* fluxtop, const TF* rhoref, const TF* rhorefh, TF tPr, TF visc, int surface) \
{ GEOM; if (surface) dk::diff_c<TF, Surface_model::Enabled >(at, a, dzi, dzhi, dxidxi, dyidyi, evisc, fluxbot, fluxtop, rhoref, rhorefh, tPr, visc, RANGE); \
        else         dk::diff_c<TF, Surface_model::Disabled>(at, a, dzi, dzhi, dxidxi, dyidyi, evisc, fluxbot, fluxtop, rhoref, rhorefh, tPr, visc, RANGE); } \
MHH_EXPORT double ref_diff_dnmul_##SFX(const TF* evisc, const TF* dzi, TF dxidxi, TF dyidyi, TF tPr) \
{ GEOM; return (double)dk::calc_dnmul<TF>(evisc, dzi, dxidxi, dyidyi, tPr, RANGE); }

DEFINE(double, f64)
DEFINE(float, f32)
