// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's Monin-Obukhov surface solver: the anonymous-namespace kernels of
// src/boundary_surface.cxx (stability :55-134, stability_neutral :136-180, surfm :182-290, surfs :292-340) and the header
// kernels of include/boundary_surface_kernels.h (prepare_lut :78-138, calc_dutot :140-186, calc_duvdz_mo :188-222,
// calc_dbdz_mo :224-243), called in the order of Boundary_surface<TF>::exec (src/boundary_surface.cxx:836-990).
// The buoyancy inputs of Thermo_dry (get_buoyancy_surf / get_buoyancy_fluxbot / get_db_ref, src/thermo_dry.cxx:700-784) are
// passed in by the caller.   bc codes: 0 Dirichlet, 1 Neumann, 2 Flux, 3 Ustar.
#include <src/boundary_surface.cxx>
#include "ref_common.h"

namespace
{
    Boundary_type bt(int c)
    {
        switch (c) { case 0: return Boundary_type::Dirichlet_type; case 1: return Boundary_type::Neumann_type;
                     case 2: return Boundary_type::Flux_type; default: return Boundary_type::Ustar_type; }
    }
    template <typename TF> Boundary_cyclic<TF>& cyc()
    {
        alignas(16) static char buf[sizeof(Boundary_cyclic<TF>)];
        return *reinterpret_cast<Boundary_cyclic<TF>*>(buf);
    }
}

#define DEFINE(TF, SFX) \
MHH_EXPORT void ref_surface_lut_##SFX(float* zL, float* f, TF z0m, TF z0h, TF zsl, int mbcbot, int thermobc) \
{ bsk::prepare_lut<TF>(zL, f, z0m, z0h, zsl, nzL_lut, bt(mbcbot), bt(thermobc)); } \
MHH_EXPORT int ref_surface_nlut_##SFX() { return nzL_lut; } \
MHH_EXPORT void ref_surface_exec_##SFX(TF* ustar, TF* obuk, int* nobuk, TF* dutot, \
        const TF* u, const TF* v, TF* ubot, TF* vbot, TF* ufluxbot, TF* vfluxbot, TF* ugradbot, TF* vgradbot, \
        TF* bfluxbot, TF* b, TF* bbot, TF db_ref, const TF* z, TF* z0m, TF* z0h, const float* zL_sl, const float* f_sl, \
        int mbcbot, int thermobc, int sw_constant_z0, int neutral, TF* dudz, TF* dvdz, TF* dbdz) \
{ \
    const Ref_geom& g = ref_geom; \
    const int ij = g.icells * g.jcells; \
    bsk::calc_dutot(dutot, u, v, ubot, vbot, g.istart, g.iend, g.jstart, g.jend, g.kstart, g.icells, g.jcells, ij, cyc<TF>()); \
    if (neutral) \
        stability_neutral(ustar, obuk, dutot, z, z0m, g.istart, g.iend, g.jstart, g.jend, g.kstart, g.icells, g.jcells, ij, bt(mbcbot), cyc<TF>()); \
    else if (sw_constant_z0) \
        stability<TF, true>(ustar, obuk, bfluxbot, b, bbot, dutot, z, z0m, z0h, zL_sl, f_sl, nobuk, db_ref, \
                g.istart, g.iend, g.jstart, g.jend, g.kstart, g.icells, g.jcells, ij, bt(mbcbot), bt(thermobc), cyc<TF>()); \
    else \
        stability<TF, false>(ustar, obuk, bfluxbot, b, bbot, dutot, z, z0m, z0h, zL_sl, f_sl, nobuk, db_ref, \
                g.istart, g.iend, g.jstart, g.jend, g.kstart, g.icells, g.jcells, ij, bt(mbcbot), bt(thermobc), cyc<TF>()); \
    surfm(ufluxbot, vfluxbot, ugradbot, vgradbot, ustar, obuk, u, ubot, v, vbot, z0m, z[g.kstart], bt(mbcbot), \
          g.istart, g.iend, g.jstart, g.jend, g.kstart, g.icells, g.jcells, ij, cyc<TF>()); \
    bsk::calc_duvdz_mo(dudz, dvdz, u, v, ubot, vbot, ufluxbot, vfluxbot, ustar, obuk, z0m, z[g.kstart], \
                       g.istart, g.iend, g.jstart, g.jend, g.kstart, g.icells, ij); \
    if (!neutral) \
        bsk::calc_dbdz_mo(dbdz, bfluxbot, ustar, obuk, z[g.kstart], g.istart, g.iend, g.jstart, g.jend, g.icells); \
} \
MHH_EXPORT void ref_surface_surfs_##SFX(TF* varbot, TF* vargradbot, TF* varfluxbot, const TF* ustar, const TF* obuk, const TF* var, \
        const TF* z0h, TF zsl, int bcbot) \
{ \
    const Ref_geom& g = ref_geom; \
    surfs(varbot, vargradbot, varfluxbot, ustar, obuk, var, z0h, zsl, bt(bcbot), g.istart, g.iend, g.jstart, g.jend, g.kstart, \
          g.icells, g.jcells, g.icells * g.jcells, cyc<TF>()); \
}

DEFINE(double, f64)
DEFINE(float, f32)
