// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's Pres_2 (src/pres_2.cxx):
//   * `tdma`, the anonymous-namespace tridiagonal solver (:202-263);
//   * the member functions init (:108-122), set_values (:125-153), input (:156-196), solve (:267-362) and
//     output (:365-387) on a stand-in object (ref_fake_pres.h), called in the order of Pres_2::exec (:66-94).
#include <src/pres_2.cxx>
#include "fields.h"
#include "ref_fake_pres.h"

#define DEFINE(TF, SFX) \
MHH_EXPORT void ref_pres_2_tdma_##SFX(TF* a, TF* b, TF* c, TF* p, TF* work2d, TF* work3d, int iblock, int jblock, int kmax) \
{ tdma<TF>(a, b, c, p, work2d, work3d, iblock, jblock, kmax); } \
MHH_EXPORT void* ref_pres_2_create_##SFX(void* fft, const TF* rhoref, const TF* rhorefh, int kcells) \
{ \
    Pres_2<TF>* p = fake_pres<Pres_2<TF>, TF>(fft, rhoref, rhorefh, kcells); \
    vec_new(p->bmati); vec_new(p->bmatj); vec_new(p->a); vec_new(p->c); vec_new(p->work2d); \
    p->Pres_2<TF>::init();    /* qualified: no virtual dispatch (the image has no vptr); also runs fft.init() */ \
    return p; \
} \
MHH_EXPORT void ref_pres_2_set_values_##SFX(void* h, TF* bmati, TF* bmatj, TF* a, TF* c) \
{ \
    Pres_2<TF>* p = static_cast<Pres_2<TF>*>(h); \
    p->Pres_2<TF>::set_values(); \
    vec_out(p->bmati, bmati); vec_out(p->bmatj, bmatj); vec_out(p->a, a); vec_out(p->c, c); \
} \
MHH_EXPORT void ref_pres_2_input_##SFX(void* h, TF* p, const TF* u, const TF* v, const TF* w, TF* ut, TF* vt, TF* wt, TF dt) \
{ \
    Pres_2<TF>* o = static_cast<Pres_2<TF>*>(h); \
    const Grid_data<TF>& gd = o->grid.get_grid_data(); \
    o->input(p, u, v, w, ut, vt, wt, gd.dzi.data(), o->fields.rhoref.data(), o->fields.rhorefh.data(), dt); \
} \
MHH_EXPORT void ref_pres_2_solve_##SFX(void* h, TF* p, TF* tmp1, TF* tmp2) \
{ \
    Pres_2<TF>* o = static_cast<Pres_2<TF>*>(h); \
    const Grid_data<TF>& gd = o->grid.get_grid_data(); \
    o->solve(p, tmp1, tmp2, gd.dz.data(), o->fields.rhoref.data()); \
} \
MHH_EXPORT void ref_pres_2_output_##SFX(void* h, TF* ut, TF* vt, TF* wt, const TF* p) \
{ \
    Pres_2<TF>* o = static_cast<Pres_2<TF>*>(h); \
    o->output(ut, vt, wt, p, o->grid.get_grid_data().dzhi.data()); \
}

DEFINE(double, f64)
DEFINE(float, f32)
