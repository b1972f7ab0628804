// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's Pres_4 (src/pres_4.cxx): the member functions init (:159-176), set_values (:179-252),
// input<dim3> (:255-317), solve (:320-529, with hdma :574-730), output<dim3> (:532-571) and calc_divergence (:733-767)
// on a stand-in object (ref_fake_pres.h), called in the order and with the work-array carving of Pres_4::exec (:77-144).
#include <src/pres_4.cxx>
#include "fields.h"
#include "ref_fake_pres.h"

#define DEFINE(TF, SFX) \
MHH_EXPORT void* ref_pres_4_create_##SFX(void* fft, int kcells) \
{ \
    std::vector<TF> ones(kcells, TF(1)); \
    Pres_4<TF>* p = fake_pres<Pres_4<TF>, TF>(fft, ones.data(), ones.data(), kcells); \
    vec_new(p->bmati); vec_new(p->bmatj); \
    vec_new(p->m1); vec_new(p->m2); vec_new(p->m3); vec_new(p->m4); vec_new(p->m5); vec_new(p->m6); vec_new(p->m7); \
    p->Pres_4<TF>::init();        /* qualified: no virtual dispatch (the image has no vptr) */ \
    return p; \
} \
MHH_EXPORT void ref_pres_4_set_values_##SFX(void* h, TF* bmati, TF* bmatj, TF* m /* 7 x kmax */) \
{ \
    Pres_4<TF>* p = static_cast<Pres_4<TF>*>(h); \
    p->Pres_4<TF>::set_values(); \
    vec_out(p->bmati, bmati); vec_out(p->bmatj, bmatj); \
    const int kmax = p->grid.get_grid_data().kmax; \
    vec_out(p->m1, m); vec_out(p->m2, m + kmax); vec_out(p->m3, m + 2*kmax); vec_out(p->m4, m + 3*kmax); \
    vec_out(p->m5, m + 4*kmax); vec_out(p->m6, m + 5*kmax); vec_out(p->m7, m + 6*kmax); \
} \
MHH_EXPORT void ref_pres_4_input_##SFX(void* h, TF* p, const TF* u, const TF* v, const TF* w, TF* ut, TF* vt, TF* wt, TF dt) \
{ \
    Pres_4<TF>* o = static_cast<Pres_4<TF>*>(h); \
    const Grid_data<TF>& gd = o->grid.get_grid_data(); \
    if (gd.jtot == 1) o->template input<false>(p, u, v, w, ut, vt, wt, gd.dzi4.data(), dt); \
    else o->template input<true>(p, u, v, w, ut, vt, wt, gd.dzi4.data(), dt); \
} \
MHH_EXPORT void ref_pres_4_solve_##SFX(void* h, TF* p, TF* tmp1) \
{ \
    Pres_4<TF>* o = static_cast<Pres_4<TF>*>(h); \
    const Grid_data<TF>& gd = o->grid.get_grid_data(); \
    const int jslice = 1; \
    const int ns = gd.iblock*jslice*(gd.kmax+4); \
    std::vector<TF> tmp2(4*(size_t)ns), tmp3(4*(size_t)ns); \
    o->solve(p, tmp1, gd.dz.data(), o->m1.data(), o->m2.data(), o->m3.data(), o->m4.data(), o->m5.data(), o->m6.data(), o->m7.data(), \
             &tmp2[0*ns], &tmp2[1*ns], &tmp2[2*ns], &tmp2[3*ns], &tmp3[0*ns], &tmp3[1*ns], &tmp3[2*ns], &tmp3[3*ns], \
             o->bmati.data(), o->bmatj.data(), jslice); \
} \
MHH_EXPORT void ref_pres_4_output_##SFX(void* h, TF* ut, TF* vt, TF* wt, TF* p) \
{ \
    Pres_4<TF>* o = static_cast<Pres_4<TF>*>(h); \
    const Grid_data<TF>& gd = o->grid.get_grid_data(); \
    if (gd.jtot == 1) o->template output<false>(ut, vt, wt, p, gd.dzhi4.data()); \
    else o->template output<true>(ut, vt, wt, p, gd.dzhi4.data()); \
} \
MHH_EXPORT double ref_pres_4_divergence_##SFX(void* h, const TF* u, const TF* v, const TF* w) \
{ \
    Pres_4<TF>* o = static_cast<Pres_4<TF>*>(h); \
    return (double)o->calc_divergence(u, v, w, o->grid.get_grid_data().dzi4.data()); \
}

DEFINE(double, f64)
DEFINE(float, f32)
