// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// The reference's own Grid<TF> (src/grid.cxx) compiled where it lies: init() (:106-198, index arithmetic and array
// sizes), calculate() (:245-376, ghost levels of z, zh, dz / dzh and their reciprocals, the 4th-order metrics dzi4 / dzhi4)
// and get_grid_data() (:379-383) run on a Grid<TF> image whose constructor inputs (the [grid] ini items, :47-93) and the
// interior z profile (Grid::create, :201-214, NetCDF in the reference) are poked in by the test.  The constructor itself
// needs Input / NetCDF and never runs.  Compiled with -fno-access-control.
#include <src/grid.cxx>
#include <cstdlib>
#include <new>
#include "ref_common.h"

// serial no-ops (reference src/grid_serial.cxx:33-44)
template<typename TF> void Grid<TF>::init_mpi() {}
template<typename TF> void Grid<TF>::exit_mpi() {}
template void Grid<double>::init_mpi();
template void Grid<float>::init_mpi();
template void Grid<double>::exit_mpi();
template void Grid<float>::exit_mpi();

// src/grid.cxx instantiates Grid<double> only (no FLOAT_SINGLE here): the single-precision members other TUs link against
template const Grid_data<float>& Grid<float>::get_grid_data() const;
template void Grid<float>::init();
template void Grid<float>::calculate();
template void Grid<float>::check_ghost_cells();

void* ref_master_image()
{
    static Master* m = [] {
        Master* m = static_cast<Master*>(std::calloc(1, sizeof(Master)));
        m->md.nprocs = 1; m->md.npx = 1; m->md.npy = 1;       // serial run (reference src/master_serial.cxx)
        return m; }();
    return m;
}

template<typename TF> Grid<TF>* grid_image()
{
    static Grid<TF>* g = [] {
        char* b = static_cast<char*>(std::calloc(1, sizeof(Grid<TF>)));
        *reinterpret_cast<void**>(b) = ref_master_image();      // Master& master is the first member (include/grid.h:182)
        Grid<TF>* g = reinterpret_cast<Grid<TF>*>(b);
        new (&g->gd) Grid_data<TF>();
        return g; }();
    return g;
}
void* ref_grid_image(int is_float) { return is_float ? static_cast<void*>(grid_image<float>()) : static_cast<void*>(grid_image<double>()); }

extern "C" void ref_set_geom(int itot, int jtot, int ktot, int igc, int jgc, int kgc);

#define DEFINE(TF, SFX) \
/* z: kcells values, interior levels filled (what Grid::create reads from the input file); order 2 or 4 */ \
MHH_EXPORT void ref_grid_setup_##SFX(int itot, int jtot, int ktot, TF xsize, TF ysize, TF zsize, int igc, int jgc, int kgc, \
        int order, const TF* z) \
{ \
    Grid<TF>* g = grid_image<TF>(); \
    Grid_data<TF>& gd = g->gd; \
    gd.itot = itot; gd.jtot = jtot; gd.ktot = ktot; gd.xsize = xsize; gd.ysize = ysize; gd.zsize = zsize; \
    gd.igc = igc; gd.jgc = jgc; gd.kgc = kgc; \
    g->spatial_order = order == 4 ? Grid_order::Fourth : Grid_order::Second; \
    g->init(); \
    for (int k = gd.kstart; k < gd.kend; ++k) gd.z[k] = z[k]; \
    g->calculate(); \
    ref_set_geom(itot, jtot, ktot, igc, jgc, kgc); \
} \
/* out: 8 x kcells = z, zh, dz, dzh, dzi, dzhi, dzi4, dzhi4; scal: dx, dy, dxi, dyi, dzhi4bot, dzhi4top */ \
MHH_EXPORT void ref_grid_get_##SFX(TF* out, TF* scal) \
{ \
    const Grid_data<TF>& gd = grid_image<TF>()->get_grid_data(); \
    const std::vector<TF>* v[8] = {&gd.z, &gd.zh, &gd.dz, &gd.dzh, &gd.dzi, &gd.dzhi, &gd.dzi4, &gd.dzhi4}; \
    for (int n = 0; n < 8; ++n) for (int k = 0; k < gd.kcells; ++k) out[n*gd.kcells + k] = (*v[n])[k]; \
    scal[0] = gd.dx; scal[1] = gd.dy; scal[2] = gd.dxi; scal[3] = gd.dyi; scal[4] = gd.dzhi4bot; scal[5] = gd.dzhi4top; \
}
DEFINE(double, f64)
DEFINE(float, f32)
