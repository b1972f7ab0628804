// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
//
// Stand-ins for the two reference *objects* the free-function kernels take by reference:
//   * Master::max/min/sum -- serial build: no-ops (reference src/master_serial.cxx:98-104).
//   * Boundary_cyclic<TF>::exec -- the serial periodic fill, RESTATED from
//     reference src/boundary_cyclic.cxx:369-443 (member function that needs live
//     Grid/Master objects, so it cannot be reached by #include alone).
#include <cstddef>
#include "defines.h"
#include "master.h"
#include "boundary_cyclic.h"
#include "grid.h"
#include "transpose.h"
#include "ref_common.h"

Ref_geom ref_geom = {};

void Master::max(double*, int) {}
void Master::max(float*, int) {}
void Master::min(double*, int) {}
void Master::min(float*, int) {}
void Master::sum(double*, int) {}
void Master::sum(float*, int) {}
void Master::sum(int*, int) {}

// Periodic ghost-cell fill, x first over all (j,k) incl. ghosts, then y over all i incl.
// ghosts (corners by ordering); jtot==1 replicates the single row over interior k only.
template<typename TF>
void Boundary_cyclic<TF>::exec(TF* const restrict data, Edge edge)
{
    const Ref_geom& g = ref_geom;
    const std::ptrdiff_t jj = g.icells;
    const std::ptrdiff_t kk = (std::ptrdiff_t)g.icells * g.jcells;

    if (edge == Edge::East_west_edge || edge == Edge::Both_edges)
    {
        for (int k = 0; k < g.kcells; ++k)
            for (int j = 0; j < g.jcells; ++j)
            {
                TF* row = data + j*jj + k*kk;
                for (int i = 0; i < g.igc; ++i)
                    row[i] = row[g.iend - g.igc + i];
                for (int i = 0; i < g.igc; ++i)
                    row[g.iend + i] = row[g.istart + i];
            }
    }
    if (edge == Edge::North_south_edge || edge == Edge::Both_edges)
    {
        if (g.jtot > 1)
        {
            for (int k = 0; k < g.kcells; ++k)
            {
                TF* slab = data + k*kk;
                for (int j = 0; j < g.jgc; ++j)
                    for (int i = 0; i < g.icells; ++i)
                        slab[i + j*jj] = slab[i + (g.jend - g.jgc + j)*jj];
                for (int j = 0; j < g.jgc; ++j)
                    for (int i = 0; i < g.icells; ++i)
                        slab[i + (g.jend + j)*jj] = slab[i + (g.jstart + j)*jj];
            }
        }
        else
        {
            for (int k = g.kstart; k < g.kend; ++k)
            {
                TF* slab = data + k*kk;
                for (int j = 0; j < g.jgc; ++j)
                    for (int i = 0; i < g.icells; ++i)
                    {
                        slab[i + j*jj]            = slab[i + g.jstart*jj];
                        slab[i + (g.jend + j)*jj] = slab[i + g.jstart*jj];
                    }
            }
        }
    }
}
// 2-D variant (reference src/boundary_cyclic.cxx:445-507): one level, jtot == 1 replicates the single row.
template<typename TF>
void Boundary_cyclic<TF>::exec_2d(TF* restrict data)
{
    const Ref_geom& g = ref_geom;
    const std::ptrdiff_t jj = g.icells;
    for (int j = 0; j < g.jcells; ++j)
    {
        TF* row = data + j*jj;
        for (int i = 0; i < g.igc; ++i)
            row[i] = row[g.iend - g.igc + i];
        for (int i = 0; i < g.igc; ++i)
            row[g.iend + i] = row[g.istart + i];
    }
    if (g.jtot > 1)
    {
        for (int j = 0; j < g.jgc; ++j)
            for (int i = 0; i < g.icells; ++i)
                data[i + j*jj] = data[i + (g.jend - g.jgc + j)*jj];
        for (int j = 0; j < g.jgc; ++j)
            for (int i = 0; i < g.icells; ++i)
                data[i + (g.jend + j)*jj] = data[i + (g.jstart + j)*jj];
    }
    else
    {
        for (int j = 0; j < g.jgc; ++j)
            for (int i = 0; i < g.icells; ++i)
            {
                data[i + j*jj]            = data[i + g.jstart*jj];
                data[i + (g.jend + j)*jj] = data[i + g.jstart*jj];
            }
    }
}
template void Boundary_cyclic<double>::exec_2d(double* restrict);
template void Boundary_cyclic<float>::exec_2d(float* restrict);
template void Boundary_cyclic<double>::exec(double* const restrict, Edge);
template void Boundary_cyclic<float>::exec(float* const restrict, Edge);

MHH_EXPORT void ref_set_geom(int itot, int jtot, int ktot, int igc, int jgc, int kgc)
{
    Ref_geom& g = ref_geom;
    g.igc = igc; g.jgc = jgc; g.jtot = jtot;
    g.icells = itot + 2*igc; g.jcells = jtot + 2*jgc; g.kcells = ktot + 2*kgc;
    g.istart = igc; g.iend = igc + itot;
    g.jstart = jgc; g.jend = jgc + jtot;
    g.kstart = kgc; g.kend = kgc + ktot;
}

MHH_EXPORT void ref_boundary_cyclic_f64(double* data, int edge)
{
    alignas(16) static char buf[sizeof(Boundary_cyclic<double>)];
    reinterpret_cast<Boundary_cyclic<double>*>(buf)->exec(data, static_cast<Edge>(edge));
}
MHH_EXPORT void ref_boundary_cyclic_f32(float* data, int edge)
{
    alignas(16) static char buf[sizeof(Boundary_cyclic<float>)];
    reinterpret_cast<Boundary_cyclic<float>*>(buf)->exec(data, static_cast<Edge>(edge));
}

// ---- stand-ins for the objects the reference's FFT / Pres_2 / Pres_4 member functions reach through (tier-2 pin of the
// pressure glue): Transpose / Boundary_cyclic::init / Master::print_message are the serial no-ops; the Grid is the
// reference's own (ref_grid.cpp).
cuda_raw_buffer::cuda_raw_buffer(size_t) {}        // CPU build: no device allocation (src/cuda_buffer.cxx, host branch)
void cuda_raw_buffer::reallocate(size_t) {}

template<typename TF> Transpose<TF>::Transpose(Master& m, Grid<TF>& g) : master(m), grid(g), mpi_types_allocated(false) {}
template<typename TF> Transpose<TF>::~Transpose() {}
template<typename TF> void Transpose<TF>::init() {}
template Transpose<double>::Transpose(Master&, Grid<double>&);
template Transpose<float>::Transpose(Master&, Grid<float>&);
template Transpose<double>::~Transpose();
template Transpose<float>::~Transpose();
template void Transpose<double>::init();
template void Transpose<float>::init();

template<typename TF> void Boundary_cyclic<TF>::init() {}
template void Boundary_cyclic<double>::init();
template void Boundary_cyclic<float>::init();

void Master::print_message(const char*, ...) {}
