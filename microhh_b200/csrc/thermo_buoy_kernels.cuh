// mhhb200 -- Thermo_buoy<TF> (the prognostic scalar IS the buoyancy; reference src/thermo_buoy.cxx).
//   calc_N2                                   :48-62      N2 = 0.5 (b[k+1] - b[k-1]) dzi[k] + bg_n2
//   calc_buoyancy_tend_2nd / _4th             :93-108, 166-183
//   calc_buoyancy_tend_u / _w / _b (slope)    :110-164, 185-246
//   calc_baroclinic_2nd / _4th                :248-282
// Thermo_buoy::exec (:345-391) applies them as separate "+=" passes over ut, wt and bt; every term here lands on its own
// element in that same order (slope term of bt before the baroclinic one), so ONE pass does the work of up to four: b, u, v,
// w are read once and each tendency is read and written once -- 10 array passes (slope + baroclinic) where the reference's
// four kernels move 13; the plain case is the reference's single pass (3 arrays).  HBM-bound, point-wise: every stencil is a
// two- or four-point interpolation along one axis, the x / y neighbours come out of L1 and the z neighbours out of L2.
#pragma once
#include "common.cuh"
#include "order4_kernels.cuh"

namespace mhh {

template <typename TF>
struct BuoyArgs
{
    TF *ut, *wt, *bt;
    const TF *b, *u, *v, *w;
    TF sinalpha, cosalpha, n2, utrans, dbdy_ls;
    int slope, baroclinic;               // slope: has_slope || has_N2 (src/thermo_buoy.cxx:352, 371)
};

template <typename TF, int ORDER>
__device__ __forceinline__ TF buoy_interp(const TF* __restrict__ q, long long o, long long s)
{
    // value half a cell below index o along stride s: interp2(q[o-s], q[o]) or interp4c(q[o-2s], q[o-s], q[o], q[o+s])
    if (ORDER == 4) return i4m(q[o - 2 * s], q[o - s], q[o], q[o + s]);
    return interp2(q[o - s], q[o]);
}

template <typename TF, int ORDER>
__global__ void __launch_bounds__(256) thermo_buoy_kernel(const BuoyArgs<TF> a, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ijk = i + j * jj + k * kk;
    if (a.slope)
    {
        a.ut[ijk] += a.sinalpha * buoy_interp<TF, ORDER>(a.b, ijk, 1);
        if (k > g.kstart) a.wt[ijk] += a.cosalpha * buoy_interp<TF, ORDER>(a.b, ijk, kk);
        // u and w at the cell centre: half a cell ABOVE index ijk along their own axis
        const TF ui = buoy_interp<TF, ORDER>(a.u, ijk + 1, 1);
        const TF wi = buoy_interp<TF, ORDER>(a.w, ijk + kk, kk);
        TF bt = a.bt[ijk];
        bt -= a.n2 * (a.sinalpha * (ui + a.utrans) + a.cosalpha * wi);
        if (a.baroclinic) bt -= a.dbdy_ls * buoy_interp<TF, ORDER>(a.v, ijk + jj, jj);
        a.bt[ijk] = bt;
    }
    else
    {
        if (k > g.kstart) a.wt[ijk] += buoy_interp<TF, ORDER>(a.b, ijk, kk);
        if (a.baroclinic) a.bt[ijk] -= a.dbdy_ls * buoy_interp<TF, ORDER>(a.v, ijk + jj, jj);
    }
}

// Thermo_buoy::get_thermo_field("N2") (src/thermo_buoy.cxx:48-62, 410-413)
template <typename TF>
__global__ void thermo_buoy_n2_kernel(TF* __restrict__ n2, const TF* __restrict__ b, const TF bg_n2, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long ijk = i + (long long)j * g.icells + k * g.ijcells;
    n2[ijk] = TF(0.5) * (b[ijk + g.ijcells] - b[ijk - g.ijcells]) * g.dzi[k] + bg_n2;
}

} // namespace mhh
