// mhhb200 -- damping layer and large-scale forcings, the cheap streaming terms between diff.exec and pres.exec
// (SURVEY 8f, N3).  Reference behaviour restated (never copied):
//   Buffer<TF>::exec / calc_buffer            src/buffer.cxx:38-58, 98-124, 170-205
//   Force<TF>::exec                           src/force.cxx:608-700: enforce_fixed_flux (:65-75), add_pressure_force (:47-62),
//                                             calc_coriolis_2nd (:78-108), calc_coriolis_4th (:110-152),
//                                             calc_large_scale_source (:154-170), advec_wls_2nd_local (:238-272)
//   Field3d_operators<TF>::calc_mean          src/field3d_operators.cxx:135-155
// All of them touch a tendency once; they run as one launch per term on the top levels (buffer) or the interior.
#pragma once
#include "common.cuh"
#include "order4_kernels.cuh"      // W4 interpolation weights

namespace mhh {

// at -= sigma ((z[k] - zstart) / (zsize - zstart))^beta (a - abuf[k]),  k = kbuf .. kend-1
template <typename TF>
__global__ void buffer_kernel(TF* __restrict__ at, const TF* __restrict__ a, const TF* __restrict__ abuf, const TF* __restrict__ sigmaz,
        const int kbuf, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = kbuf + blockIdx.z;
    if (i >= g.iend || j >= g.jend || k >= g.kend) return;
    const long long ijk = i + (long long)j * g.icells + k * g.ijcells;
    at[ijk] -= sigmaz[k] * (a[ijk] - abuf[k]);
}

// dz-weighted sums of u and ut over the interior, accumulated in double: sums[0] += sum(u dz), sums[1] += sum(ut dz)
template <typename TF>
__global__ void __launch_bounds__(256) mean_uut_kernel(const TF* __restrict__ u, const TF* __restrict__ ut, const GridDev<TF> g, double* __restrict__ sums)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    double s0 = 0., s1 = 0.;
    if (i < g.iend && j < g.jend)
    {
        const long long ijk = i + (long long)j * g.icells + k * g.ijcells;
        const double dz = (double)g.dz[k];
        s0 = (double)u[ijk] * dz; s1 = (double)ut[ijk] * dz;
    }
    for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
    __shared__ double r0[8], r1[8];
    const int tid = threadIdx.x + blockDim.x * threadIdx.y;
    if ((tid & 31) == 0) { r0[tid >> 5] = s0; r1[tid >> 5] = s1; }
    __syncthreads();
    if (tid == 0)
    {
        const int nw = (blockDim.x * blockDim.y + 31) / 32;
        for (int w = 1; w < nw; ++w) { s0 += r0[w]; s1 += r1[w]; }
        atomicAdd(&sums[0], s0); atomicAdd(&sums[1], s1);
    }
}

// enforce_fixed_flux: ut += (uflux - u_mean - utrans) / dt - ut_mean, the means taken from the device sums (no host round trip);
// FIXED = false: ut += fbody (pressure-gradient forcing)
template <typename TF, bool FIXED>
__global__ void body_force_kernel(TF* __restrict__ ut, const double* __restrict__ sums, const double inv_vol, const TF uflux, const TF utrans,
        const TF dt, const TF fbody_in, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    TF fbody = fbody_in;
    if (FIXED)
    {
        const TF u_mean = (TF)(sums[0] * inv_vol), ut_mean = (TF)(sums[1] * inv_vol);
        fbody = (uflux - u_mean - utrans) / dt - ut_mean;
    }
    ut[i + (long long)j * g.icells + k * g.ijcells] += fbody;
}

// Coriolis force with a geostrophic wind profile, 2nd- or 4th-order interpolation of the other component
template <typename TF, int ORDER>
__global__ void coriolis_kernel(TF* __restrict__ ut, TF* __restrict__ vt, const TF* __restrict__ u, const TF* __restrict__ v,
        const TF* __restrict__ ug, const TF* __restrict__ vg, const TF fc, const TF ugrid, const TF vgrid, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells;
    const long long ijk = i + j * jj + k * g.ijcells;
    if (ORDER == 2)
    {
        ut[ijk] += fc * (TF(0.25) * (v[ijk - 1] + v[ijk] + v[ijk - 1 + jj] + v[ijk + jj]) + vgrid - vg[k]);
        vt[ijk] -= fc * (TF(0.25) * (u[ijk - jj] + u[ijk] + u[ijk + 1 - jj] + u[ijk + 1]) + ugrid - ug[k]);
    }
    else
    {
        const TF c0 = W4<TF>::ci0, c1 = W4<TF>::ci1, c2 = W4<TF>::ci2, c3 = W4<TF>::ci3;
        auto rowv = [&](const long long o) { return c0 * v[o - 2] + c1 * v[o - 1] + c2 * v[o] + c3 * v[o + 1]; };
        auto rowu = [&](const long long o) { return c0 * u[o - 1] + c1 * u[o] + c2 * u[o + 1] + c3 * u[o + 2]; };
        ut[ijk] += fc * ((c0 * rowv(ijk - jj) + c1 * rowv(ijk) + c2 * rowv(ijk + jj) + c3 * rowv(ijk + 2 * jj)) + vgrid - vg[k]);
        vt[ijk] -= fc * ((c0 * rowu(ijk - 2 * jj) + c1 * rowu(ijk - jj) + c2 * rowu(ijk) + c3 * rowu(ijk + jj)) + ugrid - ug[k]);
    }
}

// scalar forcings: st += sls[k] (large-scale source) and / or first-order upwind subsidence with the profile wls[k]
template <typename TF>
__global__ void scalar_forcing_kernel(TF* __restrict__ st, const TF* __restrict__ s, const TF* __restrict__ sls, const TF* __restrict__ wls, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long kk = g.ijcells;
    const long long ijk = i + (long long)j * g.icells + k * kk;
    TF t = st[ijk];
    if (sls) t += sls[k];
    if (wls)
    {
        const TF w = wls[k];
        if (w > TF(0.)) t -= w * (s[ijk] - s[ijk - kk]) * g.dzhi[k];
        else t -= w * (s[ijk + kk] - s[ijk]) * g.dzhi[k + 1];
    }
    st[ijk] = t;
}

} // namespace mhh
