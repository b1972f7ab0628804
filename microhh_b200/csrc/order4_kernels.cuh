// mhhb200 -- 4th-order DNS schemes: Advec_4 and Diff_4 (point-wise kernels).
//
// Advec_4: the tendency is a 4th-order divergence (weights cg) of products of 4th-order interpolations (weights ci)
// of the advecting velocity and the advected quantity; in the first / last row the outermost vertical flux uses the
// one-sided bi / ti interpolation.  Diff_4: nu times a 7-point (cdg) laplacian in x and y and a div(grad) with cg
// weights and one-sided bg / tg gradients at the walls in z.  The grid is the 4th-order one (three ghost cells in every
// direction, metrics dzi4 / dzhi4, src/grid.cxx:306-375).  No base-state density: Boussinesq only, as the reference.
// One thread per point; u, v, w are fused in one launch so the three fields are read from HBM once (neighbours hit L1/L2).
//
// Reference behaviour restated (never copied):
//   Advec_4 advec_u/v/w/s<TF, dim3>, calc_cfl    src/advec_4.cxx:50-487
//   Diff_4  diff_c / diff_w<TF, dim3>            src/diff_4.cxx:40-175
//   weights                                      include/finite_difference.h:58-93
// Every direction is a separate `-=` / `+=` statement in the reference; the partial sums below are rounded in the same
// places.
#pragma once
#include "common.cuh"

namespace mhh {

template <typename TF>
struct W4
{
    static constexpr TF ci0 = TF(-1. / 16.), ci1 = TF(9. / 16.), ci2 = TF(9. / 16.), ci3 = TF(-1. / 16.);
    static constexpr TF bi0 = TF(5. / 16.), bi1 = TF(15. / 16.), bi2 = TF(-5. / 16.), bi3 = TF(1. / 16.);
    static constexpr TF ti0 = TF(1. / 16.), ti1 = TF(-5. / 16.), ti2 = TF(15. / 16.), ti3 = TF(5. / 16.);
    static constexpr TF cg0 = TF(1. / 24.), cg1 = TF(-27. / 24.), cg2 = TF(27. / 24.), cg3 = TF(-1. / 24.);
    static constexpr TF bg0 = TF(-23. / 24.), bg1 = TF(21. / 24.), bg2 = TF(3. / 24.), bg3 = TF(-1. / 24.);
    static constexpr TF tg0 = TF(1. / 24.), tg1 = TF(-3. / 24.), tg2 = TF(-21. / 24.), tg3 = TF(23. / 24.);
    static constexpr TF cdg0 = TF(-1460. / 576.), cdg1 = TF(783. / 576.), cdg2 = TF(-54. / 576.), cdg3 = TF(1. / 576.);
};

template <typename TF> __device__ __forceinline__ TF i4c(TF a, TF b, TF c, TF d)
{ return W4<TF>::ci0 * a + W4<TF>::ci1 * b + W4<TF>::ci2 * c + W4<TF>::ci3 * d; }
template <typename TF> __device__ __forceinline__ TF i4b(TF a, TF b, TF c, TF d)
{ return W4<TF>::bi0 * a + W4<TF>::bi1 * b + W4<TF>::bi2 * c + W4<TF>::bi3 * d; }
template <typename TF> __device__ __forceinline__ TF i4t(TF a, TF b, TF c, TF d)
{ return W4<TF>::ti0 * a + W4<TF>::ti1 * b + W4<TF>::ti2 * c + W4<TF>::ti3 * d; }

// 4th-order divergence along the direction with stride sq of (velocity x q).
//   VMODE 0: velocity interpolated along the direction with stride sv, 1: face velocity used directly,
//         2: the velocity IS the interpolated q (u du/dx, w dw/dz).
//   zmode 1 / 2: first / last row of the vertical direction (one-sided interpolation of the outermost flux).
template <typename TF, int VMODE>
__device__ __forceinline__ TF o4_div(const TF* __restrict__ q, const TF* __restrict__ vel, const long long ijk,
        const long long sq, const long long sv, const int zmode)
{
    TF acc = TF(0);
#pragma unroll
    for (int m = 0; m < 4; ++m)
    {
        TF qi;
        if (zmode == 1 && m == 0) qi = i4b(q[ijk - 2 * sq], q[ijk - sq], q[ijk], q[ijk + sq]);
        else if (zmode == 2 && m == 3) qi = i4t(q[ijk - sq], q[ijk], q[ijk + sq], q[ijk + 2 * sq]);
        else qi = i4c(q[ijk + (m - 3) * sq], q[ijk + (m - 2) * sq], q[ijk + (m - 1) * sq], q[ijk + m * sq]);
        TF vi;
        if (VMODE == 2) vi = qi;
        else if (VMODE == 1) vi = vel[ijk + (m - 1) * sq];
        else { const long long o = ijk + (m - 1) * sq; vi = i4c(vel[o - 2 * sv], vel[o - sv], vel[o], vel[o + sv]); }
        const TF cg = m == 0 ? W4<TF>::cg0 : m == 1 ? W4<TF>::cg1 : m == 2 ? W4<TF>::cg2 : W4<TF>::cg3;
        const TF term = cg * (vi * qi);
        acc = (m == 0) ? term : acc + term;
    }
    return acc;
}

template <typename TF>
__device__ __forceinline__ TF o4_lap7(const TF* __restrict__ a, const long long ijk, const long long s)
{
    return W4<TF>::cdg3 * a[ijk - 3 * s] + W4<TF>::cdg2 * a[ijk - 2 * s] + W4<TF>::cdg1 * a[ijk - s] + W4<TF>::cdg0 * a[ijk]
         + W4<TF>::cdg1 * a[ijk + s] + W4<TF>::cdg2 * a[ijk + 2 * s] + W4<TF>::cdg3 * a[ijk + 3 * s];
}

// vertical div(grad): din = metric of the gradient rows (dzhi4 for cell-centred fields, dzi4 shifted by one for w)
template <typename TF>
__device__ __forceinline__ TF o4_divgrad_z(const TF* __restrict__ a, const long long ijk, const long long kk,
        const TF* __restrict__ din, const int kin, const int zmode)
{
    TF acc = TF(0);
#pragma unroll
    for (int m = 0; m < 4; ++m)
    {
        TF gr;
        if (zmode == 1 && m == 0)
            gr = W4<TF>::bg0 * a[ijk - 2 * kk] + W4<TF>::bg1 * a[ijk - kk] + W4<TF>::bg2 * a[ijk] + W4<TF>::bg3 * a[ijk + kk];
        else if (zmode == 2 && m == 3)
            gr = W4<TF>::tg0 * a[ijk - kk] + W4<TF>::tg1 * a[ijk] + W4<TF>::tg2 * a[ijk + kk] + W4<TF>::tg3 * a[ijk + 2 * kk];
        else
            gr = W4<TF>::cg0 * a[ijk + (m - 3) * kk] + W4<TF>::cg1 * a[ijk + (m - 2) * kk] + W4<TF>::cg2 * a[ijk + (m - 1) * kk] + W4<TF>::cg3 * a[ijk + m * kk];
        const TF cg = m == 0 ? W4<TF>::cg0 : m == 1 ? W4<TF>::cg1 : m == 2 ? W4<TF>::cg2 : W4<TF>::cg3;
        const TF term = cg * gr * din[kin + m - 1];
        acc = (m == 0) ? term : acc + term;
    }
    return acc;
}

template <typename TF>
struct O4Args
{
    TF* ut; TF* vt; TF* wt;
    const TF* u; const TF* v; const TF* w;
    TF visc;
    TF dxidxi_c, dyidyi_c;      // diff_c: 1./(dx*dx) formed in double, narrowed (src/diff_4.cxx:55-56)
    TF dxidxi_w, dyidyi_w;      // diff_w: 1/(dx*dx) formed in TF (src/diff_4.cxx:122-123)
};

template <typename TF, bool ADV, bool DIFF, bool DIM3>
__global__ void __launch_bounds__(256) o4_uvw_kernel(const O4Args<TF> a, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ijk = i + j * jj + k * kk;
    const TF* __restrict__ u = a.u; const TF* __restrict__ v = a.v; const TF* __restrict__ w = a.w;
    const TF dxi = g.dxi, dyi = g.dyi;
    const int zc = (k == g.kstart) ? 1 : (k == g.kend - 1 ? 2 : 0);          // cell-centred rows
    const int zw = (k == g.kstart + 1) ? 1 : (k == g.kend - 1 ? 2 : 0);      // w rows (kstart+1 .. kend-1)
    const bool wrow = k > g.kstart;
    TF ut = a.ut[ijk], vt = a.vt[ijk], wt = wrow ? a.wt[ijk] : TF(0);
    if (ADV)
    {
        ut -= o4_div<TF, 2>(u, u, ijk, 1, 1, 0) * dxi;
        if (DIM3) ut -= o4_div<TF, 0>(u, v, ijk, jj, 1, 0) * dyi;
        ut -= o4_div<TF, 0>(u, w, ijk, kk, 1, zc) * g.dzi4[k];
        vt -= o4_div<TF, 0>(v, u, ijk, 1, jj, 0) * dxi;
        if (DIM3) vt -= o4_div<TF, 2>(v, v, ijk, jj, jj, 0) * dyi;
        vt -= o4_div<TF, 0>(v, w, ijk, kk, jj, zc) * g.dzi4[k];
        if (wrow)
        {
            wt -= o4_div<TF, 0>(w, u, ijk, 1, kk, 0) * dxi;
            if (DIM3) wt -= o4_div<TF, 0>(w, v, ijk, jj, kk, 0) * dyi;
            wt -= o4_div<TF, 2>(w, w, ijk, kk, kk, zw) * g.dzhi4[k];
        }
    }
    if (DIFF)
    {
        const TF visc = a.visc;
        ut += visc * o4_lap7<TF>(u, ijk, 1) * a.dxidxi_c;
        if (DIM3) ut += visc * o4_lap7<TF>(u, ijk, jj) * a.dyidyi_c;
        ut += visc * o4_divgrad_z<TF>(u, ijk, kk, g.dzhi4, k, zc) * g.dzi4[k];
        vt += visc * o4_lap7<TF>(v, ijk, 1) * a.dxidxi_c;
        if (DIM3) vt += visc * o4_lap7<TF>(v, ijk, jj) * a.dyidyi_c;
        vt += visc * o4_divgrad_z<TF>(v, ijk, kk, g.dzhi4, k, zc) * g.dzi4[k];
        if (wrow)
        {
            wt += visc * o4_lap7<TF>(w, ijk, 1) * a.dxidxi_w;
            if (DIM3) wt += visc * o4_lap7<TF>(w, ijk, jj) * a.dyidyi_w;
            wt += visc * o4_divgrad_z<TF>(w, ijk, kk, g.dzi4, k - 1, zw) * g.dzhi4[k];
        }
    }
    a.ut[ijk] = ut; a.vt[ijk] = vt;
    if (wrow) a.wt[ijk] = wt;
}

template <typename TF>
struct O4ScalArgs
{
    TF* st;
    const TF* s; const TF* u; const TF* v; const TF* w;
    TF visc, dxidxi_c, dyidyi_c;
};

template <typename TF, bool ADV, bool DIFF, bool DIM3>
__global__ void __launch_bounds__(256) o4_s_kernel(const O4ScalArgs<TF> a, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ijk = i + j * jj + k * kk;
    const TF* __restrict__ s = a.s;
    const int zc = (k == g.kstart) ? 1 : (k == g.kend - 1 ? 2 : 0);
    TF st = a.st[ijk];
    if (ADV)
    {
        st -= o4_div<TF, 1>(s, a.u, ijk, 1, 0, 0) * g.dxi;
        if (DIM3) st -= o4_div<TF, 1>(s, a.v, ijk, jj, 0, 0) * g.dyi;
        st -= o4_div<TF, 1>(s, a.w, ijk, kk, 0, zc) * g.dzi4[k];
    }
    if (DIFF)
    {
        st += a.visc * o4_lap7<TF>(s, ijk, 1) * a.dxidxi_c;
        if (DIM3) st += a.visc * o4_lap7<TF>(s, ijk, jj) * a.dyidyi_c;
        st += a.visc * o4_divgrad_z<TF>(s, ijk, kk, g.dzhi4, k, zc) * g.dzi4[k];
    }
    a.st[ijk] = st;
}

// ---- Advec_4m (src/advec_4m.cxx:88-415): the fully conservative 4th-order scheme.  Each direction is
//   - grad4( V(-1) i2(q[-3], q[0]),  V(0) i2(q[-1], q[0]),  V(+1) i2(q[0], q[+1]),  V(+2) i2(q[0], q[+3]) ) * metric
// with V the advecting velocity at the four flux points (interp4c across the direction, or the face velocity itself for a
// scalar), i2 the two-point mean, and the three directions summed in ONE `+=`.  In the first / last cell-centred row the
// outermost vertical term is mirrored over the wall: -V(+1) i2(q[-1], q[+2]) / -V(0) i2(q[-2], q[+1]).
template <typename TF> __device__ __forceinline__ TF i4m(TF a, TF b, TF c, TF d)
{ return W4<TF>::ci0 * (a + d) + W4<TF>::ci1 * (b + c); }
template <typename TF> __device__ __forceinline__ TF g4m(TF a, TF b, TF c, TF d)
{ return -W4<TF>::cg0 * (d - a) - W4<TF>::cg1 * (c - b); }

//   VMODE 0: velocity interpolated across the direction with stride sv, 1: face velocity used directly,
//         2: self-advection along the direction (vel == q, interpolated along sq).
//   zmode 1 / 2: first / last cell-centred row of the vertical direction.
template <typename TF, int VMODE>
__device__ __forceinline__ TF o4m_div(const TF* __restrict__ q, const TF* __restrict__ vel, const long long ijk,
        const long long sq, const long long sv, const int zmode)
{
    const TF h = TF(0.5);
    TF V[4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
    {
        const long long o = ijk + (m - 1) * sq;
        if (VMODE == 2) V[m] = i4m(q[o - 2 * sq], q[o - sq], q[o], q[o + sq]);
        else if (VMODE == 1) V[m] = vel[o];
        else V[m] = i4m(vel[o - 2 * sv], vel[o - sv], vel[o], vel[o + sv]);
    }
    const TF q0 = q[ijk];
    const TF f0 = (zmode == 1) ? -V[2] * (h * (q[ijk - sq] + q[ijk + 2 * sq])) : V[0] * (h * (q[ijk - 3 * sq] + q0));
    const TF f1 = V[1] * (h * (q[ijk - sq] + q0));
    const TF f2 = V[2] * (h * (q0 + q[ijk + sq]));
    const TF f3 = (zmode == 2) ? -V[1] * (h * (q[ijk - 2 * sq] + q[ijk + sq])) : V[3] * (h * (q0 + q[ijk + 3 * sq]));
    return g4m(f0, f1, f2, f3);
}

template <typename TF>
__global__ void __launch_bounds__(256) o4m_uvw_kernel(const O4Args<TF> a, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ijk = i + j * jj + k * kk;
    const TF* __restrict__ u = a.u; const TF* __restrict__ v = a.v; const TF* __restrict__ w = a.w;
    const TF dxi = g.dxi, dyi = g.dyi;
    const int zc = (k == g.kstart) ? 1 : (k == g.kend - 1 ? 2 : 0);
    a.ut[ijk] += - o4m_div<TF, 2>(u, u, ijk, 1, 1, 0) * dxi - o4m_div<TF, 0>(u, v, ijk, jj, 1, 0) * dyi
                 - o4m_div<TF, 0>(u, w, ijk, kk, 1, zc) * g.dzi4[k];
    a.vt[ijk] += - o4m_div<TF, 0>(v, u, ijk, 1, jj, 0) * dxi - o4m_div<TF, 2>(v, v, ijk, jj, jj, 0) * dyi
                 - o4m_div<TF, 0>(v, w, ijk, kk, jj, zc) * g.dzi4[k];
    if (k > g.kstart)       // w rows kstart+1 .. kend-1: no wall variants (w and its mirrored ghost levels vanish at the walls)
        a.wt[ijk] += - o4m_div<TF, 0>(w, u, ijk, 1, kk, 0) * dxi - o4m_div<TF, 0>(w, v, ijk, jj, kk, 0) * dyi
                     - o4m_div<TF, 2>(w, w, ijk, kk, kk, 0) * g.dzhi4[k];
}

template <typename TF>
__global__ void __launch_bounds__(256) o4m_s_kernel(const O4ScalArgs<TF> a, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ijk = i + j * jj + k * kk;
    const int zc = (k == g.kstart) ? 1 : (k == g.kend - 1 ? 2 : 0);
    a.st[ijk] += - o4m_div<TF, 1>(a.s, a.u, ijk, 1, 0, 0) * g.dxi - o4m_div<TF, 1>(a.s, a.v, ijk, jj, 0, 0) * g.dyi
                 - o4m_div<TF, 1>(a.s, a.w, ijk, kk, 0, zc) * g.dzi4[k];
}

// 4th-order vertical ghost cells of a cell-centred field (src/boundary.cxx:776-848): two levels at either wall from the
// wall value (Dirichlet) or the wall gradient (Neumann / flux).  gb / gt = grad4 of the z levels around the wall.
template <typename TF>
__global__ void ghost_cells_4th_kernel(TF* __restrict__ a, const GridDev<TF> g,
        const int bcbot, const TF* __restrict__ bot, const TF* __restrict__ gradbot,
        const int bctop, const TF* __restrict__ top, const TF* __restrict__ gradtop, const TF gb, const TF gt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.icells || j >= g.jcells) return;
    const long long ij = i + (long long)j * g.icells, kk = g.ijcells;
    const long long kb = ij + g.kstart * kk, kt = ij + (g.kend - 1) * kk;
    if (bcbot == 0)
    {
        a[kb - kk] = TF(8. / 3.) * bot[ij] - TF(2.) * a[kb] + TF(1. / 3.) * a[kb + kk];
        a[kb - 2 * kk] = TF(8.) * bot[ij] - TF(9.) * a[kb] + TF(2.) * a[kb + kk];
    }
    else if (bcbot == 1)
    {
        a[kb - kk] = TF(-1.) * gb * gradbot[ij] + a[kb];
        a[kb - 2 * kk] = TF(-3.) * gb * gradbot[ij] + a[kb + kk];
    }
    if (bctop == 0)
    {
        a[kt + kk] = TF(8. / 3.) * top[ij] - TF(2.) * a[kt] + TF(1. / 3.) * a[kt - kk];
        a[kt + 2 * kk] = TF(8.) * top[ij] - TF(9.) * a[kt] + TF(2.) * a[kt - kk];
    }
    else if (bctop == 1)
    {
        a[kt + kk] = TF(1.) * gt * gradtop[ij] + a[kt];
        a[kt + 2 * kk] = TF(3.) * gt * gradtop[ij] + a[kt - kk];
    }
}

// no-penetration ghost cells of w (src/boundary.cxx:850-922): conservation type (two mirrored levels) or normal type
template <typename TF>
__global__ void ghost_cells_w_4th_kernel(TF* __restrict__ w, const GridDev<TF> g, const int conservation)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.icells || j >= g.jcells) return;
    const long long ij = i + (long long)j * g.icells, kk = g.ijcells;
    const long long kb = ij + g.kstart * kk, kt = ij + g.kend * kk;
    if (conservation)
    {
        w[kb - kk] = -w[kb + kk]; w[kb - 2 * kk] = -w[kb + 2 * kk];
        w[kt + kk] = -w[kt - kk]; w[kt + 2 * kk] = -w[kt - 2 * kk];
    }
    else
    {
        w[kb - kk] = TF(-6.) * w[kb + kk] + TF(4.) * w[kb + 2 * kk] - w[kb + 3 * kk];
        w[kt + kk] = TF(-6.) * w[kt - kk] + TF(4.) * w[kt - 2 * kk] - w[kt - 3 * kk];
    }
}

// Advec_4 calc_cfl (src/advec_4.cxx:50-86): interp4c(a,b,c,d) = ci0*(a+d) + ci1*(b+c)
// Advec_4m calc_cfl (src/advec_4m.cxx:51-88) sums the four weighted terms one by one instead: SUM4 = true.
template <typename TF, bool SUM4>
__global__ void __launch_bounds__(256) o4_cfl_kernel(const TF* __restrict__ u, const TF* __restrict__ v, const TF* __restrict__ w,
        const GridDev<TF> g, double* __restrict__ out)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    TF val = TF(0);
    if (i < g.iend && j < g.jend)
    {
        const long long jj = g.icells, kk = g.ijcells;
        const long long ijk = i + j * jj + k * kk;
        auto c4 = [](TF a, TF b, TF c, TF d) { return SUM4 ? W4<TF>::ci0 * a + W4<TF>::ci1 * b + W4<TF>::ci2 * c + W4<TF>::ci3 * d
                                                           : W4<TF>::ci0 * (a + d) + W4<TF>::ci1 * (b + c); };
        val = absf(c4(u[ijk - 1], u[ijk], u[ijk + 1], u[ijk + 2])) * g.dxi
            + absf(c4(v[ijk - jj], v[ijk], v[ijk + jj], v[ijk + 2 * jj])) * g.dyi
            + absf(c4(w[ijk - kk], w[ijk], w[ijk + kk], w[ijk + 2 * kk])) * g.dzi[k];
    }
    block_max_to_global<TF>(val, out);
}

} // namespace mhh
