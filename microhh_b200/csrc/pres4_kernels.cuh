// mhhb200 -- Pres_4: the 4th-order Poisson solver (DNS configurations).
//
// Same spectral pipeline as Pres_2 (x/y transforms of fft_warp.cuh / poisson_kernels.cuh, complex [k][l][m] workspace),
// with the 4th-order divergence / gradient (cg weights), the 4-term cosine modified wavenumbers and, per horizontal
// mode, a 7-band system of kmax+4 rows (two boundary rows at either end) instead of the tridiagonal one.
// The band matrix depends only on (grid, mode): its LU factors (no pivoting) are tabulated once per context
// (`hdma_setup_kernel`, 7 x (kmax+4) reals per mode) so that a solve is two substitution sweeps without divisions chains
// of the factorisation; real and imaginary parts share the factors.
//
// Reference behaviour restated (never copied):
//   Pres_4::set_values / input / solve / hdma / output / calc_divergence   src/pres_4.cxx:178-767
#pragma once
#include "common.cuh"
#include "poisson_kernels.cuh"
#include "order4_kernels.cuh"

namespace mhh {

// wt is mirrored over the walls before the divergence (src/pres_4.cxx:290-303)
template <typename TF>
__global__ void pres4_wtbc_kernel(TF* __restrict__ wt, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.iend || j >= g.jend) return;
    const long long ij = i + (long long)j * g.icells, kk = g.ijcells;
    wt[ij + (g.kstart - 1) * kk] = -wt[ij + (g.kstart + 1) * kk];
    wt[ij + (g.kend + 1) * kk] = -wt[ij + (g.kend - 1) * kk];
}

// Pres_4::input (src/pres_4.cxx:305-316): 4th-order divergence of (ut + u/dt, ...) into compact rows of pitch `pitch`
template <typename TF, bool DIM3>
__global__ void __launch_bounds__(256) pres4_in_kernel(const TF* __restrict__ u, const TF* __restrict__ v, const TF* __restrict__ w,
        const TF* __restrict__ ut, const TF* __restrict__ vt, const TF* __restrict__ wt, TF* __restrict__ rhs,
        const long long pitch, const TF dti, const GridDev<TF> g)
{
    const int ii = blockIdx.x * blockDim.x + threadIdx.x;
    const int jq = blockIdx.y * blockDim.y + threadIdx.y;
    const int kq = blockIdx.z;
    if (ii >= g.imax || jq >= g.jmax) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ijk = (ii + g.istart) + (jq + g.jstart) * jj + (kq + g.kstart) * kk;
    auto T = [&](const TF* __restrict__ a, const TF* __restrict__ at, const long long o) { return at[o] + a[o] * dti; };
    TF p = (W4<TF>::cg0 * T(u, ut, ijk - 1) + W4<TF>::cg1 * T(u, ut, ijk) + W4<TF>::cg2 * T(u, ut, ijk + 1) + W4<TF>::cg3 * T(u, ut, ijk + 2)) * g.dxi;
    if (DIM3)
        p += (W4<TF>::cg0 * T(v, vt, ijk - jj) + W4<TF>::cg1 * T(v, vt, ijk) + W4<TF>::cg2 * T(v, vt, ijk + jj) + W4<TF>::cg3 * T(v, vt, ijk + 2 * jj)) * g.dyi;
    p += (W4<TF>::cg0 * T(w, wt, ijk - kk) + W4<TF>::cg1 * T(w, wt, ijk) + W4<TF>::cg2 * T(w, wt, ijk + kk) + W4<TF>::cg3 * T(w, wt, ijk + 2 * kk)) * g.dzi4[kq + g.kstart];
    rhs[((long long)kq * g.jmax + jq) * pitch + ii] = p;
}

template <typename TF>
struct HdmaCoef
{
    const TF* m;        // 7 x kmax band coefficients of the interior rows (set_values)
    const TF* bmati; const TF* bmatj;
};

// LU factorisation of the 7-band matrix of every mode (src/pres_4.cxx:573-667), one thread per mode column.
// lu: [7][kmax+4][ncol]
template <typename TF>
__global__ void hdma_setup_kernel(TF* __restrict__ lu, const HdmaCoef<TF> cf, const int nm, const int jtot, const int kmax, const int m_off)
{
    // nm: x-modes owned by this rank (all of them on a single GPU), the first being global mode m_off
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long ncol = (long long)nm * jtot;
    if (col >= ncol) return;
    const int l = (int)(col / nm), mx = (int)(col % nm) + m_off;
    const bool mode00 = (l == 0 && mx == 0);
    const int nr = kmax + 4;
    const long long rs = ncol;                          // row stride
    const long long bs = (long long)nr * ncol;          // band stride
    auto A = [&](int n, int r) -> TF& { return lu[n * bs + r * rs + col]; };
    // assemble the rows (src/pres_4.cxx:362-470)
    for (int r = 0; r < nr; ++r)
    {
        TF b[7] = {0, 0, 0, 0, 0, 0, 0};
        if (r == 0) { b[3] = TF(1.); b[6] = TF(-1.); }
        else if (r == 1) { b[3] = TF(1.); b[4] = TF(-1.); }
        else if (r < kmax + 2)
        {
            const int k = r - 2;
            for (int n = 0; n < 7; ++n) b[n] = cf.m[n * kmax + k];
            b[3] = b[3] + cf.bmati[mx] + cf.bmatj[l];
        }
        else if (r == kmax + 2)
        {
            if (mode00) { b[0] = TF(0.); b[1] = TF(-1 / 3.); b[2] = TF(2.); b[3] = TF(1.); }
            else { b[2] = TF(-1.); b[3] = TF(1.); }
        }
        else
        {
            if (mode00) { b[0] = TF(-2.); b[1] = TF(9.); b[2] = TF(0.); b[3] = TF(1.); }
            else { b[0] = TF(-1.); b[3] = TF(1.); }
        }
        for (int n = 0; n < 7; ++n) A(n, r) = b[n];
    }
    const TF one = TF(1.);
    // LU without pivoting, the reference's statement order
    A(0, 0) = one; A(1, 0) = one; A(2, 0) = one / A(3, 0); A(3, 0) = one;
    A(4, 0) = A(4, 0) * A(2, 0); A(5, 0) = A(5, 0) * A(2, 0); A(6, 0) = A(6, 0) * A(2, 0);
    {
        const int k = 1;
        A(0, k) = one; A(1, k) = one;
        A(2, k) = A(2, k) / A(3, k - 1);
        A(3, k) = A(3, k) - A(2, k) * A(4, k - 1);
        A(4, k) = A(4, k) - A(2, k) * A(5, k - 1);
        A(5, k) = A(5, k) - A(2, k) * A(6, k - 1);
    }
    {
        const int k = 2;
        A(0, k) = one;
        A(1, k) = A(1, k) / A(3, k - 2);
        A(2, k) = (A(2, k) - A(1, k) * A(4, k - 2)) / A(3, k - 1);
        A(3, k) = A(3, k) - A(2, k) * A(4, k - 1) - A(1, k) * A(5, k - 2);
        A(4, k) = A(4, k) - A(2, k) * A(5, k - 1) - A(1, k) * A(6, k - 2);
        A(5, k) = A(5, k) - A(2, k) * A(6, k - 1);
    }
    for (int k = 3; k < kmax + 4; ++k)
    {
        A(0, k) = A(0, k) / A(3, k - 3);
        A(1, k) = (A(1, k) - A(0, k) * A(4, k - 3)) / A(3, k - 2);
        A(2, k) = (A(2, k) - A(1, k) * A(4, k - 2) - A(0, k) * A(5, k - 3)) / A(3, k - 1);
        A(3, k) = A(3, k) - A(2, k) * A(4, k - 1) - A(1, k) * A(5, k - 2) - A(0, k) * A(6, k - 3);
        if (k < kmax + 3) A(4, k) = A(4, k) - A(2, k) * A(5, k - 1) - A(1, k) * A(6, k - 2);
        if (k < kmax + 2) A(5, k) = A(5, k) - A(2, k) * A(6, k - 1);
        if (k == kmax + 1) A(6, k) = one;
        if (k == kmax + 2) { A(5, k) = one; A(6, k) = one; }
        if (k == kmax + 3) { A(4, k) = one; A(5, k) = one; A(6, k) = one; }
    }
}

// Substitution sweeps (src/pres_4.cxx:669-729) on the complex right-hand side of every mode; rows 2..kmax+1 of the
// system are the levels of the spectral workspace, the four boundary rows (zero right-hand side) live in registers.
template <typename TF>
__global__ void __launch_bounds__(128) hdma_solve_kernel(TF* __restrict__ spec, const TF* __restrict__ lu, const SpecLayout lay, const int nm, const int jtot, const int kmax)
{
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long ncol_ = (long long)nm * jtot;
    if (col >= ncol_) return;
    const int nr = kmax + 4;
    const long long bs = (long long)nr * ncol_;
    // the y side of the workspace (single GPU: [k][l][m]; slabs: one block per source rank), a column advances by ks per level
    cplx<TF>* S = reinterpret_cast<cplx<TF>*>(spec) + lay.yidx(0, (int)(col / nm), (int)(col % nm));
    const long long ncol = lay.ykstride();
    auto Lf = [&](int n, int r) -> TF { return lu[n * bs + (long long)r * ncol_ + col]; };
    // L y = p: y0 = y1 = 0 (zero right-hand sides), then the interior rows
    cplx<TF> y1 = {0, 0}, y2 = {0, 0}, y3 = {0, 0};                  // y[r-1], y[r-2], y[r-3]
#pragma unroll 2
    for (int r = 2; r < kmax + 2; ++r)
    {
        const TF l3 = Lf(2, r), l2 = Lf(1, r), l1 = Lf(0, r);
        const cplx<TF> p = S[(long long)(r - 2) * ncol];
        const cplx<TF> y = {p.x - y1.x * l3 - y2.x * l2 - y3.x * l1, p.y - y1.y * l3 - y2.y * l2 - y3.y * l1};
        S[(long long)(r - 2) * ncol] = y;
        y3 = y2; y2 = y1; y1 = y;
    }
    cplx<TF> ya, yb;                                                 // rows kmax+2, kmax+3
    {
        int r = kmax + 2;
        ya = {-y1.x * Lf(2, r) - y2.x * Lf(1, r) - y3.x * Lf(0, r), -y1.y * Lf(2, r) - y2.y * Lf(1, r) - y3.y * Lf(0, r)};
        r = kmax + 3;
        yb = {-ya.x * Lf(2, r) - y1.x * Lf(1, r) - y2.x * Lf(0, r), -ya.y * Lf(2, r) - y1.y * Lf(1, r) - y2.y * Lf(0, r)};
    }
    // U x = y
    cplx<TF> x1, x2, x3;                                             // x[r+1], x[r+2], x[r+3]
    {
        const int r = kmax + 3;
        const TF d3 = Lf(3, r), d2 = Lf(3, r - 1), d1 = Lf(3, r - 2);
        const cplx<TF> xb = {yb.x / d3, yb.y / d3};
        const TF u5 = Lf(4, r - 1);
        const cplx<TF> xa = {(ya.x - xb.x * u5) / d2, (ya.y - xb.y * u5) / d2};
        const TF v5 = Lf(4, r - 2), v6 = Lf(5, r - 2);
        const cplx<TF> yl = S[(long long)(kmax - 1) * ncol];          // row kmax+1 = level kmax-1
        const cplx<TF> xl = {(yl.x - xa.x * v5 - xb.x * v6) / d1, (yl.y - xa.y * v5 - xb.y * v6) / d1};
        S[(long long)(kmax - 1) * ncol] = xl;
        x1 = xl; x2 = xa; x3 = xb;
    }
#pragma unroll 2
    for (int r = kmax; r >= 2; --r)
    {
        const TF u5 = Lf(4, r), u6 = Lf(5, r), u7 = Lf(6, r), d = Lf(3, r);
        const cplx<TF> y = S[(long long)(r - 2) * ncol];
        const cplx<TF> x = {(y.x - x1.x * u5 - x2.x * u6 - x3.x * u7) / d, (y.y - x1.y * u5 - x2.y * u6 - x3.y * u7) / d};
        S[(long long)(r - 2) * ncol] = x;
        x3 = x2; x2 = x1; x1 = x;
    }
}

// zero-gradient ghost levels of p, two deep at either wall, over the whole ghosted plane (src/pres_4.cxx:507-528)
template <typename TF>
__global__ void pres4_ghost_kernel(TF* __restrict__ p, const GridDev<TF> g)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.icells || j >= g.jcells) return;
    const long long ij = i + (long long)j * g.icells, kk = g.ijcells;
    const int ks = g.kstart, ke = g.kend;
    p[ij + (ks - 1) * kk] = p[ij + ks * kk];
    p[ij + (ks - 2) * kk] = p[ij + (ks + 1) * kk];
    p[ij + ke * kk] = p[ij + (ke - 1) * kk];
    p[ij + (ke + 1) * kk] = p[ij + (ke - 2) * kk];
}

// Pres_4::output (src/pres_4.cxx:531-571)
template <typename TF, bool DIM3>
__global__ void __launch_bounds__(256) pres4_out_kernel(TF* __restrict__ ut, TF* __restrict__ vt, TF* __restrict__ wt,
        const TF* __restrict__ p, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ijk = i + j * jj + k * kk;
    ut[ijk] -= (W4<TF>::cg0 * p[ijk - 2] + W4<TF>::cg1 * p[ijk - 1] + W4<TF>::cg2 * p[ijk] + W4<TF>::cg3 * p[ijk + 1]) * g.dxi;
    if (DIM3)
        vt[ijk] -= (W4<TF>::cg0 * p[ijk - 2 * jj] + W4<TF>::cg1 * p[ijk - jj] + W4<TF>::cg2 * p[ijk] + W4<TF>::cg3 * p[ijk + jj]) * g.dyi;
    if (k > g.kstart)
        wt[ijk] -= (W4<TF>::cg0 * p[ijk - 2 * kk] + W4<TF>::cg1 * p[ijk - kk] + W4<TF>::cg2 * p[ijk] + W4<TF>::cg3 * p[ijk + kk]) * g.dzhi4[k];
}

// Pres_4::calc_divergence (src/pres_4.cxx:732-767)
template <typename TF>
__global__ void __launch_bounds__(256) pres4_div_kernel(const TF* __restrict__ u, const TF* __restrict__ v, const TF* __restrict__ w,
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
        val = absf((W4<TF>::cg0 * u[ijk - 1] + W4<TF>::cg1 * u[ijk] + W4<TF>::cg2 * u[ijk + 1] + W4<TF>::cg3 * u[ijk + 2]) * g.dxi
                 + (W4<TF>::cg0 * v[ijk - jj] + W4<TF>::cg1 * v[ijk] + W4<TF>::cg2 * v[ijk + jj] + W4<TF>::cg3 * v[ijk + 2 * jj]) * g.dyi
                 + (W4<TF>::cg0 * w[ijk - kk] + W4<TF>::cg1 * w[ijk] + W4<TF>::cg2 * w[ijk + kk] + W4<TF>::cg3 * w[ijk + 2 * kk]) * g.dzi4[k]);
    }
    block_max_to_global<TF>(val, out);
}

} // namespace mhh
