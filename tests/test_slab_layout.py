"""
Host-side tests of the y-slab decomposition (no GPU): the spectral workspace layout exported by the
C ABI (mhh_slab_layout / mhh_slab_xindex / mhh_slab_yindex) and a world_size-2 gloo replay of the
distributed Poisson solve (x transform on the slab -> all-to-all -> y transform + tridiagonal solve
on the owned x-modes -> all-to-all back -> inverse x transform) that must reproduce the oracle's
single-domain Pres_2 solve.  The CUDA kernels use exactly these index functions (SpecLayout).
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from microhh_b200 import capi


def slab_info(lib, itot, jtot, ktot, P, r):
    s = capi.SlabInfo()
    assert lib.mhh_slab_layout(itot, jtot, ktot, P, r, C.byref(s)) == 0
    return s


@pytest.mark.parametrize("shape,P", [((16, 8, 6), 2), ((32, 12, 5), 4), ((20, 9, 4), 3), ((16, 8, 4), 1), ((1024, 2048, 2), 8)])
def test_layout_is_a_bijection_and_blocks_are_messages(shape, P):
    lib = capi.load()
    itot, jtot, ktot = shape
    infos = [slab_info(lib, itot, jtot, ktot, P, r) for r in range(P)]
    nm = itot//2 + 1
    assert sum(s.mcl for s in infos) == nm
    assert [s.m_off for s in infos] == list(np.cumsum([0] + [s.mcl for s in infos[:-1]]))
    assert max(s.mcl for s in infos) - min(s.mcl for s in infos) <= 1          # even deal
    if itot > 64:
        return
    for r in range(P):
        s = infos[r]
        seen = np.zeros(s.xside_elems, bool)
        for row in range(s.rows):
            for m in range(nm):
                e = lib.mhh_slab_xindex(itot, jtot, ktot, P, r, row, m)
                assert 0 <= e < s.xside_elems and not seen[e]
                seen[e] = True
                # block d of the x side == message to rank d == block r of rank d's y side, same offset inside
                d = next(q for q in range(P) if infos[q].m_off <= m < infos[q].m_off + infos[q].mcl)
                k, jl = divmod(row, s.jmax)
                ey = lib.mhh_slab_yindex(itot, jtot, ktot, P, d, k, r*s.jmax + jl, m - infos[d].m_off)
                assert e - infos[d].m_off*s.rows == ey - r*infos[d].mcl*s.rows
        assert seen.all()
        seen_y = np.zeros(s.yside_elems, bool)
        for k in range(ktot):
            for j in range(jtot):
                for ml in range(s.mcl):
                    e = lib.mhh_slab_yindex(itot, jtot, ktot, P, r, k, j, ml)
                    assert 0 <= e < s.yside_elems and not seen_y[e]
                    seen_y[e] = True
        assert seen_y.all()
    assert lib.mhh_slab_xindex(itot, jtot, ktot, P, 0, infos[0].rows, 0) == -1
    assert lib.mhh_slab_layout(itot, jtot + 1, ktot, 2, 0, C.byref(capi.SlabInfo())) != 0 or (jtot + 1) % 2 == 0


def _pres4_solve_modes(pres, g, S, m_off, mcl):
    """Pres_4::solve's 7-band system (src/pres_4.cxx:346-470, oracle.Pres4.solve) for the x-modes [m_off, m_off + mcl) of a
    real array S of shape (kmax, jtot, mcl): the matrix rows, the boundary rows and hdma, exactly as on a single domain."""
    kmax, jtot = g.kmax, g.jtot
    shp = (kmax + 4, jtot, mcl)
    M = [np.zeros(shp) for _ in range(7)]
    pt = np.zeros(shp)
    M[3][0] = 1.; M[6][0] = -1.
    M[3][1] = 1.; M[4][1] = -1.
    for n in range(7):
        M[n][2:kmax+2] = pres.m[n][:, None, None]
    M[3][2:kmax+2] = M[3][2:kmax+2] + pres.bmati[None, None, m_off:m_off+mcl] + pres.bmatj[None, :, None]
    pt[2:kmax+2] = S
    M[2][kmax+2] = -1.; M[3][kmax+2] = 1.
    M[0][kmax+3] = -1.; M[3][kmax+3] = 1.
    if m_off == 0:          # mode (0, 0) lives on the rank that owns x-mode 0
        M[0][kmax+2, 0, 0] = 0.; M[1][kmax+2, 0, 0] = -1/3.; M[2][kmax+2, 0, 0] = 2.; M[3][kmax+2, 0, 0] = 1.
        M[0][kmax+3, 0, 0] = -2.; M[1][kmax+3, 0, 0] = 9.; M[2][kmax+3, 0, 0] = 0.; M[3][kmax+3, 0, 0] = 1.
    pres.hdma(*M, pt)
    return pt[2:kmax+2]


def _slab_worker(rank, world, port, shape, out_q, order=2):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = capi.load()
        itot, jtot, ktot = shape
        zst = np.cumsum(np.linspace(0.7, 1.3, ktot))*60./np.linspace(0.7, 1.3, ktot).sum() - 0.3
        if order == 4:
            g = O.Grid(itot, jtot, ktot, 100., 80., 60., 3, 3, 3, np.float64, z=zst, order=4)
            pres = O.Pres4(g)
        else:
            g = O.Grid(itot, jtot, ktot, 100., 80., 60., 3, 3, 1, np.float64, z=zst)
            rho = np.exp(-g.z/8000.); rhoh = np.exp(-g.zh/8000.)
            pres = O.Pres2(g, rho, rhoh)
        rhs = np.random.default_rng(5).standard_normal((ktot, jtot, itot))
        rhs -= rhs.mean()
        infos = [slab_info(lib, itot, jtot, ktot, world, r) for r in range(world)]
        me = infos[rank]; nm = me.nm; jmax = me.jmax; rows = me.rows

        # ---- x side: transform my rows, write them through the layout --------------------------
        loc = rhs[:, rank*jmax:(rank+1)*jmax, :]
        X = np.fft.rfft(loc, axis=2)                                      # (k, jl, m)
        xbuf = np.zeros(me.xside_elems, np.complex128)
        xi = np.empty((ktot, jmax, nm), np.int64)
        for k in range(ktot):
            for jl in range(jmax):
                for m in range(nm):
                    xi[k, jl, m] = lib.mhh_slab_xindex(itot, jtot, ktot, world, rank, k*jmax + jl, m)
        xbuf[xi] = X

        def exchange(src, forward):
            ysz = me.mcl*rows
            dst = np.zeros(me.yside_elems if forward else me.xside_elems, np.complex128)
            reqs = []; recv = {}
            for peer in range(world):
                xs = slice(infos[peer].m_off*rows, (infos[peer].m_off + infos[peer].mcl)*rows)   # my x block for `peer`
                ys = slice(peer*ysz, (peer+1)*ysz)                                               # my y block from `peer`
                s_out = src[xs] if forward else src[ys]
                d_sl = ys if forward else xs
                if peer == rank:
                    dst[d_sl] = s_out
                    continue
                t_out = torch.from_numpy(np.ascontiguousarray(s_out).view(np.float64).copy())
                t_in = torch.empty(2*(d_sl.stop - d_sl.start), dtype=torch.float64)
                recv[peer] = (d_sl, t_in)
                reqs.append(dist.isend(t_out, peer)); reqs.append(dist.irecv(t_in, peer))
            for q in reqs:
                q.wait()
            for peer, (d_sl, t_in) in recv.items():
                dst[d_sl] = t_in.numpy().view(np.complex128)
            return dst

        ybuf = exchange(xbuf, True)
        # ---- y side: gather (k, j, ml) through the layout ---------------------------------------
        yi = np.empty((ktot, jtot, me.mcl), np.int64)
        for k in range(ktot):
            for j in range(jtot):
                for ml in range(me.mcl):
                    yi[k, j, ml] = lib.mhh_slab_yindex(itot, jtot, ktot, world, rank, k, j, ml)
        S = np.fft.fft(ybuf[yi], axis=1)                                  # (k, l, ml)
        if order == 4:
            # Pres_4: the 7-band solve on my modes, real and imaginary parts separately (the matrix is real)
            re = _pres4_solve_modes(pres, g, np.ascontiguousarray(S.real), me.m_off, me.mcl)
            im = _pres4_solve_modes(pres, g, np.ascontiguousarray(S.imag), me.m_off, me.mcl)
            ybuf[yi] = np.fft.ifft(re + 1j*im, axis=1)
            xbuf = exchange(ybuf, False)
            p_loc = np.fft.irfft(xbuf[xi], n=itot, axis=2)
            p = g.field()
            pres.solve(rhs.copy(), p)
            ref = p[g.kstart:g.kend, g.jstart:g.jend, g.istart:g.iend][:, rank*jmax:(rank+1)*jmax, :]
            out_q.put((rank, float(np.sqrt(((p_loc - ref)**2).sum()/(ref**2).sum()))))
            return
        # Pres_2::solve matrix (src/pres_2.cxx:292-324) for my modes, complex right-hand side
        kg = g.kgc
        dz = g.dz[kg:kg+ktot][:, None, None]; rr = rho[kg:kg+ktot][:, None, None]
        lam = pres.bmatj[None, :, None] + pres.bmati[None, None, me.m_off:me.m_off + me.mcl]
        b = dz*dz*rr*lam - (pres.a + pres.c)[:, None, None]
        b[0] += pres.a[0]
        top = np.full((jtot, me.mcl), pres.c[ktot-1])
        if me.m_off == 0:
            top[0, 0] = -pres.c[ktot-1]
        b[ktot-1] += top
        S = dz*dz*S
        re, im = np.ascontiguousarray(S.real), np.ascontiguousarray(S.imag)
        pres.tdma(re, b.copy()); pres.tdma(im, b.copy())
        S = np.fft.ifft(re + 1j*im, axis=1)                               # includes 1/jtot
        ybuf[yi] = S
        xbuf = exchange(ybuf, False)
        p_loc = np.fft.irfft(xbuf[xi], n=itot, axis=2)                    # includes 1/itot

        p = g.field()
        pres.solve(rhs.copy(), p)
        ref = p[g.kstart:g.kend, g.jstart:g.jend, g.istart:g.iend][:, rank*jmax:(rank+1)*jmax, :]
        err = np.sqrt(((p_loc - ref)**2).sum()/(ref**2).sum())
        out_q.put((rank, float(err)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape,order", [((16, 8, 6), 2), ((20, 12, 5), 2), ((16, 8, 8), 4), ((20, 12, 6), 4)])
def test_gloo_world2_slab_solve_matches_single_domain(shape, order):
    """order 2: Pres_2 (tridiagonal); order 4: Pres_4 (7-band hdma on the owner's modes, mode (0, 0) rows on rank 0)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 7*order
    procs = [ctx.Process(target=_slab_worker, args=(r, 2, port, shape, q, order)) for r in range(2)]
    for p in procs:
        p.start()
    res = []
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    while not q.empty():
        res.append(q.get())
    assert len(res) == 2
    for rank, err in res:
        assert err <= 1e-12, (rank, err)


def _halo_worker(rank, world, port, out_q):
    """Replays the slab ghost-row exchange (east-west local copy, then the first/last jgc interior rows over the full ghosted
    width to the south/north neighbour) with gloo and compares with the single-domain cyclic fill."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        itot, jtot, ktot, gc = 12, 8*world, 5, 3
        g = O.Grid(itot, jtot, ktot, 1., 1., 1., gc, gc, 1, np.float64)
        a = np.random.default_rng(3).standard_normal((g.kcells, g.jcells, g.icells))
        ref = a.copy()
        O.boundary_cyclic(g, ref)
        jmax = jtot//world
        gl = O.Grid(itot, jmax, ktot, 1., 1., 1., gc, gc, 1, np.float64)          # the slab as a local grid
        loc = np.ascontiguousarray(a[:, rank*jmax:rank*jmax + gl.jcells, :])     # ghost rows hold stale data
        O.boundary_cyclic(gl, loc, O.EDGE_EW)                                    # east-west: local periodic copy
        south, north = (rank - 1) % world, (rank + 1) % world
        send_s = torch.from_numpy(np.ascontiguousarray(loc[:, gl.jstart:gl.jstart+gc, :]))      # -> south's north ghosts
        send_n = torch.from_numpy(np.ascontiguousarray(loc[:, gl.jend-gc:gl.jend, :]))          # -> north's south ghosts
        recv_n = torch.empty_like(send_s); recv_s = torch.empty_like(send_n)
        reqs = [dist.isend(send_s, south, tag=1), dist.isend(send_n, north, tag=2),
                dist.irecv(recv_n, north, tag=1), dist.irecv(recv_s, south, tag=2)]
        for q in reqs:
            q.wait()
        loc[:, gl.jend:gl.jend+gc, :] = recv_n.numpy()
        loc[:, 0:gc, :] = recv_s.numpy()
        want = ref[:, rank*jmax:rank*jmax + gl.jcells, :]
        out_q.put((rank, bool(np.array_equal(loc, want))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_slab_halo_exchange_equals_global_cyclic_fill(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = [q.get() for _ in range(world)]
    assert all(ok for _, ok in res), res


@pytest.mark.parametrize("shape,P", [((16, 8, 3), 2), ((40, 12, 2), 4), ((36, 9, 2), 3)])
def test_panel_tiled_x_side_is_injective_and_panel_contiguous(shape, P):
    """The 8-mode-panel x-side layout of the fused peer transposes: injective into the padded buffer, the 8 modes of a panel
    contiguous, consecutive local rows of a panel 8 elements apart (what makes the inverse y transform's remote stores long
    contiguous runs), and every rank agrees on where block d starts."""
    lib = capi.load()
    itot, jtot, ktot = shape
    nm = itot//2 + 1
    infos = [slab_info(lib, itot, jtot, ktot, P, r) for r in range(P)]
    for r in range(P):
        s = infos[r]
        tot = C.c_longlong()
        idx = np.empty((s.rows, nm), np.int64)
        for row in range(s.rows):
            for m in range(nm):
                idx[row, m] = lib.mhh_slab_xindex_tiled(itot, jtot, ktot, P, r, row, m, C.byref(tot))
        assert idx.min() >= 0 and idx.max() < tot.value and len(np.unique(idx)) == idx.size
        assert tot.value >= nm*s.rows and tot.value <= (nm + 8*P)*s.rows              # padding: < one panel per block
        for d in range(P):
            m0, cnt = infos[d].m_off, infos[d].mcl
            for ml in range(cnt - 1):
                if (ml & 7) != 7:
                    assert np.all(idx[:, m0 + ml + 1] - idx[:, m0 + ml] == 1)          # inside a panel: contiguous modes
            jmax = s.jmax
            for k in range(ktot):
                rows = np.arange(k*jmax, (k + 1)*jmax)
                assert np.all(np.diff(idx[rows, m0]) == 8)                              # consecutive rows: next 128-byte line
        if r > 0:
            first = [lib.mhh_slab_xindex_tiled(itot, jtot, ktot, P, q, 0, infos[1].m_off, None) for q in (0, r)]
            assert first[0] == first[1]


# ---------------------------------------------------------------------------------------------------------------------
# The fused Pres_2 path: mode-major Y side, 8-row-panel X side, two-sided factorisation
# ---------------------------------------------------------------------------------------------------------------------
def slab2_info(lib, itot, jtot, ktot, P, r):
    s = capi.Slab2Info()
    assert lib.mhh_slab2_layout(itot, jtot, ktot, P, r, C.byref(s)) == 0
    return s


@pytest.mark.parametrize("shape,P", [((32, 16, 6), 2), ((64, 32, 3), 4), ((32, 8, 4), 1), ((32, 24, 2), 2)])
def test_fused_layout_is_a_bijection_with_contiguous_sequences_and_chunks(shape, P):
    lib = capi.load()
    itot, jtot, ktot = shape
    nm = itot//2 + 1
    infos = [slab2_info(lib, itot, jtot, ktot, P, r) for r in range(P)]
    assert sum(s.mcl for s in infos) == nm and all(s.ksplit == ktot//2 for s in infos)
    jmax = infos[0].jmax
    for d in range(P):
        s = infos[d]
        # Y side of rank d: every (source, mode, level, row) exactly once; rows of a sequence piece are contiguous
        seen = np.zeros(s.yside_elems, bool)
        for src in range(P):
            for ml in range(s.mcl):
                for k in range(ktot):
                    idx = np.array([lib.mhh_slab2_yindex(itot, jtot, ktot, P, d, src, ml, k, jl) for jl in range(jmax)])
                    assert np.all(np.diff(idx) == 1) and not seen[idx].any()
                    seen[idx] = True
            # block src is one contiguous message
            first = lib.mhh_slab2_yindex(itot, jtot, ktot, P, d, src, 0, 0, 0)
            assert first == src*s.mcl*ktot*jmax
        assert seen.all()
    # X side (the same on every rank): every (mode owner, level, row, mode) once; 8 consecutive rows of a mode are contiguous
    seen = np.zeros(infos[0].xside_elems, bool)
    for d in range(P):
        for k in range(ktot):
            for ml in range(infos[d].mcl):
                idx = np.array([lib.mhh_slab2_xindex(itot, jtot, ktot, P, d, k, jl, ml) for jl in range(jmax)])
                assert idx.min() >= 0 and not seen[idx].any()
                seen[idx] = True
                for j0 in range(0, jmax - 7, 8):
                    assert np.all(np.diff(idx[j0:j0+8]) == 1) and idx[j0] % 8 == 0
    assert seen.sum() == nm*ktot*jmax                      # the rest is panel padding (jmax not a multiple of 8)
    assert lib.mhh_slab2_xindex(itot, jtot, ktot, P, 0, ktot, 0, 0) == -1


def twisted_tdma(a, b, c, d, ks):
    """The two-sided factorisation of the fused kernels (poisson_fused.cuh), in numpy, vectorised over the trailing axes:
    levels < ks eliminated upwards, the others downwards, 2 x 2 interface, substitution outwards from the interface."""
    K = d.shape[0]
    T = np.empty_like(b); pp = np.empty_like(d)
    w = b[0].copy(); T[0] = 1./w; pp[0] = d[0]*T[0]
    for k in range(1, ks):
        w = b[k] - a[k]*(c[k-1]/w); T[k] = 1./w
        pp[k] = (d[k] - a[k]*pp[k-1])*T[k]
    v = b[K-1].copy(); T[K-1] = 1./v; pp[K-1] = d[K-1]*T[K-1]
    for k in range(K-2, ks-1, -1):
        v = b[k] - c[k]*(a[k+1]/v); T[k] = 1./v
        pp[k] = (d[k] - c[k]*pp[k+1])*T[k]
    alpha = c[ks-1]*T[ks-1]; beta = a[ks]*T[ks]
    x = np.empty_like(d)
    x[ks] = (pp[ks] - beta*pp[ks-1])/(1. - alpha*beta)
    x[ks-1] = pp[ks-1] - alpha*x[ks]
    for k in range(ks-2, -1, -1):
        x[k] = pp[k] - (c[k]*T[k])*x[k+1]
    for k in range(ks+1, K):
        x[k] = pp[k] - (a[k]*T[k])*x[k-1]
    return x


@pytest.mark.parametrize("K", [6, 7, 16, 33])
def test_two_sided_factorisation_equals_the_reference_tdma(K):
    """The fused kernels' two-sided sweep solves the reference's tridiagonal systems (src/pres_2.cxx:202-324) to rounding."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as O
    itot, jtot = 16, 8
    g = O.Grid(itot, jtot, K, 100., 80., 60., 3, 3, 1, np.float64,
               z=np.cumsum(np.linspace(0.7, 1.3, K))*60./np.linspace(0.7, 1.3, K).sum() - 0.3)
    rho = np.exp(-g.z/8000.); rhoh = np.exp(-g.zh/8000.)
    pres = O.Pres2(g, rho, rhoh)
    kg = g.kgc
    dz = g.dz[kg:kg+K][:, None, None]; rr = rho[kg:kg+K][:, None, None]
    lam = pres.bmatj[None, :, None] + pres.bmati[None, None, :itot//2+1]
    b = dz*dz*rr*lam - (pres.a + pres.c)[:, None, None]
    b[0] += pres.a[0]
    top = np.full((jtot, itot//2+1), pres.c[K-1]); top[0, 0] = -pres.c[K-1]
    b[K-1] += top
    d = np.random.default_rng(K).standard_normal(b.shape)
    ref = d.copy()
    pres.tdma(ref, b.copy())
    a3 = pres.a[:, None, None]*np.ones_like(b); c3 = pres.c[:, None, None]*np.ones_like(b)
    x = twisted_tdma(a3, b, c3, d, K//2)
    assert np.sqrt(((x - ref)**2).sum()/(ref**2).sum()) <= 1e-13
