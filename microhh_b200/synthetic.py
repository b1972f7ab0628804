"""
Synthetic dry-CBL-LES-shaped inputs for tests and bench.py:
smooth divergence-free-ish modes plus seeded noise for u, v, w; a stably stratified `th`
profile with near-surface noise; Boussinesq (rhoref = 1) or anelastic base state; smooth
positive 2-D Monin-Obukhov gradients as the surface-model inputs.  Pure numpy, host side.
"""
import numpy as np


def make_case(gd, seed=2, anelastic=False, noise=0.05, ns=1):
    TF = gd.TF
    rng = np.random.default_rng(seed)
    kc, jc, ic = gd.shape
    x = (np.arange(ic) - gd.igc + 0.5)*float(gd.dx)
    xh = (np.arange(ic) - gd.igc)*float(gd.dx)
    # a y slab (npy > 1) sees its own stretch of the global coordinate; the noise is seeded per rank
    joff = gd.mpicoordy*gd.jmax
    if gd.npy > 1:
        rng = np.random.default_rng(seed + 7919*gd.mpicoordy)
    y = (np.arange(jc) - gd.jgc + 0.5 + joff)*float(gd.dy)
    yh = (np.arange(jc) - gd.jgc + joff)*float(gd.dy)
    z = gd.z.astype(np.float64); zh = gd.zh.astype(np.float64)
    Lx, Ly, Lz = float(gd.xsize), float(gd.ysize), float(gd.zsize)
    twopi = 2.*np.pi

    def modes(xx, yy, zz, phase):
        X = xx[None, None, :]; Y = yy[None, :, None]; Z = zz[:, None, None]
        return (np.sin(twopi*X/Lx + phase)*np.cos(twopi*2*Y/Ly)*np.cos(np.pi*Z/Lz)
                + 0.5*np.cos(twopi*3*X/Lx)*np.sin(twopi*Y/Ly + phase)*np.sin(twopi*Z/Lz))

    def interior_noise(amp):
        a = np.zeros((kc, jc, ic))
        a[gd.kstart:gd.kend, gd.jstart:gd.jend, gd.istart:gd.iend] = amp*(
            rng.random((gd.kmax, gd.jmax, gd.imax)) - 0.5)
        return a

    u = modes(xh, y, z, 0.3) + interior_noise(noise)
    v = modes(x, yh, z, 1.1) + interior_noise(noise)
    w = 0.5*modes(x, y, zh, 2.0)*np.sin(np.pi*zh/Lz)[:, None, None] + interior_noise(noise)
    w[:gd.kstart+1] = 0.
    w[gd.kend:] = 0.
    th = 300. + 0.003*z[:, None, None] + 0.1*modes(x, y, z, 0.7) + interior_noise(0.1)*(z[:, None, None] < 0.1*Lz)
    fields = dict(u=u, v=v, w=w, th=th)
    scal = ["th"]
    for n in range(1, ns):
        name = f"s{n}"
        fields[name] = 1. + 0.5*modes(x, y, z, 0.5*n) + interior_noise(noise)
        scal.append(name)

    if anelastic:
        rhoref = np.exp(-z/8000.); rhorefh = np.exp(-zh/8000.)
    else:
        rhoref = np.ones(kc); rhorefh = np.ones(kc)
    thref = np.full(kc, 300.); threfh = np.full(kc, 300.)

    X2 = x[None, :]; Y2 = y[:, None]
    smooth2 = lambda ph: 1. + 0.3*np.sin(twopi*X2/Lx + ph)*np.cos(twopi*Y2/Ly)
    two_d = dict(
        dudz_mo=0.05*smooth2(0.2), dvdz_mo=0.03*smooth2(1.3), dbdz_mo=-1e-4*smooth2(0.6),
        z0m=np.full((jc, ic), 0.1),
        u_fluxbot=-0.02*smooth2(0.9), v_fluxbot=-0.01*smooth2(2.1),
        u_fluxtop=np.zeros((jc, ic)), v_fluxtop=np.zeros((jc, ic)),
        u_gradbot=0.05*smooth2(0.2), v_gradbot=0.03*smooth2(1.3),
        u_gradtop=np.zeros((jc, ic)), v_gradtop=np.zeros((jc, ic)),
    )
    for s in scal:
        two_d[f"{s}_fluxbot"] = 0.1*smooth2(0.4)
        two_d[f"{s}_fluxtop"] = np.zeros((jc, ic))
        two_d[f"{s}_gradbot"] = -0.01*smooth2(0.4)
        two_d[f"{s}_gradtop"] = np.full((jc, ic), 0.003)

    cast = lambda a: np.ascontiguousarray(a.astype(TF))
    out = {k: cast(a) for k, a in fields.items()}
    out.update({k: cast(a) for k, a in two_d.items()})
    out.update(rhoref=cast(rhoref), rhorefh=cast(rhorefh), thref=cast(thref), threfh=cast(threfh))
    out["scalars"] = scal
    for name in ["u", "v", "w"] + scal:
        out[name + "t"] = np.zeros(gd.shape, TF)
    out["evisc"] = np.zeros(gd.shape, TF)
    out["p"] = np.zeros(gd.shape, TF)
    return out


def slab_of(case, gd_global, gd_local):
    """The y slab of rank `gd_local.mpicoordy` cut out of a global case (ghost rows included: local row jl is
    global row mpicoordy*jmax + jl of the ghosted global array).  1-D profiles and names are shared."""
    r = gd_local.mpicoordy
    j0, j1 = r*gd_local.jmax, r*gd_local.jmax + gd_local.jcells
    out = {}
    for k, a in case.items():
        if isinstance(a, np.ndarray) and a.ndim == 3 and a.shape == gd_global.shape:
            out[k] = np.ascontiguousarray(a[:, j0:j1, :])
        elif isinstance(a, np.ndarray) and a.ndim == 2 and a.shape == gd_global.shape2d:
            out[k] = np.ascontiguousarray(a[j0:j1, :])
        else:
            out[k] = a
    return out


def fill_fields_device(f, gd, seed=2, noise=0.01):
    """The same synthetic dry-CBL-shaped state as make_case, generated ON THE DEVICE straight into the tensors of a
    dycore.Fields `f` (u, v, w, th and the 2-D surface-model companions), level by level, so that grids whose host copy
    would not fit the box (1024^3 fp64: 8.7 GB per field) can be benchmarked.  Returns the 1-D base-state profiles
    (host numpy) for Context.set_basestate.  Values follow make_case's formulas; the noise stream differs (torch)."""
    import torch
    dev = f["u"].device
    tdt = f["u"].dtype
    gen = torch.Generator(device=dev); gen.manual_seed(seed + 7919*gd.mpicoordy)
    kc, jc, ic = gd.shape
    f64 = torch.float64
    joff = gd.mpicoordy*gd.jmax
    ar = lambda n: torch.arange(n, device=dev, dtype=f64)
    x = (ar(ic) - gd.igc + 0.5)*float(gd.dx); xh = (ar(ic) - gd.igc)*float(gd.dx)
    y = (ar(jc) - gd.jgc + 0.5 + joff)*float(gd.dy); yh = (ar(jc) - gd.jgc + joff)*float(gd.dy)
    z = gd.z.astype(np.float64); zh = gd.zh.astype(np.float64)
    Lx, Ly, Lz = float(gd.xsize), float(gd.ysize), float(gd.zsize)
    twopi = 2.*np.pi

    def modes(xx, yy, zk, phase):
        X = xx[None, :]; Y = yy[:, None]
        return (torch.sin(twopi*X/Lx + phase)*torch.cos(twopi*2*Y/Ly)*np.cos(np.pi*zk/Lz)
                + 0.5*torch.cos(twopi*3*X/Lx)*torch.sin(twopi*Y/Ly + phase)*np.sin(twopi*zk/Lz))

    def inoise(amp):
        a = torch.zeros((jc, ic), device=dev, dtype=f64)
        a[gd.jstart:gd.jend, gd.istart:gd.iend] = amp*(torch.rand((gd.jmax, gd.imax), device=dev, dtype=f64, generator=gen) - 0.5)
        return a

    for k in range(kc):
        f["u"][k].copy_((modes(xh, y, z[k], 0.3) + inoise(noise)).to(tdt))
        f["v"][k].copy_((modes(x, yh, z[k], 1.1) + inoise(noise)).to(tdt))
        if k <= gd.kstart or k >= gd.kend:
            f["w"][k].zero_()
        else:
            f["w"][k].copy_((0.5*modes(x, y, zh[k], 2.0)*np.sin(np.pi*zh[k]/Lz) + inoise(noise)).to(tdt))
        th = 300. + 0.003*z[k] + 0.1*modes(x, y, z[k], 0.7)
        if z[k] < 0.1*Lz:
            th = th + inoise(0.1)
        f["th"][k].copy_(th.to(tdt))
        for n, s in enumerate(f.scalars[1:], start=1):
            f[s][k].copy_((1. + 0.5*modes(x, y, z[k], 0.5*n) + inoise(noise)).to(tdt))
    X2 = x[None, :]; Y2 = y[:, None]
    smooth2 = lambda ph: 1. + 0.3*torch.sin(twopi*X2/Lx + ph)*torch.cos(twopi*Y2/Ly)
    two_d = dict(dudz_mo=0.05*smooth2(0.2), dvdz_mo=0.03*smooth2(1.3), dbdz_mo=-1e-4*smooth2(0.6),
                 z0m=torch.full((jc, ic), 0.1, device=dev, dtype=f64),
                 u_fluxbot=-0.02*smooth2(0.9), v_fluxbot=-0.01*smooth2(2.1),
                 u_gradbot=0.05*smooth2(0.2), v_gradbot=0.03*smooth2(1.3))
    for s in f.scalars:
        two_d[f"{s}_fluxbot"] = 0.1*smooth2(0.4)
        two_d[f"{s}_gradbot"] = -0.01*smooth2(0.4)
        two_d[f"{s}_gradtop"] = torch.full((jc, ic), 0.003, device=dev, dtype=f64)
    for n, a in two_d.items():
        f[n].copy_(a.to(tdt))
    TF = gd.TF
    ones = np.ones(kc, TF)
    return dict(rhoref=ones, rhorefh=ones.copy(), thref=np.full(kc, 300., TF), threfh=np.full(kc, 300., TF))
