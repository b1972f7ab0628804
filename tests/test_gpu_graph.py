"""
CUDA-graph replay of the full RK3 step (mhh_dycore_step on a single GPU: second call with identical arguments is captured,
later calls replay it): replayed steps must be bit-identical to eager ones -- same kernels, same arguments, same order -- for
the LES path (incl. the forked scalar update), the Deardorff closure with registered forcing, and the 4th-order DNS path; a
change of dt or of the registered forcing falls back to eager steps and re-captures.
"""
import numpy as np
import pytest

from util import make_pair, add_sgstke
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def run(monkeypatch, graph, build, nsteps=5, dts=None, dt=0.5):
    monkeypatch.setenv("MHH_GRAPH", "1" if graph else "0")
    D, ctx, f, dyc = build()
    for n in range(nsteps):
        dyc.step(f, dts[n] if dts else dt)
    ctx.sync()
    out = {n: f[n].cpu().numpy() for n in ["u", "v", "w", "p"] + f.scalars}
    rep = ctx.graph_replays
    launches = ctx.launch_count
    ctx.close()
    return out, rep, launches


def les(dtype, swdiff="smag2", forcing=False, igc=4):
    def build():
        from microhh_b200 import dycore as D
        g, gd, case = make_pair(64, 32, 24, dtype, stretched=True, anelastic=True, ns=2, igc=igc)
        if swdiff == "tke2":
            add_sgstke(g, case)
            O.tke2_enforce_min(g, case["sgstke"])
        ctx = D.Context(gd, 0)
        ctx.set_basestate(case["rhoref"], case["rhorefh"], case["thref"], case["threfh"])
        f = D.Fields(ctx, case, scalars=case["scalars"])
        prm = D.make_params(swdiff=swdiff)
        keep = []
        if swdiff == "tke2":
            T = D.Diff_tke2(ctx, prm, f); T.register(); keep.append(T)
        if forcing:
            rng = np.random.default_rng(4)
            p = lambda s=1.: (s*rng.standard_normal(gd.kcells)).astype(dtype)
            F = D.Forcing(ctx, f, swbuffer=True, zstart=float(0.7*g.zsize), sigma=2., beta=2.,
                          bufferprofs=dict(u=p(), v=p(), w=p(0.1), th=(300. + p()).astype(dtype)), swlspres="geo", fc=1e-4, ug=p(), vg=p())
            # (not swlspres = uflux: its domain means are atomic sums, whose rounding differs from run to run even eagerly)
            F.register(); keep.append(F)
        f._keep_alive = keep
        return D, ctx, f, D.Dycore(ctx, prm)
    return build


def dns(dtype):
    def build():
        from microhh_b200 import dycore as D
        from microhh_b200.grid import GridData
        from microhh_b200.synthetic import make_case
        gd = GridData(64, 48, 32, 2*np.pi, np.pi, 2., 3, 3, 3, dtype, order=4)
        case = make_case(gd)
        ctx = D.Context(gd, 0)
        ctx.set_basestate(case["rhoref"], case["rhorefh"], case["thref"], case["threfh"])
        f = D.Fields(ctx, case, visc=1e-3, svisc=1e-3)
        prm = D.make_params(swadvec="4m", swdiff="4", swthermo=None, surface_model=False, mbcbot=0, mbctop=0)
        return D, ctx, f, D.Dycore(ctx, prm)
    return build


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("config", ["les", "les_igc3", "tke2_forcing", "dns"])
def test_graph_replay_is_bit_identical(dtype, config, monkeypatch):
    build = {"les": les(dtype), "les_igc3": les(dtype, igc=3), "tke2_forcing": les(dtype, "tke2", True), "dns": dns(dtype)}[config]
    dt = 1e-3 if config == "dns" else 0.5
    eager, rep0, l0 = run(monkeypatch, False, build, dt=dt)
    graph, rep1, l1 = run(monkeypatch, True, build, dt=dt)
    assert rep0 == 0
    assert rep1 == 4, rep1                       # step 1 eager, step 2 captured + replayed, 3..5 replayed
    assert l0 == l1                              # the launch count reports the replayed kernels too
    for n in eager:
        assert np.array_equal(eager[n], graph[n]), n
        assert np.isfinite(eager[n]).all(), n


def test_graph_follows_changing_arguments(monkeypatch):
    """adaptive dt: steps whose dt differs from the captured one run eagerly; a dt seen twice in a row is captured again"""
    dts = [0.5, 0.5, 0.5, 0.25, 0.3, 0.3, 0.3]
    eager, _, _ = run(monkeypatch, False, les(np.float64), len(dts), dts)
    graph, rep, _ = run(monkeypatch, True, les(np.float64), len(dts), dts)
    assert rep == 4, rep                         # 0.5: steps 2, 3; 0.3: steps 6, 7
    for n in eager:
        assert np.array_equal(eager[n], graph[n]), n
