"""
MicroHH `.ini` case files -> the hot path's configuration (host logic, no GPU).

The `.ini` files stay unchanged (BASELINE.json north_star): this module reads them with the reference's rules
(`src/input.cxx`: `[section]`, `key=value`, `#` comments, `key[sub]=value`) and applies the reference's defaults for the
switches of the dynamical core:

  * `[grid] swspatialorder` selects the grid order; `[advec] swadvec`, `[diff] swdiff`, `[pres] swpres` default to it
    (src/advec.cxx:54-59, src/diff.cxx:56-61, src/pres.cxx:67-71);
  * ghost cells as the scheme constructors request them (src/advec_2.cxx:40-43, src/advec_2i5.cxx:39-46,
    src/grid.cxx:87-92);
  * `[boundary] mbcbot/mbctop` noslip -> Dirichlet, freeslip / neumann -> Neumann; `sbcbot/sbctop` dirichlet | neumann |
    flux (src/boundary.cxx:197-260); `swboundary != default` switches the surface model on (src/diff_smag2.cxx:495-506);
  * `[diff] cs, tPr, swmason, dnmax` (src/diff_smag2.cxx:275-290), `[advec] cflmax, fluxlimit_list`,
    `[fields] visc, svisc`, `[thermo] swthermo`.

`CaseConfig.unsupported()` says which switches of a case lie outside the path this library accelerates (those parts keep
running in MicroHH; the fused `mhh_dycore_substep` needs an empty list).
"""
import numpy as np

from . import capi
from .grid import GridData


def read_ini(path):
    """{section: {key: value-string}} with the reference's syntax (src/input.cxx)."""
    out = {}
    sec = None
    with open(path) as fh:
        for raw in fh:
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            if line.startswith("[") and line.endswith("]"):
                sec = line[1:-1].strip()
                out.setdefault(sec, {})
                continue
            if "=" not in line or sec is None:
                raise ValueError(f"{path}: cannot parse line {raw!r}")
            k, v = line.split("=", 1)
            out[sec][k.strip()] = v.strip()
    return out


def _get(ini, sec, key, default=None, conv=str):
    v = ini.get(sec, {}).get(key)
    if v is None:
        if default is None:
            raise KeyError(f"[{sec}] {key} is required")
        return default
    if conv is bool:
        return v.lower() in ("1", "true", "yes")
    return conv(v)


def _get_sub(ini, sec, key, sub, default=None, conv=str):
    """Input::get_item(section, key, subitem) (src/input.cxx): `key[sub]` wins over the plain `key`; no default -> required."""
    d = ini.get(sec, {})
    v = d.get(f"{key}[{sub}]", d.get(key))
    if v is None:
        if default is None:
            raise KeyError(f"[{sec}] {key} (or {key}[{sub}]) is required")
        return default
    return conv(v)


_MBC = {"noslip": capi.BC_DIRICHLET, "freeslip": capi.BC_NEUMANN, "neumann": capi.BC_NEUMANN}
_SBC = {"dirichlet": capi.BC_DIRICHLET, "neumann": capi.BC_NEUMANN, "flux": capi.BC_NEUMANN}
_SWADVEC = {"2": 2, "2i5": 25, "2i4": 24, "2i62": 262, "4": 4, "4m": 41}
_SWDIFF = {"smag2": 1, "2": 2, "tke2": 3, "4": 4}


class CaseConfig:
    def __init__(self, ini):
        self.ini = ini
        g = lambda k, conv: _get(ini, "grid", k, conv=conv)
        self.itot, self.jtot, self.ktot = g("itot", int), g("jtot", int), g("ktot", int)
        self.xsize, self.ysize, self.zsize = g("xsize", float), g("ysize", float), g("zsize", float)
        self.order = _get(ini, "grid", "swspatialorder", conv=str)
        if self.order not in ("2", "4"):
            raise ValueError(f"{self.order} is an illegal value for swspatialorder")
        self.npx = _get(ini, "master", "npx", 1, int)
        self.npy = _get(ini, "master", "npy", 1, int)
        self.swadvec = _get(ini, "advec", "swadvec", self.order)
        self.swdiff = _get(ini, "diff", "swdiff", self.order)
        self.swpres = _get(ini, "pres", "swpres", self.order)
        self.swthermo = _get(ini, "thermo", "swthermo", "0")
        self.swboundary = _get(ini, "boundary", "swboundary", "default")
        self.cflmax = _get(ini, "advec", "cflmax", 1.0, float)
        fl = _get(ini, "advec", "fluxlimit_list", "")
        self.fluxlimit_list = tuple(x.strip() for x in fl.split(",") if x.strip())
        self.dnmax = _get(ini, "diff", "dnmax", 0.4, float)
        self.cs = _get(ini, "diff", "cs", 0.23, float)
        self.tPr = _get(ini, "diff", "tPr", 1./3., float)
        self.swmason = _get(ini, "diff", "swmason", True, bool)
        self.visc = _get(ini, "fields", "visc", 0., float)
        # src/boundary.cxx:188-189: mbcbot / mbctop are required keys
        self.mbcbot = _get(ini, "boundary", "mbcbot")
        self.mbctop = _get(ini, "boundary", "mbctop")
        sl = _get(ini, "fields", "slist", "")
        # the thermo class registers its prognostic scalars first (Thermo_dry: th, src/thermo_dry.cxx:377; Thermo_buoy: b,
        # src/thermo_buoy.cxx:316; Thermo_moist: thl and qt, src/thermo_moist.cxx:1089-1090)
        thermo_scalars = {"dry": ["th"], "buoy": ["b"], "moist": ["thl", "qt"]}.get(self.swthermo, [])
        self.scalars = thermo_scalars + [x.strip() for x in sl.split(",") if x.strip()]
        self.swmicro = _get(ini, "micro", "swmicro", "0")
        if self.swdiff == "tke2":
            self.scalars.append("sgstke")       # Diff_tke2's constructor adds the prognostic SGS TKE (src/diff_tke2.cxx:544)
        # per scalar, `key[name]` before `key` (src/boundary.cxx:234-235, src/fields.cxx:412); required when there are scalars
        self.sbcbot = {n: _get_sub(ini, "boundary", "sbcbot", n) for n in self.scalars}
        self.sbctop = {n: _get_sub(ini, "boundary", "sbctop", n) for n in self.scalars}
        # Thermo_buoy reads the diffusivity of b under the name th (src/thermo_buoy.cxx:320)
        self.svisc = {n: _get_sub(ini, "fields", "svisc", "th" if (n == "b" and self.swthermo == "buoy") else n, conv=float) for n in self.scalars}

    @classmethod
    def from_file(cls, path):
        return cls(read_ini(path))

    def ghost_cells(self):
        """(igc, jgc, kgc) as the reference's constructors set them."""
        if self.order == "4":
            return 3, 3, 3
        if self.swadvec == "2i5":
            return 3, 3, (2 if self.fluxlimit_list else 1)
        if self.swadvec == "2i4":
            return 2, 2, 2                                    # src/advec_2i4.cxx:38-41
        if self.swadvec == "2i62":
            return 3, 3, (2 if self.fluxlimit_list else 1)    # src/advec_2i62.cxx:42-45
        return 1, 1, 1

    def unsupported(self):
        """Switches of this case outside the accelerated path (empty list = the fused sub-step applies)."""
        bad = []
        if self.swadvec not in _SWADVEC:
            bad.append(f"swadvec={self.swadvec}")
        if self.swdiff not in _SWDIFF:
            bad.append(f"swdiff={self.swdiff}")
        if self.swpres not in ("2", "4"):
            bad.append(f"swpres={self.swpres}")
        if self.swthermo not in ("0", "dry", "buoy", "moist"):
            bad.append(f"swthermo={self.swthermo}")
        if self.swthermo == "buoy" and self.swdiff not in ("2", "4"):
            bad.append("swthermo=buoy with an LES closure (the closures are wired to Thermo_dry / Thermo_moist)")
        if self.swthermo == "moist" and (self.order != "2" or self.swdiff not in ("smag2", "2")):
            bad.append("swthermo=moist needs a 2nd-order grid and swdiff=smag2 or 2")
        if self.swmicro != "0":
            bad.append(f"swmicro={self.swmicro}")
        if self.order == "4" and (self.swadvec not in ("4", "4m") or (self.swdiff, self.swpres) != ("4", "4")) and not bad:
            bad.append("4th-order grid with mixed schemes")
        if self.swdiff == "tke2" and (self.order != "2" or self.swboundary == "default"):
            bad.append("swdiff=tke2 needs a 2nd-order grid and a surface model (src/diff_tke2.cxx:555-558)")
        if self.order == "4" and self.swthermo not in ("0", "buoy"):
            bad.append("4th-order grid with thermo other than buoy")
        if self.mbcbot not in _MBC or self.mbctop not in _MBC:
            bad.append(f"mbcbot/mbctop={self.mbcbot}/{self.mbctop}")
        for n in self.scalars:
            if self.sbcbot[n] not in _SBC or self.sbctop[n] not in _SBC:
                bad.append(f"sbcbot/sbctop[{n}]={self.sbcbot[n]}/{self.sbctop[n]}")
        if len(self.scalars) > capi.MHH_MAX_SCALARS:
            bad.append(f"{len(self.scalars)} scalars (max {capi.MHH_MAX_SCALARS})")
        if self.npx != 1:
            bad.append(f"npx={self.npx} (the decomposition is y slabs: npx=1)")
        if self.fluxlimit_list and self.ghost_cells()[2] != 1:
            pass        # the limiter itself is supported; kgc=2 grids run the point-wise kernels
        return bad

    def grid_data(self, z=None, dtype=np.float64, mpicoordy=0):
        igc, jgc, kgc = self.ghost_cells()
        return GridData(self.itot, self.jtot, self.ktot, self.xsize, self.ysize, self.zsize, igc, jgc, kgc, dtype, z=z,
                        npx=1, npy=self.npy, mpicoordy=mpicoordy, order=int(self.order))

    def make_params(self):
        """mhh_params of the case (raises when the case is outside the accelerated path)."""
        bad = self.unsupported()
        if bad:
            raise ValueError("outside the accelerated path: " + ", ".join(bad))
        p = capi.ParamsC()
        p.swadvec = _SWADVEC[self.swadvec]; p.swdiff = _SWDIFF[self.swdiff]
        p.swthermo = {"0": 0, "dry": 1, "buoy": 2, "moist": 3}[self.swthermo]
        p.surface_model = int(self.swboundary != "default")
        p.sw_mason = int(self.swmason)
        p.cs = self.cs; p.tPr = self.tPr
        p.mbcbot = _MBC[self.mbcbot]; p.mbctop = _MBC[self.mbctop]
        for i, n in enumerate(self.scalars):
            p.sbcbot[i] = _SBC[self.sbcbot[n]]; p.sbctop[i] = _SBC[self.sbctop[n]]
        return p

    def thermo_buoy_params(self):
        """Keyword arguments of dycore.Thermo_buoy ([thermo] alpha, N2, swbaroclinic, dbdy_ls; [grid] utrans; src/thermo_buoy.cxx:318-329)."""
        ini = self.ini
        sw = _get(ini, "thermo", "swbaroclinic", False, bool)
        return dict(alpha=_get(ini, "thermo", "alpha", 0., float), n2=_get(ini, "thermo", "N2", 0., float),
                    utrans=_get(ini, "grid", "utrans", 0., float), swbaroclinic=sw,
                    dbdy_ls=_get(ini, "thermo", "dbdy_ls", conv=float) if sw else 0.)

    def thermo_moist_params(self):
        """pbot and swupdatebasestate of dycore.Thermo_moist ([thermo] pbot is required, src/thermo_moist.cxx:1105, 1115)."""
        return dict(pbot=_get(self.ini, "thermo", "pbot", conv=float), swupdatebasestate=_get(self.ini, "thermo", "swupdatebasestate", True, bool))

