// mhhb200 -- host drivers of the tendency stages: eddy viscosity, fused advection + diffusion (+ buoyancy) of every scheme family,
// the time-step limiter reductions.
#include "host_common.cuh"

namespace mhhhost {

template <typename TF>
int vec_width(const GridDev<TF>& g, std::initializer_list<const void*> ptrs)
{
    int v = 2;
    if (g.icells % 2 != 0 || (g.igc - TILE_H) % 2 != 0 || (g.ijcells % 2) != 0) v = 1;
    for (const void* p : ptrs)
        if (p && (reinterpret_cast<uintptr_t>(p) % (2 * sizeof(TF))) != 0) v = 1;
    return v;
}

inline int pick_kchunk_waves(int ntiles_xy, int kmax, int slots, int warm);

inline int pick_kchunk(int ntiles_xy, int kmax, int num_sms)
{
    // enough CTAs to fill the machine twice, but chunks of at least 16 levels (warm-up level amortised)
    int nz = (2 * num_sms + ntiles_xy - 1) / ntiles_xy;
    nz = std::max(1, std::min(nz, std::max(1, kmax / 16)));
    return (kmax + nz - 1) / nz;
}

// (defined with the other TMA helpers below)
template <typename TF> bool make_field_tmap(CUtensorMap* m, const void* fld, const GridDev<TF>& g, int bx, int by);
template <typename TF> bool tma_ok(const GridDev<TF>& g, std::initializer_list<const void*> ptrs);
inline int pick_kchunk_waves(int ntiles_xy, int kmax, int slots, int warm);

template <typename TF>
int evisc_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, const TF* n2)
{
    NEED_BASE(c);
    const GridDev<TF>& g = c->g;
    NEED(c, f->evisc, "evisc"); NEED(c, f->u, "u"); NEED(c, f->v, "v"); NEED(c, f->w, "w");
    EviscArgs<TF> a{};
    a.evisc = P<TF>(f->evisc); a.u = P<TF>(f->u); a.v = P<TF>(f->v); a.w = P<TF>(f->w);
    a.n2 = n2; a.th = nullptr;
    // Thermo_type::Disabled (src/diff_smag2.cxx:507-545): no stability correction, calc_evisc_neutral
    const bool neutral = !n2 && prm->swthermo == 0;
    if (!n2 && !neutral)
    {
        if (f->ns < 1 || !f->s[0]) { c->err = "exec_viscosity: no N2 field and no scalar 0 (th) to derive it from"; return MHH_E_INVALID; }
        a.th = P<TF>(f->s[0]);
    }
    a.n2mode = n2 ? 0 : 1;
    a.surface = prm->surface_model; a.mason = prm->sw_mason;
    a.cs = (TF)prm->cs; a.tPr = (TF)prm->tPr;
    if (a.surface)
    {
        NEED(c, f->dudz_mo, "dudz_mo"); NEED(c, f->dvdz_mo, "dvdz_mo"); NEED(c, f->z0m, "z0m");
        if (!neutral) NEED(c, f->dbdz_mo, "dbdz_mo");
        a.dudz = P<TF>(f->dudz_mo); a.dvdz = P<TF>(f->dvdz_mo); a.dbdz = P<TF>(f->dbdz_mo); a.z0m = P<TF>(f->z0m);
    }
    if (neutral)
    {
        evisc_neutral_kernel<TF><<<c->grd_interior(), c->blk(), 0, c->stream>>>(a, g, c->d_mlen0, (TF)f->visc);
        KCHECKN(c, "evisc_neutral_kernel");
    }
    else if (!c->force_plain && c->evisc_tma && e3_pick_hl(g.istart, (int)sizeof(TF)) > 0 && g.igc >= e3_pick_hl(g.istart, (int)sizeof(TF))
             && tma_ok<TF>(g, {a.u, a.v, a.w, a.n2mode == 1 ? (const void*)a.th : (const void*)a.u}))
    {
        // TMA-staged, warp-specialised kernel (evisc3_kernels.cuh): (32 | 64) x 8 tiles, eight consumer warps + one producer warp
        constexpr int TY = 8;
        const int hl = e3_pick_hl(g.istart, (int)sizeof(TF));
        // points per lane (tile 32 or 64 wide) and resident CTAs per SM the kernel is compiled for (register cap); measured at
        // 512^3 in ms/step (cp.async kernel: fp64 6.09, fp32 3.96):
        //   fp64  NPL=1: MB 2 | 3 | 4 = 4.82 | 4.59 | 5.73      NPL=2: MB 2 = 5.33 (3, 4 spill)
        //   fp32  NPL=2: MB 3 | 4 = 2.86 | 3.00                  NPL=1: MB 4 = 3.04
        // (deeper rings -- 6 or 8 planes per field -- and 128-wide tiles with four points per lane were slower: the kernel is
        // bound by the latency of its dependent fp64 chains, so resident warps count, not bytes in flight)
        const int npl = (c->evisc3_npl == 1 || c->evisc3_npl == 2) ? c->evisc3_npl : (sizeof(TF) == 8 ? 1 : 2);
        const int mb = (c->evisc3_mb >= 2 && c->evisc3_mb <= 4) ? c->evisc3_mb : 3;
        const int ring = 4;
        const int ew = e3_w(npl);
        const int ntx = (g.imax + ew - 1) / ew, nty = (g.jmax + TY - 1) / TY;
        EviscTileArgs<TF> t{a, c->d_mlen0, pick_kchunk_waves(ntx * nty, g.kmax, c->num_sms * mb, 1)};
        dim3 grid(ntx, nty, (g.kmax + t.kchunk - 1) / t.kchunk);
        const size_t smem = evisc3_smem(sizeof(TF), t.kchunk, TY, hl, npl, ring);
        CUtensorMap tu, tv, tw, tth;
        if (!make_field_tmap<TF>(&tu, a.u, g, e3_px(hl, npl), TY + 2) || !make_field_tmap<TF>(&tv, a.v, g, e3_px(hl, npl), TY + 2) ||
            !make_field_tmap<TF>(&tw, a.w, g, e3_px(hl, npl), TY + 2) ||
            !make_field_tmap<TF>(&tth, a.n2mode == 1 ? (const void*)a.th : (const void*)a.u, g, e3_px(hl, npl), TY + 2))
        { c->err = "cuTensorMapEncodeTiled failed"; return MHH_E_CUDA; }
#define E3(S, H, MB, N, R) do { \
            static size_t attr_smem_dev[64] = {0}; size_t& attr_smem = attr_smem_dev[c->device & 63];   /* the attribute is per device */ \
            if (attr_smem < smem) { CUDA_TRY(c, cudaFuncSetAttribute(evisc3_kernel<TF, S, TY, H, MB, N, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; } \
            evisc3_kernel<TF, S, TY, H, MB, N, R><<<grid, 32 * (TY + 1), smem, c->stream>>>(tu, tv, tw, tth, t, g); } while (0)
#define E3R(S, H, MB, N) E3(S, H, MB, N, 4)       /* deeper rings (6, 8 planes per field) measured slower: 5.96 - 6.94 vs 5.66 ms/step */
#define E3N(S, H, MB) do { if (npl == 1) E3R(S, H, MB, 1); else E3R(S, H, MB, 2); } while (0)
#define E3M(S, H) do { if (mb == 2) E3N(S, H, 2); else if (mb == 4) E3N(S, H, 4); else E3N(S, H, 3); } while (0)
#define E3H(S) do { if (hl == 1) E3M(S, 1); else if (hl == 2) E3M(S, 2); else E3M(S, 4); } while (0)
        if (a.surface) E3H(true); else E3H(false);
#undef E3H
#undef E3M
#undef E3N
#undef E3R
#undef E3
        KCHECKN(c, "evisc3_kernel");
    }
    else if (!c->force_plain)
    {
        const int ty = c->tile_y;
        const int ntx = (g.imax + TILE_X - 1) / TILE_X, nty = (g.jmax + ty - 1) / ty;
        const int mb = (ty == 16) ? 1 : (c->evisc_mb == 3 || c->evisc_mb == 4 ? c->evisc_mb : 2);
        EviscTileArgs<TF> t{a, c->d_mlen0, pick_kchunk_waves(ntx * nty, g.kmax, c->num_sms * mb, 1)};
        dim3 grid(ntx, nty, (g.kmax + t.kchunk - 1) / t.kchunk);
        const size_t smem = evisc_tile_smem(sizeof(TF), t.kchunk, ty);
        int vec = 2;
        if (g.icells % 2 != 0 || (g.igc - EH) % 2 != 0 || (g.ijcells % 2) != 0) vec = 1;
        for (const void* p : {(const void*)a.u, (const void*)a.v, (const void*)a.w})
            if (reinterpret_cast<uintptr_t>(p) % (2 * sizeof(TF)) != 0) vec = 1;
#define ET4(S, V, Y, MB) do { \
            static size_t attr_smem_dev[64] = {0}; size_t& attr_smem = attr_smem_dev[c->device & 63];   /* the attribute is per device */ \
            if (attr_smem < smem) { CUDA_TRY(c, cudaFuncSetAttribute(evisc_tile_kernel<TF, S, V, Y, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; } \
            evisc_tile_kernel<TF, S, V, Y, MB><<<grid, TILE_X * Y, smem, c->stream>>>(t, g); } while (0)
#define ET3(S, V, Y) do { if (c->evisc_mb == 3) ET4(S, V, Y, 3); else if (c->evisc_mb == 4) ET4(S, V, Y, 4); else ET4(S, V, Y, 512 / (TILE_X * Y)); } while (0)
#define ET(S, V) do { if (ty == 16) ET4(S, V, 16, 1); else ET3(S, V, 8); } while (0)
        if (a.surface) { if (vec == 2) ET(true, 2); else ET(true, 1); }
        else { if (vec == 2) ET(false, 2); else ET(false, 1); }
#undef ET
#undef ET3
#undef ET4
        KCHECKN(c, "evisc_tile_kernel");
    }
    else
    {
        evisc_kernel<TF><<<c->grd_interior(), c->blk(), 0, c->stream>>>(a, g, c->d_mlen0);
        KCHECKN(c, "evisc_kernel");
    }
    if (!a.surface)
    {
        dim3 b(64, 4), gr((g.icells + 63) / 64, (g.jcells + 3) / 4);
        evisc_mirror_kernel<TF><<<gr, b, 0, c->stream>>>(a.evisc, g);
        KCHECKN(c, "evisc_mirror_kernel");
    }
    return cyclic_impl<TF>(c, a.evisc, MHH_EDGE_BOTH, false);
}
template <typename TF>
MomArgs<TF> mom_args(const mhh_fields* f)
{
    MomArgs<TF> a{};
    a.ut = P<TF>(f->ut); a.vt = P<TF>(f->vt); a.wt = P<TF>(f->wt);
    a.u = P<TF>(f->u); a.v = P<TF>(f->v); a.w = P<TF>(f->w);
    a.evisc = P<TF>(f->evisc);
    a.th = (f->ns > 0) ? P<TF>(f->s[0]) : nullptr;
    a.u_fluxbot = P<TF>(f->u_fluxbot); a.u_fluxtop = P<TF>(f->u_fluxtop);
    a.v_fluxbot = P<TF>(f->v_fluxbot); a.v_fluxtop = P<TF>(f->v_fluxtop);
    a.visc = (TF)f->visc;
    return a;
}

template <typename TF>
ScalArgs<TF> scal_args(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, int n, const mhh_tke2* tke = nullptr)
{
    const GridDev<TF>& g = c->g;
    ScalArgs<TF> a{};
    a.st = P<TF>(f->st[n]); a.s = P<TF>(f->s[n]);
    a.u = P<TF>(f->u); a.v = P<TF>(f->v); a.w = P<TF>(f->w);
    a.evisc = P<TF>(f->evisc);
    a.fluxbot = P<TF>(f->s_fluxbot[n]); a.fluxtop = P<TF>(f->s_fluxtop[n]);
    a.visc = (TF)f->svisc[n];
    a.tPr = prm ? (TF)prm->tPr : TF(1);
    if (tke)
    {
        // Diff_tke2::exec (src/diff_tke2.cxx:733-789): tPr_dummy = 1; sgstke diffuses with the eddy viscosity for momentum,
        // every other scalar with the one for heat when there is buoyancy
        a.tPr = TF(1);
        if (n != tke->isgstke && prm && prm->swthermo != 0) a.evisc = P<TF>(tke->eviscs);
    }
    // 1./(dx*dx) is formed in double in the reference and narrowed to TF (src/diff_smag2.cxx:445)
    a.dxidxi = (TF)(1. / ((double)g.dx * (double)g.dx));
    a.dyidyi = (TF)(1. / ((double)g.dy * (double)g.dy));
    return a;
}

template <typename TF>
int check_mom(Ctx<TF>* c, const mhh_fields* f, bool need_evisc, bool surface)
{
    NEED(c, f, "fields");
    NEED(c, f->u, "u"); NEED(c, f->v, "v"); NEED(c, f->w, "w");
    NEED(c, f->ut, "ut"); NEED(c, f->vt, "vt"); NEED(c, f->wt, "wt");
    if (f->ns < 0 || f->ns > MHH_MAX_SCALARS) { c->err = "ns out of range"; return MHH_E_INVALID; }
    for (int n = 0; n < f->ns; ++n) { NEED(c, f->s[n], "scalar"); NEED(c, f->st[n], "scalar tendency"); }
    if (need_evisc) NEED(c, f->evisc, "evisc");
    if (need_evisc && surface)
    {
        NEED(c, f->u_fluxbot, "u_fluxbot"); NEED(c, f->u_fluxtop, "u_fluxtop");
        NEED(c, f->v_fluxbot, "v_fluxbot"); NEED(c, f->v_fluxtop, "v_fluxtop");
        for (int n = 0; n < f->ns; ++n) { NEED(c, f->s_fluxbot[n], "s_fluxbot"); NEED(c, f->s_fluxtop[n], "s_fluxtop"); }
    }
    return MHH_OK;
}

// ---- TMA tensor maps (driver entry point fetched through the runtime; libcuda is not linked) ----
inline PFN_cuTensorMapEncodeTiled tmap_encoder()
{
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    if (!fn)
    {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    }
    return fn;
}

// 3-D map over a ghosted field (icells, jcells, kcells) with a box of (bx, by, 1) elements
template <typename TF>
bool make_field_tmap(CUtensorMap* m, const void* fld, const GridDev<TF>& g, int bx, int by)
{
    PFN_cuTensorMapEncodeTiled enc = tmap_encoder();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)g.icells, (cuuint64_t)g.jcells, (cuuint64_t)g.kcells};
    const cuuint64_t strides[2] = {(cuuint64_t)g.icells * sizeof(TF), (cuuint64_t)g.ijcells * sizeof(TF)};
    const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUtensorMapDataType dt = sizeof(TF) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    return enc(m, dt, 3, const_cast<void*>(fld), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// TMA needs 16-byte aligned base pointers and row/plane pitches that are multiples of 16 bytes
template <typename TF>
bool tma_ok(const GridDev<TF>& g, std::initializer_list<const void*> ptrs)
{
    if (((size_t)g.icells * sizeof(TF)) % 16 != 0 || (g.imax % 2) != 0 || g.igc < 3 || g.jgc < 3) return false;
    for (const void* p : ptrs)
        if (!p || (reinterpret_cast<uintptr_t>(p) % 16) != 0) return false;
    return tmap_encoder() != nullptr;
}

// z-chunk of the marching kernels: fill whole waves of resident CTAs, pay (warm-up levels)/kchunk per chunk
inline int pick_kchunk_waves(int ntiles_xy, int kmax, int slots, int warm)
{
    int best_nz = 1; double best = -1.;
    for (int nz = 1; nz <= std::max(1, kmax / 16); ++nz)
    {
        const int kchunk = (kmax + nz - 1) / nz;
        const int nzz = (kmax + kchunk - 1) / kchunk;
        const long long ctas = (long long)ntiles_xy * nzz;
        const long long waves = (ctas + slots - 1) / slots;
        const double eff = (double)ctas / (double)(waves * slots) * (double)kmax / (double)(nzz * (kchunk + warm));
        if (eff > best * 1.02) { best = eff; best_nz = nz; }
    }
    return (kmax + best_nz - 1) / best_nz;
}

// fused advection + diffusion (+ buoyancy) of u, v, w with the z-marching tile kernel
template <typename TF>
int mom_tile_launch(Ctx<TF>* c, const MomArgs<TF>& a, bool surface, bool buoy)
{
    const GridDev<TF>& g = c->g;
    const int ty = c->tile_y;
    const int ntx = (g.imax + TILE_X - 1) / TILE_X, nty = (g.jmax + ty - 1) / ty;
    MomTileArgs<TF> t{a, pick_kchunk(ntx * nty, g.kmax, c->num_sms)};
    dim3 grid(ntx, nty, (g.kmax + t.kchunk - 1) / t.kchunk);
    const size_t smem = mom_tile_smem(sizeof(TF), t.kchunk, ty);
    const int vec = vec_width<TF>(g, {a.u, a.v, a.w, a.evisc});
#define MT(S, B, V, Y) do { \
        static size_t attr_smem_dev[64] = {0}; size_t& attr_smem = attr_smem_dev[c->device & 63];   /* the attribute is per device */ \
        if (attr_smem < smem) { CUDA_TRY(c, cudaFuncSetAttribute(mom_tile_kernel<TF, S, B, V, Y>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; } \
        mom_tile_kernel<TF, S, B, V, Y><<<grid, TILE_X * Y, smem, c->stream>>>(t, g); } while (0)
#define MT2(S, B) do { if (vec == 2) { if (ty == 16) MT(S, B, 2, 16); else MT(S, B, 2, 8); } \
                       else { if (ty == 16) MT(S, B, 1, 16); else MT(S, B, 1, 8); } } while (0)
    if (surface && buoy) MT2(true, true);
    else if (surface) MT2(true, false);
    else if (buoy) MT2(false, true);
    else MT2(false, false);
#undef MT2
#undef MT
    KCHECKN(c, "mom_tile_kernel");
    return MHH_OK;
}

// warp-specialised variant (tile3_kernels.cuh): one CTA of (3+nsc)*ty+1 warps per SM; the first scalar rides along
template <typename TF>
int mom3_launch(Ctx<TF>* c, const MomArgs<TF>& a, const ScalArgs<TF>* sc, bool surface, bool buoy, int hl, bool adv2 = false)
{
    const GridDev<TF>& g = c->g;
    const int nsc = sc ? 1 : 0;
    const int ty = (c->tile3_y && !adv2) ? c->tile3_y : (nsc ? 3 : 4);       // the Advec_2 variant exists for the default tile heights
    const int ntx = (g.imax + T2_W - 1) / T2_W, nty = (g.jmax + ty - 1) / ty;
    Tend3Args<TF> t{};
    t.m = a; if (sc) t.sc = *sc;
    t.kchunk = pick_kchunk_waves(ntx * nty, g.kmax, c->num_sms * mom3_min_blocks<TF>(), 2);
    t.prefetch = c->prefetch;
    dim3 grid(ntx, nty, (g.kmax + t.kchunk - 1) / t.kchunk);
    const size_t smem = mom3_smem(sizeof(TF), t.kchunk, ty, nsc, hl);
    CUtensorMap tu, tv, tw, te, ts, tut, tvt, twt, tst;
    const int by = ty + 2 * T2_H;
    const int px = t2_px(hl);
    if (!make_field_tmap<TF>(&tu, a.u, g, px, by) || !make_field_tmap<TF>(&tv, a.v, g, px, by) ||
        !make_field_tmap<TF>(&tw, a.w, g, px, by) || !make_field_tmap<TF>(&te, a.evisc, g, px, by) ||
        !make_field_tmap<TF>(&ts, sc ? (const void*)sc->s : (const void*)a.u, g, px, by) ||
        !make_field_tmap<TF>(&tut, a.ut, g, px, ty) || !make_field_tmap<TF>(&tvt, a.vt, g, px, ty) ||
        !make_field_tmap<TF>(&twt, a.wt, g, px, ty) || !make_field_tmap<TF>(&tst, sc ? (const void*)sc->st : (const void*)a.ut, g, px, ty))
    { c->err = "cuTensorMapEncodeTiled failed"; return MHH_E_CUDA; }
#define M3A(S, B, N, Y, H, A) do { \
        static size_t attr_smem_dev[64] = {0}; size_t& attr_smem = attr_smem_dev[c->device & 63];   /* the attribute is per device */ \
        if (attr_smem < smem) { CUDA_TRY(c, cudaFuncSetAttribute(mom3_kernel<TF, S, B, N, Y, H, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; } \
        mom3_kernel<TF, S, B, N, Y, H, A><<<grid, 32 * ((3 + N) * Y + 1), smem, c->stream>>>(tu, tv, tw, te, ts, tut, tvt, twt, tst, t, g); } while (0)
#define M3(S, B, N, Y, H) M3A(S, B, N, Y, H, false)
    // the odd-aligned variant (halo 3) exists for fp64 only
#define M3H(S, B, N, Y) do { if (hl == 4) M3(S, B, N, Y, 4); else if (sizeof(TF) == 8) M3(S, B, N, Y, (sizeof(TF) == 8 ? 3 : 4)); } while (0)
#define M3Y(S, B, N) do { if (adv2) { if (hl == 4) M3A(S, B, N, (N ? 3 : 4), 4, true); else if (sizeof(TF) == 8) M3A(S, B, N, (N ? 3 : 4), (sizeof(TF) == 8 ? 3 : 4), true); } \
                          else if (ty == 3) M3H(S, B, N, 3); else if (ty == 5) M3H(S, B, N, 5); else M3H(S, B, N, 4); } while (0)
    if (nsc)
    {
        if (surface && buoy) M3Y(true, true, 1);
        else if (surface) M3Y(true, false, 1);
        else if (buoy) M3Y(false, true, 1);
        else M3Y(false, false, 1);
    }
    else
    {
        if (surface && buoy) M3Y(true, true, 0);
        else if (surface) M3Y(true, false, 0);
        else if (buoy) M3Y(false, true, 0);
        else M3Y(false, false, 0);
    }
#undef M3Y
#undef M3H
#undef M3
#undef M3A
    if (adv2) KCHECKN(c, "mom3_kernel_advec2"); else KCHECKN(c, "mom3_kernel");
    return MHH_OK;
}

// One further scalar with the TMA-staged kernel: the scalar group of mom3_kernel alone (SONLY), 64 x (8 | 12) tiles.  Returns
// MHH_NOT_FUSED when the shared memory of the tile does not fit (the caller then runs the cp.async tile kernel).
template <typename TF>
int scal3_launch(Ctx<TF>* c, const ScalArgs<TF>& sc, bool surface, int hl)
{
    const GridDev<TF>& g = c->g;
    constexpr int TYS = t3_sonly_ty<TF>();
    if (hl != 4 && !(hl == 3 && sizeof(TF) == 8)) return MHH_NOT_FUSED;
    const int ntx = (g.imax + T2_W - 1) / T2_W, nty = (g.jmax + TYS - 1) / TYS;
    Tend3Args<TF> t{};
    t.m.u = sc.u; t.m.v = sc.v; t.m.w = sc.w; t.m.evisc = sc.evisc; t.m.visc = sc.visc;
    t.sc = sc;
    t.kchunk = pick_kchunk_waves(ntx * nty, g.kmax, c->num_sms * mom3_min_blocks<TF>(), 2);
    // the per-level profiles live in shared memory beside the planes: shorten the z-chunk until the CTA fits
    const size_t limit = (size_t)227 * 1024;
    while (t.kchunk > 16 && mom3_smem(sizeof(TF), t.kchunk, TYS, 1, hl) > limit) t.kchunk = (t.kchunk + 1) / 2;
    const size_t smem = mom3_smem(sizeof(TF), t.kchunk, TYS, 1, hl);
    if (smem > limit) return MHH_NOT_FUSED;
    t.prefetch = c->prefetch;
    dim3 grid(ntx, nty, (g.kmax + t.kchunk - 1) / t.kchunk);
    CUtensorMap tu, tv, tw, te, ts, tst;
    const int by = TYS + 2 * T2_H, px = t2_px(hl);
    if (!make_field_tmap<TF>(&tu, sc.u, g, px, by) || !make_field_tmap<TF>(&tv, sc.v, g, px, by) ||
        !make_field_tmap<TF>(&tw, sc.w, g, px, by) || !make_field_tmap<TF>(&te, sc.evisc, g, px, by) ||
        !make_field_tmap<TF>(&ts, sc.s, g, px, by) || !make_field_tmap<TF>(&tst, sc.st, g, px, TYS))
    { c->err = "cuTensorMapEncodeTiled failed"; return MHH_E_CUDA; }
#define S3(S, H) do { \
        static size_t attr_smem_dev[64] = {0}; size_t& attr_smem = attr_smem_dev[c->device & 63];   /* the attribute is per device */ \
        if (attr_smem < smem) { CUDA_TRY(c, cudaFuncSetAttribute(mom3_kernel<TF, S, false, 1, TYS, H, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; } \
        mom3_kernel<TF, S, false, 1, TYS, H, false, true><<<grid, 32 * (TYS + 1), smem, c->stream>>>(tu, tv, tw, te, ts, tst, tst, tst, tst, t, g); } while (0)
#define S3H(S) do { if (hl == 4) S3(S, 4); else S3(S, (sizeof(TF) == 8 ? 3 : 4)); } while (0)
    if (surface) S3H(true); else S3H(false);
#undef S3H
#undef S3
    KCHECKN(c, "scal3_kernel");
    return MHH_OK;
}

template <typename TF>
int scal_tile_launch(Ctx<TF>* c, const ScalArgs<TF>& a, bool surface)
{
    const GridDev<TF>& g = c->g;
    const int ty = c->tile_y;
    const int ntx = (g.imax + TILE_X - 1) / TILE_X, nty = (g.jmax + ty - 1) / ty;
    ScalTileArgs<TF> t{a, pick_kchunk(ntx * nty, g.kmax, c->num_sms)};
    dim3 grid(ntx, nty, (g.kmax + t.kchunk - 1) / t.kchunk);
    const size_t smem = scal_tile_smem(sizeof(TF), t.kchunk, ty);
    const int vec = vec_width<TF>(g, {a.s, a.evisc});
#define ST3(S, V, Y) do { \
        static size_t attr_smem_dev[64] = {0}; size_t& attr_smem = attr_smem_dev[c->device & 63];   /* the attribute is per device */ \
        if (attr_smem < smem) { CUDA_TRY(c, cudaFuncSetAttribute(scal_tile_kernel<TF, S, V, Y>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; } \
        scal_tile_kernel<TF, S, V, Y><<<grid, TILE_X * Y, smem, c->stream>>>(t, g); } while (0)
#define ST(S, V) do { if (ty == 16) ST3(S, V, 16); else ST3(S, V, 8); } while (0)
    if (surface) { if (vec == 2) ST(true, 2); else ST(true, 1); }
    else { if (vec == 2) ST(false, 2); else ST(false, 1); }
#undef ST
#undef ST3
    KCHECKN(c, "scal_tile_kernel");
    return MHH_OK;
}

// tendencies: adv / diff / buoyancy in any combination (templates keep the unused parts out)
template <typename TF>
int tend_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, bool adv, bool diff, bool buoy, const mhh_tke2* tke, bool adv2)
{
    NEED_BASE(c);
    const GridDev<TF>& g = c->g;
    if (!diff) tke = nullptr;
    if (adv2)
    {
        // Advec_2 + Diff_smag2 | Diff_tke2 fused (cases/drycblles as shipped): only the TMA-staged kernel has the variant, and only
        // the scalar that rides along in it is covered; MHH_NOT_FUSED tells the caller to run the two schemes separately
        const int hl2 = t2_pick_hl(g.igc, (int)sizeof(TF));
        bool ok = adv && diff && !c->force_plain && !c->no_tma && c->fuse_scalar && hl2 != 0 && g.igc >= TILE_H && g.jgc >= TILE_H && f->ns <= 1
                  && tma_ok<TF>(g, {f->u, f->v, f->w, f->evisc, f->ut, f->vt, f->wt, (buoy && f->ns > 0) ? (const void*)f->s[0] : (const void*)f->u});
        if (ok && f->ns == 1)
        {
            const ScalArgs<TF> s0 = scal_args<TF>(c, f, prm, 0, tke);
            ok = !f->s_fluxlimit[0] && tma_ok<TF>(g, {s0.s, s0.st}) && s0.evisc == P<TF>(f->evisc);
        }
        if (!ok) return MHH_NOT_FUSED;
    }
    const bool surface = diff && prm && (prm->surface_model || tke);
    int rc = check_mom<TF>(c, f, diff, surface);
    if (rc != MHH_OK) return rc;
    if (adv && (g.igc < 3 || g.jgc < 3)) { c->err = "advec_2i5 needs igc, jgc >= 3"; return MHH_E_INVALID; }
    if (buoy && (f->ns < 1)) { c->err = "buoyancy needs scalar 0 (th)"; return MHH_E_INVALID; }
    const MomArgs<TF> a = mom_args<TF>(f);
    dim3 gr = c->grd_interior(), b = c->blk();
#define LAUNCH_MOM(A, D, S, B) tend_uvw_kernel<TF, A, D, S, B><<<gr, b, 0, c->stream>>>(a, g)
    const bool tiles = adv && diff && g.igc >= TILE_H && g.jgc >= TILE_H && !c->force_plain;
    int first_scalar = 0;       // scalars [0, first_scalar) were handled by the fused momentum kernel
    // TMA-staged path: needs a 16-byte aligned box origin istart - halo (halo 4 with igc = 4, what the adapters request;
    // halo 3 for fp64 fields with igc = 3); everything else runs the cp.async tile kernels
    const int hl = tiles ? t2_pick_hl(g.igc, (int)sizeof(TF)) : 0;
    const bool tma = tiles && !c->no_tma && hl != 0
                     && tma_ok<TF>(g, {a.u, a.v, a.w, a.evisc, a.ut, a.vt, a.wt, buoy ? (const void*)a.th : (const void*)a.u});
    if (tiles)
    {
        if (tma)
        {
            // scalar 0 rides along as the fourth warp group when its arrays qualify for TMA too
            ScalArgs<TF> s0{};
            // measured on B200 fp64 (512^3): 5.8 ms fused (3 rows, 13 warps) vs 4.3 + 2.9 ms as two kernels
            bool fuse = f->ns > 0 && c->fuse_scalar && !f->s_fluxlimit[0];
            // (the fused scalar group reads the momentum kernel's evisc planes: with Diff_tke2 only a scalar that diffuses
            // with evisc qualifies)
            if (fuse) { s0 = scal_args<TF>(c, f, prm, 0, tke); fuse = tma_ok<TF>(g, {s0.s, s0.st}) && s0.evisc == a.evisc; }
            rc = mom3_launch<TF>(c, a, fuse ? &s0 : nullptr, surface, buoy, hl, adv2);
            if (fuse) first_scalar = 1;
        }
        else rc = mom_tile_launch<TF>(c, a, surface, buoy);
        if (rc != MHH_OK) return rc;
    }
    else if (!adv && buoy) { c->err = "tend_impl: the diffusion-only kernel carries no buoyancy (run the buoyancy on its own first)"; return MHH_E_INVALID; }
    else if (adv && diff && surface && buoy) LAUNCH_MOM(true, true, true, true);
    else if (adv && diff && surface) LAUNCH_MOM(true, true, true, false);
    else if (adv && diff && buoy) LAUNCH_MOM(true, true, false, true);
    else if (adv && diff) LAUNCH_MOM(true, true, false, false);
    else if (adv) LAUNCH_MOM(true, false, false, false);
    else if (diff && surface) LAUNCH_MOM(false, true, true, false);
    else if (diff) LAUNCH_MOM(false, true, false, false);
    else { c->err = "tend_impl: nothing to do"; return MHH_E_INVALID; }
#undef LAUNCH_MOM
    if (!tiles) KCHECKN(c, "tend_uvw_kernel");
    for (int n = first_scalar; n < f->ns; ++n)
    {
        const ScalArgs<TF> s = scal_args<TF>(c, f, prm, n, tke);
        if (adv && f->s_fluxlimit[n])
        {
            // `fluxlimit_list` scalar (src/advec_2i5.cxx:1046-1056): Koren-limited advection, then the diffusion alone
            advec_s_lim_kernel<TF><<<gr, b, 0, c->stream>>>(s.st, s.s, s.u, s.v, s.w, g);
            KCHECKN(c, "advec_s_lim_kernel");
            if (diff)
            {
                if (surface) tend_s_kernel<TF, false, true, true><<<gr, b, 0, c->stream>>>(s, g);
                else tend_s_kernel<TF, false, true, false><<<gr, b, 0, c->stream>>>(s, g);
                KCHECKN(c, "tend_s_kernel");
            }
            continue;
        }
        if (tiles)
        {
            // further scalars: the scalar group of the TMA kernel on its own, else the cp.async tile kernel
            rc = (tma && c->scal_tma && tma_ok<TF>(g, {s.s, s.st, s.evisc})) ? scal3_launch<TF>(c, s, surface, hl) : MHH_NOT_FUSED;
            if (rc == MHH_NOT_FUSED) rc = scal_tile_launch<TF>(c, s, surface);
            if (rc != MHH_OK) return rc;
            continue;
        }
#define LAUNCH_S(A, D, S) tend_s_kernel<TF, A, D, S><<<gr, b, 0, c->stream>>>(s, g)
        if (adv && diff && surface) LAUNCH_S(true, true, true);
        else if (adv && diff) LAUNCH_S(true, true, false);
        else if (adv) LAUNCH_S(true, false, false);
        else if (diff && surface) LAUNCH_S(false, true, true);
        else LAUNCH_S(false, true, false);
#undef LAUNCH_S
        KCHECKN(c, "tend_s_kernel");
    }
    return MHH_OK;
}

// Advec_2 / Diff_2 / thermo_dry buoyancy in any combination (order2_kernels.cuh)
template <typename TF>
int o2_impl(Ctx<TF>* c, const mhh_fields* f, bool adv, bool diff, bool buoy)
{
    NEED_BASE(c);
    const GridDev<TF>& g = c->g;
    int rc = check_mom<TF>(c, f, false, false);
    if (rc != MHH_OK) return rc;
    if (buoy && f->ns < 1) { c->err = "buoyancy needs scalar 0 (th)"; return MHH_E_INVALID; }
    if (!adv && !diff && !buoy) { c->err = "o2_impl: nothing to do"; return MHH_E_INVALID; }
    O2Args<TF> a{};
    a.ut = P<TF>(f->ut); a.vt = P<TF>(f->vt); a.wt = P<TF>(f->wt);
    a.u = P<TF>(f->u); a.v = P<TF>(f->v); a.w = P<TF>(f->w);
    a.th = f->ns > 0 ? P<TF>(f->s[0]) : nullptr;
    a.visc = (TF)f->visc;
    // `const double dxidxi = 1/(dx*dx);` (src/diff_2.cxx:44-45): the division itself is done in TF, then widened
    a.dxidxi = (double)(TF(1) / (g.dx * g.dx)); a.dyidyi = (double)(TF(1) / (g.dy * g.dy));
    dim3 gr = c->grd_interior(), b = c->blk();
#define O2(A, D, B) o2_uvw_kernel<TF, A, D, B><<<gr, b, 0, c->stream>>>(a, g)
    if (adv && diff && buoy) O2(true, true, true);
    else if (adv && diff) O2(true, true, false);
    else if (adv && buoy) O2(true, false, true);
    else if (adv) O2(true, false, false);
    else if (diff && buoy) O2(false, true, true);
    else if (diff) O2(false, true, false);
    else O2(false, false, true);
#undef O2
    KCHECKN(c, "o2_uvw_kernel");
    if (!adv && !diff) return MHH_OK;
    for (int n = 0; n < f->ns; ++n)
    {
        O2ScalArgs<TF> s{P<TF>(f->st[n]), P<TF>(f->s[n]), a.u, a.v, a.w, (TF)f->svisc[n], a.dxidxi, a.dyidyi};
        if (adv && diff) o2_s_kernel<TF, true, true><<<gr, b, 0, c->stream>>>(s, g);
        else if (adv) o2_s_kernel<TF, true, false><<<gr, b, 0, c->stream>>>(s, g);
        else o2_s_kernel<TF, false, true><<<gr, b, 0, c->stream>>>(s, g);
        KCHECKN(c, "o2_s_kernel");
    }
    return MHH_OK;
}

// Advec_2i4 (adv_sw = 24) / Advec_2i62 (adv_sw = 262): the advection alone (order2i_kernels.cuh); flux-limited scalars of 2i62
// take the Koren-limited kernel Advec_2i5 uses (src/advec_2i62.cxx:455-476 calls the same advec_s_lim)
template <typename TF>
int adv2i_impl(Ctx<TF>* c, const mhh_fields* f, int adv_sw)
{
    NEED_BASE(c);
    const GridDev<TF>& g = c->g;
    int rc = check_mom<TF>(c, f, false, false);
    if (rc != MHH_OK) return rc;
    const bool i4 = adv_sw == 24;
    if (g.dzi4) { c->err = "advec 2i4 / 2i62: second-order grids only"; return MHH_E_INVALID; }
    // the ghost cells the reference's constructors ask for (src/advec_2i4.cxx:38-41, src/advec_2i62.cxx:42-45)
    if (i4 ? (g.igc < 2 || g.jgc < 2 || g.kgc < 2) : (g.igc < 3 || g.jgc < 3))
    { c->err = i4 ? "advec_2i4 needs igc, jgc, kgc >= 2" : "advec_2i62 needs igc, jgc >= 3"; return MHH_E_INVALID; }
    if (g.kmax < 4) { c->err = "advec 2i4 / 2i62 need ktot >= 4"; return MHH_E_INVALID; }
    Adv2iArgs<TF> a{P<TF>(f->ut), P<TF>(f->vt), P<TF>(f->wt), P<TF>(f->u), P<TF>(f->v), P<TF>(f->w)};
    dim3 gr = c->grd_interior(), b = c->blk();
    if (i4) adv2i_uvw_kernel<TF, 4><<<gr, b, 0, c->stream>>>(a, g);
    else    adv2i_uvw_kernel<TF, 6><<<gr, b, 0, c->stream>>>(a, g);
    KCHECKN(c, "adv2i_uvw_kernel");
    for (int n = 0; n < f->ns; ++n)
    {
        NEED(c, f->s[n], "scalar"); NEED(c, f->st[n], "scalar tendency");
        if (!i4 && f->s_fluxlimit[n])
        {
            if (g.kgc < 2) { c->err = "flux-limited scalars need kgc >= 2 (src/advec_2i62.cxx:44)"; return MHH_E_INVALID; }
            advec_s_lim_kernel<TF><<<gr, b, 0, c->stream>>>(P<TF>(f->st[n]), P<TF>(f->s[n]), a.u, a.v, a.w, g);
            KCHECKN(c, "advec_s_lim_kernel");
            continue;
        }
        if (i4) adv2i_s_kernel<TF, 4><<<gr, b, 0, c->stream>>>(P<TF>(f->st[n]), P<TF>(f->s[n]), a.u, a.v, a.w, g);
        else    adv2i_s_kernel<TF, 6><<<gr, b, 0, c->stream>>>(P<TF>(f->st[n]), P<TF>(f->s[n]), a.u, a.v, a.w, g);
        KCHECKN(c, "adv2i_s_kernel");
    }
    return MHH_OK;
}

// Advec_4 (adv_sw = 4) or Advec_4m (adv_sw = 41) / Diff_4 in any combination (order4_kernels.cuh)
template <typename TF>
int o4_impl(Ctx<TF>* c, const mhh_fields* f, int adv_sw, bool diff)
{
    const bool adv = adv_sw == 4;
    const GridDev<TF>& g = c->g;
    if (!g.dzi4) { c->err = "4th-order schemes need a 4th-order grid (dzi4 / dzhi4 in mhh_grid_desc, three ghost cells)"; return MHH_E_INVALID; }
    if (g.kmax < 4) { c->err = "4th-order schemes need ktot >= 4"; return MHH_E_INVALID; }
    int rc = check_mom<TF>(c, f, false, false);
    if (rc != MHH_OK) return rc;
    O4Args<TF> a{};
    a.ut = P<TF>(f->ut); a.vt = P<TF>(f->vt); a.wt = P<TF>(f->wt);
    a.u = P<TF>(f->u); a.v = P<TF>(f->v); a.w = P<TF>(f->w);
    a.visc = (TF)f->visc;
    a.dxidxi_c = (TF)(1. / (double)(g.dx * g.dx)); a.dyidyi_c = (TF)(1. / (double)(g.dy * g.dy));
    a.dxidxi_w = TF(1) / (g.dx * g.dx); a.dyidyi_w = TF(1) / (g.dy * g.dy);
    const bool dim3 = g.jtot > 1;
    {
        ::dim3 gr = c->grd_interior(), b = c->blk();
        if (adv_sw == 41)
        {
            o4m_uvw_kernel<TF><<<gr, b, 0, c->stream>>>(a, g);
            KCHECKN(c, "o4m_uvw_kernel");
            for (int n = 0; n < f->ns; ++n)
            {
                O4ScalArgs<TF> s{P<TF>(f->st[n]), P<TF>(f->s[n]), a.u, a.v, a.w, (TF)f->svisc[n], a.dxidxi_c, a.dyidyi_c};
                o4m_s_kernel<TF><<<gr, b, 0, c->stream>>>(s, g);
                KCHECKN(c, "o4m_s_kernel");
            }
            if (!diff) return MHH_OK;
        }
        else if (adv_sw != 0 && adv_sw != 4) { c->err = "o4_impl: the 4th-order advection schemes are 4 and 41 (4m)"; return MHH_E_INVALID; }
#define O4(A, D, T) o4_uvw_kernel<TF, A, D, T><<<gr, b, 0, c->stream>>>(a, g)
        if (adv && diff) { if (dim3) O4(true, true, true); else O4(true, true, false); }
        else if (adv) { if (dim3) O4(true, false, true); else O4(true, false, false); }
        else if (diff) { if (dim3) O4(false, true, true); else O4(false, true, false); }
        else { c->err = "o4_impl: nothing to do"; return MHH_E_INVALID; }
#undef O4
        KCHECKN(c, "o4_uvw_kernel");
        for (int n = 0; n < f->ns; ++n)
        {
            O4ScalArgs<TF> s{P<TF>(f->st[n]), P<TF>(f->s[n]), a.u, a.v, a.w, (TF)f->svisc[n], a.dxidxi_c, a.dyidyi_c};
#define O4S(A, D, T) o4_s_kernel<TF, A, D, T><<<gr, b, 0, c->stream>>>(s, g)
            if (adv && diff) { if (dim3) O4S(true, true, true); else O4S(true, true, false); }
            else if (adv) { if (dim3) O4S(true, false, true); else O4S(true, false, false); }
            else { if (dim3) O4S(false, true, true); else O4S(false, true, false); }
#undef O4S
            KCHECKN(c, "o4_s_kernel");
        }
    }
    return MHH_OK;
}

template <typename TF>
int o2_cfl_impl(Ctx<TF>* c, const mhh_fields* f, double* out, int order)
{
    const GridDev<TF>& g = c->g;
    NEED(c, f->u, "u"); NEED(c, f->v, "v"); NEED(c, f->w, "w");
    CUDA_TRY(c, cudaMemsetAsync(c->d_red, 0, sizeof(double), c->stream));
    if (order == 4) o4_cfl_kernel<TF, false><<<c->grd_interior(), c->blk(), 0, c->stream>>>(P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), g, c->d_red);
    else if (order == 41) o4_cfl_kernel<TF, true><<<c->grd_interior(), c->blk(), 0, c->stream>>>(P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), g, c->d_red);
    else if (order == 24) adv2i_cfl_kernel<TF, 4><<<c->grd_interior(), c->blk(), 0, c->stream>>>(P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), g, c->d_red);
    else if (order == 262) adv2i_cfl_kernel<TF, 6><<<c->grd_interior(), c->blk(), 0, c->stream>>>(P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), g, c->d_red);
    else o2_cfl_kernel<TF><<<c->grd_interior(), c->blk(), 0, c->stream>>>(P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), g, c->d_red);
    const char* cfl_name = order == 2 ? "o2_cfl_kernel" : (order == 24 || order == 262) ? "adv2i_cfl_kernel" : "o4_cfl_kernel";
    KCHECKN(c, cfl_name);
    if (c->nranks > 1)
    {
        if (!c->comm) { c->err = "slab context without communicator: call mhh_comm_init first"; return MHH_E_INVALID; }
        NcclApi* api = nccl_api(c->err);
        if (!api) return MHH_E_CUDA;
        NCCL_TRY(c, api, api->AllReduce(c->d_red, c->d_red, 1, ncclFloat64, ncclMax, c->comm, c->stream));
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->h_red, c->d_red, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *out = *c->h_red;
    return MHH_OK;
}

template <typename TF, int MODE>
int reduce_impl(Ctx<TF>* c, const TF* u, const TF* v, const TF* w, TF p0, TF p1, TF p2, double* out)
{
    const GridDev<TF>& g = c->g;
    CUDA_TRY(c, cudaMemsetAsync(c->d_red, 0, sizeof(double), c->stream));
    reduce_kernel<TF, MODE><<<c->grd_interior(), c->blk(), 0, c->stream>>>(u, v, w, g, p0, p1, p2, c->d_red);
    KCHECKN(c, "reduce_kernel");
    if (c->nranks > 1)
    {
        if (!c->comm) { c->err = "slab context without communicator: call mhh_comm_init first"; return MHH_E_INVALID; }
        NcclApi* api = nccl_api(c->err);
        if (!api) return MHH_E_CUDA;
        NCCL_TRY(c, api, api->AllReduce(c->d_red, c->d_red, 1, ncclFloat64, ncclMax, c->comm, c->stream));
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->h_red, c->d_red, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *out = *c->h_red;
    return MHH_OK;
}

template <typename TF>
int reduce_mode_impl(Ctx<TF>* c, int mode, const TF* u, const TF* v, const TF* w, TF p0, TF p1, TF p2, double* out)
{
    if (mode == 0) return reduce_impl<TF, 0>(c, u, v, w, p0, p1, p2, out);
    if (mode == 1) return reduce_impl<TF, 1>(c, u, v, w, p0, p1, p2, out);
    return reduce_impl<TF, 2>(c, u, v, w, p0, p1, p2, out);
}

#define INST(TF) \
    template int check_mom<TF>(Ctx<TF>*, const mhh_fields*, bool, bool); \
    template int evisc_impl<TF>(Ctx<TF>*, const mhh_fields*, const mhh_params*, const TF*); \
    template int tend_impl<TF>(Ctx<TF>*, const mhh_fields*, const mhh_params*, bool, bool, bool, const mhh_tke2*, bool); \
    template int o2_impl<TF>(Ctx<TF>*, const mhh_fields*, bool, bool, bool); template int o4_impl<TF>(Ctx<TF>*, const mhh_fields*, int, bool); \
    template int o2_cfl_impl<TF>(Ctx<TF>*, const mhh_fields*, double*, int); \
    template int adv2i_impl<TF>(Ctx<TF>*, const mhh_fields*, int); \
    template int reduce_mode_impl<TF>(Ctx<TF>*, int, const TF*, const TF*, const TF*, TF, TF, TF, double*);

INST(double)
INST(float)
#undef INST

} // namespace mhhhost
