// mhh_adapters.hpp -- drop-in adapters between MicroHH's scheme classes and libmhhb200's C ABI.
//
// This header is compiled INSIDE the MicroHH source tree (it includes MicroHH's own headers; none of
// them is copied here).  Each adapter derives from the reference's abstract scheme class, keeps the
// reference's constructor / virtual signatures, and forwards the device work to the C entry points
// of include/mhhb200.h.  The factories (src/advec.cxx:54-98, src/diff.cxx:56-92, src/pres.cxx:67-86)
// return these classes for the same .ini switches, so `microhh init/run`, the .ini/.nc case files and
// the Field3d ghost-cell layout stay unchanged.  See INTEGRATION.md for the three factory edits.
//
//   reference class (interface)                     adapter                      C ABI
//   Advec<TF>        include/advec.h:45-71          Advec_b200<TF, SW, TYPE>     mhh_advec_exec / mhh_advec_get_cfl (2i5, 2, 2i4, 2i62, 4, 4m)
//   Diff<TF>         include/diff.h:38-71           Diff_smag2_b200<TF>          mhh_diff_smag2_exec_viscosity / _exec / _get_dn
//                                                   Diff_{2,4}_b200<TF>          mhh_diff_2_exec / mhh_diff_4_exec / mhh_diff_2_get_dn
//   Pres<TF>         include/pres.h:41-92           Pres_{2,4}_b200<TF>          mhh_pres_exec / mhh_pres_check_divergence
//   Boundary_cyclic  include/boundary_cyclic.h:35   Boundary_cyclic_b200<TF>     mhh_boundary_cyclic / _2d
//   Timeloop::exec   include/timeloop.h:65          timeloop_exec_b200()         mhh_timeloop_rk3
//   Model::exec loop src/model.cxx:356-504          dycore_substep_b200()        mhh_dycore_substep (fused fast path)
//   Thermo_buoy<TF>  include/thermo_buoy.h:47       Thermo_buoy_b200<TF>         mhh_thermo_buoy_exec / _n2, mhh_dycore_set_thermo_buoy
//   Thermo_moist<TF> include/thermo_moist.h:54      Thermo_moist_b200<TF>        mhh_thermo_moist_exec / _set_profiles / _get_profiles /
//                                                                                _get_thermo_field, mhh_dycore_set_thermo_moist
//
// Errors: the C ABI never throws; MHH_CHECK rethrows as std::runtime_error, which main() already
// turns into "message + exit code 1" (main/microhh.cxx:59-68).
#pragma once

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "mhhb200.h"

#include "master.h"
#include "grid.h"
#include "fields.h"
#include "advec.h"
#include "diff.h"
#include "pres.h"
#include "boundary.h"
#include "boundary_cyclic.h"
#include "thermo.h"
#include "thermo_buoy.h"
#include "thermo_moist.h"
#include "stats.h"
#include "constants.h"

namespace mhhb200
{
    inline void check(mhh_ctx* ctx, const int rc, const char* what)
    {
        if (rc != MHH_OK)
            throw std::runtime_error(std::string("mhhb200: ") + what + ": " + mhh_last_error(ctx));
    }
    #define MHH_CHECK(ctx, call) ::mhhb200::check((ctx), (call), #call)

    template<typename TF> constexpr int dtype_of() { return sizeof(TF) == 8 ? MHH_F64 : MHH_F32; }

    // One context per process (= per GPU), shared by all adapters of a Model.  Created after Grid::create
    // and Fields::create, because the metric arrays and the base state are inputs.
    template<typename TF>
    class Context
    {
        public:
            Context(Master& master, Grid<TF>& grid, Fields<TF>& fields, const int device=0)
            {
                const Grid_data<TF>& gd = grid.get_grid_data();
                const MPI_data& md = master.get_MPI_data();
                mhh_grid_desc d{};
                d.itot = gd.itot; d.jtot = gd.jtot; d.ktot = gd.ktot;
                d.imax = gd.imax; d.jmax = gd.jmax; d.kmax = gd.kmax;
                d.igc = gd.igc; d.jgc = gd.jgc; d.kgc = gd.kgc;
                d.xsize = gd.xsize; d.ysize = gd.ysize; d.zsize = gd.zsize;
                d.z = gd.z.data(); d.zh = gd.zh.data(); d.dz = gd.dz.data(); d.dzh = gd.dzh.data();
                d.dzi = gd.dzi.data(); d.dzhi = gd.dzhi.data();
                d.npx = md.npx; d.npy = md.npy; d.mpicoordx = md.mpicoordx; d.mpicoordy = md.mpicoordy;
                if (grid.get_spatial_order() == Grid_order::Fourth) { d.dzi4 = gd.dzi4.data(); d.dzhi4 = gd.dzhi4.data(); }
                const int rc = mhh_ctx_create(&d, dtype_of<TF>(), device, &ctx);
                if (rc != MHH_OK)
                {
                    const std::string msg = mhh_last_error(ctx);
                    mhh_ctx_destroy(ctx);
                    throw std::runtime_error("mhhb200: mhh_ctx_create: " + msg);
                }
                // MicroHH drives everything on the legacy default stream
                MHH_CHECK(ctx, mhh_set_stream(ctx, nullptr));
                MHH_CHECK(ctx, mhh_set_basestate(ctx, fields.rhoref.data(), fields.rhorefh.data(), nullptr, nullptr));

                #ifdef USEMPI
                // y slabs: rank 0 makes the NCCL id, MPI broadcasts it (Master owns the communicator)
                if (md.npy > 1)
                {
                    unsigned char id[MHH_COMM_ID_BYTES] = {};
                    if (md.mpiid == 0) MHH_CHECK(ctx, mhh_comm_get_unique_id(id, MHH_COMM_ID_BYTES));
                    master.broadcast(reinterpret_cast<char*>(id), MHH_COMM_ID_BYTES);
                    MHH_CHECK(ctx, mhh_comm_init(ctx, id, MHH_COMM_ID_BYTES));
                    // fused transposes / ghost rows over NVLink peer memory: all-gather the CUDA IPC handles in rank order
                    std::vector<unsigned char> mine(MHH_IPC_BYTES), all((size_t)MHH_IPC_BYTES * md.nprocs);
                    MHH_CHECK(ctx, mhh_comm_get_ipc_handles(ctx, mine.data(), MHH_IPC_BYTES));
                    MPI_Allgather(mine.data(), MHH_IPC_BYTES, MPI_BYTE, all.data(), MHH_IPC_BYTES, MPI_BYTE, md.commxy);
                    // the transport is a collective property: if the mapping failed on ANY rank, every rank goes back to NCCL
                    // (a throw on one rank alone would leave the others hanging in their next collective)
                    int ok = mhh_comm_open_peers(ctx, all.data(), (int)all.size()) == MHH_OK ? 1 : 0, all_ok = 0;
                    MPI_Allreduce(&ok, &all_ok, 1, MPI_INT, MPI_MIN, md.commxy);
                    if (!all_ok)
                        MHH_CHECK(ctx, mhh_comm_disable_peers(ctx));
                }
                #endif
            }
            ~Context() { mhh_ctx_destroy(ctx); }
            Context(const Context&) = delete;
            Context& operator=(const Context&) = delete;

            // Thermo_dry's reference profiles are only known after Thermo::create
            void set_thermo_basestate(Fields<TF>& fields, const std::vector<TF>& thref, const std::vector<TF>& threfh)
            { MHH_CHECK(ctx, mhh_set_basestate(ctx, fields.rhoref.data(), fields.rhorefh.data(), thref.data(), threfh.data())); }

            mhh_ctx* ctx = nullptr;
    };

    // Device pointers of the Fields maps in the C ABI's POD (borrowed per call, never owned).
    template<typename TF>
    mhh_fields fields_view(Fields<TF>& fields, Boundary<TF>* boundary=nullptr, const std::vector<std::string>* fluxlimit_list=nullptr)
    {
        mhh_fields f{};
        f.u = fields.mp.at("u")->fld_g;   f.v = fields.mp.at("v")->fld_g;   f.w = fields.mp.at("w")->fld_g;
        f.ut = fields.mt.at("u")->fld_g;  f.vt = fields.mt.at("v")->fld_g;  f.wt = fields.mt.at("w")->fld_g;
        if (fields.sd.count("evisc")) f.evisc = fields.sd.at("evisc")->fld_g;
        if (fields.sd.count("p"))     f.p = fields.sd.at("p")->fld_g;
        f.visc = fields.visc;
        f.u_fluxbot = fields.mp.at("u")->flux_bot_g; f.u_fluxtop = fields.mp.at("u")->flux_top_g;
        f.v_fluxbot = fields.mp.at("v")->flux_bot_g; f.v_fluxtop = fields.mp.at("v")->flux_top_g;
        f.u_bot = fields.mp.at("u")->fld_bot_g; f.u_gradbot = fields.mp.at("u")->grad_bot_g;
        f.u_top = fields.mp.at("u")->fld_top_g; f.u_gradtop = fields.mp.at("u")->grad_top_g;
        f.v_bot = fields.mp.at("v")->fld_bot_g; f.v_gradbot = fields.mp.at("v")->grad_bot_g;
        f.v_top = fields.mp.at("v")->fld_top_g; f.v_gradtop = fields.mp.at("v")->grad_top_g;
        int n = 0;
        // the thermodynamic scalar must be scalar 0 (buoyancy / N2 source; Thermo_buoy's "b"): std::map order puts "th" / "thl"
        // after e.g. "qt", so it is moved to the front explicitly
        std::vector<std::string> names;
        for (auto& it : fields.sp) names.push_back(it.first);
        for (const char* thname : {"th", "thl", "b"})
            for (size_t i = 0; i < names.size(); ++i)
                if (names[i] == thname) { std::swap(names[0], names[i]); }
        for (const std::string& name : names)
        {
            if (n == MHH_MAX_SCALARS) throw std::runtime_error("mhhb200: more than MHH_MAX_SCALARS prognostic scalars");
            auto& s = fields.sp.at(name);
            f.s[n] = s->fld_g; f.st[n] = fields.st.at(name)->fld_g; f.svisc[n] = s->visc;
            f.s_fluxbot[n] = s->flux_bot_g; f.s_fluxtop[n] = s->flux_top_g;
            f.s_bot[n] = s->fld_bot_g; f.s_gradbot[n] = s->grad_bot_g;
            f.s_top[n] = s->fld_top_g; f.s_gradtop[n] = s->grad_top_g;
            if (fluxlimit_list)
                for (const std::string& lim : *fluxlimit_list) if (lim == name) f.s_fluxlimit[n] = 1;
            ++n;
        }
        f.ns = n;
        if (boundary && boundary->get_switch() != "default")
        {
            f.dudz_mo = boundary->get_dudz_g(); f.dvdz_mo = boundary->get_dvdz_g();
            f.dbdz_mo = boundary->get_dbdz_g(); f.z0m = boundary->get_z0m_g();
        }
        return f;
    }

    // ---- Advec_2i5 (src/advec_2i5.cxx:955-1063), Advec_2 (src/advec_2.cxx:288-345), Advec_4 (src/advec_4.cxx:573-684):
    // Advec_4m (src/advec_4m.cxx:511-615), Advec_2i4 (src/advec_2i4.cxx:664-738), Advec_2i62 (src/advec_2i62.cxx:397-480).
    // SW = the C ABI's swadvec code (25, 2, 24, 262, 4, 41)
    template<typename TF, int SW, Advection_type TYPE>
    class Advec_b200 : public Advec<TF>
    {
        public:
            Advec_b200(Master& m, Grid<TF>& g, Fields<TF>& f, Input& in, std::shared_ptr<Context<TF>> c) :
                Advec<TF>(m, g, f, in), c(std::move(c))
            {
                // ghost cells as the reference constructors ask for them (src/advec_2i5.cxx:42-45, src/advec_2.cxx:40-43,
                // src/advec_4.cxx:41-48)
                if (SW == 25 || SW == 262) fluxlimit_list = in.get_list<std::string>("advec", "fluxlimit_list", "", std::vector<std::string>());   // src/advec_2i5.cxx:39-40, src/advec_2i62.cxx:39-40
                // :42-46 asks for (3, 3, 1|2); FOUR ghost cells in x make the interior start at a 16-byte aligned element (fp64 and
                // fp32) and, for USESP, the row pitch a multiple of 16 bytes: the TMA-staged kernels then use aligned vector
                // accesses throughout (with 3 the fp64 path still works, ~3 % slower; fp32 falls back to the cp.async kernels)
                if (SW == 25) g.set_minimum_ghost_cells(4, 3, fluxlimit_list.empty() ? 1 : 2);
                else if (SW == 2) g.set_minimum_ghost_cells(1, 1, 1);
                else if (SW == 24) g.set_minimum_ghost_cells(2, 2, 2);                                        // src/advec_2i4.cxx:38-41
                else if (SW == 262) g.set_minimum_ghost_cells(3, 3, fluxlimit_list.empty() ? 1 : 2);          // src/advec_2i62.cxx:42-45
                else g.set_minimum_ghost_cells(3, 3, 3);
            }

            void create(Stats<TF>&) override {}
            void exec(Stats<TF>&) override
            {
                const mhh_fields f = fields_view(this->fields, static_cast<Boundary<TF>*>(nullptr), &fluxlimit_list);
                MHH_CHECK(c->ctx, mhh_advec_exec(c->ctx, SW, &f));
            }
            double get_cfl(double dt) override
            {
                const mhh_fields f = fields_view(this->fields);
                double cfl = 0.;
                MHH_CHECK(c->ctx, mhh_advec_get_cfl(c->ctx, SW, &f, dt, &cfl));
                return cfl;
            }
            unsigned long get_time_limit(unsigned long idt, double dt) override
            {
                // src/advec_2i5.cxx:984-992
                double cfl = get_cfl(dt);
                cfl = std::max(this->cflmin, cfl);
                return idt * this->cflmax / cfl;
            }
            void get_advec_flux(Field3d<TF>&, const Field3d<TF>&) override
            { throw std::runtime_error("mhhb200: get_advec_flux is a statistics path (out of scope)"); }
            Advection_type get_switch() const override { return TYPE; }

        private:
            std::shared_ptr<Context<TF>> c;
            std::vector<std::string> fluxlimit_list;
    };
    template<typename TF> using Advec_2i5_b200 = Advec_b200<TF, 25, Advection_type::Advec_2i5>;
    template<typename TF> using Advec_2_b200   = Advec_b200<TF, 2,  Advection_type::Advec_2>;
    template<typename TF> using Advec_4_b200   = Advec_b200<TF, 4,  Advection_type::Advec_4>;
    template<typename TF> using Advec_4m_b200  = Advec_b200<TF, 41, Advection_type::Advec_4m>;
    template<typename TF> using Advec_2i4_b200  = Advec_b200<TF, 24,  Advection_type::Advec_2i4>;
    template<typename TF> using Advec_2i62_b200 = Advec_b200<TF, 262, Advection_type::Advec_2i62>;

    // ---- Diff_smag2 (src/diff_smag2.cxx:312-607) ------------------------------------------------
    template<typename TF>
    class Diff_smag2_b200 : public Diff<TF>
    {
        public:
            Diff_smag2_b200(Master& m, Grid<TF>& g, Fields<TF>& f, Boundary<TF>& b, Input& in, std::shared_ptr<Context<TF>> c) :
                Diff<TF>(m, g, f, b, in), c(std::move(c))
            {
                // same .ini keys as src/diff_smag2.cxx:275-290
                dnmax = in.get_item<double>("diff", "dnmax", "", 0.4);
                prm.cs = in.get_item<TF>("diff", "cs", "", 0.23);
                prm.tPr = in.get_item<TF>("diff", "tPr", "", 1./3.);
                prm.sw_mason = in.get_item<bool>("diff", "swmason", "", true);
                this->tPr = prm.tPr;
                prm.swadvec = 25; prm.swdiff = 1;
                f.init_diagnostic_field("evisc", "Eddy viscosity", "m2 s-1", "thermo", g.get_grid_data().sloc);
            }
            Diffusion_type get_switch() const override { return Diffusion_type::Diff_smag2; }
            void init() override {}
            void create(Stats<TF>&, const bool) override {}
            void exec_viscosity(Stats<TF>&, Thermo<TF>& thermo) override
            {
                prm.surface_model = this->boundary.get_switch() != "default";
                // three ways, as the reference dispatches on thermo.get_switch() (src/diff_smag2.cu:131,181):
                //   Disabled -> calc_evisc_neutral (no N2 at all: Thermo_disabled::get_thermo_field_g throws, include/thermo_disabled.h:81)
                //   Dry      -> N2 derived from th inside the fused kernel
                //   other    -> N2 fetched from the thermo class
                const Thermo_type sw = thermo.get_switch();
                prm.swthermo = sw == Thermo_type::Dry ? 1 : 0;
                const mhh_fields f = fields_view(this->fields, &this->boundary);
                if (sw == Thermo_type::Disabled || sw == Thermo_type::Dry)
                    MHH_CHECK(c->ctx, mhh_diff_smag2_exec_viscosity(c->ctx, &f, &prm, nullptr));
                else
                {
                    auto n2 = this->fields.get_tmp_g();
                    thermo.get_thermo_field_g(*n2, "N2", false);
                    const int rc = mhh_diff_smag2_exec_viscosity(c->ctx, &f, &prm, n2->fld_g);
                    this->fields.release_tmp_g(n2);
                    MHH_CHECK(c->ctx, rc);
                }
            }
            void exec(Stats<TF>&) override
            {
                const mhh_fields f = fields_view(this->fields, &this->boundary);
                MHH_CHECK(c->ctx, mhh_diff_smag2_exec(c->ctx, &f, &prm));
            }
            void exec_stats(Stats<TF>&, Thermo<TF>&) override {}
            void diff_flux(Field3d<TF>&, const Field3d<TF>&) override
            { throw std::runtime_error("mhhb200: diff_flux is a statistics path (out of scope)"); }
            double get_dn(double dt) override
            {
                const mhh_fields f = fields_view(this->fields);
                double dn = 0.;
                MHH_CHECK(c->ctx, mhh_diff_smag2_get_dn(c->ctx, &f, &prm, dt, &dn));
                return dn;
            }
            unsigned long get_time_limit(unsigned long idt, double dt) override
            {
                // src/diff_smag2.cxx:312-331
                double dn = get_dn(dt);
                dn = std::max(Constants::dsmall, dn);
                return idt * dnmax / dn;
            }
            void prepare_device(Boundary<TF>&) override {}
            void clear_device() override {}

        private:
            std::shared_ptr<Context<TF>> c;
            mhh_params prm{};
            double dnmax;
    };

    // ---- Diff_tke2 (src/diff_tke2.cxx:514-983): Deardorff SGS-TKE closure.  Same constructor work as the reference (ini keys,
    // the prognostic "sgstke", the diagnostics "evisc" / "eviscs"); the limiter on sgstke (Limiter<TF>::exec,
    // src/limiter.cxx:117-129) is limiter_exec_b200 below.
    template<typename TF>
    class Diff_tke2_b200 : public Diff<TF>
    {
        public:
            Diff_tke2_b200(Master& m, Grid<TF>& g, Fields<TF>& f, Boundary<TF>& b, Input& in, std::shared_ptr<Context<TF>> c) :
                Diff<TF>(m, g, f, b, in), c(std::move(c))
            {
                // same .ini keys as src/diff_tke2.cxx:522-557
                dnmax = in.get_item<double>("diff", "dnmax", "", 0.4);
                tke.ap  = in.get_item<TF>("diff", "ap",  "", 1.5);  tke.cf  = in.get_item<TF>("diff", "cf",  "", 2.5);
                tke.ce1 = in.get_item<TF>("diff", "ce1", "", 0.19); tke.ce2 = in.get_item<TF>("diff", "ce2", "", 0.51);
                tke.cm  = in.get_item<TF>("diff", "cm",  "", 0.12); tke.ch1 = in.get_item<TF>("diff", "ch1", "", 1.);
                tke.ch2 = in.get_item<TF>("diff", "ch2", "", 2.);   tke.cn  = in.get_item<TF>("diff", "cn",  "", 0.76);
                prm.sw_mason = in.get_item<bool>("diff", "swmason", "", true);
                sw_buoy = in.get_item<std::string>("thermo", "swthermo", "", "0") != "0";
                prm.swadvec = 25; prm.swdiff = 3; prm.tPr = 1.; prm.surface_model = 1;
                const std::string group_name = "sgstke";
                f.init_prognostic_field("sgstke", "SGS TKE", "m2 s-2", group_name, g.get_grid_data().sloc, false);
                f.sp.at("sgstke")->visc = in.get_item<TF>("fields", "svisc", "sgstke");
                f.init_diagnostic_field("evisc", "Eddy viscosity for momentum", "m2 s-1", group_name, g.get_grid_data().sloc);
                if (sw_buoy)
                    f.init_diagnostic_field("eviscs", "Eddy viscosity for scalars", "m2 s-1", group_name, g.get_grid_data().sloc);
                if (g.get_spatial_order() != Grid_order::Second)
                    throw std::runtime_error("Diff_tke2 only runs with second order grids.");
                if (b.get_switch() == "default")
                    throw std::runtime_error("Diff_tke2 does not support resolved walls.");
            }
            Diffusion_type get_switch() const override { return Diffusion_type::Diff_tke2; }
            void init() override {}
            void create(Stats<TF>&, const bool cold_start) override
            {
                if (cold_start)
                    MHH_CHECK(c->ctx, mhh_diff_tke2_create(c->ctx, this->fields.sp.at("sgstke")->fld_g));
            }
            // sgstke's position in the C ABI's scalar list follows fields_view's ordering
            const mhh_tke2& closure(const mhh_fields& f)
            {
                tke.isgstke = -1;
                const void* e = this->fields.sp.at("sgstke")->fld_g;
                for (int n = 0; n < f.ns; ++n)
                    if (f.s[n] == e) tke.isgstke = n;
                tke.eviscs = nullptr;
                if (sw_buoy) tke.eviscs = this->fields.sd.at("eviscs")->fld_g;
                return tke;
            }
            void exec_viscosity(Stats<TF>&, Thermo<TF>& thermo) override
            {
                const Thermo_type sw = thermo.get_switch();
                prm.swthermo = sw == Thermo_type::Disabled ? 0 : 1;
                const mhh_fields f = fields_view(this->fields, &this->boundary);
                if (sw == Thermo_type::Disabled || sw == Thermo_type::Dry)
                    MHH_CHECK(c->ctx, mhh_diff_tke2_exec_viscosity(c->ctx, &f, &prm, &closure(f), nullptr));
                else
                {
                    auto n2 = this->fields.get_tmp_g();
                    thermo.get_thermo_field_g(*n2, "N2", false);
                    const int rc = mhh_diff_tke2_exec_viscosity(c->ctx, &f, &prm, &closure(f), n2->fld_g);
                    this->fields.release_tmp_g(n2);
                    MHH_CHECK(c->ctx, rc);
                }
            }
            void exec(Stats<TF>&) override
            {
                const mhh_fields f = fields_view(this->fields, &this->boundary);
                MHH_CHECK(c->ctx, mhh_diff_tke2_exec(c->ctx, &f, &prm, &closure(f)));
            }
            void exec_stats(Stats<TF>&, Thermo<TF>&) override {}
            void diff_flux(Field3d<TF>&, const Field3d<TF>&) override
            { throw std::runtime_error("mhhb200: diff_flux is a statistics path (out of scope)"); }
            double get_dn(double dt) override
            {
                const mhh_fields f = fields_view(this->fields, &this->boundary);
                double dn = 0.;
                MHH_CHECK(c->ctx, mhh_diff_tke2_get_dn(c->ctx, &f, &prm, &closure(f), dt, &dn));
                return dn;
            }
            unsigned long get_time_limit(unsigned long idt, double dt) override
            { return idt * dnmax / std::max(Constants::dsmall, get_dn(dt)); }         // src/diff_tke2.cxx:580-608
            void prepare_device(Boundary<TF>&) override {}
            void clear_device() override {}

        private:
            std::shared_ptr<Context<TF>> c;
            mhh_params prm{};
            mhh_tke2 tke{};
            bool sw_buoy = false;
            double dnmax;
    };

    // ---- Limiter<TF>::exec (src/limiter.cu, CPU: src/limiter.cxx:98-130) on one prognostic field; for Diff_tke2:
    //   limiter_exec_b200(*b200, fields, "sgstke", Constants::sgstke_min<TF>, sub_dt)
    template<typename TF>
    void limiter_exec_b200(Context<TF>& c, Fields<TF>& fields, const std::string& name, const double min_value, const double sub_dt)
    { MHH_CHECK(c.ctx, mhh_limiter_exec(c.ctx, fields.at.at(name)->fld_g, fields.ap.at(name)->fld_g, min_value, sub_dt)); }

    // ---- Diff_2 (src/diff_2.cxx:120-190) and Diff_4 (src/diff_4.cxx:200-310): ORDER = 2 | 4 --------------------
    template<typename TF, int ORDER>
    class Diff_const_b200 : public Diff<TF>
    {
        public:
            Diff_const_b200(Master& m, Grid<TF>& g, Fields<TF>& f, Boundary<TF>& b, Input& in, std::shared_ptr<Context<TF>> c) :
                Diff<TF>(m, g, f, b, in), c(std::move(c))
            { dnmax = in.get_item<double>("diff", "dnmax", "", 0.4); }
            Diffusion_type get_switch() const override { return ORDER == 2 ? Diffusion_type::Diff_2 : Diffusion_type::Diff_4; }
            void init() override {}
            void create(Stats<TF>&, const bool) override {}
            void exec_viscosity(Stats<TF>&, Thermo<TF>&) override {}
            void exec(Stats<TF>&) override
            {
                const mhh_fields f = fields_view(this->fields);
                MHH_CHECK(c->ctx, ORDER == 2 ? mhh_diff_2_exec(c->ctx, &f) : mhh_diff_4_exec(c->ctx, &f));
            }
            void exec_stats(Stats<TF>&, Thermo<TF>&) override {}
            void diff_flux(Field3d<TF>&, const Field3d<TF>&) override
            { throw std::runtime_error("mhhb200: diff_flux is a statistics path (out of scope)"); }
            double get_dn(double dt) override
            {
                const mhh_fields f = fields_view(this->fields);
                double dn = 0.;
                MHH_CHECK(c->ctx, mhh_diff_2_get_dn(c->ctx, &f, dt, &dn));     // Diff_4 uses the same formula (src/diff_4.cxx:226-245)
                return dn;
            }
            unsigned long get_time_limit(unsigned long idt, double dt) override
            { return idt * dnmax / std::max(Constants::dsmall, get_dn(dt)); }         // src/diff_2.cxx:127-136
            void prepare_device(Boundary<TF>&) override {}
            void clear_device() override {}

        private:
            std::shared_ptr<Context<TF>> c;
            double dnmax;
    };
    template<typename TF> using Diff_2_b200 = Diff_const_b200<TF, 2>;
    template<typename TF> using Diff_4_b200 = Diff_const_b200<TF, 4>;

    // ---- Pres_2 (src/pres_2.cxx:66-105) and Pres_4 (src/pres_4.cxx:76-156): SW = swpres (2 | 4) -------------------
    template<typename TF, int SW>
    class Pres_b200 : public Pres<TF>
    {
        public:
            Pres_b200(Master& m, Grid<TF>& g, Fields<TF>& f, FFT<TF>& fft, Input& in, std::shared_ptr<Context<TF>> c) :
                Pres<TF>(m, g, f, fft, in), c(std::move(c)) {}
            void init() override {}
            void set_values() override {}       // the tables are built by mhh_set_basestate
            void create(Stats<TF>&) override {}
            void exec(double sub_dt, Stats<TF>&) override
            {
                const mhh_fields f = fields_view(this->fields);
                MHH_CHECK(c->ctx, mhh_pres_exec(c->ctx, SW, &f, sub_dt));
            }
            TF check_divergence() override
            {
                const mhh_fields f = fields_view(this->fields);
                double div = 0.;
                MHH_CHECK(c->ctx, mhh_pres_check_divergence(c->ctx, SW, &f, &div));
                return static_cast<TF>(div);
            }
            void prepare_device() override {}
            void clear_device() override {}

        private:
            std::shared_ptr<Context<TF>> c;
    };
    template<typename TF> using Pres_2_b200 = Pres_b200<TF, 2>;
    template<typename TF> using Pres_4_b200 = Pres_b200<TF, 4>;

    // ---- Boundary_cyclic::exec_g / exec_2d_g (src/boundary_cyclic.cu:98-128) --------------------
    template<typename TF>
    struct Boundary_cyclic_b200
    {
        std::shared_ptr<Context<TF>> c;
        void exec_g(TF* fld, Edge edge=Edge::Both_edges)
        {
            const int e = edge == Edge::East_west_edge ? MHH_EDGE_EAST_WEST : edge == Edge::North_south_edge ? MHH_EDGE_NORTH_SOUTH : MHH_EDGE_BOTH;
            MHH_CHECK(c->ctx, mhh_boundary_cyclic(c->ctx, fld, e));
        }
        void exec_2d_g(TF* fld) { MHH_CHECK(c->ctx, mhh_boundary_cyclic_2d(c->ctx, fld)); }
    };

    // ---- Timeloop::exec (src/timeloop.cu:82-112): rk3 of every prognostic (field, tendency) pair ---
    template<typename TF>
    void timeloop_exec_b200(Context<TF>& c, Fields<TF>& fields, const int substep, const double dt)
    {
        for (auto& it : fields.at)
            MHH_CHECK(c.ctx, mhh_timeloop_rk3(c.ctx, fields.ap.at(it.first)->fld_g, it.second->fld_g, substep, dt));
    }

    // ---- fused sub-step: replaces the boundary -> diff.exec_viscosity -> thermo/advec/diff.exec ->
    // pres.exec -> timeloop.exec sequence of Model::exec (src/model.cxx:356-504) by ONE call when the
    // case uses swadvec=2i5, swdiff=smag2, swpres=2 and thermo_dry (or no thermo).
    template<typename TF>
    void dycore_substep_b200(Context<TF>& c, Fields<TF>& fields, Boundary<TF>& boundary, const mhh_params& prm,
                             const int substep, const double dt)
    {
        const mhh_fields f = fields_view(fields, &boundary);
        MHH_CHECK(c.ctx, mhh_dycore_substep(c.ctx, &f, &prm, substep, dt));
    }

    // ---- "next" rows of the scope table: the bodies of three member functions whose classes are not polymorphic per scheme
    // (Boundary_surface / Buffer / Force keep their device arrays private), so the binding is a one-line body, not a subclass:
    //
    //   Boundary_surface<TF>::exec (src/boundary_surface.cu, CPU: src/boundary_surface.cxx:836-990), constant z0:
    //       mhhb200::surface_exec_b200(*b200, fields, *this, prm,
    //                                  {ustar_g, obuk_g, nobuk_g, z0m_g, z0h_g, dutot_tmp, {sbc codes}});
    //   Buffer<TF>::exec (src/buffer.cu, CPU: src/buffer.cxx:170-205)  and  Force<TF>::exec (src/force.cu, CPU: src/force.cxx:608-700):
    //       mhhb200::buffer_exec_b200(*b200, fields, forcing);   mhhb200::force_exec_b200(*b200, fields, forcing, sub_dt);
    //   with `forcing` filled once from the classes' own members (bufferprofs_g, ug_g, vg_g, lsprofs_g, wls_g, the [buffer] /
    //   [force] ini values) -- or registered with mhh_dycore_set_forcing so that the fused sub-step runs both in place.
    // mhh_boundary_surface_init (lookup table of the Obukhov solver, include/boundary_surface_kernels.h:78-138) is called once
    // from Boundary_surface<TF>::init_solver.
    template<typename TF>
    void surface_exec_b200(Context<TF>& c, Fields<TF>& fields, Boundary<TF>& boundary, const mhh_params& prm, const mhh_surface& s)
    {
        const mhh_fields f = fields_view(fields, &boundary);
        MHH_CHECK(c.ctx, mhh_boundary_surface_exec(c.ctx, &f, &prm, &s));
    }
    template<typename TF>
    void buffer_exec_b200(Context<TF>& c, Fields<TF>& fields, const mhh_forcing& forcing)
    {
        const mhh_fields f = fields_view(fields);
        MHH_CHECK(c.ctx, mhh_buffer_exec(c.ctx, &f, &forcing));
    }
    template<typename TF>
    void force_exec_b200(Context<TF>& c, Fields<TF>& fields, const mhh_forcing& forcing, const double sub_dt)
    {
        const mhh_fields f = fields_view(fields);
        MHH_CHECK(c.ctx, mhh_force_exec(c.ctx, &f, &forcing, sub_dt));
    }

    // ---- Field3d_io<TF>::save_field3d / load_field3d (src/field3d_io.cxx): restart IO straight from / into the device field, same
    // file layout; the bodies of Fields<TF>::save / load (src/fields.cxx:1243-1320) become, per prognostic field,
    //     nerror += mhhb200::field3d_save_b200(*b200, f.second->fld_g, filename, no_offset, gd.kstart, gd.kend);
    // Return value as the reference's: 0 = ok, 1 = failed (the message is in mhh_last_error).
    template<typename TF>
    int field3d_save_b200(Context<TF>& c, const TF* fld_g, const char* filename, const TF offset, const int kstart, const int kend)
    { return mhh_field3d_save(c.ctx, fld_g, filename, offset, kstart, kend) == MHH_OK ? 0 : 1; }
    template<typename TF>
    int field3d_load_b200(Context<TF>& c, TF* fld_g, const char* filename, const TF offset, const int kstart, const int kend)
    { return mhh_field3d_load(c.ctx, fld_g, filename, offset, kstart, kend) == MHH_OK ? 0 : 1; }

    // ---- staged sub-step: the fused path with MicroHH's own stages kept in between (surface model, statistics, ...), in the
    // order of Model::exec (src/model.cxx:368-504): pre = cyclic + ghost cells + diff.exec_viscosity; [MicroHH's surface model];
    // set_ghost_cells again (:401); post = thermo + advec + diff fused, (buffer, force if registered), pres, rk3.
    template<typename TF>
    void dycore_substep_pre_b200(Context<TF>& c, Fields<TF>& fields, Boundary<TF>& boundary, const mhh_params& prm)
    { const mhh_fields f = fields_view(fields, &boundary); MHH_CHECK(c.ctx, mhh_dycore_substep_pre(c.ctx, &f, &prm)); }
    template<typename TF>
    void dycore_set_ghost_cells_b200(Context<TF>& c, Fields<TF>& fields, Boundary<TF>& boundary, const mhh_params& prm)
    { const mhh_fields f = fields_view(fields, &boundary); MHH_CHECK(c.ctx, mhh_dycore_set_ghost_cells(c.ctx, &f, &prm)); }
    template<typename TF>
    void dycore_tendencies_b200(Context<TF>& c, Fields<TF>& fields, Boundary<TF>& boundary, const mhh_params& prm)
    { const mhh_fields f = fields_view(fields, &boundary); MHH_CHECK(c.ctx, mhh_dycore_tendencies(c.ctx, &f, &prm)); }
    template<typename TF>
    void dycore_substep_post_b200(Context<TF>& c, Fields<TF>& fields, Boundary<TF>& boundary, const mhh_params& prm,
                                  const int substep, const double dt)
    { const mhh_fields f = fields_view(fields, &boundary); MHH_CHECK(c.ctx, mhh_dycore_substep_post(c.ctx, &f, &prm, substep, dt)); }
    // index of a prognostic scalar in fields_view's ordering
    template<typename TF>
    int scalar_index(Fields<TF>& fields, const std::string& name)
    {
        std::vector<std::string> names;
        for (auto& it : fields.sp) names.push_back(it.first);
        for (const char* thname : {"th", "thl", "b"})
            for (size_t i = 0; i < names.size(); ++i)
                if (names[i] == thname) { std::swap(names[0], names[i]); }
        for (size_t i = 0; i < names.size(); ++i)
            if (names[i] == name) return (int)i;
        throw std::runtime_error("mhhb200: no prognostic scalar " + name);
    }

    // ---- Thermo_buoy (src/thermo_buoy.cxx:306-391): the reference class keeps everything but the device work.  `bs` is private
    // there, so the constructor reads the same .ini keys again (:318-329).  Thermo<TF>::factory (src/thermo.cxx) returns this class
    // for swthermo = buoy.
    template<typename TF>
    class Thermo_buoy_b200 : public Thermo_buoy<TF>
    {
        public:
            Thermo_buoy_b200(Master& m, Grid<TF>& g, Fields<TF>& f, Input& in, std::shared_ptr<Context<TF>> c) :
                Thermo_buoy<TF>(m, g, f, in), flds(f), c(std::move(c))
            {
                tb.alpha = in.get_item<TF>("thermo", "alpha", "", 0.);
                tb.n2 = in.get_item<TF>("thermo", "N2", "", 0.);
                tb.utrans = g.get_grid_data().utrans;
                tb.swbaroclinic = in.get_item<bool>("thermo", "swbaroclinic", "", false);
                tb.dbdy_ls = tb.swbaroclinic ? in.get_item<TF>("thermo", "dbdy_ls", "") : TF(0.);
            }
            void exec(const double, Stats<TF>& stats) override
            {
                const mhh_fields f = fields_view(flds);
                MHH_CHECK(c->ctx, mhh_thermo_buoy_exec(c->ctx, &f, &tb));
                stats.calc_tend(*flds.mt.at("w"), "buoy");
            }
            // run thermo.exec inside mhh_dycore_substep (prm.swthermo = 2) instead
            void register_fused() { MHH_CHECK(c->ctx, mhh_dycore_set_thermo_buoy(c->ctx, &tb)); }
            // get_thermo_field_g("N2") (src/thermo_buoy.cu): N2 into a device field
            void get_N2_g(TF* n2_g) { MHH_CHECK(c->ctx, mhh_thermo_buoy_n2(c->ctx, n2_g, flds.sp.at("b")->fld_g, tb.n2)); }
            mhh_thermo_buoy tb{};
        private:
            Fields<TF>& flds;          // the base classes re-declare Thermo<TF>::fields private
            std::shared_ptr<Context<TF>> c;
    };

    // ---- Thermo_moist (src/thermo_moist.cxx:1080-1447, GPU twin src/thermo_moist.cu:902-959): create_basestate / load stay the
    // reference's (host calc_base_state from the input profiles or the restart file); its eight profiles go to the context once
    // (get_basestate_vector is the public accessor).  exec then runs on the device INCLUDING the base-state update, where the
    // reference's GPU build copies two mean profiles to the host, integrates there and copies eight profiles back, every sub-step.
    // sync_host (default on) mirrors the updated profiles back into the reference object's host vectors and its own device copies
    // (forward_device) after every exec, for the parts of MicroHH that read them (statistics, radiation, microphysics); a run
    // that only needs them at output time can switch it off and call backward_basestate() there.
    template<typename TF>
    class Thermo_moist_b200 : public Thermo_moist<TF>
    {
        public:
            Thermo_moist_b200(Master& m, Grid<TF>& g, Fields<TF>& f, Input& in, const Sim_mode sim_mode, std::shared_ptr<Context<TF>> c) :
                Thermo_moist<TF>(m, g, f, in, sim_mode), flds(f), c(std::move(c))
            {
                tm.pbot = in.get_item<TF>("thermo", "pbot", "");
                tm.swupdatebasestate = in.get_item<bool>("thermo", "swupdatebasestate", "", true);
            }
            // call after create_basestate / load (the profiles exist) and Fields::create (the scalar order is final)
            void forward_basestate()
            {
                tm.ithl = scalar_index(flds, "thl"); tm.iqt = scalar_index(flds, "qt");
                auto p = [&](const char* n) { return static_cast<const void*>(this->get_basestate_vector(n).data()); };
                MHH_CHECK(c->ctx, mhh_thermo_moist_set_profiles(c->ctx, p("p"), p("ph"), p("rho"), p("rhoh"), p("thv"), p("thvh"), p("exner"), p("exnerh")));
                uploaded = true;
            }
            void backward_basestate()
            {
                auto p = [&](const char* n) { return static_cast<void*>(const_cast<TF*>(this->get_basestate_vector(n).data())); };
                MHH_CHECK(c->ctx, mhh_thermo_moist_get_profiles(c->ctx, p("p"), p("ph"), p("rho"), p("rhoh"), p("thv"), p("thvh"), p("exner"), p("exnerh")));
                this->forward_device();
            }
            void exec(const double, Stats<TF>& stats) override
            {
                if (!uploaded) forward_basestate();
                const mhh_fields f = fields_view(flds);
                MHH_CHECK(c->ctx, mhh_thermo_moist_exec(c->ctx, &f, &tm));
                if (tm.swupdatebasestate && sync_host) backward_basestate();
                long long bad = 0;
                if (check_convergence)
                {
                    MHH_CHECK(c->ctx, mhh_thermo_moist_nonconverged(c->ctx, &bad));
                    if (bad) throw std::runtime_error("Non-converging saturation adjustment");        // functions.h:252-263
                }
                stats.calc_tend(*flds.mt.at("w"), "buoy");
            }
            void register_fused()
            {
                if (!uploaded) forward_basestate();
                MHH_CHECK(c->ctx, mhh_dycore_set_thermo_moist(c->ctx, &tm));
            }
            // get_thermo_field_g for "b", "ql", "N2" (MHH_MOIST_*) into a device field
            void get_thermo_field_b200(TF* out_g, const int which)
            {
                const mhh_fields f = fields_view(flds);
                MHH_CHECK(c->ctx, mhh_thermo_moist_get_thermo_field(c->ctx, which, out_g, &f, &tm));
            }
            mhh_thermo_moist tm{};
            bool sync_host = true, check_convergence = true;
        private:
            Fields<TF>& flds;
            std::shared_ptr<Context<TF>> c;
            bool uploaded = false;
    };
}
