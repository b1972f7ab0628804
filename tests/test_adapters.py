"""The C++ adapter classes (microhh_b200/host/mhh_adapters.hpp) must compile against the reference's own
headers: every adapter template is explicitly instantiated with g++ -fsyntax-only.  Needs /root/reference
(this container); skipped on the GPU box."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

SRC = r'''
#include "microhh_b200/host/mhh_adapters.hpp"
#include "input.h"
template class mhhb200::Context<double>;
template class mhhb200::Context<float>;
template class mhhb200::Advec_b200<double, 25, Advection_type::Advec_2i5>;
template class mhhb200::Advec_b200<float, 2, Advection_type::Advec_2>;
template class mhhb200::Advec_b200<double, 4, Advection_type::Advec_4>;
template class mhhb200::Diff_smag2_b200<double>;
template class mhhb200::Diff_tke2_b200<double>;
template class mhhb200::Diff_tke2_b200<float>;
template void mhhb200::limiter_exec_b200<double>(mhhb200::Context<double>&, Fields<double>&, const std::string&, double, double);
template int mhhb200::field3d_save_b200<double>(mhhb200::Context<double>&, const double*, const char*, double, int, int);
template int mhhb200::field3d_load_b200<float>(mhhb200::Context<float>&, float*, const char*, float, int, int);
template class mhhb200::Diff_const_b200<double, 2>;
template class mhhb200::Diff_const_b200<float, 4>;
template class mhhb200::Pres_b200<float, 2>;
template class mhhb200::Pres_b200<double, 4>;
template struct mhhb200::Boundary_cyclic_b200<double>;
template void mhhb200::timeloop_exec_b200<double>(mhhb200::Context<double>&, Fields<double>&, int, double);
template void mhhb200::dycore_substep_b200<float>(mhhb200::Context<float>&, Fields<float>&, Boundary<float>&, const mhh_params&, int, double);
template class mhhb200::Advec_b200<double, 41, Advection_type::Advec_4m>;
template void mhhb200::surface_exec_b200<double>(mhhb200::Context<double>&, Fields<double>&, Boundary<double>&, const mhh_params&, const mhh_surface&);
template void mhhb200::buffer_exec_b200<float>(mhhb200::Context<float>&, Fields<float>&, const mhh_forcing&);
template void mhhb200::force_exec_b200<double>(mhhb200::Context<double>&, Fields<double>&, const mhh_forcing&, double);
template void mhhb200::dycore_substep_pre_b200<double>(mhhb200::Context<double>&, Fields<double>&, Boundary<double>&, const mhh_params&);
template void mhhb200::dycore_set_ghost_cells_b200<double>(mhhb200::Context<double>&, Fields<double>&, Boundary<double>&, const mhh_params&);
template void mhhb200::dycore_tendencies_b200<float>(mhhb200::Context<float>&, Fields<float>&, Boundary<float>&, const mhh_params&);
template void mhhb200::dycore_substep_post_b200<double>(mhhb200::Context<double>&, Fields<double>&, Boundary<double>&, const mhh_params&, int, double);
template class mhhb200::Advec_b200<double, 24, Advection_type::Advec_2i4>;
template class mhhb200::Advec_b200<float, 262, Advection_type::Advec_2i62>;
template class mhhb200::Thermo_buoy_b200<double>;
template class mhhb200::Thermo_buoy_b200<float>;
template class mhhb200::Thermo_moist_b200<double>;
template class mhhb200::Thermo_moist_b200<float>;
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "include")), reason="reference tree not mounted")
def test_adapters_instantiate_against_reference_headers():
    with tempfile.NamedTemporaryFile("w", suffix=".cpp", delete=False) as fh:
        fh.write(SRC)
        path = fh.name
    try:
        cmd = ["g++", "-std=c++20", "-fsyntax-only", "-DUSECUDA", "-DRESTRICTKEYWORD=__restrict__", "-w",
               f"-I{ROOT}", f"-I{ROOT}/include", f"-I{ROOT}/oracle/ref/shim", f"-I{REF}/include",
               "-I/usr/local/cuda/include", path]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-3000:]
    finally:
        os.unlink(path)
