// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
//
// A Pres_2<TF> / Pres_4<TF> object good enough to run the reference's own member functions init(), set_values(),
// input(), solve(), output() (and hdma) WITHOUT the Model around it.  The real constructors need Input, Fields (NetCDF,
// Stats, ...) -- far outside the path -- so the object is a zeroed memory image in which only what those member
// functions touch is brought to life:
//   * the four reference members of Pres<TF> (include/pres.h:61-64; stored as pointers right after the vptr):
//     Master (zeroed image: serial MPI coordinates 0), Grid (the reference's own Grid<TF> of ref_grid.cpp), Fields (zeroed image with rhoref / rhorefh constructed), FFT (a REAL FFT<TF> of ref_fft*.cpp);
//   * the std::vector members, placement-constructed.
// No virtual function is ever called and no constructor / destructor of the class runs.  Compiled with
// -fno-access-control so that the private members can be reached.
#pragma once
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include <stdexcept>
#include "ref_common.h"

template<typename P, typename TF>
P* fake_pres(void* fft, const TF* rhoref, const TF* rhorefh, int kcells)
{
    void* mbuf = ref_master_image();
    void* gbuf = ref_grid_image(sizeof(TF) == 4);
    char* fbuf = static_cast<char*>(std::calloc(1, sizeof(Fields<TF>)));
    Fields<TF>* F = reinterpret_cast<Fields<TF>*>(fbuf);
    new (&F->rhoref) std::vector<TF>(rhoref, rhoref + kcells);
    new (&F->rhorefh) std::vector<TF>(rhorefh, rhorefh + kcells);
    char* pbuf = static_cast<char*>(std::calloc(1, sizeof(P)));
    P* p = reinterpret_cast<P*>(pbuf);
    // the member that follows the four references pins their position
    if (reinterpret_cast<char*>(&p->field3d_operators) - pbuf != 5 * (std::ptrdiff_t)sizeof(void*))
        throw std::runtime_error("fake_pres: unexpected Pres<TF> layout");
    void** slot = reinterpret_cast<void**>(pbuf);
    slot[1] = mbuf; slot[2] = gbuf; slot[3] = fbuf; slot[4] = fft;
    return p;
}
template<typename TF> void vec_new(std::vector<TF>& v) { new (&v) std::vector<TF>(); }
template<typename TF> void vec_out(const std::vector<TF>& v, TF* out) { std::memcpy(out, v.data(), v.size() * sizeof(TF)); }
