// mhhb200 -- y-slab decomposition helpers: north/south ghost-row exchange staging.
//
// The reference has no multi-GPU path (CMakeLists.txt:50-52); its CPU-MPI code is the semantic
// model: Boundary_cyclic::exec with MPI neighbours (src/boundary_cyclic.cxx:115-176) sends the
// first/last `jgc` interior rows -- over the full ghosted width in x, all levels -- to the
// south/north neighbour after the east-west fill, which also completes the corners.
// Here x is never split, so east-west stays the local periodic copy; only these rows travel.
//
// Staging layout (one buffer per direction, all fields of a batch in one message):
//   [field][k = 0..kcells)[row = 0..w)[i = 0..icells)
#pragma once
#include "common.cuh"

namespace mhh {

constexpr int HALO_MAX_FIELDS = 3 + MAX_SCALARS;

template <typename TF>
struct HaloFields
{
    TF* f[HALO_MAX_FIELDS];
    int nf;
};

// sendS <- rows [jstart, jstart+w)  (become the south neighbour's north ghost rows)
// sendN <- rows [jend-w, jend)      (become the north neighbour's south ghost rows)
template <typename TF>
__global__ void __launch_bounds__(256) halo_pack_kernel(const HaloFields<TF> h, const GridDev<TF> g, const int w,
        TF* __restrict__ sendS, TF* __restrict__ sendN)
{
    const long long per = (long long)w * g.icells * g.kcells;
    const long long n = per * h.nf;
    const int rowlen = w * g.icells;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    {
        const int f = (int)(e / per);
        const long long r = e - f * per;
        const int k = (int)(r / rowlen);
        const int r2 = (int)(r - (long long)k * rowlen);         // row*icells + i: contiguous in the field too
        const TF* __restrict__ a = h.f[f];
        const long long lev = (long long)k * g.ijcells;
        sendS[e] = a[lev + (long long)g.jstart * g.icells + r2];
        sendN[e] = a[lev + (long long)(g.jend - w) * g.icells + r2];
    }
}

// north ghost rows [jend, jend+w) <- recvN (the north neighbour's sendS);  south ghost rows [jstart-w, jstart) <- recvS
template <typename TF>
__global__ void __launch_bounds__(256) halo_unpack_kernel(const HaloFields<TF> h, const GridDev<TF> g, const int w,
        const TF* __restrict__ recvN, const TF* __restrict__ recvS)
{
    const long long per = (long long)w * g.icells * g.kcells;
    const long long n = per * h.nf;
    const int rowlen = w * g.icells;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    {
        const int f = (int)(e / per);
        const long long r = e - f * per;
        const int k = (int)(r / rowlen);
        const int r2 = (int)(r - (long long)k * rowlen);
        TF* __restrict__ a = h.f[f];
        const long long lev = (long long)k * g.ijcells;
        a[lev + (long long)g.jend * g.icells + r2] = recvN[e];
        a[lev + (long long)(g.jstart - w) * g.icells + r2] = recvS[e];
    }
}

// Peer-memory variant of the pack: the strips are stored straight into the neighbours' receive buffers over NVLink
// (dstN = the SOUTH neighbour's "from north" buffer gets my first rows, dstS = the NORTH neighbour's "from south" buffer
// gets my last rows); no send staging, no NCCL transfer.
template <typename TF>
__global__ void __launch_bounds__(256) halo_push_kernel(const HaloFields<TF> h, const GridDev<TF> g, const int w,
        TF* __restrict__ south_recvN, TF* __restrict__ north_recvS)
{
    const long long per = (long long)w * g.icells * g.kcells;
    const long long n = per * h.nf;
    const int rowlen = w * g.icells;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    {
        const int f = (int)(e / per);
        const long long r = e - f * per;
        const int k = (int)(r / rowlen);
        const int r2 = (int)(r - (long long)k * rowlen);
        const TF* __restrict__ a = h.f[f];
        const long long lev = (long long)k * g.ijcells;
        south_recvN[e] = a[lev + (long long)g.jstart * g.icells + r2];
        north_recvS[e] = a[lev + (long long)(g.jend - w) * g.icells + r2];
    }
    __threadfence_system();
}

} // namespace mhh
