// mhhb200 -- restart IO of one 3-D field: Field3d_io<TF>::save_field3d / load_field3d of the reference (serial build
// src/field3d_io.cxx:669-751; MPI build, same single-file layout through MPI-IO subarrays, :57-160).  The file is the interior
// [kstart, kend) x jtot x itot as raw TF, no header -- unchanged, so restart files move freely between MicroHH and this
// library.  The field stays on the device: a kernel packs the interior rows (+ offset) of a batch of levels, the batch goes to
// pinned host memory and from there to its place in the file with pwrite; a y slab writes / reads rows
// [mpicoordy*jmax, (mpicoordy+1)*jmax) of every level of the same file (what the MPI subarray view does), so N ranks produce
// one file without gathering the field anywhere.
#include "host_common.cuh"
#include <fcntl.h>
#include <unistd.h>
#include <cerrno>

namespace mhh {

// compact[(k - k0) * jmax * imax + j * imax + i] = fld[interior] + offset  (SAVE)  |  fld[interior] = compact - offset  (LOAD)
template <typename TF, bool SAVE>
__global__ void __launch_bounds__(256) io_pack_kernel(TF* __restrict__ compact, TF* __restrict__ fld, const TF offset, const GridDev<TF> g,
        const int k0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int kl = blockIdx.z;
    if (i >= g.imax || j >= g.jmax) return;
    const long long ijk = (i + g.istart) + (long long)(j + g.jstart) * g.icells + (long long)(k0 + kl) * g.ijcells;
    const long long n = i + (long long)j * g.imax + (long long)kl * g.imax * g.jmax;
    if (SAVE) compact[n] = fld[ijk] + offset;
    else fld[ijk] = compact[n] - offset;
}

} // namespace mhh

namespace mhhhost {

template <typename TF>
int io_buffers(Ctx<TF>* c, size_t bytes)
{
    if (c->io_cap >= bytes) return MHH_OK;
    if (c->io_dev) cudaFree(c->io_dev);
    if (c->io_host) cudaFreeHost(c->io_host);
    c->io_dev = nullptr; c->io_host = nullptr; c->io_cap = 0;
    CUDA_TRY(c, cudaMalloc(&c->io_dev, bytes));
    CUDA_TRY(c, cudaMallocHost(&c->io_host, bytes));
    c->io_cap = bytes;
    return MHH_OK;
}

static bool full_io(int fd, char* buf, size_t n, off_t off, bool write)
{
    while (n > 0)
    {
        const ssize_t r = write ? pwrite(fd, buf, n, off) : pread(fd, buf, n, off);
        if (r < 0 && errno == EINTR) continue;
        if (r <= 0) return false;          // error, or a file shorter than the field
        buf += r; n -= (size_t)r; off += r;
    }
    return true;
}

template <typename TF, bool SAVE>
int field3d_io_impl(Ctx<TF>* c, TF* fld, const char* filename, double offset, int kstart, int kend)
{
    const GridDev<TF>& g = c->g;
    NEED(c, fld, "field"); NEED(c, filename, "filename");
    if (kstart < 0 || kend > g.kcells || kend <= kstart) { c->err = "field3d io: bad level range"; return MHH_E_INVALID; }
    const int npy = c->desc.npy > 1 ? c->desc.npy : 1, ry = npy > 1 ? c->desc.mpicoordy : 0;
    // serial build: fopen(filename, "wbx") -- an existing file is an error; slabs: every rank opens the one file
    const int flags = SAVE ? (O_WRONLY | O_CREAT | (npy == 1 ? O_EXCL : 0)) : O_RDONLY;
    const int fd = open(filename, flags, 0644);
    if (fd < 0) { c->err = std::string(SAVE ? "save_field3d: cannot create " : "load_field3d: cannot open ") + filename + ": " + strerror(errno); return MHH_E_IO; }
    const size_t level = (size_t)g.imax * g.jmax * sizeof(TF);                       // this rank's rows of one level
    const size_t glevel = (size_t)g.itot * g.jtot * sizeof(TF);                      // one level of the file
    const int nk = kend - kstart;
    const int batch = (int)std::max<size_t>(1, std::min<size_t>((size_t)nk, ((size_t)256 << 20) / level));
    int rc = io_buffers<TF>(c, (size_t)batch * level);
    if (rc != MHH_OK) { close(fd); return rc; }
    TF* d = static_cast<TF*>(c->io_dev); char* h = static_cast<char*>(c->io_host);
    for (int k0 = 0; k0 < nk && rc == MHH_OK; k0 += batch)
    {
        const int nb = std::min(batch, nk - k0);
        dim3 b(64, 4), gr((g.imax + 63) / 64, (g.jmax + 3) / 4, nb);
        auto file_io = [&]() -> bool {
            if (npy == 1) return full_io(fd, h, (size_t)nb * level, (off_t)((size_t)k0 * glevel), SAVE);
            for (int kl = 0; kl < nb; ++kl)
                if (!full_io(fd, h + (size_t)kl * level, level, (off_t)((size_t)(k0 + kl) * glevel + (size_t)ry * level), SAVE)) return false;
            return true; };
        if (SAVE)
        {
            io_pack_kernel<TF, true><<<gr, b, 0, c->stream>>>(d, fld, (TF)offset, g, kstart + k0);
            c->launches++;
            cudaError_t e = cudaMemcpyAsync(h, d, (size_t)nb * level, cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) { c->err = std::string("save_field3d: ") + cudaGetErrorString(e); rc = MHH_E_CUDA; break; }
            if (!file_io()) { c->err = std::string("save_field3d: write failed: ") + strerror(errno); rc = MHH_E_IO; }
        }
        else
        {
            if (!file_io()) { c->err = std::string("load_field3d: short read from ") + filename; rc = MHH_E_IO; break; }
            cudaError_t e = cudaMemcpyAsync(d, h, (size_t)nb * level, cudaMemcpyHostToDevice, c->stream);
            if (e != cudaSuccess) { c->err = std::string("load_field3d: ") + cudaGetErrorString(e); rc = MHH_E_CUDA; break; }
            io_pack_kernel<TF, false><<<gr, b, 0, c->stream>>>(d, fld, (TF)offset, g, kstart + k0);
            c->launches++;
            e = cudaStreamSynchronize(c->stream);                                    // the staging buffers are reused by the next batch
            if (e != cudaSuccess) { c->err = std::string("load_field3d: ") + cudaGetErrorString(e); rc = MHH_E_CUDA; }
        }
    }
    if (close(fd) != 0 && rc == MHH_OK && SAVE) { c->err = "save_field3d: close failed"; rc = MHH_E_IO; }
    return rc;
}

} // namespace mhhhost

using namespace mhhhost;

#define DISPATCH1(ctx, expr) \
    do { if (!(ctx)) return MHH_E_INVALID; \
         cudaError_t e_ = cudaSetDevice((ctx)->device); \
         if (e_ != cudaSuccess) { (ctx)->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e_); return MHH_E_CUDA; } \
         if ((ctx)->dtype == MHH_F64) { typedef double TF; Ctx<TF>* c = static_cast<Ctx<TF>*>(ctx); (void)c; return (expr); } \
         else { typedef float TF; Ctx<TF>* c = static_cast<Ctx<TF>*>(ctx); (void)c; return (expr); } } while (0)

extern "C" {

int mhh_field3d_save(mhh_ctx* ctx, const void* fld, const char* filename, double offset, int kstart, int kend)
{ DISPATCH1(ctx, (field3d_io_impl<TF, true>(c, const_cast<TF*>(P<TF>(fld)), filename, offset, kstart, kend))); }

int mhh_field3d_load(mhh_ctx* ctx, void* fld, const char* filename, double offset, int kstart, int kend)
{ DISPATCH1(ctx, (field3d_io_impl<TF, false>(c, P<TF>(fld), filename, offset, kstart, kend))); }

} // extern "C"
