// Probe: which 3-D tiled tensor maps / boxes does TMA accept for fp64 fields?  (development tool)
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void probe(const __grid_constant__ CUtensorMap tm, double* out, int nbox, int x, int y, int z)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem);
    double* dst = reinterpret_cast<double*>(smem + 128);
    const unsigned b = smem_u32(bar), d = smem_u32(dst);
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(nbox * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                     :: "r"(d), "l"(reinterpret_cast<unsigned long long>(&tm)), "r"(b), "r"(x), "r"(y), "r"(z) : "memory");
    }
    unsigned ok;
    do { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(b), "r"(0) : "memory"); } while (!ok);
    for (int t = threadIdx.x; t < nbox; t += blockDim.x) out[t] = dst[t];
}
int main(int argc, char** argv)
{
    int ic = atoi(argv[1]), jc = atoi(argv[2]), kc = atoi(argv[3]), bx = atoi(argv[4]), by = atoi(argv[5]);
    int x = atoi(argv[6]), y = atoi(argv[7]), z = atoi(argv[8]);
    // optional 9th argument: element stride along x (2 = every other column: the even/odd de-interleaved planes planned
    // for mom3, DESIGN.md "next steps"); the box then delivers ceil(bx / sx) elements per row
    const int sx = argc > 9 ? atoi(argv[9]) : 1;
    void* p; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    size_t n = (size_t)ic * jc * kc;
    std::vector<double> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (double)i;
    double *d, *o; cudaMalloc(&d, n * 8); cudaMalloc(&o, (size_t)bx * by * 8);
    cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)ic, (cuuint64_t)jc, (cuuint64_t)kc};
    cuuint64_t str[2] = {(cuuint64_t)ic * 8, (cuuint64_t)ic * jc * 8};
    cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1}, es[3] = {(cuuint32_t)sx, 1, 1};
    const int nx = (bx + sx - 1) / sx;          // elements per row that land in shared memory
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d  dims %d %d %d box %d %d at %d %d %d\n", (int)r, ic, jc, kc, bx, by, x, y, z);
    if (r != CUDA_SUCCESS) return 0;
    size_t smem = 128 + (size_t)nx * by * 8;
    cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<0><<<1, 128, smem>>>(tm, o, nx * by, x, y, z);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  run: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 0;
    std::vector<double> ho((size_t)bx * by);
    cudaMemcpy(ho.data(), o, ho.size() * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int jj = 0; jj < by; ++jj) for (int ii = 0; ii < nx; ++ii)
    {
        int gi = x + ii * sx, gj = y + jj;
        double exp = (gi >= 0 && gi < ic && gj >= 0 && gj < jc && z >= 0 && z < kc) ? h[(size_t)gi + (size_t)gj * ic + (size_t)z * ic * jc] : 0.;
        if (ho[(size_t)jj * nx + ii] != exp) ++bad;
    }
    printf("  mismatches: %d of %d (element stride %d)\n", bad, nx * by, sx);
    return 0;
}
