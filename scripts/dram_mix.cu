// DRAM bandwidth of plain streaming kernels with the read : write mixes of the library's launches, as the yardstick
// next to MEASURED_PEAKS.json's copy figure (1 : 1):  copy (d/dx, d/dy, d/dz: 8 B read + 8 B written per point),
// one read + two writes (the fused x/y launch: 8 + 16 B per point).  Not product code; scripts/dram_mix.py drives it.
#include <cuda_runtime.h>
extern "C" {
__global__ void k_copy(const double2 *__restrict__ a, double2 *__restrict__ b, long n2)
{
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n2; i += (long)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void k_r1w2(const double2 *__restrict__ a, double2 *__restrict__ b, double2 *__restrict__ c, long n2)
{
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n2; i += (long)gridDim.x * blockDim.x) {
        const double2 v = a[i];
        b[i] = v;
        c[i] = make_double2(v.y, v.x);
    }
}
int mix_copy(const void *a, void *b, long n, int blocks, int threads, void *stream)
{
    k_copy<<<blocks, threads, 0, (cudaStream_t)stream>>>((const double2 *)a, (double2 *)b, n / 2);
    return (int)cudaGetLastError();
}
int mix_r1w2(const void *a, void *b, void *c, long n, int blocks, int threads, void *stream)
{
    k_r1w2<<<blocks, threads, 0, (cudaStream_t)stream>>>((const double2 *)a, (double2 *)b, (double2 *)c, n / 2);
    return (int)cudaGetLastError();
}
}
