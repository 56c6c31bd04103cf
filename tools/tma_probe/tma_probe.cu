// standalone probe of the 3-D TMA window load used by the plain search kernel (experiment support; not part of the library)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <vector>
struct alignas(64) Map { unsigned long long o[16]; };
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ Map tmap, int c0, int c1, int c2, uint8_t* out, int bytes)
{
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned long long* bar = (unsigned long long*)(sm + 4096);
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(bar)), "r"(1) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     :: "r"(s32(sm)), "l"(&tmap), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    uint32_t ok;
    do { asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s32(bar)), "r"(0) : "memory"); } while (!ok);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv)
{
    const int bw = argc > 1 ? atoi(argv[1]) : 16, c0 = argc > 2 ? atoi(argv[2]) : 37;
    const int stride = 352, rows = 256, planes = 8;
    std::vector<uint8_t> h((size_t)stride * rows * planes);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 8));
    uint8_t *d, *dout; cudaMalloc(&d, h.size()); cudaMalloc(&dout, 4096);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = NULL; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry %d %p %d\n", (int)e, fn, (int)q);
    cuuint64_t dims[3] = { (cuuint64_t)stride, (cuuint64_t)rows, (cuuint64_t)planes };
    cuuint64_t strides[2] = { (cuuint64_t)stride, (cuuint64_t)stride * rows };
    cuuint32_t box[3] = { (cuuint32_t)bw, 13, 4 }, es[3] = { 1, 1, 1 };
    CUtensorMap m;
    CUresult r = ((EncodeFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    Map mm; memcpy(&mm, &m, 128);
    const int c1 = 21, c2 = 4, bytes = bw * 13 * 4;
    probe<<<1, 64, 8192>>>(mm, c0, c1, c2, dout, bytes);
    e = cudaDeviceSynchronize();
    printf("sync %s\n", cudaGetErrorString(e));
    std::vector<uint8_t> o(bytes); cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int p = 0; p < 4; p++) for (int y = 0; y < 13; y++) for (int x = 0; x < bw; x++)
        if (o[(p * 13 + y) * bw + x] != h[((size_t)(c2 + p) * rows + c1 + y) * stride + c0 + x]) bad++;
    printf("mismatches %d\n", bad);
    return 0;
}
