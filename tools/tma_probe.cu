// Development probe: 3-D TMA tile load with OOB zero fill; variants selected by argv.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <bool PARAM>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, const CUtensorMap* gmap, float* out, int bw, int bh, int x, int y, int z, int bz)
{
    extern __shared__ __align__(128) unsigned char raw[];
    float* s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(raw) + 127) & ~uintptr_t(127));
    uint64_t* bar = reinterpret_cast<uint64_t*>(s + bw * bh * bz);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bw * bh * bz * 4) : "memory");
        const CUtensorMap* m = PARAM ? &tmap : gmap;
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(s)), "l"(reinterpret_cast<uint64_t>(m)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bw * bh * bz; i += blockDim.x) out[i] = s[i];
}
int main(int argc, char** argv)
{
    int w = atoi(argv[1]), h = atoi(argv[2]), pitch = atoi(argv[3]), nb = atoi(argv[4]);
    int bw = atoi(argv[5]), bh = atoi(argv[6]), x = atoi(argv[7]), y = atoi(argv[8]), z = atoi(argv[9]);
    int use_param = atoi(argv[10]);
    int bz = argc > 11 ? atoi(argv[11]) : 1;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    std::vector<float> host((size_t)pitch * h * nb);
    for (size_t i = 0; i < host.size(); ++i) host[i] = (float)(i % 1000) + 1;
    float* d; cudaMalloc(&d, host.size() * 4); cudaMemcpy(d, host.data(), host.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)nb};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * h * 4};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bz}, es[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d q=%d\n", (int)r, (int)q);
    if (r != CUDA_SUCCESS) return 1;
    CUtensorMap* gmap; cudaMalloc(&gmap, sizeof(map)); cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice);
    float* out; cudaMalloc(&out, bw * bh * bz * 4);
    int smem = bw * bh * bz * 4 + 16 + 128;
    cudaFuncSetAttribute(probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (use_param) probe<true><<<1, 128, smem>>>(map, gmap, out, bw, bh, x, y, z, bz);
    else probe<false><<<1, 128, smem>>>(map, gmap, out, bw, bh, x, y, z, bz);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<float> res(bw * bh * bz);
    cudaMemcpy(res.data(), out, bw * bh * bz * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int p2 = 0; p2 < bz; ++p2) for (int r2 = 0; r2 < bh; ++r2) for (int c = 0; c < bw; ++c) {
        int gx = x + c, gy = y + r2, gz = z + p2;
        float want = (gx >= 0 && gx < w && gy >= 0 && gy < h && gz >= 0 && gz < nb) ? host[((size_t)gz * h + gy) * pitch + gx] : 0.f;
        if (res[(p2 * bh + r2) * bw + c] != want) { if (bad < 5) printf("mismatch r=%d c=%d got %f want %f\n", r2, c, res[(p2 * bh + r2) * bw + c], want); ++bad; }
    }
    printf("mismatches: %d\n", bad);
    return 0;
}
