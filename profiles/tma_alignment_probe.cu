#include <cuda_runtime.h>
#include <cuda.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <stdlib.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct Maps { CUtensorMap m[2]; };
template <int MODE>
__global__ void k(const __grid_constant__ Maps tm, const CUtensorMap *gm, float *out, int bw, int bh, int c0, int c1, int c2) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t *bar = (uint64_t *)(sm + 8192);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const CUtensorMap *map = MODE == 0 ? &tm.m[0] : gm;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bw * bh * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(sm)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    uint32_t ok = 0; int spins = 0;
    while (!ok && spins++ < 1000000)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
    if (threadIdx.x == 0) out[0] = ok ? 1.f : -1.f;
    for (int t = threadIdx.x; t < bw * bh; t += blockDim.x) out[1 + t] = ((float *)sm)[t];
}
int main(int argc, char **argv) {
    void *f = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)f;
    printf("entry %p q=%d\n", f, (int)q);
    const int n3 = 56, n2 = 44, np = 8, pitch = 64;
    std::vector<float> h((size_t)np * n2 * pitch);
    for (int i = 0; i < np; i++) for (int j = 0; j < n2; j++) for (int kk = 0; kk < pitch; kk++) h[((size_t)i * n2 + j) * pitch + kk] = i * 10000 + j * 100 + kk;
    float *d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    {
        const int test = 0;
        const int bw = atoi(argv[1]), bh = atoi(argv[2]), c0 = atoi(argv[3]), c1 = atoi(argv[4]), promo = atoi(argv[5]);
        Maps M;
        cuuint64_t dims[3] = { n3, n2, np }; cuuint64_t str[2] = { pitch * 4, (cuuint64_t)n2 * pitch * 4 };
        cuuint32_t box[3] = { (cuuint32_t)bw, (cuuint32_t)bh, 1 }, es[3] = { 1, 1, 1 };
        CUresult r = enc(&M.m[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        M.m[1] = M.m[0];
        CUtensorMap *gm; cudaMalloc(&gm, sizeof(CUtensorMap)); cudaMemcpy(gm, &M.m[0], sizeof(CUtensorMap), cudaMemcpyHostToDevice);
        float *out; cudaMalloc(&out, 4 * (1 + bw * bh)); cudaMemset(out, 0, 4 * (1 + bw * bh));
        k<0><<<1, 128, 8192 + 64>>>(M, gm, out, bw, bh, c0, c1, 3);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> o(1 + bw * bh);
        cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
        printf("enc=%d box %dx%d at (%d,%d) promo %d: %s ok=%g first row: %g %g %g %g ... row2: %g %g %g\n", (int)r, bw, bh, c0, c1, promo,
               cudaGetErrorString(e), o[0], o[1], o[2], o[3], o[4], o[1 + 2 * bw], o[1 + 2 * bw + 1], o[1 + 2 * bw + 2]);
    }
    return 0;
}
