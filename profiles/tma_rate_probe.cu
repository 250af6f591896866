// Probe: TMA load throughput as a function of the box shape.  One CTA per SM streams boxes of
// (W floats x H rows x 1 plane x C components) from a large 4-D tensor into a ring of shared-memory slots; a consumer
// warp waits for each slot and releases it (no arithmetic, no stores).  If the TMA unit were byte-bound
// every shape would reach the same GB/s; a per-row (or per-instruction) cost shows up as a dependence
// on W and H.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate_probe tma_rate_probe.cu
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok;
}
constexpr int NS = 8;
__global__ void __launch_bounds__(64, 1) stream(const __grid_constant__ CUtensorMap tm, int W, int H, int C, int n3, int n2, int n1, int boxbytes, int slotbytes, long long *sink) {
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t *full = (uint64_t *)(sm + NS * slotbytes), *empty = full + NS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(full + s)) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(empty + s)) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // this CTA's boxes: tiles (k, j) round-robin over CTAs, marching all planes
    const int ntk = n3 / W, ntj = n2 / H, ntiles = ntk * ntj;
    long long count = 0;
    if (warp == 0) {
        if (lane != 0) return;
        int slot = 0; unsigned ph = 0xFFFFFFFFu;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int k0 = (t % ntk) * W, j0 = (t / ntk) * H;
            for (int i = 0; i < n1; i++) {
                while (!try_wait(empty + slot, (ph >> slot) & 1)) { }
                ph ^= 1u << slot;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(full + slot)), "r"(boxbytes) : "memory");
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                             ::"r"(smem_u32(sm + slot * slotbytes)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(full + slot)), "r"(k0), "r"(j0), "r"(i), "r"(0) : "memory");
                slot = slot + 1 == NS ? 0 : slot + 1;
            }
        }
    } else {
        int slot = 0; unsigned ph = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            for (int i = 0; i < n1; i++) {
                while (!try_wait(full + slot, (ph >> slot) & 1)) { }
                ph ^= 1u << slot;
                count += ((volatile int *)(sm + slot * slotbytes))[lane];
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + slot)) : "memory");
                slot = slot + 1 == NS ? 0 : slot + 1;
            }
        }
        if (count == 0x7fffffffffffll) *sink = count;
    }
}
int main() {
    void *f = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)f;
    const int n1 = 256, n2 = 512, n3 = 1024, NC = 3;   // 3 components x 512 MB
    const size_t vol = (size_t)n1 * n2 * n3;
    float *d; cudaMalloc(&d, vol * 4 * NC); cudaMemset(d, 0, vol * 4 * NC);
    long long *sink; cudaMalloc(&sink, 8);
    const int shapes[][3] = { {32, 8, 1}, {64, 8, 1}, {128, 8, 1}, {256, 8, 1}, {64, 16, 1}, {64, 32, 1}, {128, 16, 1}, {64, 8, 3}, {128, 8, 3}, {128, 4, 3}, {256, 4, 3}, {72, 12, 3} };
    cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (auto &sh : shapes) {
        const int W = sh[0], H = sh[1], C = sh[2];
        CUtensorMap tm;
        cuuint64_t dims[4] = { (cuuint64_t)n3, (cuuint64_t)n2, (cuuint64_t)n1, (cuuint64_t)NC }; cuuint64_t str[3] = { (cuuint64_t)n3 * 4, (cuuint64_t)n2 * n3 * 4, (cuuint64_t)vol * 4 };
        cuuint32_t box[4] = { (cuuint32_t)W, (cuuint32_t)H, 1, (cuuint32_t)C }, es[4] = { 1, 1, 1, 1 };
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int boxbytes = W * H * C * 4, slotbytes = (boxbytes + 1023) / 1024 * 1024;
        const int smem = NS * slotbytes + 2 * NS * 8;
        const int wq = W == 72 ? 64 : W, hq = H == 12 ? 8 : H;   // the halo-shaped box strides like its 64 x 8 tile
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        stream<<<148, 64, smem>>>(tm, wq, hq, C, n3, n2, n1, boxbytes, slotbytes, sink);
        cudaEventRecord(e0);
        stream<<<148, 64, smem>>>(tm, wq, hq, C, n3, n2, n1, boxbytes, slotbytes, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double nbox = (double)(n3 / wq) * (n2 / hq) * n1;
        printf("enc=%d box %3d x %2d x %d (%5d B, %3d rows): %.3f ms  %.0f GB/s  %.1f ns/box/SM  %.1f cycles/row @1.9GHz  (%s)\n", (int)r, W, H, C, boxbytes, H * C, ms,
               nbox * boxbytes / ms / 1e6, ms * 1e6 / (nbox / 148), ms * 1e6 / (nbox / 148) / (H * C) * 1.9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
