// Probe: what DRAM bandwidth does the *access pattern* of the FDTD half-step allow, independent of
// the arithmetic?  A CTA owns a (8 x W) tile of a (n2 x n3) plane and marches planes, reading NIN
// field arrays and writing NOUT of them (the stress half-step of a solid tile: 13 in, 13 out; fluid: 7/7).
// Every thread issues all loads of a plane before the first store (maximum memory-level parallelism),
// so the number printed is the ceiling for a tile width of W floats (W*4-byte contiguous row segments).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tile_bw_probe tile_bw_probe.cu && ./tile_bw_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>
template <int W, int NIN, int NOUT>
__global__ void __launch_bounds__(256) probe(float *const *in, float *const *out, int n1, int n2, int pitch, int chunk) {
    constexpr int V = W / 32;                      // floats per thread per row
    const int k = blockIdx.x * W + threadIdx.x * V;   // V consecutive floats per thread (vector access)
    const int j = blockIdx.y * 8 + threadIdx.y;
    const int i0 = blockIdx.z * chunk, i1 = min(i0 + chunk, n1);
    if (j >= n2) return;
    for (int i = i0; i < i1; i++) {
        const long long q = ((long long)i * n2 + j) * pitch + k;
        float v[NIN][V];
#pragma unroll
        for (int f = 0; f < NIN; f++) {
            if (V == 4) *reinterpret_cast<float4 *>(v[f]) = *reinterpret_cast<const float4 *>(in[f] + q);
            else if (V == 2) *reinterpret_cast<float2 *>(v[f]) = *reinterpret_cast<const float2 *>(in[f] + q);
            else v[f][0] = in[f][q];
        }
#pragma unroll
        for (int f = 0; f < NOUT; f++) {
#pragma unroll
            for (int e = 0; e < V; e++) v[f][e] += 1.0f;
            if (V == 4) *reinterpret_cast<float4 *>(out[f] + q) = *reinterpret_cast<float4 *>(v[f]);
            else if (V == 2) *reinterpret_cast<float2 *>(out[f] + q) = *reinterpret_cast<float2 *>(v[f]);
            else out[f][q] = v[f][0];
        }
        if (NIN > NOUT) {   // keep the extra loads alive
            float s = 0.f;
#pragma unroll
            for (int f = NOUT; f < NIN; f++) for (int e = 0; e < V; e++) s += v[f][e];
            if (s == 123.456f) out[0][q] = s;
        }
    }
}
template <int W, int NIN, int NOUT>
void run(const char *name, float **din, float **dout, int n1, int n2, int n3, int chunk) {
    dim3 blk(32, 8), grid(n3 / W, (n2 + 7) / 8, (n1 + chunk - 1) / chunk);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 3; it++) probe<W, NIN, NOUT><<<grid, blk>>>(din, dout, n1, n2, n3, chunk);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int it = 0; it < reps; it++) probe<W, NIN, NOUT><<<grid, blk>>>(din, dout, n1, n2, n3, chunk);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    const double bytes = (double)n1 * n2 * n3 * 4 * (NIN + NOUT);
    printf("%-28s W=%3d in=%2d out=%2d chunk=%3d grid=%5d  %.3f ms  %.0f GB/s  (%s)\n", name, W, NIN, NOUT, chunk, grid.x * grid.y * grid.z, ms, bytes / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int n1 = 240, n2 = 240, n3 = 384;   // k padded to a multiple of 128
    const int NF = 13;
    std::vector<float *> in(NF), out(NF);
    const size_t vol = (size_t)n1 * n2 * n3;
    for (int f = 0; f < NF; f++) { cudaMalloc(&in[f], vol * 4); cudaMalloc(&out[f], vol * 4); cudaMemset(in[f], 0, vol * 4); cudaMemset(out[f], 0, vol * 4); }
    float **din, **dout;
    cudaMalloc(&din, NF * sizeof(float *)); cudaMalloc(&dout, NF * sizeof(float *));
    cudaMemcpy(din, in.data(), NF * sizeof(float *), cudaMemcpyHostToDevice);
    // in place, like the solver: outputs alias the inputs
    cudaMemcpy(dout, in.data(), NF * sizeof(float *), cudaMemcpyHostToDevice);
    for (int chunk : { 60, 240 }) {
        run<32, 13, 13>("solid stress (in place)", din, dout, n1, n2, n3, chunk);
        run<64, 13, 13>("solid stress (in place)", din, dout, n1, n2, n3, chunk);
        run<128, 13, 13>("solid stress (in place)", din, dout, n1, n2, n3, chunk);
        run<32, 10, 7>("att fluid stress (in place)", din, dout, n1, n2, n3, chunk);
        run<64, 10, 7>("att fluid stress (in place)", din, dout, n1, n2, n3, chunk);
        run<128, 10, 7>("att fluid stress (in place)", din, dout, n1, n2, n3, chunk);
        run<32, 6, 3>("particle fluid (in place)", din, dout, n1, n2, n3, chunk);
        run<64, 6, 3>("particle fluid (in place)", din, dout, n1, n2, n3, chunk);
        run<128, 6, 3>("particle fluid (in place)", din, dout, n1, n2, n3, chunk);
    }
    return 0;
}
