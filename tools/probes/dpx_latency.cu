// dpx_latency.cu — measurement aid: dependent-issue latency (cycles per step of a dependent chain, one warp) and throughput (cycles per
// instruction with 8 independent chains per thread, 8 warps per scheduler) of the min-plus step  v = min(v + 1, w)  written three ways:
//   dpx   : VIADDMNMX.U16x2 (__viaddmin_u16x2), two voxels per register
//   int   : IADD + VIMNMX.U16x2 (__vminu2)
//   half  : HADD2 + HMNMX2 on values biased by 1024.0 (exact for integers below 2048)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dpx_latency dpx_latency.cu && ./dpx_latency
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int N = 4096;

template <int MODE>
__device__ __forceinline__ unsigned step(unsigned v, unsigned w) {
    if (MODE == 0) return __viaddmin_u16x2(v, 0x00010001u, w);
    if (MODE == 1) return __vminu2(v + 0x00010001u, w);
    __half2 hv = *reinterpret_cast<__half2*>(&v), hw = *reinterpret_cast<__half2*>(&w);
    hv = __hmin2(__hadd2(hv, __floats2half2_rn(1.0f, 1.0f)), hw);
    return *reinterpret_cast<unsigned*>(&hv);
}

template <int MODE, int CHAINS>
__global__ void probe(const unsigned* in, unsigned* out, long long* cycles) {
    unsigned v[CHAINS], w[CHAINS];
    for (int c = 0; c < CHAINS; ++c) { v[c] = in[threadIdx.x + 32 * c]; w[c] = in[threadIdx.x + 32 * c + 1024]; }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N / 16; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) v[c] = step<MODE>(v[c], w[c]);
    }
    const long long t1 = clock64();
    unsigned s = 0;
    for (int c = 0; c < CHAINS; ++c) s ^= v[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE, int CHAINS>
double run(int threads, const unsigned* in, unsigned* out, long long* cyc) {
    probe<MODE, CHAINS><<<1, threads>>>(in, out, cyc);
    probe<MODE, CHAINS><<<1, threads>>>(in, out, cyc);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    return (double)h / N;
}

int main() {
    unsigned *in, *out;
    long long* cyc;
    cudaMalloc(&in, 8192 * 4);
    cudaMalloc(&out, 4096 * 4);
    cudaMalloc(&cyc, 8);
    unsigned h[8192];
    for (int i = 0; i < 8192; ++i) h[i] = 0x64FE64FEu;  // 254 in both lanes (integer view: large; half view: 1024 + 254)
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    const char* names[3] = {"dpx  VIADDMNMX.U16x2       ", "int  IADD + VIMNMX.U16x2    ", "half HADD2 + HMNMX2         "};
    printf("cycles per min-plus step (chain of %d)\n", N);
    printf("%s 1 warp x 1 chain: %.2f   1 warp x 4 chains: %.2f per chain-step   32 warps x 4 chains (8 per scheduler): %.2f per warp-step\n", names[0],
           run<0, 1>(32, in, out, cyc), run<0, 4>(32, in, out, cyc) , run<0, 4>(1024, in, out, cyc));
    printf("%s 1 warp x 1 chain: %.2f   1 warp x 4 chains: %.2f per chain-step   32 warps x 4 chains (8 per scheduler): %.2f per warp-step\n", names[1],
           run<1, 1>(32, in, out, cyc), run<1, 4>(32, in, out, cyc), run<1, 4>(1024, in, out, cyc));
    printf("%s 1 warp x 1 chain: %.2f   1 warp x 4 chains: %.2f per chain-step   32 warps x 4 chains (8 per scheduler): %.2f per warp-step\n", names[2],
           run<2, 1>(32, in, out, cyc), run<2, 4>(32, in, out, cyc), run<2, 4>(1024, in, out, cyc));
    return 0;
}
