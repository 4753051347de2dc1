// Instruction-cache capacity probe (sm_100a): a loop whose body is N KB of straight-line FFMA, run by 32 warps per SM on every SM.
// Prints cycles per instruction per SMSP for each body size; run under ncu for sm__icc_request_hit_rate / gcc requests.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache_probe icache_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define F1 a = fmaf(a, b, c); c = fmaf(c, b, a);
#define F4 F1 F1 F1 F1
#define F16 F4 F4 F4 F4
#define F64 F16 F16 F16 F16          // 128 FFMA = 2 KB
#define K2 F64
#define K4 K2 K2
#define K8 K4 K4
#define K16 K8 K8
#define K32 K16 K16
#define K64 K32 K32
template <int KB> __device__ __forceinline__ void body(float &a, float b, float &c);
template <> __device__ __forceinline__ void body<4>(float &a, float b, float &c) { K4 }
template <> __device__ __forceinline__ void body<8>(float &a, float b, float &c) { K8 }
template <> __device__ __forceinline__ void body<16>(float &a, float b, float &c) { K16 }
template <> __device__ __forceinline__ void body<24>(float &a, float b, float &c) { K16 K8 }
template <> __device__ __forceinline__ void body<32>(float &a, float b, float &c) { K32 }
template <> __device__ __forceinline__ void body<40>(float &a, float b, float &c) { K32 K8 }
template <> __device__ __forceinline__ void body<48>(float &a, float b, float &c) { K32 K16 }
template <> __device__ __forceinline__ void body<56>(float &a, float b, float &c) { K32 K16 K8 }
template <> __device__ __forceinline__ void body<64>(float &a, float b, float &c) { K64 }
template <> __device__ __forceinline__ void body<80>(float &a, float b, float &c) { K64 K16 }
template <> __device__ __forceinline__ void body<96>(float &a, float b, float &c) { K64 K32 }
template <> __device__ __forceinline__ void body<128>(float &a, float b, float &c) { K64 K64 }
template <> __device__ __forceinline__ void body<192>(float &a, float b, float &c) { K64 K64 K64 }
// skew > 0: warp w starts its first pass (w * skew) iterations of a dummy spin later, so the warps of an SM are NOT in lockstep
template <int KB> __global__ void __launch_bounds__(1024, 1) k_probe(float *out, int iters, int skew, long long *cyc) {
    float a = threadIdx.x * 1e-3f, b = 0.999f, c = 0.5f;
    if (skew) { for (int i = 0; i < (int)(threadIdx.x >> 5) * skew; ++i) a = fmaf(a, b, c); }
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) body<KB>(a, b, c);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    if (a == 12345.0f) out[0] = a + c;
}
template <int KB> void run(float *d, long long *dc, int skew) {
    const int iters = 4096 / KB > 8 ? 4096 / KB : 8;
    k_probe<KB><<<148, 1024>>>(d, 2, skew, dc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_probe<KB><<<148, 1024>>>(d, iters, skew, dc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc; cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
    const double inst = (double)iters * KB * 64.0;  // per warp
    printf("body %3d KB  skew %4d  iters %5d  %8.3f ms  cycles/inst/warp %.3f  -> per SMSP (8 warps) %.3f cyc/inst\n", KB, skew, iters, ms, cyc / inst, cyc / inst / 8.0);
}
int main(int argc, char **argv) {
    float *d; long long *dc; cudaMalloc(&d, 4); cudaMalloc(&dc, 8);
    for (int skew : {0, 997}) {
        run<4>(d, dc, skew); run<8>(d, dc, skew); run<16>(d, dc, skew); run<24>(d, dc, skew); run<32>(d, dc, skew); run<40>(d, dc, skew); run<48>(d, dc, skew);
        run<56>(d, dc, skew); run<64>(d, dc, skew); run<80>(d, dc, skew); run<96>(d, dc, skew); run<128>(d, dc, skew); run<192>(d, dc, skew);
    }
    return 0;
}
