// Can a small kernel on a second stream run BESIDE a persistent kernel that holds one large-shared-memory CTA on every SM
// (the situation of the tcgen05 kernel and the next batch's filter / linking kernels)? Measures, for several launch
// configurations, how long the small kernel takes from launch to completion while the big one spins for ~5 ms.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o coresidency_probe coresidency_probe.cu && ./coresidency_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void big_kernel(long long cycles, int *sink) {
    extern __shared__ int sm[];
    sm[threadIdx.x] = threadIdx.x;
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {
        __nanosleep(200);
    }
    if (sm[threadIdx.x] == -1) {
        *sink = 1;
    }
}
__global__ void __cluster_dims__(2, 1, 1) big_cluster_kernel(long long cycles, int *sink) {
    extern __shared__ int sm[];
    sm[threadIdx.x] = threadIdx.x;
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {
        __nanosleep(200);
    }
    if (sm[threadIdx.x] == -1) {
        *sink = 1;
    }
}
template <int STATIC_KB>
__global__ void small_kernel(float *out, int iters) {
    __shared__ float s[STATIC_KB * 256];
    s[threadIdx.x] = threadIdx.x;
    __syncthreads();
    float v = s[(threadIdx.x + 1) & 255];
    for (int i = 0; i < iters; ++i) {
        v = v * 1.0001f + 0.5f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = v;
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int least, greatest;
    CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    cudaStream_t s_big, s_big_hi, s_small;
    CK(cudaStreamCreateWithFlags(&s_big, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithPriority(&s_big_hi, cudaStreamNonBlocking, greatest));
    CK(cudaStreamCreateWithFlags(&s_small, cudaStreamNonBlocking));
    float *out;
    int *sink;
    CK(cudaMalloc(&out, 4096 * 256 * sizeof(float)));
    CK(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1, b0, b1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventCreate(&b0));
    CK(cudaEventCreate(&b1));
    const long long cycles = 5LL * 1900000; // ~5 ms at 1.9 GHz
    printf("SMs %d, stream priorities %d..%d\n", sms, least, greatest);
    for (int big_kb : { 161, 193 }) {
        CK(cudaFuncSetAttribute(big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big_kb * 1024));
        CK(cudaFuncSetAttribute(big_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big_kb * 1024));
        for (int variant = 0; variant < 6; ++variant) {
            // 0: plain; 1: small kernel with carveout 100; 2: device cache config PreferShared; 3: big kernel as clusters of 2;
            // 4: big kernel on the high-priority stream; 5: clusters + high priority + small carveout 100
            const bool carve = variant == 1 || variant == 5, prefer = variant == 2, cluster = variant == 3 || variant == 5,
                       hi = variant == 4 || variant == 5;
            CK(cudaDeviceSetCacheConfig(prefer ? cudaFuncCachePreferShared : cudaFuncCachePreferNone));
            CK(cudaFuncSetAttribute(small_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, carve ? 100 : -1));
            CK(cudaFuncSetAttribute(small_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, carve ? 100 : -1));
            for (int small_kb : { 1, 16 }) {
                CK(cudaDeviceSynchronize());
                cudaStream_t sb = hi ? s_big_hi : s_big;
                CK(cudaEventRecord(b0, sb));
                if (cluster) {
                    big_cluster_kernel<<<sms, 192, big_kb * 1024, sb>>>(cycles, sink);
                } else {
                    big_kernel<<<sms, 192, big_kb * 1024, sb>>>(cycles, sink);
                }
                CK(cudaGetLastError());
                CK(cudaEventRecord(b1, sb));
                // give the big kernel time to occupy the SMs
                cudaStreamQuery(sb);
                for (volatile int spin = 0; spin < 2000000; ++spin) {
                }
                CK(cudaEventRecord(e0, s_small));
                if (small_kb == 1) {
                    small_kernel<1><<<2048, 256, 0, s_small>>>(out, 2000);
                } else {
                    small_kernel<16><<<2048, 256, 0, s_small>>>(out, 2000);
                }
                CK(cudaGetLastError());
                CK(cudaEventRecord(e1, s_small));
                CK(cudaDeviceSynchronize());
                float ms_small = 0, ms_big = 0, ms_gap = 0;
                CK(cudaEventElapsedTime(&ms_small, e0, e1));
                CK(cudaEventElapsedTime(&ms_big, b0, b1));
                CK(cudaEventElapsedTime(&ms_gap, b0, e1));
                printf("big %3d KB %-8s %-4s | small %2d KB static %-12s %-13s : small kernel %.3f ms (big %.3f ms, small done %.3f ms after big start) -> %s\n",
                       big_kb, cluster ? "cluster2" : "plain", hi ? "hi" : "norm", small_kb, carve ? "carveout=100" : "carveout=def",
                       prefer ? "PreferShared" : "PreferNone", ms_small, ms_big, ms_gap, ms_gap < ms_big - 0.5f ? "RAN BESIDE" : "WAITED");
            }
        }
    }
    return 0;
}
