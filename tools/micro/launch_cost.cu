// What does a launch of a 148-CTA kernel with ~194 KB of dynamic shared memory cost on B200, and
// which part of the stats scan's fixed ~15 us is it?  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(384, 1) k_empty(int* sink) {
    extern __shared__ unsigned char smem[];
    if (sink == nullptr && threadIdx.x == 9999) smem[0] = 1;
}

__global__ void __launch_bounds__(128) k_small(int* sink) {
    if (sink && threadIdx.x == 0 && blockIdx.x == 0) sink[0] += 1;
}

// table fill like fast_scan_kernel<6, stats>: 12 KB from global, replicated 8x into 96 KB
__global__ void __launch_bounds__(384, 1) k_fill(const uint4* __restrict__ tab, int* sink, int hist_words) {
    extern __shared__ unsigned char smem[];
    uint4 v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int i = threadIdx.x + k * 384;
        v[k] = __ldg(tab + ((i >> 3) & 767));
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int i = threadIdx.x + k * 384;
        *reinterpret_cast<uint4*>(smem + (i >> 3) * 128 + (i & 7) * 16) = v[k];
    }
    unsigned int* hist = reinterpret_cast<unsigned int*>(smem + 98304);
    for (int i = threadIdx.x; i < hist_words; i += 384) hist[i] = 0u;
    __syncthreads();
    if (sink && smem[threadIdx.x] == 77 && hist[threadIdx.x] == 5) sink[1] = 1;
}

// + histogram flush: 256 bins x 32 words summed, one global atomic per bin and CTA
__global__ void __launch_bounds__(384, 1) k_fill_flush(const uint4* __restrict__ tab, unsigned long long* ghist,
                                                       int* sink, int do_atomics) {
    extern __shared__ unsigned char smem[];
    uint4 v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int i = threadIdx.x + k * 384;
        v[k] = __ldg(tab + ((i >> 3) & 767));
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int i = threadIdx.x + k * 384;
        *reinterpret_cast<uint4*>(smem + (i >> 3) * 128 + (i & 7) * 16) = v[k];
    }
    unsigned int* hist = reinterpret_cast<unsigned int*>(smem + 98304);
    for (int i = threadIdx.x; i < 257 * 32; i += 384) hist[i] = 1u;
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += 384) {
        unsigned int s = 0;
        for (int l = 0; l < 32; ++l) s += hist[i * 32 + ((l + i) & 31)];
        if (do_atomics && s) atomicAdd(ghist + i, (unsigned long long)s);
        else if (s == 12345) sink[2] = 1;
    }
}

template <class F>
static float time_us(F launch, int reps = 300) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 20; ++i) launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    return ms * 1e3f / reps;
}

int main() {
    int* sink; cudaMalloc(&sink, 64); cudaMemset(sink, 0, 64);
    uint4* tab; cudaMalloc(&tab, 768 * 16); cudaMemset(tab, 1, 768 * 16);
    unsigned long long* gh; cudaMalloc(&gh, 256 * 8); cudaMemset(gh, 0, 256 * 8);
    const int big = 98304 + 63360 + 257 * 128;     // tables + strip + histogram = 194 560 B
    cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_fill_flush, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    printf("empty, no smem            : %6.2f us\n", time_us([&] { k_empty<<<148, 384, 0>>>(sink); }));
    printf("empty, 190 KB smem        : %6.2f us\n", time_us([&] { k_empty<<<148, 384, big>>>(sink); }));
    printf("small + empty 190 KB      : %6.2f us (pair)\n",
           time_us([&] { k_small<<<148, 128>>>(sink); k_empty<<<148, 384, big>>>(sink); }));
    printf("small + empty no smem     : %6.2f us (pair)\n",
           time_us([&] { k_small<<<148, 128>>>(sink); k_empty<<<148, 384, 0>>>(sink); }));
    printf("memset + empty 190 KB     : %6.2f us (pair)\n",
           time_us([&] { cudaMemsetAsync(sink, 0, 64); k_empty<<<148, 384, big>>>(sink); }));
    printf("fill tables               : %6.2f us\n", time_us([&] { k_fill<<<148, 384, big>>>(tab, sink, 0); }));
    printf("fill tables + zero hist   : %6.2f us\n", time_us([&] { k_fill<<<148, 384, big>>>(tab, sink, 257 * 32); }));
    printf("fill + hist sum (no atom) : %6.2f us\n", time_us([&] { k_fill_flush<<<148, 384, big>>>(tab, gh, sink, 0); }));
    printf("fill + hist flush atomics : %6.2f us\n", time_us([&] { k_fill_flush<<<148, 384, big>>>(tab, gh, sink, 1); }));
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
