// Does a kernel whose straight-line code is ~30 KB pay for cold instruction fetch on every launch?
// (fast_scan_kernel's fully unrolled 32-position body is 1 600 - 2 200 SASS instructions.)
#include <cstdio>
#include <cuda_runtime.h>

template <int N>
__global__ void __launch_bounds__(384, 1) k_straight(unsigned* out, unsigned seed) {
    unsigned a = seed + threadIdx.x, b = a * 3u + 1u, c = b ^ 0x9e3779b9u, d = c + 7u;
#pragma unroll
    for (int i = 0; i < N; ++i) {           // 4 dependent-free integer ops per step, all distinct constants
        a = a * 1664525u + (unsigned)(i * 2654435761u + 1013904223u);
        b ^= a >> ((i & 15) + 1);
        c += b * (unsigned)(2 * i + 3);
        d = (d << 1) ^ c ^ (unsigned)(i * 40503u);
    }
    if ((a ^ b ^ c ^ d) == 0x12345678u) out[0] = a;
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

__global__ void k_other(unsigned* out) { if (threadIdx.x == 9999) out[1] = 1; }

int main() {
    unsigned* out; cudaMalloc(&out, 64);
    printf("straight-line   16 steps: %6.2f us/launch\n", time_us([&] { k_straight<16><<<148, 384>>>(out, 1); }));
    printf("straight-line  128 steps: %6.2f us/launch\n", time_us([&] { k_straight<128><<<148, 384>>>(out, 1); }));
    printf("straight-line  512 steps: %6.2f us/launch\n", time_us([&] { k_straight<512><<<148, 384>>>(out, 1); }));
    printf("straight-line 1024 steps: %6.2f us/launch\n", time_us([&] { k_straight<1024><<<148, 384>>>(out, 1); }));
    printf("straight-line 2048 steps: %6.2f us/launch\n", time_us([&] { k_straight<2048><<<148, 384>>>(out, 1); }));
    printf("1024 steps, 1 warp/CTA  : %6.2f us/launch\n", time_us([&] { k_straight<1024><<<148, 32>>>(out, 1); }));
    printf("1024 steps alternating with another kernel: %6.2f us/pair\n",
           time_us([&] { k_other<<<148, 128>>>(out); k_straight<1024><<<148, 384>>>(out, 1); }));
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
