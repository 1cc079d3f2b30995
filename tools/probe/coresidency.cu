// Probe: when do CTAs of two kernels on different streams share an SM?  Kernel A needs a lot of shared memory, kernel B none.
// Each kernel spins for a fixed number of clocks with few CTAs per SM; if they overlap, the pair takes ~max, else ~sum.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__global__ void spinA(long long cycles, int *sink) {
    extern __shared__ int sm[];
    sm[threadIdx.x] = threadIdx.x;
    long long t0 = clock64();
    while (clock64() - t0 < cycles) { }
    if (sm[threadIdx.x] == -1) *sink = 1;
}
__global__ void spinB(long long cycles, int *sink) {
    long long t0 = clock64();
    while (clock64() - t0 < cycles) { }
    if (t0 == -1) *sink = 1;
}
__global__ void spinB2(long long cycles, int *sink) {     // same, with a token dynamic shared allocation
    extern __shared__ int sm[];
    sm[threadIdx.x] = 1;
    long long t0 = clock64();
    while (clock64() - t0 < cycles) { }
    if (sm[threadIdx.x] == -1) *sink = 1;
}
static float run(int order, int smemA, bool useB2, int smemB, int ctasA, int ctasB, int sms) {
    cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    cudaEvent_t e0, e1, ea, eb; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&ea); cudaEventCreate(&eb);
    int *sink; cudaMalloc(&sink, 4);
    const long long cyc = 2000000;   // ~1 ms
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaDeviceSynchronize();
        cudaEventRecord(e0, s1);
        cudaStreamWaitEvent(s2, e0, 0);
        if (order == 0) {
            spinA<<<ctasA * sms, 128, smemA, s1>>>(cyc, sink);
            if (useB2) spinB2<<<ctasB * sms, 128, smemB, s2>>>(cyc, sink); else spinB<<<ctasB * sms, 128, 0, s2>>>(cyc, sink);
        } else {
            if (useB2) spinB2<<<ctasB * sms, 128, smemB, s2>>>(cyc, sink); else spinB<<<ctasB * sms, 128, 0, s2>>>(cyc, sink);
            spinA<<<ctasA * sms, 128, smemA, s1>>>(cyc, sink);
        }
        cudaEventRecord(eb, s2);
        cudaStreamWaitEvent(s1, eb, 0);
        cudaEventRecord(e1, s1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    cudaFuncSetAttribute(spinA, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(spinB2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("SMs %d; single kernel ~ %.3f ms\n", sms, run(0, 100 * 1024, false, 0, 1, 0 + 1, sms) / 1);   // (pair, see below)
    struct { const char *name; int hintA, hintB; bool b2; int smemB; int smemA; } cases[] = {
        {"A 100KB, B none, no hints", -1, -1, false, 0, 100 * 1024},
        {"A 100KB, B none, both hint MaxShared", 100, 100, false, 0, 100 * 1024},
        {"A 100KB, B 1KB dyn, no hints", -1, -1, true, 1024, 100 * 1024},
        {"A 100KB, B 1KB dyn, both hint MaxShared", 100, 100, true, 1024, 100 * 1024},
        {"A 100KB, B none, A hint 50 B hint 50", 50, 50, false, 0, 100 * 1024},
        {"A 24KB, B none, no hints", -1, -1, false, 0, 24 * 1024},
        {"A 24KB, B none, both hint 50", 50, 50, false, 0, 24 * 1024},
        {"A 24KB, B none, both hint MaxShared", 100, 100, false, 0, 24 * 1024},
        {"A 24KB, B none, both hint MaxL1(0)", 0, 0, false, 0, 24 * 1024},
    };
    for (auto &c : cases) {
        cudaFuncSetAttribute(spinA, cudaFuncAttributePreferredSharedMemoryCarveout, c.hintA);
        cudaFuncSetAttribute(spinB, cudaFuncAttributePreferredSharedMemoryCarveout, c.hintB);
        cudaFuncSetAttribute(spinB2, cudaFuncAttributePreferredSharedMemoryCarveout, c.hintB);
        for (int order = 0; order < 2; order++) {
            float ms = run(order, c.smemA, c.b2, c.smemB, 1, 1, sms);
            printf("%-44s %s first: %.3f ms  (%s)\n", c.name, order == 0 ? "A" : "B", ms, ms < 1.6 ? "overlap" : "serial");
        }
    }
    return 0;
}
