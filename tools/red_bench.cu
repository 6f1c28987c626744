// Microbenchmark: throughput of global reductions (RED) on B200 for the access pattern of the scatter kernels:
// random runs of `run` consecutive slots in an array of `n_slots` elements.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <class T>
__global__ void k_red(T* arr, long long n_slots, long long n_items, int run, unsigned seed) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_items * run; i += (long long)gridDim.x * blockDim.x) {
        const long long item = i / run;
        const int j = (int)(i % run);
        unsigned long long h = (unsigned long long)item * 0x9E3779B97F4A7C15ull + seed;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        const long long s = (long long)(h % (unsigned long long)(n_slots - run)) + j;
        atomicAdd(&arr[s], (T)1);
    }
}
template <class T>
void run_one(const char* name, long long n_slots, long long n_items, int run) {
    T* arr; cudaMalloc(&arr, n_slots * sizeof(T)); cudaMemset(arr, 0, n_slots * sizeof(T));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_red<T><<<148 * 8, 256>>>(arr, n_slots, n_items, run, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int it = 0; it < 3; ++it) k_red<T><<<148 * 8, 256>>>(arr, n_slots, n_items, run, 7 + it);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
    printf("%-6s slots=%lldM run=%d: %.3f ms for %lldM REDs -> %.1f G RED/s\n", name, n_slots >> 20, run, ms, (n_items * run) >> 20,
           n_items * run / (ms * 1e6));
    cudaFree(arr);
}
int main() {
    const long long n_items = 40ll << 20;
    for (long long slots : {7ll << 20, 64ll << 20}) {
        for (int run : {1, 6}) {
            run_one<double>("f64", slots, n_items / run, run);
            run_one<float>("f32", slots, n_items / run, run);
            run_one<unsigned>("u32", slots, n_items / run, run);
            run_one<unsigned long long>("u64", slots, n_items / run, run);
        }
    }
    return 0;
}
