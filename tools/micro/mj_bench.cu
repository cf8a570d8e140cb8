// Micro-benchmark: where the Maxwell-Juttner electron sampling of one scattering spends its cycles
// (single thread, as on the scattering lane): bessel_K2(1/theta) vs the rejection loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I mcrat_b200/csrc tools/micro/mj_bench.cu -o /tmp/mj_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "device_math.cuh"
using namespace mcrat;

__global__ void bench(const double *theta, int n, long long *out)
{
    if (threadIdx.x || blockIdx.x) return;
    for (int k = 0; k < n; ++k) {
        const double factor = theta[k];
        long long t0 = clock64();
        double k2 = bessel_K2(1.0 / factor);
        long long t1 = clock64();
        long long trials = 0, cyc = 0;
        for (int rep = 0; rep < 64; ++rep) {
            EventRng rng;
            rng.replay = 0; rng.k0 = 1; rng.k1 = 2; rng.iter = rep; rng.draw = 0; rng.buf = nullptr; rng.pos = 0; rng.n = 0;
            rng.exhausted = 0; rng.pre = nullptr; rng.npre = 0;
            long long a = clock64();
            double y = 1, f = 0, x = 0;
            while (((f != f) || (y > f))) {
                x = rng.uniform_pos() * (1 + 100 * factor);
                double bx = sqrt(1 - (1 / (x * x)));
                y = rng.uniform() / 2.0;
                f = x * x * (bx / k2) * exp(-1 * x / factor);
                trials++;
            }
            cyc += clock64() - a;
            if (x == 12345.678) out[0] = 1;
        }
        out[3 * k + 0] = t1 - t0;
        out[3 * k + 1] = cyc / 64;
        out[3 * k + 2] = trials;
        if (k2 == 0.12345) out[0] = 2;
    }
}

int main()
{
    const int n = 9;
    double th[n] = {1.7e-3, 3e-3, 1e-2, 3e-2, 0.1, 0.3, 0.5, 0.7, 1.0};
    double *d; long long *o, h[3 * n];
    cudaMalloc(&d, sizeof(th)); cudaMalloc(&o, sizeof(h));
    cudaMemcpy(d, th, sizeof(th), cudaMemcpyHostToDevice);
    bench<<<1, 32>>>(d, n, o);
    cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost);
    for (int k = 0; k < n; ++k)
        printf("theta %-8g K2: %8lld cycles   rejection loop: %8lld cycles per electron, %.1f trials per electron\n", th[k], h[3 * k],
               h[3 * k + 1], h[3 * k + 2] / 64.0);
    return cudaDeviceSynchronize() != cudaSuccess;
}
