// fp64_issue.cu -- what the FP64 pipe of one B200 SM really issues (round 2, VERDICT "What's weak" #3).
//
// The roofline of K1 (photon x cell scan: DADD + DSETP per dimension) is quoted against the hardware FP64 issue peak,
// 148 SMs x 64 lanes x SM clock.  The DFMA probe of round 1 reached 91 % of that figure; this tool separates the
// two possible reasons -- a lower SM clock under FP64 load, or an issue limit of the pipe -- by timing each kernel
// with both the SM cycle counter (clock64) and the nanosecond global timer, for several instruction mixes and
// numbers of resident warps:
//     dfma   : 16 independent DFMA chains per thread           (3 source operands)
//     dadd   : 16 independent DADD chains per thread           (2 source operands)
//     dmul   : 16 independent DMUL chains
//     scan3  : K1's own 3-D mix: per "eval" 3 DADD (d = x - c) + 3 DSETP (|d| <= h, chained), 8 photons per thread,
//              cell operands from shared memory (broadcast LDS.128), rare predicated hit
// Output: one line per (mix, warps/SM): thread-instr per clock per SM (64 = nominal), SM MHz seen by the kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/fp64_issue tools/micro/fp64_issue.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct Stamp {
    long long c0, c1;
    unsigned long long t0, t1;
};

template <int OP>
__global__ void __launch_bounds__(256) chains(double *out, int iters, double x, double y, Stamp *st)
{
    double a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = 1.0 + 1e-3 * (threadIdx.x + j);
    long long c0 = clock64();
    unsigned long long t0 = gtimer();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (OP == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[j]) : "d"(x), "d"(y));
            if (OP == 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[j]) : "d"(y));
            if (OP == 2) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[j]) : "d"(x));
        }
    }
    long long c1 = clock64();
    unsigned long long t1 = gtimer();
    double s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += a[j];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0) {
        st[blockIdx.x].c0 = c0; st[blockIdx.x].c1 = c1; st[blockIdx.x].t0 = t0; st[blockIdx.x].t1 = t1;
    }
}

// K1's 3-D inner loop on a synthetic tile: P photons per thread, cells broadcast from shared memory.
// PTX = 1: the containment test as the library writes it (chained DSETP + one predicated minimum);
// PTX = 0: the C++ form (three independent DSETP + VIMNMX + two SEL per test).
template <int P, int PTX>
__global__ void __launch_bounds__(128) scan3(int *out, int tiles, Stamp *st)
{
    __shared__ double4 sA[256];
    __shared__ double2 sB[256];
    for (int c = threadIdx.x; c < 256; c += blockDim.x) {
        sA[c] = make_double4(1.0 + c, 2.0 + c, 3.0 + c, 1e-9);
        sB[c] = make_double2(1e-9, 1e-9);
    }
    __syncthreads();
    double x0[P], x1[P], x2[P];
    int best[P];
#pragma unroll
    for (int p = 0; p < P; ++p) { // every coordinate differs per photon and thread: nothing to share between tests
        x0[p] = 0.5 + threadIdx.x + 1000.0 * p;
        x1[p] = 0.25 + 3.0 * threadIdx.x + 17.0 * p;
        x2[p] = 0.125 + 5.0 * threadIdx.x + 29.0 * p;
        best[p] = 0x7fffffff;
    }
    long long c0 = clock64();
    unsigned long long t0 = gtimer();
    for (int t = 0; t < tiles; ++t) {
#pragma unroll 8
        for (int c = 0; c < 256; ++c) {
            const double4 a = sA[c];
            const double2 b = sB[c];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                if (PTX) {
                    const double d0 = x0[p] - a.x, d1 = x1[p] - a.y, d2 = x2[p] - a.z;
                    asm("{\n\t.reg .pred p;\n\t.reg .f64 e0, e1, e2;\n\tabs.f64 e0, %1;\n\tabs.f64 e1, %2;\n\tabs.f64 e2, %3;\n\t"
                        "setp.le.f64 p, e0, %4;\n\tsetp.le.and.f64 p, e1, %5, p;\n\tsetp.le.and.f64 p, e2, %6, p;\n\t"
                        "@p min.s32 %0, %0, %7;\n\t}"
                        : "+r"(best[p])
                        : "d"(d0), "d"(d1), "d"(d2), "d"(a.w), "d"(b.x), "d"(b.y), "r"(t * 256 + c));
                } else {
                    bool hit = (fabs(x0[p] - a.x) <= a.w) & (fabs(x1[p] - a.y) <= b.x) & (fabs(x2[p] - a.z) <= b.y);
                    if (hit) best[p] = min(best[p], t * 256 + c);
                }
            }
        }
    }
    long long c1 = clock64();
    unsigned long long t1 = gtimer();
    int s = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) s ^= best[p];
    if (s == 12345) out[0] = s;
    if (threadIdx.x == 0) {
        st[blockIdx.x].c0 = c0; st[blockIdx.x].c1 = c1; st[blockIdx.x].t0 = t0; st[blockIdx.x].t1 = t1;
    }
}

static void report(const char *name, int warps_per_sm, double instr_per_thread, int threads_per_sm, Stamp *hst, int nblocks,
                   float ms, int nsm)
{
    // SM clock the kernel saw: cycles / nanoseconds of every block's own stamps
    double cyc = 0, ns = 0;
    for (int b = 0; b < nblocks; ++b) {
        cyc += (double)(hst[b].c1 - hst[b].c0);
        ns += (double)(hst[b].t1 - hst[b].t0);
    }
    const double mhz = cyc / ns * 1e3;
    const double ginstr = instr_per_thread * threads_per_sm * nsm / (ms * 1e-3) / 1e9; // whole kernel, CUDA events
    const double nominal = nsm * 64.0 * mhz * 1e6 / 1e9;
    printf("%-12s warps/SM %3d : %8.1f G FP64-pipe thread-instr/s (%.3f ms) = %5.1f %% of %d SMs x 64 lanes x %.0f MHz (clock seen by the kernel)\n",
           name, warps_per_sm, ginstr, ms, 100.0 * ginstr / nominal, nsm, mhz);
}

int main(int argc, char **argv)
{
    int dev = 0;
    cudaSetDevice(dev);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    const int nsm = prop.multiProcessorCount;
    printf("%s, %d SMs, max SM clock %.0f MHz\n", prop.name, nsm, prop.clockRate / 1e3);
    double *out;
    int *iout;
    Stamp *st, *hst;
    cudaMalloc(&out, 8);
    cudaMalloc(&iout, 4);
    cudaMalloc(&st, sizeof(Stamp) * nsm * 32);
    hst = (Stamp *)malloc(sizeof(Stamp) * nsm * 32);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 1 << 15;
    const char *names[3] = {"dfma", "dadd", "dmul"};
    for (int op = 0; op < 3; ++op)
        for (int bps = 1; bps <= 8; bps *= 2) { // blocks of 256 threads per SM: 8, 16, 32, 64 warps
            const int nblocks = nsm * bps;
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (op == 0) chains<0><<<nblocks, 256>>>(out, iters, 1.0000001, 1e-9, st);
                if (op == 1) chains<1><<<nblocks, 256>>>(out, iters, 1.0000001, 1e-9, st);
                if (op == 2) chains<2><<<nblocks, 256>>>(out, iters, 1.0000001, 1e-9, st);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            cudaMemcpy(hst, st, sizeof(Stamp) * nblocks, cudaMemcpyDeviceToHost);
            report(names[op], bps * 8, 16.0 * iters, bps * 256, hst, nblocks, best, nsm);
        }
    // K1's mix: 6 FP64-pipe instructions per photon-cell eval
    for (int variant = 0; variant < 4; ++variant)
        for (int bps = 2; bps <= 8; bps += 2) {
            const int nblocks = nsm * bps, tiles = 64;
            const int P = (variant & 1) ? 9 : 8;
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (variant == 0) scan3<8, 0><<<nblocks, 128>>>(iout, tiles, st);
                if (variant == 1) scan3<9, 0><<<nblocks, 128>>>(iout, tiles, st);
                if (variant == 2) scan3<8, 1><<<nblocks, 128>>>(iout, tiles, st);
                if (variant == 3) scan3<9, 1><<<nblocks, 128>>>(iout, tiles, st);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            cudaMemcpy(hst, st, sizeof(Stamp) * nblocks, cudaMemcpyDeviceToHost);
            const char *nm[4] = {"scan3 P8 C++", "scan3 P9 C++", "scan3 P8 PTX", "scan3 P9 PTX"};
            report(nm[variant], bps * 4, 6.0 * P * 256 * tiles, bps * 128, hst, nblocks, best, nsm);
        }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        printf("CUDA error: %s\n", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
