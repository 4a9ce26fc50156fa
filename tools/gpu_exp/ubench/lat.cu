// Dependent-chain latencies on B200 (one warp, clock64 around N dependent operations).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int OP>
__global__ void chain(double a, double b, unsigned* out, long long* cyc, const unsigned long long* ptrs) {
    double x = a; unsigned u = (unsigned)a; unsigned long long p = (unsigned long long)ptrs;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = x + b;
        if (OP == 1) x = x * b;
        if (OP == 2) x = fma(x, b, a);
        if (OP == 3) { u = __double2uint_rz(x); x = __longlong_as_double(__double_as_longlong(x) + (u & 1)); }   // F2I + 1 int op + reinterpret
        if (OP == 4) u = u * 2069u + 7u;                       // IMAD
        if (OP == 5) u = (u >> 3) ^ 0x9e3779b9u;               // SHF/LOP3
        if (OP == 6) p = *(const unsigned long long*)p;                                      // pointer chase, ld.global
        if (OP == 7) p = __ldg((const unsigned long long*)p);                                // pointer chase, ld.global.nc
        if (OP == 8) { asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(p) : "l"(p)); }      // L2 only
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; out[0] = u + (unsigned)x + (unsigned)p; }
}
int main() {
    unsigned* out; long long* cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
    // pointer-chase rings: small (L1-resident, 16 KB), medium (L2-resident, 16 MB, stride 4 KB+128), large (1 GB, DRAM)
    const char* names[] = {"DADD", "DMUL", "DFMA", "F2I.U32.F64 (+IADD64)", "IMAD", "SHF+LOP3", "ld.global", "ld.global.nc", "ld.global.cg"};
    for (int op = 0; op < 6; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (op) {
                case 0: chain<0><<<1, 32>>>(1.5, 1e-9, out, cyc, nullptr); break;
                case 1: chain<1><<<1, 32>>>(1.5, 1.0000001, out, cyc, nullptr); break;
                case 2: chain<2><<<1, 32>>>(1.5, 0.999, out, cyc, nullptr); break;
                case 3: chain<3><<<1, 32>>>(123456.7, 1.0, out, cyc, nullptr); break;
                case 4: chain<4><<<1, 32>>>(3.0, 1.0, out, cyc, nullptr); break;
                case 5: chain<5><<<1, 32>>>(3.0, 1.0, out, cyc, nullptr); break;
            }
            cudaDeviceSynchronize();
        }
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-24s %.1f cycles per op\n", names[op], (double)h / N);
    }
    struct Ring { const char* name; size_t bytes; size_t stride; } rings[] = {
        {"16 KB ring (L1)", 16 << 10, 136}, {"8 MB ring (L2)", 8 << 20, 4224}, {"64 MB ring (L2, both dies)", 64 << 20, 33 * 1024 + 128},
        {"2 GB ring (DRAM)", (size_t)2 << 30, (size_t)1 << 20 | 4224}};
    for (auto& r : rings) {
        size_t n = r.bytes / 8, step = r.stride / 8;
        unsigned long long* d; cudaMalloc(&d, r.bytes);
        // ring: element i points to element (i + step) % n  (only the visited elements matter)
        size_t hops = 4 * N; unsigned long long* h = (unsigned long long*)malloc(hops * 16);
        size_t idx = 0;
        for (size_t k = 0; k < hops; ++k) { size_t nxt = (idx + step) % n; unsigned long long v = (unsigned long long)(d + nxt); cudaMemcpy(d + idx, &v, 8, cudaMemcpyHostToDevice); idx = nxt; }
        for (int op = 6; op <= 8; ++op) {
            for (int rep = 0; rep < 3; ++rep) {
                if (op == 6) chain<6><<<1, 32>>>(0, 0, out, cyc, d);
                if (op == 7) chain<7><<<1, 32>>>(0, 0, out, cyc, d);
                if (op == 8) chain<8><<<1, 32>>>(0, 0, out, cyc, d);
                cudaDeviceSynchronize();
            }
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-28s %-14s %.1f cycles per load (warm)\n", r.name, names[op], (double)c / N);
        }
        cudaFree(d); free(h);
    }
    return 0;
}
