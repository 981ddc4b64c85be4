// Issue-rate microbenchmark for the sm_100a pipes the field arithmetic and Blake2s lean on.
// Each kernel runs ITER iterations of U independent dependent-chains per thread; result = lane-ops / clk / SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench tools/pipe_bench.cu && ./pipe_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048
#define U 8

template <int MODE>
__global__ void __launch_bounds__(256) k(uint64_t *out, unsigned long long *cyc, uint32_t seed) {
    uint32_t a[U], b[U];
    uint64_t w[U];
    double d[U], e[U];
#pragma unroll
    for (int i = 0; i < U; i++) {
        a[i] = seed * (i + 1) + threadIdx.x;
        b[i] = seed ^ (i * 77 + threadIdx.x);
        w[i] = ((uint64_t)a[i] << 32) | b[i];
        d[i] = 1.0 + a[i] * 1e-9;
        e[i] = 1.0 + b[i] * 1e-9;
    }
    double dm = 1.0 + seed * 1e-12, da = seed * 1e-13;
    uint32_t m = seed | 1, c0 = seed >> 3;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < U; i++) {
                if (MODE == 0) asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w[i]) : "r"(m));
                if (MODE == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(c0));
                if (MODE == 2) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(c0));
                if (MODE == 3) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dm), "d"(da));
                if (MODE == 4) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dm), "d"(da));
                if (MODE == 5) asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(a[i]) : "r"(m), "r"(b[i]));  // -> IADD3
                if (MODE == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(m), "r"(b[i]));
                if (MODE == 7) asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((uint64_t)m << 20 | c0));
                if (MODE == 8) asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(a[i]));
                if (MODE == 9) {  // 1 DFMA : 1 IMAD.WIDE
                    asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dm), "d"(da));
                    asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w[i]) : "r"(m));
                }
                if (MODE == 10) {  // 1 DFMA : 2 IADD3
                    asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dm), "d"(da));
                    asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(a[i]) : "r"(m), "r"(b[i]));
                    asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(b[i]) : "r"(m), "r"(c0));
                }
                if (MODE == 11) {  // 1 IMAD.WIDE : 2 IADD3
                    asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w[i]) : "r"(m));
                    asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(a[i]) : "r"(m), "r"(b[i]));
                    asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(b[i]) : "r"(m), "r"(c0));
                }
                if (MODE == 12) {  // carry chain pair as the Montgomery code emits: mad.lo.cc + madc.hi.cc
                    uint32_t lo = (uint32_t)w[i], hi = (uint32_t)(w[i] >> 32);
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a[i]), "r"(m));
                    w[i] = ((uint64_t)hi << 32) | lo;
                }
                if (MODE == 13) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float *)&a[i]) : "f"(*(float *)&m), "f"(*(float *)&c0));
                if (MODE == 14) {  // 64-bit add of a DFMA result (the Emmart accumulate): 1 DFMA + 1 add.u64
                    asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(e[i]) : "d"(d[i]), "d"(dm), "d"(da));
                    asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(__double_as_longlong(e[i])));
                }
                if (MODE == 15) asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; xor.b32 lo, lo, hi; mul.wide.u32 %0, lo, %1;}" : "+l"(w[i]) : "r"(m));
                if (MODE == 16) {  // 2 DFMA : 1 IMAD.WIDE : 2 IADD3
                    asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dm), "d"(da));
                    asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(e[i]) : "d"(dm), "d"(da));
                    asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w[i]) : "r"(m));
                    asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(a[i]) : "r"(m), "r"(b[i]));
                    asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(b[i]) : "r"(m), "r"(c0));
                }
            }
        }
    }
    long long t1 = clock64();
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < U; i++) acc += a[i] + b[i] + w[i] + (uint64_t)__double_as_longlong(d[i]) + (uint64_t)__double_as_longlong(e[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int MODE>
void run(const char *name, double ops_per_slot, int warps_per_sm) {
    int sms = 148, threads = 256;
    int blocks = sms * (warps_per_sm * 32 / threads);
    uint64_t *out;
    unsigned long long *cyc;
    cudaMalloc(&out, (size_t)blocks * threads * 8);
    cudaMalloc(&cyc, blocks * 8);
    k<MODE><<<blocks, threads>>>(out, cyc, 12345);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, cyc, 12345);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long *h = new unsigned long long[blocks];
    cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < blocks; i++) mx = h[i] > mx ? h[i] : mx;
    double slots = (double)ITER * 4 * U * threads * (blocks / sms);   // per SM
    printf("%-44s warps/SM %2d  %7.2f slots/clk/SM  (%6.2f lane-ops/clk/SM)  %.3f ms  clk %.0f MHz\n", name, warps_per_sm,
           slots / mx, slots * ops_per_slot / mx, ms, mx / ms / 1e3);
    cudaFree(out);
    cudaFree(cyc);
    delete[] h;
}

int main() {
    for (int w : {16, 32}) {
        run<0>("IMAD.WIDE.U32 (mad.wide.u32 acc)", 1, w);
        run<15>("mul.wide.u32", 1, w);
        run<1>("IMAD lo (mad.lo.u32)", 1, w);
        run<2>("IMAD.HI (mad.hi.u32)", 1, w);
        run<12>("mad.lo.cc+madc.hi.cc pair", 1, w);
        run<3>("DFMA rn", 1, w);
        run<4>("DFMA rz", 1, w);
        run<13>("FFMA", 1, w);
        run<5>("IADD3", 1, w);
        run<6>("LOP3", 1, w);
        run<8>("SHF", 1, w);
        run<7>("add.u64", 1, w);
        run<9>("1 DFMA + 1 IMAD.WIDE", 2, w);
        run<10>("1 DFMA + 2 IADD3", 3, w);
        run<11>("1 IMAD.WIDE + 2 IADD3", 3, w);
        run<14>("1 DFMA + 1 add.u64", 2, w);
        run<16>("2 DFMA + 1 IMAD.WIDE + 2 IADD3", 5, w);
    }
    return 0;
}
