// Microbenchmark: sustained issue rate per SM of the packed FP32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2, two
// float lanes per thread per instruction) next to their scalar forms, including the operand forms the AoS complex FFT
// codelets use (swapped halves with per-lane sign, splat immediate, broadcast scalar register).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_issue f32x2_issue.cu && ./f32x2_issue
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float2* out, int iters, float a, float b) {
    float2 x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float2 y = x[(i + 5) & 15];
                if (MODE == 0) x[i] = __fadd2_rn(x[i], y);                                              // FADD2 r,r
                if (MODE == 1) x[i] = __ffma2_rn(x[i], make_float2(a, a), y);                           // FFMA2 r, r.F32 (broadcast), r
                if (MODE == 2) x[i] = __ffma2_rn(x[i], make_float2(0.99991f, 0.99991f), y);             // FFMA2 r, imm, r
                if (MODE == 3) x[i] = __ffma2_rn(make_float2(-y.y, y.x), make_float2(0.4142f, 0.4142f), x[i]);  // FFMA2 r.LO_HI.NP, imm, r
                if (MODE == 4) x[i] = __fmul2_rn(x[i], make_float2(b, b));                              // FMUL2 r, r.F32
                if (MODE == 5) x[i] = __fadd2_rn(x[i], make_float2(-y.y, y.x));                         // FADD2 r, r.LO_HI.NP
                if (MODE == 6) { x[i].x = x[i].x + y.x; x[i].y = x[i].y + y.y; }                        // 2 x FADD (counted as 2)
                if (MODE == 7) { x[i].x = fmaf(x[i].x, a, y.x); x[i].y = fmaf(x[i].y, a, y.y); }        // 2 x FFMA
                if (MODE == 8) x[i] = __ffma2_rn(x[i], y, x[(i + 9) & 15]);                             // FFMA2 r,r,r (3 register pairs)
            }
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 16; ++i) s = make_float2(s.x + x[i].x, s.y + x[i].y);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int per) {
    float2* out;
    cudaMalloc(&out, 148 * 1024 * sizeof(float2));
    const int iters = 2000;
    for (int warps : {4, 8, 12, 16, 32}) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        k<MODE><<<148, warps * 32>>>(out, 10, 1.0001f, 0.9999f);
        cudaEventRecord(e0);
        k<MODE><<<148, warps * 32>>>(out, iters, 1.0001f, 0.9999f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double instr = double(iters) * 8 * 16 * warps * per;  // warp-instructions per SM
        printf("%-34s warps/SM %2d : %.3f warp-instr/clk/SM = %.1f float-lane-ops/clk/SM (at 1.965 GHz)\n", name, warps,
               instr / (ms * 1e-3 * 1.965e9), instr / (ms * 1e-3 * 1.965e9) * 32 * (per == 1 ? 2 : 1));
    }
    cudaFree(out);
}

int main() {
    run<0>("FADD2 r,r", 1);
    run<1>("FFMA2 r,r.F32,r", 1);
    run<2>("FFMA2 r,imm,r", 1);
    run<3>("FFMA2 r.LO_HI.NP,imm,r", 1);
    run<4>("FMUL2 r,r.F32", 1);
    run<5>("FADD2 r,r.LO_HI.NP", 1);
    run<8>("FFMA2 r,r,r", 1);
    run<6>("2 x FADD r,r", 2);
    run<7>("2 x FFMA r,r,r", 2);
    return 0;
}
