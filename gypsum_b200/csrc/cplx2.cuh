// Complex arithmetic on float2 = (re, im) with the PACKED FP32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2: two float
// lanes per thread per instruction).  A complex add is ONE instruction, a rotation by +-j or a real-times-complex
// multiply-add ONE (the half swap and the per-lane sign ride on the operand: `R.F32x2.LO_HI.NP`, a splat immediate or a
// broadcast scalar register `R.F32` cost nothing), a complex product TWO -- half the issue slots of the scalar forms at the
// same lane throughput (tools/microbench/f32x2_issue.cu: 1.97 packed vs 3.85 scalar warp-instructions / clk / SM).
//
// Every lane of every function is one IEEE round-to-nearest add / mul / fma, so the host versions below (used by the lane
// emulator, tests/emu) give bit-identical results, and so did the scalar re[] / im[] code these replaced.
#pragma once
#include "gb_common.cuh"

namespace gb {

#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
#define GB_ADD2(a, b) __fadd2_rn(a, b)
#define GB_MUL2(a, b) __fmul2_rn(a, b)
#define GB_FMA2(a, b, c) __ffma2_rn(a, b, c)
#else
GB_HD GB_INLINE float2 gb_host_add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
GB_HD GB_INLINE float2 gb_host_mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
GB_HD GB_INLINE float2 gb_host_fma2(float2 a, float2 b, float2 c) {
    return make_float2(__builtin_fmaf(a.x, b.x, c.x), __builtin_fmaf(a.y, b.y, c.y));
}
#define GB_ADD2(a, b) gb_host_add2(a, b)
#define GB_MUL2(a, b) gb_host_mul2(a, b)
#define GB_FMA2(a, b, c) gb_host_fma2(a, b, c)
#endif

GB_HD GB_INLINE float2 c_add(float2 a, float2 b) { return GB_ADD2(a, b); }                                  // a + b
GB_HD GB_INLINE float2 c_sub(float2 a, float2 b) { return GB_ADD2(a, make_float2(-b.x, -b.y)); }            // a - b
GB_HD GB_INLINE float2 c_add_mj(float2 a, float2 b) { return GB_ADD2(a, make_float2(b.y, -b.x)); }          // a + (-j) b
GB_HD GB_INLINE float2 c_sub_mj(float2 a, float2 b) { return GB_ADD2(a, make_float2(-b.y, b.x)); }          // a - (-j) b = a + j b
GB_HD GB_INLINE float2 c_fma(float c, float2 x, float2 y) { return GB_FMA2(x, make_float2(c, c), y); }      // y + c x   (c real)
GB_HD GB_INLINE float2 c_fma_j(float c, float2 x, float2 y) {                                               // y + c (j x)
    return GB_FMA2(make_float2(-x.y, x.x), make_float2(c, c), y);
}
GB_HD GB_INLINE float2 c_scale(float c, float2 x) { return GB_MUL2(x, make_float2(c, c)); }                 // c x
// Complex products, two packed instructions each: FMUL2 of the rotated operand (half swap + per-lane sign ride on the
// operand as `R.F32x2.LO_HI.NP`) by the broadcast imaginary part (`R.F32`), then FFMA2 by the broadcast real part.  The
// modifiers sit on the DATA operand, the twiddle / spectrum factor is only ever read as two broadcast scalars, so no
// operand has to be materialised.
// z * w = (fma(zr, wr, -(zi wi)), fma(zi, wr, zr wi))
GB_HD GB_INLINE float2 cmul(float2 z, float2 w) {
    const float2 p = GB_MUL2(make_float2(-z.y, z.x), make_float2(w.y, w.y));
    return GB_FMA2(z, make_float2(w.x, w.x), p);
}
// z * conj(w) = (fma(zr, wr, zi wi), fma(zi, wr, -(zr wi)))
GB_HD GB_INLINE float2 cmulc(float2 z, float2 w) {
    const float2 p = GB_MUL2(make_float2(z.y, -z.x), make_float2(w.y, w.y));
    return GB_FMA2(z, make_float2(w.x, w.x), p);
}
// e + o * conj(w) = (fma(or, wr, fma(oi, wi, er)), fma(oi, wr, fma(-or, wi, ei)))
GB_HD GB_INLINE float2 cfmac(float2 e, float2 o, float2 w) {
    const float2 t = GB_FMA2(make_float2(o.y, -o.x), make_float2(w.y, w.y), e);
    return GB_FMA2(o, make_float2(w.x, w.x), t);
}

}  // namespace gb
