// Lane-level building blocks of the warp FFT-1024 and of the polyphase correlation.
//
// Every function here is a pure function of (lane, registers, shared tile) with no warp intrinsics, so the
// same code is compiled for the device kernels (kernels.cu) and for the host lane emulator (tests/emu), which
// runs the 32 lanes of a warp as a loop.  Synchronisation is the caller's job.
//
// Math.  A 1-ms replica at N = s*1023 samples is chips repeated s times
// (reference satellite.py:20-31), so the length-N circular correlation of utils.py:59-73 splits exactly into
// s circular correlations of length 1023 against the +-1 chip sequence:
//     corr[s*q + r] = sum_m z_r[m] * c[(m - q) mod 1023],   z_r[m] = sum_{t<s} y[(s*m + r + t) mod N].
// Each length-1023 circular correlation is carried, exactly, by a zero-padded length-2048 linear one, and the
// 2048-point transforms are split radix-2 into two 1024-point warp transforms:
//     forward (input half zero):  Z[2f] = FFT1024(z)[f],  Z[2f+1] = FFT1024(z * W2048^n)[f]
//     inverse (only k < 1024):    out[k] = IFFT1024(Y_even)[k] + W2048^-k * IFFT1024(Y_odd)[k]
// A warp transform holds x[lane + 32 j] in registers: FFT-32 over j, twiddle W1024^(lane*k1), 32x32 transpose
// through a padded shared tile, FFT-32 over lane.  Output X[lane + 32 k2] lands in the same layout.
#pragma once
#include "fft32_gen.cuh"
#include "gb_common.cuh"

namespace gb {

GB_HD GB_INLINE float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
GB_HD GB_INLINE float2 cmulc(float2 a, float2 b) {  // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

// Phase 1 of the forward warp FFT-1024.  re/im[j] = x[lane + 32 j].  Writes u[lane][k1]*W1024^(lane k1) to the
// tile, row k1, column lane.  tw1[k1*32 + lane] = exp(-2 pi i lane k1 / 1024).
GB_HD GB_INLINE void wfft_phase1(float (&re)[32], float (&im)[32], int lane, const float2* tw1, float2* tile) {
    fft32_fwd(re, im);
    tile[lane] = make_float2(re[0], im[0]);
#pragma unroll
    for (int k1 = 1; k1 < 32; ++k1) {
        const float2 w = tw1[k1 * 32 + lane];
        tile[k1 * kTStride + lane] = make_float2(re[k1] * w.x - im[k1] * w.y, re[k1] * w.y + im[k1] * w.x);
    }
}
// Phase 2: thread `lane` owns column k1 = lane: reads u[l][lane], l = 0..31, FFT-32 over l.  Afterwards
// re/im[k2] = X[lane + 32 k2].
GB_HD GB_INLINE void wfft_phase2(float (&re)[32], float (&im)[32], int lane, const float2* tile) {
#pragma unroll
    for (int l = 0; l < 32; ++l) {
        const float2 v = tile[lane * kTStride + l];
        re[l] = v.x;
        im[l] = v.y;
    }
    fft32_fwd(re, im);
}

// Carrier wipe-off of one sample (utils.py:93-97 / tracker.py:278-281): x * exp(-j 2 pi cycles), with the
// phase reduced to [-0.5, 0.5] cycles in float64 BEFORE going to float32.
GB_HD GB_INLINE float2 wipeoff(float2 x, double cycles) {
#if defined(__CUDA_ARCH__)
    const double fr = cycles - rint(cycles);
    float s, c;
    sincospif(2.0f * static_cast<float>(fr), &s, &c);
#else
    const double fr = cycles - __builtin_rint(cycles);
    const double a = 6.283185307179586476925 * static_cast<double>(static_cast<float>(fr));
    const float s = static_cast<float>(__builtin_sin(a)), c = static_cast<float>(__builtin_cos(a));
#endif
    return make_float2(x.x * c + x.y * s, x.y * c - x.x * s);  // x * (c - j s)
}

// Polyphase boxcar: z_r[m] for m = lane + 32 j, from the wiped-off millisecond stored as ypoly[t][m'] =
// y[s*m' + t] with row length 1024 and ypoly[t][1023] = ypoly[t][0] (the circular wrap).
GB_HD GB_INLINE void build_z(float (&re)[32], float (&im)[32], int lane, int r, int s, const float2* ypoly) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int m = lane + 32 * j;
        float ar = 0.f, ai = 0.f;
        if (m < kChips) {
            for (int t = 0; t < s; ++t) {
                const int rt = r + t;
                const int row = rt >= s ? rt - s : rt;
                const float2 v = ypoly[row * kFft + m + (rt >= s ? 1 : 0)];
                ar += v.x;
                ai += v.y;
            }
        }
        re[j] = ar;
        im[j] = ai;
    }
}

// x[n] *= W2048^n (forward odd half) for n = lane + 32 j;  tw2[n] = exp(-2 pi i n / 2048), n < 1024.
GB_HD GB_INLINE void mul_tw2(float (&re)[32], float (&im)[32], int lane, const float2* tw2) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float2 w = tw2[lane + 32 * j];
        const float a = re[j], b = im[j];
        re[j] = a * w.x - b * w.y;
        im[j] = a * w.y + b * w.x;
    }
}
// Fast magnitude: MUFU.SQRT (sqrt.approx, ~1 ulp) on the device instead of the IEEE sequence with its slow path.
GB_HD GB_INLINE float gb_sqrt(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return __builtin_sqrtf(x);
#endif
}

// Radix-2 recombination of the two inverse half transforms, out[k] = E[k] + conj(W2048^k) O[k], split so both
// warps of a pair do the same amount of work: the even-bin warp finishes lags k = lane + 32 jj (jj < 16) from
// its own E and the partner's raw O; the odd-bin warp finishes k = lane + 32 (16 + jj) from its own O and the
// partner's E.  `theirs` is the partner's exchange tile ([jj*32 + lane]).
GB_HD GB_INLINE void combine_even(const float (&re)[32], const float (&im)[32], int lane, const float2* tw2,
                                  const float2* theirs, float (&xr)[16], float (&xi)[16]) {
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
        const float2 o = theirs[jj * 32 + lane];
        const float2 w = tw2[lane + 32 * jj];
        xr[jj] = re[jj] + (o.x * w.x + o.y * w.y);
        xi[jj] = im[jj] + (o.y * w.x - o.x * w.y);
    }
}
GB_HD GB_INLINE void combine_odd(const float (&re)[32], const float (&im)[32], int lane, const float2* tw2,
                                 const float2* theirs, float (&xr)[16], float (&xi)[16]) {
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
        const float2 e = theirs[jj * 32 + lane];
        const float2 w = tw2[lane + 32 * (16 + jj)];
        xr[jj] = e.x + (re[16 + jj] * w.x + im[16 + jj] * w.y);
        xi[jj] = e.y + (im[16 + jj] * w.x - re[16 + jj] * w.y);
    }
}
// What each warp hands to its partner: the even-bin warp its E[k] for the upper lags, the odd-bin warp its raw
// O[k] for the lower lags.
GB_HD GB_INLINE void exchange_store(const float (&re)[32], const float (&im)[32], int lane, int h, float2* mine) {
    if (h == 0) {
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) mine[jj * 32 + lane] = make_float2(re[16 + jj], im[16 + jj]);
    } else {
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) mine[jj * 32 + lane] = make_float2(re[jj], im[jj]);
    }
}

// Branch-free reduction of one thread's 16 finished lags (q = lane + 32 (16 h + jj), profile index s q + r) into
// (max, first index of max, count of max, sum).  Lag 1023 does not exist (only lane 31, h = 1, jj = 15).
GB_HD GB_INLINE void thread_peak16(const float (&v)[16], int lane, int h, int s, int r, Peak& out, float& fsum) {
    const bool last_invalid = (h == 1) && (lane == 31);
    float m = v[0];
    float sm = v[0];
#pragma unroll
    for (int jj = 1; jj < 15; ++jj) {
        m = m > v[jj] ? m : v[jj];
        sm += v[jj];
    }
    const float v15 = last_invalid ? -1.0f : v[15];
    m = m > v15 ? m : v15;
    sm += last_invalid ? 0.0f : v[15];
    int first = 15, c = 0;
    {
        const bool eq = v15 == m;
        c += eq ? 1 : 0;
    }
#pragma unroll
    for (int jj = 14; jj >= 0; --jj) {
        const bool eq = v[jj] == m;
        first = eq ? jj : first;
        c += eq ? 1 : 0;
    }
    out.mx = m;
    out.idx = s * (lane + 32 * (16 * h + first)) + r;
    out.cnt = c;
    out.sum = 0.0;
    fsum = sm;
}

GB_HD GB_INLINE void mul_tw2_conj(float (&re)[32], float (&im)[32], int lane, const float2* tw2) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float2 w = tw2[lane + 32 * j];
        const float a = re[j], b = im[j];
        re[j] = a * w.x + b * w.y;
        im[j] = b * w.x - a * w.y;
    }
}

}  // namespace gb

namespace gb {

// Padded +-1 chip sequence c'[m], m in [0, 2048): c'[m] = c[m] for m < 1023, c'[2048 - t] = c[1023 - t] for
// t = 1..1022, zero at 1023..1025.  A circular length-2048 correlation against c' of a signal supported on
// [0, 1023) equals the circular length-1023 correlation against c (lags 0..1022).
GB_HD GB_INLINE int padded_chip(const uint8_t* chips, int m) {
    if (m < kChips) return chips[m] ? 1 : -1;
    if (m >= kPad - (kChips - 1)) return chips[m - (kPad - kChips)] ? 1 : -1;
    return 0;
}

// conj(FFT2048(c'))[g] / 2048 in float64 from an exact-phase table cs[t] = (cos, sin)(2 pi t / 2048).
// The 1/2048 is the ifft scaling of utils.py:73 folded in.
GB_HD GB_INLINE void replica_spectrum_bin(const uint8_t* chips, int g, const double2* cs, double& re, double& im) {
    double ar = 0.0, ai = 0.0;
    for (int m = 0; m < kPad; ++m) {
        const int c = padded_chip(chips, m);
        if (c == 0) continue;
        const double2 w = cs[(g * m) & (kPad - 1)];  // exp(-2 pi i g m/2048) = cos - j sin; conj -> cos + j sin
        ar += c * w.x;
        ai += c * w.y;
    }
    re = ar / kPad;
    im = ai / kPad;
}

}  // namespace gb
