// Shared between device code and the host-side lane emulator (tests/emu).  No reference code here.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GB_HD __host__ __device__
#define GB_INLINE __forceinline__
#else
#define GB_HD
#define GB_INLINE inline
#endif

namespace gb {

constexpr int kChips = 1023;   // constants.py:7  PRN_CHIP_COUNT
constexpr int kFft = 1024;     // one warp-level transform
constexpr int kPad = 2048;     // zero-padded length carrying a length-1023 circular correlation (>= 2*1023-1)
constexpr int kTStride = 34;   // padded row of the 32x32 transpose tile (float2 units): 16-byte aligned rows, conflict-free
                               // for the 64-bit column writes and the 128-bit row reads
constexpr int kTileF2 = 32 * kTStride;

// Result of reducing one correlation profile; mirrors include/gypsum_b200.h gb200_cell_record (32 bytes).
struct CellRecord {
    float peak;      // max of the (non-coherent) profile, or of |coherent profile|
    int32_t argmax;  // first index attaining it (np.argmax rule, acquisition.py:184)
    double sum;      // sum over all N profile values
    int32_t count;   // how many values equal the max (utils.py:113 excludes all of them)
    float probe_re;  // coherent profile value at the requested index (acquisition.py:136)
    float probe_im;
    int32_t pad_;
};
static_assert(sizeof(CellRecord) == 32, "record must stay 32 bytes");

struct Peak {
    float mx;
    int idx;
    int cnt;
    double sum;
};

GB_HD GB_INLINE void peak_init(Peak& p) {
    p.mx = -1.0f;
    p.idx = 0x7fffffff;
    p.cnt = 0;
    p.sum = 0.0;
}
// Profile values are >= 0.  First index wins ties; cnt counts elements equal to the running max.
GB_HD GB_INLINE void peak_push(Peak& p, float v, int n) {
    if (v > p.mx) {
        p.mx = v;
        p.idx = n;
        p.cnt = 1;
    } else if (v == p.mx) {
        p.cnt += 1;
        p.idx = n < p.idx ? n : p.idx;
    }
}
GB_HD GB_INLINE void peak_merge(Peak& a, const Peak& b) {
    if (b.mx > a.mx) {
        a.mx = b.mx;
        a.idx = b.idx;
        a.cnt = b.cnt;
    } else if (b.mx == a.mx) {
        a.cnt += b.cnt;
        a.idx = b.idx < a.idx ? b.idx : a.idx;
    }
    a.sum += b.sum;
}

}  // namespace gb
