// Persistent per-channel tracking kernel (reference gypsum/tracker.py:264-389).
//
// One CTA per channel walks the stream one millisecond at a time: the next 1-ms IQ chunk is prefetched into a
// second shared buffer with cp.async while the current one is processed (double buffering); the carrier is wiped
// off with the channel's current (Doppler, phase); the polyphase forward warp FFTs, the product with the PRN's
// replica spectrum and the inverse warp FFTs run back to back in registers (nothing goes to global memory);
// the early / late correlations of tracker.py:293-295 are just lags p-1 / p+1 of that same circular correlation,
// and the prompt profile of :307-313 is the correlation rolled by p; one thread then runs the float64 loop
// filters (tracker_core.cuh) and the next millisecond starts.  Feedback makes time strictly sequential per
// channel; channels are independent.
#include "kernels.cuh"
#include "ptx_helpers.cuh"
#include "tracker_core.cuh"
#include "warp_fft.cuh"

namespace gb {

constexpr int kTrackThreads = 256;

struct TrackPartial {
    float mx;
    int key;  // index in the rolled prompt profile
    int cnt;
    float sum;
    float re, im;
    int pad[2];
};

__device__ __forceinline__ int pymod_int(int a, int m) {
    int r = a % m;
    return r < 0 ? r + m : r;
}

// The float64 loop filters run on one thread; keeping them out of line keeps the per-millisecond instruction
// footprint of the other warps small (the loop body otherwise overflows the instruction cache).
__device__ __noinline__ void track_update_device(TrackState* st, float2 E, float2 L, float2 peak, float strength, int key,
                                                 double t0, const TrackConsts* tc, const double* rtab, TrackMsRecord* out) {
    TrackMsRecord rec;
    track_update(*st, E, L, peak, strength, key, t0, *tc, rtab, rec);
    *out = rec;
}

template <int S>
__global__ void __launch_bounds__(kTrackThreads, 1) k_track_channels(const TrackArgs a) {
    extern __shared__ __align__(16) float2 smem[];
    constexpr int n_fft_warps = 2 * S;
    float2* iqbuf = smem;                 // [2][N]
    float2* ypoly = iqbuf + 2 * a.N;      // [S][1024]
    float2* crep_s = ypoly + S * kFft;    // [2][1024]
    float2* tw1_s = crep_s + 2 * kFft;
    float2* tw2_s = tw1_s + kFft;
    float2* tiles = tw2_s + kFft;                            // [2s][kTileF2]
    TrackState* st = reinterpret_cast<TrackState*>(tiles + static_cast<size_t>(n_fft_warps) * kTileF2);
    TrackPartial* partial = reinterpret_cast<TrackPartial*>(st + 1);  // [8]
    float2* el = reinterpret_cast<float2*>(partial + 8);              // [2] early, late
    float2* coarse = el + 2;                                          // [16] carrier at samples 0, 256, ... (+ phase)
    uint64_t* mbar = reinterpret_cast<uint64_t*>(coarse + 16);
    double* rtab = reinterpret_cast<double*>(mbar + 2);               // [256] 1 / n for the lock-window counts
    TrackConsts* tc = reinterpret_cast<TrackConsts*>(rtab + 256);

    const int slot = blockIdx.x;                                     // where this CTA's records go
    const int ch = a.channel_idx ? a.channel_idx[slot] : slot;       // which channel's state it advances
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    TrackState* gst = a.states + ch;

    // ---- load the channel state (keeping a copy for a later rollback when asked to), the twiddles and this PRN's
    //      replica spectrum ----
    {
        const int* src = reinterpret_cast<const int*>(gst);
        int* dst = reinterpret_cast<int*>(st);
        int* shd = a.shadow ? reinterpret_cast<int*>(a.shadow + ch) : nullptr;
        for (int i = tid; i < static_cast<int>(sizeof(TrackState) / 4); i += kTrackThreads) {
            const int v = src[i];
            dst[i] = v;
            if (shd) shd[i] = v;
        }
    }
    if (tid == 0) {
        mbar_init(mbar, 1);
        *tc = track_consts(a.fs);
    }
    rtab[tid] = tid ? 1.0 / static_cast<double>(tid) : 0.0;  // kTrackThreads == 256 entries
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(mbar, 4 * kFft * sizeof(float2));
        bulk_g2s(tw1_s, a.tw1, kFft * sizeof(float2), mbar);
        bulk_g2s(tw2_s, a.tw2, kFft * sizeof(float2), mbar);
        bulk_g2s(crep_s, a.crep + static_cast<size_t>(st->prn) * 2 * kFft, 2 * kFft * sizeof(float2), mbar);
    }
    mbar_wait(mbar, 0);

    const int chunk16 = a.N / 2;  // 16-byte pieces per 1-ms chunk (N is even for every supported rate)
    auto prefetch = [&](int k) {
        const float2* src = a.iq + static_cast<size_t>(k) * a.N;
        float2* dst = iqbuf + (k & 1) * a.N;
        for (int i = tid; i < chunk16; i += kTrackThreads) cp_async16(dst + 2 * i, src + 2 * i);
        cp_async_commit();
    };
    if (a.n_ms > 0) prefetch(0);

    const int r = warp >> 1, h = warp & 1;
    float2* tile = tiles + warp * kTileF2;
    const float2* ptile = tiles + (warp ^ 1) * kTileF2;
    TrackMsRecord* out = a.out + static_cast<size_t>(slot) * a.n_ms;

    for (int k = 0; k < a.n_ms; ++k) {
        if (k + 1 < a.n_ms) {
            prefetch(k + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();  // chunk k landed; the previous millisecond's loop-filter update is visible
        if (st->lost) {   // tracker.py:378: the channel stopped; later milliseconds are not processed
            if (tid == 0) {
                TrackMsRecord rec = {};
                rec.lost = 2;
                rec.doppler = rec.doppler_hist = st->doppler;
                rec.carrier_phase = rec.carrier_phase_hist = st->carrier_phase;
                rec.code_phase = st->code_phase;
                out[k] = rec;
            }
            continue;
        }
        const double f = st->doppler, phi_cycles = st->carrier_phase * (1.0 / kTau), t0 = a.start_times ? a.start_times[k] : a.t0_single;
        const int p = st->code_phase;
        const int pm = pymod_int(p, a.N);
        const int kE = pymod_int(p - 1, a.N), kL = pymod_int(p + 1, a.N);

        // ---- carrier wipe-off (tracker.py:278-281): exp(-j(2 pi f (n/fs + t0) + phi)) = coarse[n/256] * fine[n%256],
        //      each factor from a float64-reduced phase; polyphase de-interleave ----
        if (tid < (a.N + kTrackThreads - 1) / kTrackThreads)
            coarse[tid] = wipeoff(make_float2(1.f, 0.f), f * (static_cast<double>(tid * kTrackThreads) * a.inv_fs + t0) + phi_cycles);
        const float2 fine = carrier_at(f, static_cast<double>(tid), a.inv_fs);
        __syncthreads();
        const float2* buf = iqbuf + (k & 1) * a.N;
        for (int kk = 0, n = tid; n < a.N; ++kk, n += kTrackThreads)
            ypoly[(n % S) * kFft + n / S] = cmul(buf[n], cmul(coarse[kk], fine));
        __syncthreads();
        if (tid < S) ypoly[tid * kFft + (kFft - 1)] = ypoly[tid * kFft];
        __syncthreads();

        if (warp < n_fft_warps) {
            float2 x[32];
            // forward transform, spectrum product, inverse transform
            build_z(x, lane, r, S, ypoly);
            if (h) mul_tw2(x, lane, tw2_s);
            wfft_phase1<false>(x, lane, tw1_s, tile);
            __syncwarp();
            wfft_phase2<false>(x, lane, tile);
            __syncwarp();
            mul_vec(x, lane, crep_s + h * kFft);
            wfft_phase1<true>(x, lane, tw1_s, tile);
            __syncwarp();
            wfft_phase2<true>(x, lane, tile);
            __syncwarp();
            exchange_store(x, lane, h, tile);
            pair_barrier(r);
            float2 out16[16];
            if (h == 0) combine_even(x, lane, tw2_s, ptile, out16);
            else combine_odd(x, lane, tw2_s, ptile, out16);

            // ---- prompt profile statistics in rolled order (tracker.py:308-313), early / late taps ----
            float mx = -1.f, sum = 0.f, bre = 0.f, bim = 0.f;
            int key = 0x7fffffff, cnt = 0;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int q = lane + 32 * (16 * h + jj);
                if (q < kChips) {
                    const int n = S * q + r;
                    const float v = gb_mag(out16[jj]);
                    int kk = n - pm;
                    kk = kk < 0 ? kk + a.N : kk;
                    if (v > mx || (v == mx && kk < key)) {
                        cnt = v > mx ? 1 : cnt + 1;
                        mx = v;
                        key = kk;
                        bre = out16[jj].x;
                        bim = out16[jj].y;
                    } else if (v == mx) {
                        cnt++;
                    }
                    sum += v;
                    if (n == kE) el[0] = out16[jj];
                    if (n == kL) el[1] = out16[jj];
                    if (a.profiles) a.profiles[(static_cast<size_t>(slot) * a.n_ms + k) * a.N + kk] = v;
                }
            }
            const int bits = __float_as_int(mx);
            const int mb = __reduce_max_sync(0xffffffffu, bits);
            const bool is = bits == mb;
            const int kmin = __reduce_min_sync(0xffffffffu, is ? key : 0x7fffffff);
            const int ctot = __reduce_add_sync(0xffffffffu, is ? cnt : 0);
            const unsigned owner = __ballot_sync(0xffffffffu, is && key == kmin);
            const int src = __ffs(owner) - 1;
            bre = __shfl_sync(0xffffffffu, bre, src);
            bim = __shfl_sync(0xffffffffu, bim, src);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            if (lane == 0) {
                TrackPartial pp;
                pp.mx = __int_as_float(mb);
                pp.key = kmin;
                pp.cnt = ctot;
                pp.sum = sum;
                pp.re = bre;
                pp.im = bim;
                pp.pad[0] = pp.pad[1] = 0;
                partial[warp] = pp;
            }
        }
        __syncthreads();
        if (tid == 0) {
            float mx = -1.f, pre = 0.f, pim = 0.f;
            int key = 0x7fffffff, cnt = 0;
            double sum = 0.0;
            for (int w = 0; w < n_fft_warps; ++w) {
                const TrackPartial pp = partial[w];
                if (pp.mx > mx || (pp.mx == mx && pp.key < key)) {
                    cnt = pp.mx > mx ? pp.cnt : cnt + pp.cnt;
                    mx = pp.mx;
                    key = pp.key;
                    pre = pp.re;
                    pim = pp.im;
                } else if (pp.mx == mx) {
                    cnt += pp.cnt;
                }
                sum += static_cast<double>(pp.sum);
            }
            const double m = static_cast<double>(mx);
            const float strength = static_cast<float>(m / ((sum - cnt * m) / (a.N - cnt)));  // utils.py:111-116
            track_update_device(st, el[0], el[1], make_float2(pre, pim), strength, key, t0, tc, rtab, &out[k]);
        }
        // the __syncthreads at the top of the next millisecond publishes st / frees el, partial
    }
    __syncthreads();
    {
        const int* src = reinterpret_cast<const int*>(st);
        int* dst = reinterpret_cast<int*>(gst);
        for (int i = tid; i < static_cast<int>(sizeof(TrackState) / 4); i += kTrackThreads) dst[i] = src[i];
    }
}

size_t track_smem_bytes(int N, int s) {
    return (2 * static_cast<size_t>(N) + static_cast<size_t>(s) * kFft + 4 * kFft + 2 * static_cast<size_t>(s) * kTileF2) *
               sizeof(float2) +
           sizeof(TrackState) + 8 * sizeof(TrackPartial) + (2 + 16) * sizeof(float2) + 16 + 256 * sizeof(double) + sizeof(TrackConsts);
}

cudaError_t configure_track_kernel() {
    cudaError_t e = cudaFuncSetAttribute(k_track_channels<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_track_channels<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t launch_track_channels(const TrackArgs& a, cudaStream_t st) {
    const size_t sm = track_smem_bytes(a.N, a.s);
    if (a.s == 2) k_track_channels<2><<<a.n_channels, kTrackThreads, sm, st>>>(a);
    else if (a.s == 4) k_track_channels<4><<<a.n_channels, kTrackThreads, sm, st>>>(a);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace gb
