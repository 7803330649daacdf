// Kernel argument blocks and launch wrappers (implemented in kernels.cu, used by engine.cu).
#pragma once
#include "gb_common.cuh"
#include "tracker_core.cuh"

namespace gb {

constexpr int kKindCoherent = 1;     // utils.py:23-25: IntegrationType.Coherent = auto() -> 1
constexpr int kKindNonCoherent = 2;  // IntegrationType.NonCoherent -> 2

// doppler_spectra: one CTA per (unique Doppler u, millisecond i).
struct SpectraArgs {
    const float2* iq;        // complex64 samples, block b starts at b * block_stride
    const double* doppler;   // [n_doppler] Hz
    float2* spec;            // [n_blocks*n_doppler][M][s][2][1024]
    const float2* tw1;       // [32][32]  exp(-2 pi i lane k1 / 1024)
    const float2* tw2;       // [1024]    exp(-2 pi i n / 2048)
    long long block_stride;  // samples between consecutive blocks (= M*N)
    double inv_fs;
    int N, s, M, n_doppler, n_units;  // n_units = n_blocks * n_doppler
};

// correlate_cells: one warp pair per (cell, r-range); NP pairs per CTA all on the same PRN.
struct CorrelateArgs {
    const float2* spec;
    const float2* crep;  // [n_prn][2][1024]  conj(FFT2048(c'))/2048, even / odd bins
    const float2* tw1;   // [32][32]
    const float2* tw2;   // [1024]
    CellRecord* records;
    float* profile;      // optional: full profile of the single cell (N floats, or 2N when coherent)
    int N, s, M, kind;
    int rsplit;          // pairs cooperating on one cell (divides s and NP)
    int n_groups;
    // grid mode (cells = blocks x prn list x doppler list)
    int grid_mode, P, D, n_blocks, chunks;  // chunks = ceil(n_blocks * D / cells_per_group): groups per PRN
    int win_chunks;  // k_correlate_w2048, grid mode: chunks per L2 window (0 = the whole batch is one window), see the kernel
    const int* prn_idx;           // [P]
    // list mode (cells sorted by PRN)
    const int* grp_first;
    const int* grp_count;
    const int* grp_prn;
    const int* cell_u;      // spectrum unit of each sorted cell
    const int* cell_out;    // where its record goes
    const int* cell_probe;  // coherent probe index or -1 (both modes, indexed by output slot; may be null)
    int stag_a, stag_b;       // k_correlate_w2048 start stagger in ns: (warp / 4) * stag_a + (warp % 4) * stag_b
    const double* cell_gate;  // optional, indexed by output slot: NaN = this cell is switched off (device-planned lists)
};

// On-device Doppler refinement (reference acquisition.py:70-152): per-satellite search state and result.
constexpr int kRefineMaxBins = 32;  // acquisition.py:163-167 yields 20..28 bins per pass
struct RefineState {
    double center;         // center_doppler_shift_estimation
    double kept_doppler;   // best_..._across_all_search_space.doppler_shift
    double kept_strength;  // .correlation_strength
    int kept_index;        // .sample_offset_of_correlation_peak
    int have_kept;
};
struct RefineResult {  // 32 bytes, mirrored by gb200_acquisition_result
    double doppler;
    double strength;
    float probe_re, probe_im;  // coherent profile value at the kept peak index
    int code_phase;
    int pad_;
};
struct BestRecord {  // 32 bytes, mirrored by gb200_best_record
    double doppler;
    double strength;
    float peak;
    int code_phase;
    int bin;
    int pad_;
};
cudaError_t launch_best_bins(int n_rows, int D, int N, const CellRecord* rec, const double* doppler, BestRecord* out,
                             cudaStream_t s);
cudaError_t launch_refine_plan(int n_sv, double spread, const RefineState* st, double* doppler, cudaStream_t s);
cudaError_t launch_refine_select(int n_sv, int N, const CellRecord* rec, const double* doppler, RefineState* st, cudaStream_t s);
cudaError_t launch_refine_coherent_plan(int n_sv, const RefineState* st, double* doppler, int* probe, cudaStream_t s);
cudaError_t launch_refine_finalize(int n_sv, const RefineState* st, const CellRecord* rec, RefineResult* out, cudaStream_t s);
cudaError_t launch_refine_init(int n_sv, RefineState* st, cudaStream_t s);

// track_channels: one persistent CTA per channel.
struct TrackArgs {
    const float2* iq;           // [n_ms * N] the stream every channel consumes
    const double* start_times;  // [n_ms] chunk start timestamps (antenna_sample_provider.py:88-89)
    TrackState* states;         // [n_channels]
    TrackMsRecord* out;         // [n_channels][n_ms]
    float* profiles;            // optional [n_channels][n_ms][N]: |prompt profile| (tracker.py:308-309)
    const float2* crep;
    const float2* tw1;
    const float2* tw2;
    double fs, inv_fs;
    int N, s, n_ms, n_channels;
    double t0_single;           // start time used when start_times is null (single-millisecond launches: no upload)
    const int* channel_idx;     // optional [n_channels]: CTA b runs channel channel_idx[b] (records still go to out[b]); null = b
    TrackState* shadow;         // optional [capacity]: every launched channel's state as it was BEFORE this launch (rollback)
};

// integrate_bits: one warp per tracking channel (bits.cu, bits_core.cuh).
struct BitState;
struct BitEvent;
struct BitArgs {
    const TrackMsRecord* records;  // [n_channels][n_ms] as written by k_track_channels
    const double* start_times;     // [n_ms] chunk start / end timestamps
    const double* end_times;
    BitState* states;              // [n_channels]
    BitEvent* events;              // [n_channels][max_events]
    int* counts;                   // [n_channels] events produced (may exceed max_events: truncated)
    int n_ms, n_channels, max_events;
};

// acquire_fused: one CTA per (PRN, Doppler) cell, the whole pipeline in one kernel (fused.cu).
struct FusedArgs {
    const float2* iq;       // [M*N] one block
    const double* doppler;  // [n_cells]
    const int* prn;         // [n_cells] replica row
    const int* probe;       // [n_cells] coherent probe index or -1 (may be null)
    CellRecord* records;    // [n_cells]
    const float2* crep;
    const float2* tw1;
    const float2* tw2;
    double inv_fs;
    int N, M, n_cells;
};
bool fused_supports(int s);
cudaError_t configure_fused_kernel();
cudaError_t launch_acquire_fused(const FusedArgs& a, int s, int kind, cudaStream_t st);

size_t track_smem_bytes(int N, int s);
cudaError_t configure_track_kernel();
cudaError_t launch_track_channels(const TrackArgs& a, cudaStream_t st);
cudaError_t launch_integrate_bits(const BitArgs& a, cudaStream_t st);
size_t spectra_smem_bytes(int s);
bool spectra_supports(int s);
size_t correlate_smem_bytes(int np);
cudaError_t launch_init_tables(float2* tw1, float2* tw2, cudaStream_t st);
cudaError_t launch_replica_spectra(const uint8_t* chips_dev, int n_prn, float2* crep, cudaStream_t st);
cudaError_t launch_doppler_spectra(const SpectraArgs& a, cudaStream_t st);
cudaError_t launch_correlate_cells(const CorrelateArgs& a, int np, int grid, cudaStream_t st);
cudaError_t launch_correlate_w2048(const CorrelateArgs& a, int nw, int grid, cudaStream_t st);
cudaError_t launch_correlate_generic(const float2* iq, const float2* replica, int N, int n_ms, double doppler, double inv_fs,
                                     int kind, float* out, cudaStream_t st);
cudaError_t configure_kernels();

}  // namespace gb
