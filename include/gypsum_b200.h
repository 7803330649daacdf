/* gypsum_b200 -- C ABI of the B200 acquisition / tracking correlation engine.
 *
 * The reference (codyd51/gypsum) has no FFI: its boundary for this path is two Python classes and one pure
 * function.  Each entry point below names the reference interface it stands behind (paths relative to the
 * reference checkout).  Plain pointers and sizes only; the library owns every device allocation behind the
 * handle; host arrays belong to the caller and are only read/written during the call.  All functions return 0
 * on success or a GB200_E* code; gb200_last_error() gives the message.  There is no CPU fallback: without a
 * CUDA device gb200_create fails.
 *
 * Threading: one caller thread per engine (the reference is single-threaded, receiver.py:85-146).  Host-pointer
 * entry points return after the results are on the host; *_device entry points only enqueue work on the
 * engine's stream (gb200_set_stream).
 */
#ifndef GYPSUM_B200_H
#define GYPSUM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB200_ABI_VERSION 2

#define GB200_OK 0
#define GB200_EINVAL 1  /* bad argument            -> ValueError  (utils.py:106) */
#define GB200_ECUDA 2   /* CUDA runtime failure     -> RuntimeError */
#define GB200_ESTATE 3  /* replicas / IQ not loaded -> RuntimeError (acquisition.py:112) */

/* utils.py:23-25  IntegrationType(Enum): Coherent = auto() (1), NonCoherent = auto() (2) */
#define GB200_COHERENT 1
#define GB200_NON_COHERENT 2

typedef struct gb200_engine gb200_engine;

/* One reduced correlation profile (32 bytes).  Replaces the f64[N] profile that utils.py:77-108 returns and
 * that acquisition.py:180-189 immediately reduces with np.max / np.argmax /
 * get_normalized_correlation_peak_strength (utils.py:111-116):
 *     strength = peak / ((sum - count*peak) / (N - count)).                                             */
typedef struct gb200_cell_record {
    float peak;      /* np.max(profile); for coherent integration, max |profile|                */
    int32_t argmax;  /* np.argmax(profile): first index attaining the max, 0..N-1               */
    double sum;      /* sum of all N profile values                                             */
    int32_t count;   /* number of values equal to peak (utils.py:113 drops every one of them)   */
    float probe_re;  /* coherent only: profile[probe_idx] (acquisition.py:136 np.angle input)    */
    float probe_im;
    int32_t reserved;
} gb200_cell_record;

int gb200_abi_version(void);

/* receiver.py:46-66: one engine per stream format.  samples_per_ms must be a multiple of 1023
 * (antenna_sample_provider.py:131-136, SampleProviderAttributes).                                  */
int gb200_create(int device_ordinal, int samples_per_second, int samples_per_ms, gb200_engine** out);
int gb200_destroy(gb200_engine* e);
/* e may be NULL: message of the last failed gb200_create on this thread. */
const char* gb200_last_error(const gb200_engine* e);

/* Work is enqueued on this cudaStream_t (0 / NULL = the engine's own stream). */
int gb200_set_stream(gb200_engine* e, void* cuda_stream);

/* satellite.py:20-31 GpsSatellite.prn_as_complex + gps_ca_prn_codes.py:120-131: chips[n_prn][1023] in {0,1};
 * replica index p in later calls refers to row p.  Builds conj(FFT(replica)) on the device once
 * (the reference redoes np.fft.fft(prn_replica) on every call, utils.py:66).                        */
int gb200_set_replicas(gb200_engine* e, const uint8_t* chips, int n_prn);

/* receiver.py:219 antenna_data: complex64[n_samples] (interleaved float32 I,Q -- the on-disk format of
 * antenna_sample_provider.py:112-119).  upload copies from the host; bind uses a device buffer in place.    */
int gb200_upload_iq(gb200_engine* e, const float* iq_host, int64_t n_samples);
/* iq_device must be 16-byte aligned (the fused and tracking kernels stage it with bulk / cp.async copies). */
int gb200_bind_iq_device(gb200_engine* e, const void* iq_device, int64_t n_samples);

/* receiver.py:68,100,219 rolling_samples_buffer (deque(maxlen=ACQUISITION_INTEGRATION_PERIOD_MS)) on the device: every
 * new millisecond of antenna_sample_provider.py:94-124 is uploaded ONCE (one 8*N-byte copy) and both the detector's
 * 10-ms window and every tracking channel read it in place.  The newest n_ms <= capacity_ms milliseconds are always
 * contiguous in device memory (each millisecond is stored twice, capacity_ms apart).                                  */
typedef struct gb200_ring gb200_ring;
int gb200_ring_create(gb200_engine* e, int capacity_ms, gb200_ring** out);
int gb200_ring_destroy(gb200_ring* r);
/* Append n_ms whole milliseconds (complex64[n_ms * N]) from host memory. */
int gb200_ring_append(gb200_ring* r, const float* iq_host, int n_ms);
/* The engine's IQ binding := the newest n_ms milliseconds (zero copy); what gb200_detect / gb200_acquire_* /
 * gb200_tracker_process* then read.  n_ms <= min(capacity_ms, milliseconds appended so far).                          */
int gb200_ring_bind_newest(gb200_ring* r, int n_ms);
int gb200_ring_appended(const gb200_ring* r, int64_t* total_ms);

/* Benchmark-shaped search grid: the loaded IQ holds n_blocks independent blocks of
 * ms_per_block milliseconds; every (block, prn_idx[a], doppler_hz[b]) cell is one
 * utils.py:77 integrate_correlation_with_doppler_shifted_prn evaluation reduced to a record.
 * out[(block*n_prn + a)*n_doppler + b].                                                             */
int gb200_acquire_grid(gb200_engine* e, int n_blocks, int ms_per_block, const int32_t* prn_idx, int n_prn,
                       const double* doppler_hz, int n_doppler, int integration_type, gb200_cell_record* out_host);
int gb200_acquire_grid_device(gb200_engine* e, int n_blocks, int ms_per_block, const int32_t* prn_idx, int n_prn,
                              const double* doppler_hz, int n_doppler, int integration_type, void* out_device);
/* The same grid host to host in ONE call for latency-bound callers (receiver.py:219-224 hands over one window per
 * scan): copy-in, both kernels and copy-out are replayed as one CUDA graph per grid shape, one host synchronisation.
 * iq_host: complex64[n_blocks * ms_per_block * N].  Pageable and pinned buffers are both accepted; with pinned ones
 * (cudaHostAlloc / cudaHostRegister) inputs above 64 KB are read by the copy engine in place and small record sets
 * (<= 256 KB) are stored by the kernel straight into out_host, which saves the two staging copies.                    */
int gb200_acquire_grid_host(gb200_engine* e, const float* iq_host, int n_blocks, int ms_per_block, const int32_t* prn_idx,
                            int n_prn, const double* doppler_hz, int n_doppler, int integration_type,
                            gb200_cell_record* out_host);

/* acquisition.py:179-189 applied to every (block, prn) row of a grid on the device: the first Doppler bin with the
 * largest profile maximum, its code phase and strength -- 32 bytes per (block, prn) instead of 32 bytes per cell
 * (what a multi-GPU gather has to move).  out[(block * n_prn + a)].                                    */
typedef struct gb200_best_record {
    double doppler_hz;  /* BestNonCoherentCorrelationProfile.doppler_shift                  (acquisition.py:186) */
    double strength;    /* .correlation_strength                                            (acquisition.py:189) */
    float peak;         /* np.max of the winning bin's profile                                                    */
    int32_t code_phase; /* .sample_offset_of_correlation_peak                               (acquisition.py:184) */
    int32_t bin;        /* index of the winning bin in doppler_hz                                                */
    int32_t reserved;
} gb200_best_record;
int gb200_acquire_grid_best(gb200_engine* e, int n_blocks, int ms_per_block, const int32_t* prn_idx, int n_prn,
                            const double* doppler_hz, int n_doppler, int integration_type, gb200_best_record* out_host);
int gb200_acquire_grid_best_device(gb200_engine* e, int n_blocks, int ms_per_block, const int32_t* prn_idx, int n_prn,
                                   const double* doppler_hz, int n_doppler, int integration_type, void* out_device);

/* acquisition.py:154-190 get_best_doppler_shift_estimation / :122-136: an arbitrary list of (prn, Doppler)
 * cells over the first n_ms milliseconds of the loaded IQ.  probe_idx (may be NULL) gives, per cell, the
 * profile index whose complex value is wanted for coherent integration, or -1.                      */
int gb200_acquire_cells(gb200_engine* e, int n_cells, const int32_t* prn_idx, const double* doppler_hz,
                        const int32_t* probe_idx, int n_ms, int integration_type, gb200_cell_record* out_host);

/* acquisition.py:70-152 for a batch of satellites, entirely on the device: the ten refinement passes
 * (spread 7000 Hz halved while >= 10, bins range(int(c-s), int(c+s), int(s/10)), first bin with the largest
 * profile maximum, kept result = strictly greatest strength), then one coherent integration at the kept Doppler.
 * No host round trip between passes.  out[i] belongs to prn_idx[i].                                    */
typedef struct gb200_acquisition_result {
    double doppler_hz;  /* SatelliteAcquisitionAttemptResult.doppler_shift        (acquisition.py:120)       */
    double strength;    /* .correlation_strength                                   (acquisition.py:138)       */
    float probe_re;     /* coherent profile at the kept peak index; np.angle of it */
    float probe_im;     /*   is .carrier_wave_phase_shift                          (acquisition.py:136)       */
    int32_t code_phase; /* .prn_phase_shift                                        (acquisition.py:137)       */
    int32_t reserved;
} gb200_acquisition_result;
int gb200_detect(gb200_engine* e, int n_sv, const int32_t* prn_idx, int n_ms, gb200_acquisition_result* out_host);

/* utils.py:77-108 in full: the N-value profile of one cell.  out_host holds N floats (non-coherent) or
 * 2N floats (coherent, interleaved re,im).                                                          */
int gb200_correlation_profile(gb200_engine* e, int prn_idx, double doppler_hz, int n_ms, int integration_type,
                              float* out_host);
/* The same for ANY replica (utils.py:59-73 / :77-108 accept an arbitrary complex prn_replica of N samples, not only the
 * chips-repeated form GpsSatellite.prn_as_complex produces): replica_host is complex64[N]; the circular correlation is
 * evaluated directly (N^2 multiply-adds, float64 accumulation).  Slow path for the public helpers, never used by the
 * receiver's own calls.                                                                                              */
int gb200_correlation_profile_replica(gb200_engine* e, const float* replica_host, double doppler_hz, int n_ms,
                                      int integration_type, float* out_host);

/* ---------------------------------------------------------------------------------------------------------
 * Tracking (gypsum/tracker.py).  A tracker is a bank of channels sharing the engine's loaded IQ stream; every
 * channel is one GpsSatelliteTracker (tracker.py:206-389): state {Doppler, carrier phase, code phase} seeded
 * from an acquisition result (satellite_signal_processing_pipeline.py:56-62).
 * --------------------------------------------------------------------------------------------------------- */
typedef struct gb200_tracker gb200_tracker;

/* One millisecond of one channel (112 bytes): what GpsSatelliteTracker.process_samples (tracker.py:331-389)
 * leaves in tracking_params' histories plus the emitted pseudosymbol.                                        */
typedef struct gb200_track_record {
    double doppler;        /* current_doppler_shift when process_samples returns (tracker.py:260, :385)      */
    double carrier_phase;  /* current_carrier_wave_phase_shift when it returns   (tracker.py:258-259, :386)   */
    double error;          /* Costas discriminator I*Q                       (tracker.py:249, :261)           */
    double disc;           /* (|E|^2 - |L|^2) / 2                            (tracker.py:297, :300)           */
    double phase_acc;      /* self.phase after the update                    (tracker.py:298-303)             */
    double doppler_hist;       /* what :352 appends to doppler_shifts: the value BEFORE the 6-s adjustment of :380-387 */
    double carrier_phase_hist; /* what :353 appends to carrier_wave_phases, likewise                          */
    float peak_re, peak_im; /* coherent prompt correlation peak              (tracker.py:313, :346)           */
    float strength;        /* get_normalized_correlation_peak_strength       (tracker.py:311, :347)           */
    float early_re, early_im, late_re, late_im; /* np.correlate taps         (tracker.py:293-295)             */
    int32_t code_phase;    /* current_prn_code_phase_shift after this ms     (tracker.py:299)                 */
    int32_t symbol;        /* sign(Re peak): +1 / -1 (0 only if Re peak == 0) (tracker.py:316)                */
    int32_t locked;        /* is_locked() used for this ms's loop bandwidth  (tracker.py:251)                 */
    int32_t lost;          /* 1: LostSatelliteLockError raised at this ms (tracker.py:378); 2: channel already stopped */
    int32_t peak_offset;   /* np.argmax of the prompt profile                (tracker.py:310)                 */
    int32_t reserved[2];
} gb200_track_record;

/* satellite_signal_processing_pipeline.py:56-63: one channel per (replica row, Doppler, carrier phase, code
 * phase).  Needs samples_per_ms == 2046 or 4092 (the reference hard-wires 2046, tracker.py:301-303,319).     */
int gb200_tracker_create(gb200_engine* e, int n_channels, const int32_t* prn_idx, const double* doppler_hz,
                         const double* carrier_phase, const int32_t* code_phase, gb200_tracker** out);
int gb200_tracker_destroy(gb200_tracker* t);
/* tracker.py:331 process_samples for n_ms consecutive 1-ms chunks of the engine's loaded IQ, every channel.
 * start_times[n_ms]: AntennaSampleChunk.start_time of each chunk.  out_host[channel*n_ms + ms].
 * profiles_host (may be NULL): [channel][ms][N] |prompt correlation profile| (tracker.py:309), float32.      */
int gb200_tracker_process(gb200_tracker* t, int n_ms, const double* start_times, gb200_track_record* out_host,
                          float* profiles_host);
/* Enqueue only; records stay on the device (out_device: n_channels*n_ms records). */
int gb200_tracker_process_device(gb200_tracker* t, int n_ms, const double* start_times, void* out_device);
/* Read / overwrite the loop state of one channel (tracking_params.current_* and tracker.phase). */
int gb200_tracker_get_state(gb200_tracker* t, int channel, double* doppler_hz, double* carrier_phase, double* phase_acc,
                            int32_t* code_phase, int32_t* lost);
/* set_state also clears the channel's `lost` flag (the reference tracker object keeps working after it raised). */
int gb200_tracker_set_state(gb200_tracker* t, int channel, double doppler_hz, double carrier_phase, double phase_acc,
                            int32_t code_phase);

/* A pool of channel slots for callers that create and drop trackers one at a time (receiver.py:226-267: one
 * GpsSatelliteTracker per acquired satellite, dropped on LostSatelliteLockError).  gb200_tracker_create_pool makes
 * `capacity` idle slots; gb200_tracker_reset_channel seeds one like satellite_signal_processing_pipeline.py:56-63 does
 * (fresh histories, lost = 0).  gb200_tracker_process_channels is gb200_tracker_process for a chosen subset in ONE
 * launch: out_host[i * n_ms + ms] belongs to channels[i].  With keep_undo != 0 every launched channel's state before
 * the call is kept, and gb200_tracker_undo_channel puts it back -- used by the drop-in GpsSatelliteTracker objects,
 * which advance all channels that share a chunk with one launch when the first of them is asked
 * (receiver.py:103-106 loops the same chunk over every pipeline) and take the step back for a channel that turns
 * out not to be asked.                                                                                                */
int gb200_tracker_create_pool(gb200_engine* e, int capacity, gb200_tracker** out);
int gb200_tracker_reset_channel(gb200_tracker* t, int channel, int32_t prn_idx, double doppler_hz, double carrier_phase,
                                int32_t code_phase);
int gb200_tracker_process_channels(gb200_tracker* t, int n_sel, const int32_t* channels, int n_ms, const double* start_times,
                                   int keep_undo, gb200_track_record* out_host, float* profiles_host);
int gb200_tracker_undo_channel(gb200_tracker* t, int channel);

/* A pipelined stream of equally shaped grid batches -- the receiver's steady state (receiver.py:85-146 hands over one
 * block after another): `submit` copies a batch of n_blocks*M*N complex64 samples from host memory, runs the grid of
 * gb200_acquire_grid on it and sends the n_blocks*P*D records to out_host; `collect` waits for the OLDEST batch in
 * flight.  Up to `depth` batches are in flight: the host->device copy of batch k+1 and the device->host copy of batch
 * k-1 run on their own streams under the kernels of batch k.  iq_host / out_host are DMA'd directly when they are
 * pinned, staged otherwise; both must stay valid until the batch is collected.  While a stream exists it owns the
 * engine's IQ binding (gb200_upload_iq / gb200_bind_iq_device must be called again before other acquire calls). */
typedef struct gb200_grid_stream gb200_grid_stream;
int gb200_grid_stream_create(gb200_engine* e, int n_blocks, int n_ms, const int32_t* prn_idx, int n_prn,
                             const double* doppler_hz, int n_doppler, int kind, int depth, gb200_grid_stream** out);
int gb200_grid_stream_submit(gb200_grid_stream* g, const float* iq_host, gb200_cell_record* out_host);
int gb200_grid_stream_collect(gb200_grid_stream* g);
int gb200_grid_stream_destroy(gb200_grid_stream* g);

/* Pseudosymbol -> navigation bit integration (gypsum/navigation_bit_intergrator.py), the consumer of the tracker's
 * +-1 stream (satellite_signal_processing_pipeline.py:77-79).  One EmitNavigationBitEvent (:29-39), 32 bytes. */
typedef struct gb200_bit_event {
    double receiver_timestamp;               /* start_of_pseudosymbol of the bit's first symbol      (:188) */
    double trailing_edge_receiver_timestamp; /* end_of_pseudosymbol of its last symbol               (:189) */
    int32_t ms_index;   /* millisecond (within this call) whose symbol completed the bit                    */
    int32_t bit_value;  /* 1 = BitValue.ONE, 0 = BitValue.ZERO, -1 = BitValue.UNKNOWN                (:149-161) */
    int32_t slide;      /* NavigationBitIntegrator.slide when the bit was emitted                           */
    int32_t pad_;
} gb200_bit_event;

/* NavigationBitIntegrator.process_pseudosymbol (:278-288) for every channel over n_ms millisecond records that are
 * in device memory: records_device ([channel][n_ms] gb200_track_record), or NULL for the records the last
 * gb200_tracker_process call of this tracker left on the device (same n_ms).  start_times / end_times: the chunk
 * timestamps (antenna_sample_provider.py:88-91); receiver_timestamp of :278 is the chunk start.  Each channel keeps
 * one integrator (bit phase, queue, health history) across calls; a channel whose record says `lost` stops there.
 * events_host: [channel][max_events]; counts_host: [channel] events produced (> max_events means truncated). */
int gb200_tracker_integrate_bits(gb200_tracker* t, int n_ms, const double* start_times, const double* end_times,
                                 const void* records_device, gb200_bit_event* events_host, int32_t max_events,
                                 int32_t* counts_host);
/* history of one channel's integrator: out[0..7] = emitted_bit_count, failed_bit_count, processed_pseudosymbol_count,
 * slide, determined_bit_phase (-1 = None), previous_bit_phase_decision (-1 = None), pseudosymbol_cursor_within_queue,
 * stopped. */
int gb200_tracker_bit_state(gb200_tracker* t, int channel, int64_t out[8]);

/* Kernel selection for gb200_acquire_cells.  Two implementations of the same arithmetic exist:
 *   0  doppler_spectra + correlate_cells: the PRN-independent half of the pipeline (wipe-off, forward transform) is
 *      computed once per distinct Doppler bin and shared by every PRN -- the grid shape (gb200_acquire_grid always
 *      uses it);
 *   1  the single fused block-per-(PRN, Doppler) kernel: IQ chunk and replica spectrum staged by TMA, whole
 *      pipeline in one CTA (2046 / 4092 samples per ms only) -- best when every cell has its own Doppler, as in
 *      the refinement passes of acquisition.py:81-101 (gb200_detect uses it);
 *  -1  (default) choose per call from the number of distinct Doppler values.
 * Results agree to float32 rounding.                                                                   */
int gb200_set_fused(gb200_engine* e, int mode);

/* Kernels launched by this engine so far (bench.py's gpu_launches). */
int gb200_launch_count(const gb200_engine* e, int64_t* out);

/* Measurement aid (bench.py roofline): when enabled, every doppler_spectra (which = 0) / correlate_cells
 * (which = 1) launch is bracketed by CUDA events on the engine's stream; gb200_kernel_timing synchronises and
 * returns the summed device time and the number of launches since timing was enabled.                 */
int gb200_enable_kernel_timing(gb200_engine* e, int on);
int gb200_kernel_timing(gb200_engine* e, int which, double* total_ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* GYPSUM_B200_H */
