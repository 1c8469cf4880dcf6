/*
 * bauklank_stretch.h -- C ABI of libbauklank_stretch.so, the B200-native drop-in for the Signalsmith-Stretch hot path
 * of hanskerkhof/BAUKLANK-audio-stretch.
 *
 * Part 1 mirrors, name for name, the 18 exports the reference's JS glue binds from its WASM engine
 * (app/SignalsmithStretch.mjs:462-479: wasm exports "h".."y").  Same argument meaning, same conventions: planar
 * f32, the ENGINE owns the I/O buffer returned by setBuffers (HOST memory here, as in the reference where it is wasm
 * linear memory), no return codes -- an unrecoverable error aborts the process like the reference's wasm trap
 * (abort() import, :454-459).  One "current" engine instance per process, like one wasm module instance per node.
 *
 * Part 2 is the batched, handle-based form of the same operator for many independent streams with DEVICE-resident
 * audio: the worklet's per-quantum drive (app/SignalsmithStretch.mjs:826-954) is compiled on the host into a block
 * table and executed by a short pipeline of CUDA kernels per time chunk.  No torch types cross this boundary: plain pointers, sizes
 * and a cudaStream_t passed as void*.
 *
 * There is no CPU fallback: every entry point that computes requires a CUDA device.
 */
#ifndef BAUKLANK_STRETCH_H
#define BAUKLANK_STRETCH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- Part 1: the reference's own 18 entry points */
/* wasm export "h"  (app/SignalsmithStretch.mjs:463, used at :806-815): allocates 2*channels*length floats and
 * returns the base; in[c] = base + length*c, out[c] = base + length*(c+channels). */
float *setBuffers(int channels, int length);
int blockSamples(void);    /* "i" :464 */
int intervalSamples(void); /* "j" :465 */
int inputLatency(void);    /* "k" :466, used at :805, :800 */
int outputLatency(void);   /* "l" :467 */
void reset(void);          /* "m" :468, used at :792 */
void presetDefault(int channels, float sampleRate); /* "n" :469, used at :796 */
void presetCheaper(int channels, float sampleRate); /* "o" :470, used at :794 */
void configure(int channels, int blockSamples, int intervalSamples, int splitComputation); /* "p" :471, used at :791 */
void setTransposeFactor(float multiplier, float tonalityLimit);    /* "q" :472 */
void setTransposeSemitones(float semitones, float tonalityLimit);  /* "r" :473, used at :847 */
void setFormantFactor(float multiplier, int compensatePitch);      /* "s" :474 */
void setFormantSemitones(float semitones, int compensatePitch);    /* "t" :475, used at :848 */
void setFormantBase(float baseFreq);                               /* "u" :476, used at :849 */
void seek(int inputSamples, double playbackRate);                  /* "v" :477, used at :935 */
void process(int inputSamples, int outputSamples);                 /* "w" :478, used at :869, :882, :936 */
void flush(int outputSamples); /* "x" :479 -- bound but never called by the reference JS; implemented (W#46), bit-exact */
int stretch_main(int argc, char **argv); /* "y" :480 `_main`, a no-op (a shared object cannot export `main`) */
/* the reference seeds its RNG from crypto.getRandomValues (:382-394); a drop-in needs a way to pin it */
void stretch_set_seed(uint32_t seed);

/* ---------------------------------------------------------------- Part 2: batched streams on the device */
typedef struct bsb_engine bsb_engine;

/* one entry of the worklet's time map (app/SignalsmithStretch.mjs:587-600); JS numbers are doubles */
typedef struct bsb_segment {
  double output, input, rate;
  double semitones, tonality_hz, formant_semitones, formant_base_hz, loop_start, loop_end;
  int32_t active, formant_compensation;
  /* not NaN: the driver calls setTransposeFactor / setFormantFactor (wasm exports "q" / "s", :472, :474) with this
   * multiplier instead of the semitone setters the worklet uses; NaN = unset (the worklet's own behaviour) */
  double transpose_factor, formant_factor;
} bsb_segment;

/* configure()/presetDefault()/presetCheaper() for a batch; all streams of one engine share the configuration */
bsb_engine *bsb_create(int channels, int block_samples, int interval_samples, int split_computation, double sample_rate);
bsb_engine *bsb_create_preset(int channels, double sample_rate, int cheaper);
void bsb_destroy(bsb_engine *e);
int bsb_block_samples(const bsb_engine *e);
int bsb_interval_samples(const bsb_engine *e);
int bsb_input_latency(const bsb_engine *e);
int bsb_output_latency(const bsb_engine *e);
int bsb_fft_samples(const bsb_engine *e);
int bsb_bands(const bsb_engine *e);
const char *bsb_last_error(const bsb_engine *e);

/* Start describing a batch of n_streams independent streams.  d_clip / d_out are DEVICE pointers to planar f32
 * [channels][len]. */
int bsb_begin(bsb_engine *e, int n_streams);
/* buffer-playback drive: per render quantum `seek(bufferLength, rate); process(0, quantum)` (:883-943).  Inactive
 * segments (stop(), :861-869) run `process(quantum, quantum)` on a zeroed input buffer like the worklet does, including
 * the engine's silence gate: after 2*blockSamples of inactive output every call returns zeros.  A stream that is started
 * again after the gate has closed re-arms its block phase at an arbitrary output sample -- that one case is refused here
 * (bsb_last_error says so) and left to Part 1. */
int bsb_add_kiosk(bsb_engine *e, int stream, const float *d_clip, long long clip_len, float *d_out, long long n_out,
                  int quantum, const bsb_segment *segments, int n_segments, uint32_t seed);
/* The same drive, already resolved quantum by quantum (what a host-side mirror of the worklet's time map,
 * app/SignalsmithStretch.mjs:603-744 + :840-897, turns a control trace into): for render quantum k the f32 values the
 * three setters receive (:847-849), the segment's rate and Math.round(inputTime * sampleRate) (:897), and the range
 * [valid_start, valid_end) of clip samples the worklet's buffer store holds at that moment (zero outside). */
typedef struct bsb_quantum {
  double rate;
  long long input_samples_end, valid_start, valid_end;
  float semitones, tonality_limit, formant_semitones, formant_base;
  int32_t formant_compensation, active;
  float transpose_factor, formant_factor;   /* NaN = unset, see bsb_segment */
} bsb_quantum;
int bsb_add_kiosk_table(bsb_engine *e, int stream, const float *d_clip, long long clip_len, float *d_out, long long n_out,
                        int quantum, const bsb_quantum *table, long long n_quanta, uint32_t seed);
/* The same drive from a control trace: schedule() calls (remoteMethods.schedule, app/SignalsmithStretch.mjs:656-701), each
 * applied before render quantum `quantum`, edit the worklet's time map; every render quantum then looks its segment up like
 * process() does (:840-844).  This is what a controller's `set` messages become after the kiosk app's mapping
 * (app/multi/app.mjs:478-616, mirrored by the Python ControllerMapper).  NaN / negative = "not in the call's object":
 * output_time defaults to currentTime, input is extrapolated, rate / semitones / loop bounds / active are inherited from
 * the latest segment; tonality_hz, formant_semitones, formant_base_hz and formant_compensation are NOT inherited by the
 * reference (it would hand NaN to the engine) and must be given.  One stored audio buffer = the whole clip. */
typedef struct bsb_trace_event {
  long long quantum;
  double output_time, input, rate, semitones, loop_start, loop_end;
  double tonality_hz, formant_semitones, formant_base_hz;
  int32_t active, formant_compensation;
  double transpose_factor, formant_factor;   /* NaN = unset, see bsb_segment */
} bsb_trace_event;
int bsb_add_kiosk_trace(bsb_engine *e, int stream, const float *d_clip, long long clip_len, float *d_out, long long n_out,
                        int quantum, const bsb_trace_event *events, long long n_events, uint32_t seed);
/* configure() arithmetic without a device (W#25): out = {fftSamples, bands, inputLatency, outputLatency, longStep, inner*16+outer} */
int bsb_query_geometry(int block_samples, int interval_samples, int split_computation, int out[6]);
/* streaming drive: `process(n_in, n_out)` n_calls times over a contiguous input (:870-882 generalised) */
int bsb_add_streaming(bsb_engine *e, int stream, const float *d_clip, long long clip_len, float *d_out, int n_in,
                      int n_out, long long n_calls, const bsb_segment *segments, int n_segments, uint32_t seed);
/* upload the block tables, allocate per-stream state; chunk_blocks = blocks per kernel launch (0 = automatic) */
int bsb_commit(bsb_engine *e, int chunk_blocks);
/* run every block of every stream on `cuda_stream` (a cudaStream_t); state is reset first */
int bsb_run(bsb_engine *e, void *cuda_stream);
/* Same as bsb_run for audio that lives in HOST memory (ideally pinned): h_clips[s] / h_outs[s] are planar f32
 * [channels][clip_len] / [channels][n_out] of stream s.  The device buffers given to bsb_add_* are used as staging; the
 * copies are pipelined with the kernels time chunk by time chunk (input of chunk i+1 and output of chunk i-1 move
 * while chunk i computes).  Everything is ordered after, and joined back into, `cuda_stream`. */
int bsb_run_host(bsb_engine *e, const float *const *h_clips, float *const *h_outs, void *cuda_stream);
/* (h_outs may be NULL: the inputs come from the host, the outputs stay in the device buffers given to bsb_add_*) */
/* block until everything the last bsb_run / bsb_run_host queued has finished (the host outputs of bsb_run_host are only
 * complete after this, or after the caller's own synchronisation of `cuda_stream`) */
int bsb_synchronize(bsb_engine *e);
/* rebind the device I/O pointers of an already planned batch (same shapes) without re-planning */
int bsb_rebind(bsb_engine *e, int stream, const float *d_clip, float *d_out);
long long bsb_total_blocks(const bsb_engine *e);
long long bsb_stream_blocks(const bsb_engine *e, int stream);
int bsb_chunk_blocks(const bsb_engine *e);
/* kernel launches issued by the last bsb_run, and a read-out of one block record (for the indexing tests):
 * out[0]=flags out[1]=timeFactor bits, out[2..4]=cur window {start lo hi}, out[5..7]=prev window */
long long bsb_launch_count(const bsb_engine *e);
/* The reference's process() stops running blocks after 2*blockSamples silent input samples (its "silence gate", W#48
 * 7838-7943).  Where that depends on the audio the ahead-of-time block plan cannot follow it, so every run re-derives
 * from the clips as they are on the device: (streaming drives) how many process() calls the reference would have
 * gated; (kiosk drives with inactive segments) how many of the seeks the plan took for loud -- the buffer held clip
 * samples -- were in fact digital silence, which would have left the gate's counter running.  0 = the batch result is
 * the reference's.  Synchronises with the stream of the last run. */
long long bsb_gate_events(bsb_engine *e);
int bsb_block_info(const bsb_engine *e, int stream, long long block, long long out[8]);
/* Per-kernel accounting of the last bsb_run.  launches and units (analysis: window x channel transforms actually
 * computed; other kernels: channel-blocks) are always counted; device milliseconds only while profiling is on (one
 * CUDA event pair per launch on the run's stream, read back when a stat is queried -- the run is not serialised). */
void bsb_set_profiling(bsb_engine *e, int on);
/* chunk pipelining (default on): the chain + synthesis kernels of time chunk i run on a second internal CUDA stream
 * beside the analysis / map / term kernels of chunk i+1; off = every kernel in order on the caller's stream (used to
 * time kernels in isolation).  Results are identical either way. */
void bsb_set_overlap(bsb_engine *e, int on);
/* STFT kernels specialised for the preset geometries (default on where one exists: 48 kHz presetDefault / presetCheaper,
 * the kiosk's blockMs 200, 96 kHz presetDefault, block 960); off = the run-time-geometry kernels every other
 * configuration uses.  Results are identical either way (same butterflies, same order). */
void bsb_set_fast_fft(bsb_engine *e, int on);
int bsb_fast_fft_active(const bsb_engine *e);   /* 1 = this engine's runs use the specialised STFT kernels */
int bsb_kernel_count(const bsb_engine *e);
int bsb_kernel_stat(bsb_engine *e, int i, const char **name, double *ms, long long *launches, long long *units);
/* the launches of kernel i one by one (profiling on): device ms and units of up to `max` launches in launch order; returns
 * how many the last run made */
int bsb_kernel_launches(bsb_engine *e, int i, double *ms, long long *units, int max);

/* Self-test of the branch-free divide / square-root helpers of the chain kernel (device pointers, n elements):
 * q = x / d, r = sqrt(x) for positive d; flags bit0/bit1 = the operand pair was outside the helpers' safe range (the
 * kernel recomputes such steps with the plain IEEE operators).  Used by the GPU parity tests only. */
int bsb_selftest_arith(const float *d_x, const float *d_d, float *d_q, float *d_r, int *d_flags, int n);

#ifdef __cplusplus
}
#endif
#endif
