/*
 * adsbdec_b200.h - C ABI of the B200-native Mode S demodulation path.
 *
 * This library replaces, for the `adsbdec -f FILE` path only, the reference's
 * per-buffer calls into air.c / demod.c / valid.c.  Everything here is plain C:
 * pointers, sizes, integer return codes (0 = ok, <0 = error, the reference's
 * convention at air.c:117 and output.c:278-279).  No CUDA or torch types cross
 * the boundary; device pointers travel as `const void *`.
 *
 * Reference interface each entry point stands in for (file:line into the
 * reference tree, TLeconte/adsbdec):
 *
 *   adsbdec_b200_push          static void decodeiq(const unsigned short *r, const int len)   air.c:54
 *                              (called by fileInput, air.c:239, with 1 Mi-sample buffers)
 *                              + everything below it: deqframe (adsbdec.h:5, demod.c:84),
 *                              getdf/getabyte (demod.c:31-81), validShort/validLong
 *                              (demod.c:28-29, valid.c:39-82)
 *   adsbdec_b200_finish        end of fileInput's read loop -> handlerExit(0)                  air.c:243-244
 *   adsbdec_b200_flush / _poll bounded latency for a live rx_callback -> decodeiq source          air.c:173-177
 *   adsbdec_b200_sink_fn       void netout(const uint8_t *frame, const int len,
 *                                          const uint64_t ts, const uint32_t pw)                output.c:159
 *   adsbdec_b200_stats         stat_try / stat_ok read by print_stats()                         valid.c:30-31, 84-101
 *   adsbdec_b200_config.accept_df18   the global `df` set by -a                                 demod.c:26, :59, main.c:76
 *   adsbdec_b200_format_packet int formatpkt(blk_t *blk, char *pkt) (WITH_AIR build)            output.c:204-262
 *   adsbdec_b200_decode_file   void *fileInput(void *arg)                                       air.c:219-246
 *
 * Semantics that are kept bit-for-bit: float32 envelope
 * arithmetic incl. FMA contraction of the reference build, int-truncated
 * preamble sums, DF11/17/18 gate, syndrome==0 acceptance, greedy skip after an
 * accepted frame, the iteration-count timestamp, and the tail of the capture
 * that the reference never scans.
 *
 * There is NO CPU fallback: every decode entry point fails with
 * ADSB_B200_ENODEV when no CUDA device can be used.
 */
#ifndef ADSBDEC_B200_H
#define ADSBDEC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADSB_B200_ABI_VERSION 1

/* error codes (all negative) */
#define ADSB_B200_OK        0
#define ADSB_B200_EINVAL   -1   /* bad argument (NULL handle, n not a multiple of 4, ...) */
#define ADSB_B200_ENODEV   -2   /* no usable CUDA device / kernels not built for it      */
#define ADSB_B200_ECUDA    -3   /* a CUDA call failed; see adsbdec_b200_last_error()     */
#define ADSB_B200_ENOMEM   -4
#define ADSB_B200_ESTATE   -5   /* push after finish, ...                                */
#define ADSB_B200_EIO      -6   /* decode_file: cannot open / read                       */
#define ADSB_B200_ERANGE   -7   /* output capacity exhausted even after regrow           */

typedef struct adsbdec_b200 adsbdec_b200_t;

/* Frame sink: netout()'s signature (output.c:159) plus a user pointer.  Called on
 * the thread that calls push()/finish()/decode_*(), in non-decreasing ts order.
 * `frame` is only valid during the call (netout copies it, output.c:165). */
typedef void (*adsbdec_b200_sink_fn)(const uint8_t *frame, int len, uint64_t ts,
                                     uint32_t pw, void *user);

/* One accepted frame, as handed to the sink, plus its absolute 10 MS/s offset. */
typedef struct {
	uint64_t ts;        /* deqframe's iteration counter (demod.c:99)              */
	uint64_t n;         /* absolute output-sample offset of the preamble start    */
	uint32_t pw;        /* (p1+p2)/4 (demod.c:127)                                */
	int32_t len;        /* 7 or 14                                                */
	uint8_t bytes[14];
	uint8_t pad_[2];    /* [0]: bits repaired (always 0 unless the repair extension is on) */
} adsbdec_b200_frame_t;

typedef struct {
	int32_t abi_version;             /* ADSB_B200_ABI_VERSION                                    */
	int32_t device;                  /* CUDA device ordinal                                      */
	int32_t accept_df18;             /* -a: also decode DF18 (demod.c:59)                        */
	int32_t emulate_reference_tail;  /* 1: drop the tail the reference never scans (air.c:94-99) */
	int32_t collect_stats;           /* 1: maintain Try/Ok counters (valid.c:46,53,68,75)        */
	int32_t kernel;                  /* 0 = auto (fused sm_100a kernel), 1 = generic kernels only */
	uint64_t batch_samples;          /* host-push staging batch, 0 = default (32 Mi samples)     */
	uint64_t start_sample;           /* absolute index of the first pushed sample (0 for a file;
	                                    a time-chunk of a longer capture passes its offset so the
	                                    FIR rotation / fidx wrap (air.c:34,65) line up)           */
	uint64_t reserved[4];            /* [0], [1]: initial candidate / try list capacities (0 = default;
	                                    lists regrow on demand), [2]: threads decode_file reads a regular
	                                    file with (0 = up to 16, 1 = the calling thread only),
	                                    [3]: EXTENSION, error repair by syndrome lookup: 0 = off = the
	                                    reference (valid.c:52,:74 want syndrome 0), 1 = one flipped bit,
	                                    2 = also two flipped bits of a 112-bit frame; never the DF field;
	                                    | 0x100: EXTENSION, address/parity frames: DF4 / DF5 / DF20 / DF21, whose
	                                    parity field carries CRC ^ ICAO address, are accepted when a clean DF11 /
	                                    DF17 / DF18 frame announced that address at an earlier offset of the
	                                    stream (stream mode only; the Try/Ok counters do not count them);
	                                    | 0x200: EXTENSION, I/Q input: the pushed 16-bit values are SIGNED I/Q pairs
	                                    of a complex baseband stream at 10 MS/s (libairspy's INT16_IQ sample type)
	                                    instead of real samples at 20 MS/s: one pair per envelope sample, envelope
	                                    = I*I + Q*Q, no fs/4 mixer or FIR; everything behind the envelope as usual */
} adsbdec_b200_config_t;

/* Fills *cfg with defaults (device 0, reference-compatible behaviour). */
void adsbdec_b200_default_config(adsbdec_b200_config_t *cfg);

/* Creates a per-stream decoder on cfg->device.  Returns 0 and *out, or <0. */
int adsbdec_b200_open(const adsbdec_b200_config_t *cfg, adsbdec_b200_t **out);
void adsbdec_b200_close(adsbdec_b200_t *h);

/* Rewinds the handle to an empty stream (fidx = start_sample, ts = 0, counters 0) and keeps
 * its buffers: what restarting the reference process on another file does. */
int adsbdec_b200_reset(adsbdec_b200_t *h, uint64_t start_sample);

int adsbdec_b200_set_sink(adsbdec_b200_t *h, adsbdec_b200_sink_fn fn, void *user);

/* decodeiq(r, len) replacement.  `samples` is caller-owned HOST memory and may be
 * reused on return.  n must be a multiple of 4 (the reference consumes groups of
 * four and over-reads otherwise, air.c:59-82; decode_file emulates that case). */
int adsbdec_b200_push(adsbdec_b200_t *h, const uint16_t *samples, size_t n);

/* Same stream, other wire format: 12-bit samples packed eight to twelve bytes, the way the
 * airspy delivers them with packing on (the reference leaves it off, air.c:128-152; layout of libairspy's
 * unpack_samples: three little-endian 32-bit words, most significant bits first).  A quarter less to move
 * over PCIe; unpacked on the GPU into the same sample window, so everything behind it is unchanged.
 * `packed` is HOST memory, 4-byte aligned; n counts SAMPLES and must be a multiple of 8.  May be mixed
 * with adsbdec_b200_push() on one stream. */
int adsbdec_b200_push_packed12(adsbdec_b200_t *h, const void *packed, size_t n);

/* Same, with the samples already in DEVICE memory of cfg->device.  The pointer
 * must stay valid until the call returns. */
int adsbdec_b200_push_device(adsbdec_b200_t *h, const void *dev_samples, size_t n);

/* Live sources (the airspy callback, air.c:173-177, hands decodeiq 64 Ki..1 Mi samples at a time and
 * the reference decodes them before it returns).  push() only scans once a batch is full and hands
 * frames out two batches later; these two bound the latency instead:
 *   flush: submit whatever is staged as a scan job now (no-op if nothing is staged);
 *   poll:  deliver the frames of every scan job that has finished, without blocking.
 * Both return the number of frames delivered to the sink / frame list by the call, or < 0.
 * The decoded output does not depend on when, or whether, they are called. */
int adsbdec_b200_flush(adsbdec_b200_t *h);
int adsbdec_b200_poll(adsbdec_b200_t *h);

/* End of input: flushes staged samples, resolves and emits the remaining frames. */
int adsbdec_b200_finish(adsbdec_b200_t *h);

/* fileInput() replacement: 2 MiB reads, ragged-tail emulation, then finish().
 * A file that cannot be opened yields no frames and returns ADSB_B200_EIO
 * (the reference prints stats and exits 0, air.c:225-228; the CLI maps it). */
int adsbdec_b200_decode_file(adsbdec_b200_t *h, const char *path);

/* stat_try / stat_ok for DF 11, 17, 18, in print_stats order (valid.c:88-97). */
int adsbdec_b200_stats(adsbdec_b200_t *h, uint32_t try_[3], uint32_t ok[3]);

/* Frames accepted so far are also kept in the handle (in ts order, page-locked memory the GPU
 * writes them to; the pointer is valid until the next push / flush / poll / finish / clear). */
size_t adsbdec_b200_num_frames(const adsbdec_b200_t *h);
const adsbdec_b200_frame_t *adsbdec_b200_frames(const adsbdec_b200_t *h);
/* Drop ALL kept frames (long-running streams); a no-op while frames wait for the sink.  With a
 * sink installed the library does this itself once a million delivered frames have piled up. */
void adsbdec_b200_clear_frames(adsbdec_b200_t *h);

/* Device time (ms, CUDA events on the library's stream) spent in the demodulation
 * kernels since open / since the last call, and how many kernels were launched. */
int adsbdec_b200_kernel_time(adsbdec_b200_t *h, double *ms, uint64_t *launches, int reset);

/* Host-side telemetry since open / the last reset: where the caller's thread spent its time and
 * what the kernels reported. */
typedef struct {
	double wait_job_ms;      /* blocked waiting for a scan job to finish                      */
	double resolve_ms;       /* timestamps + deqframe call schedule over accepted frames       */
	double stats_issue_ms;   /* issuing the statistics pass                                    */
	double sink_ms;          /* sink calls and Ok counters                                     */
	double wait_upload_ms;   /* blocked at the end of push() until the caller's buffer is free */
	uint64_t jobs;           /* scan jobs completed                                            */
	uint64_t candidates;     /* CRC-valid candidates reported by the kernels (before the greedy rule) */
	uint64_t tries;          /* gate-passing offsets reported (valid.c:46, :68 before exclusion)     */
} adsbdec_b200_telemetry_t;
int adsbdec_b200_telemetry(adsbdec_b200_t *h, adsbdec_b200_telemetry_t *out, int reset);

/* Per-kernel timing of the fused kernel alone (for the roofline line). */
int adsbdec_b200_fused_time(adsbdec_b200_t *h, double *ms, uint64_t *launches,
                            uint64_t *samples, int reset);
/* Time of the two light kernels behind it (k_gate + k_slice: bit arrays -> tries -> candidates),
 * summed over the same launches. */
int adsbdec_b200_gate_time(adsbdec_b200_t *h, double *ms, int reset);

/* formatpkt(): fmt 0 = avr "*HEX;\n", 1 = avrmlat "@TTTTTTTTTTTT HEX;\n" (12 MHz
 * timestamp = ts*12/10, output.c:218), 2 = beast binary.  pkt must hold 256 bytes
 * (output.c:267).  Returns the packet length. */
int adsbdec_b200_format_packet(int fmt, const uint8_t *frame, int len, uint64_t ts,
                               uint32_t pw, char *pkt);
/* The same for an array of frames, packets back to back - the byte stream runOutput()
 * (output.c:286-336) writes for them.  Returns the bytes written (<= 48 per frame) or <0. */
int64_t adsbdec_b200_format_frames(int fmt, const adsbdec_b200_frame_t *frames, size_t count,
                                   char *out, size_t cap);

/* Debug / parity tap: envelope samples amp[first_out .. first_out+count) of a capture
 * held in device or host memory, computed by the generic kernel (is_device = 0/1).
 * start_sample as in the config.  Output is HOST float32. */
int adsbdec_b200_envelope(int device, const void *samples, size_t n_samples, int is_device,
                          uint64_t start_sample, float *amp_out, size_t count);

/* Debug / parity tap of the FUSED kernel: runs it over a host capture with one CTA and returns,
 * for the first `count` (multiple of 32) offsets from *j_begin_out, the envelope it computed in
 * shared memory and its three decision bit arrays (MSB of word 0 = offset j_begin), with
 * f[n] = (int)(a[n]+a[n+10]):
 *   p bit n = f[n] > 2*f[n+5]      (demod.c:102-107: p1 > 2*s1 at offset n)
 *   q bit n = f[n+5] > 2*f[n]      (demod.c:102-107: p2 > 2*s2 at offset n-30)
 *   b bit n = a[n] > a[n+5]        (demod.c:34)
 * Bits are only complete for offsets the kernel's stage 2 reached (all but the last ~2000). */
int adsbdec_b200_debug_fused(adsbdec_b200_t *h, const uint16_t *samples, size_t n_samples,
                             uint64_t start_sample, float *amp_out, uint32_t *p_out, uint32_t *q_out,
                             uint32_t *b_out, size_t count, int64_t *j_begin_out);

/* ---- host-only helpers (no CUDA call; usable without a GPU) -------------- */

/* Sequential semantics of demod.c:84-144 + air.c:94-99 applied to a sorted list of
 * CRC-valid candidates (greedy skip, ts drift, tail cut-off).  Used by the decoder
 * itself and exposed so that multi-GPU hosts can merge per-chunk candidate lists. */
typedef struct adsbdec_b200_resolver adsbdec_b200_resolver_t;
typedef struct {
	uint64_t n;          /* absolute offset              */
	uint32_t pw;
	uint32_t len;        /* 7 or 14                       */
	uint8_t bytes[16];   /* first `len` bytes significant */
} adsbdec_b200_candidate_t;

adsbdec_b200_resolver_t *adsbdec_b200_resolver_new(int emulate_reference_tail);
void adsbdec_b200_resolver_free(adsbdec_b200_resolver_t *r);
/* Candidates must arrive in non-decreasing n over the life of the resolver and be
 * complete for all n < scanned_upto; total_out = envelope samples produced so far. */
int adsbdec_b200_resolver_feed(adsbdec_b200_resolver_t *r, const adsbdec_b200_candidate_t *c,
                               size_t count, uint64_t scanned_upto, uint64_t total_out,
                               adsbdec_b200_sink_fn fn, void *user);
/* The same schedule over frames that already went through the greedy rule and carry their timestamps -
 * what the stream decoder feeds itself from kernels_resolve.cu; only the deqframe call schedule (which frames
 * the reference's scan ever reaches, air.c:94-99) is applied.  ok[0..2] (may be NULL) += accepted DF11 /
 * DF17 / DF18 frames.  A resolver takes candidates or stamped frames, not both. */
int adsbdec_b200_resolver_feed_frames(adsbdec_b200_resolver_t *r, const adsbdec_b200_frame_t *f, size_t count,
                                      uint64_t scanned_upto, uint64_t total_out, uint32_t ok[3],
                                      adsbdec_b200_sink_fn fn, void *user);
/* The in-place form the stream decoder itself runs (api.cu, complete_job): the caller keeps ALL stamped frames of
 * the stream in one array and hands in only what the call schedule looks at, keys[i] = frame i's start << 1 |
 * (frame i has 112 bits).  keys[0 .. n_vis) have been reached by the schedule before, keys[n_vis .. n_total) are new
 * or still in front of it.  Returns the new n_vis (finish != 0: the last calls at the end of the stream,
 * air.c:243-244), or < 0.  Nothing is copied, no sink is called; do not mix with feed / feed_frames on one resolver. */
int64_t adsbdec_b200_resolver_advance_keys(adsbdec_b200_resolver_t *r, const uint64_t *keys, size_t n_vis,
                                           size_t n_total, uint64_t scanned_upto, uint64_t total_out, int finish);
int adsbdec_b200_resolver_finish(adsbdec_b200_resolver_t *r, adsbdec_b200_sink_fn fn, void *user);
/* Number of spans accepted so far / copy them ([start, start+skip) pairs). */
uint64_t adsbdec_b200_resolver_scan_end(const adsbdec_b200_resolver_t *r);
/* Without a sink (fn == NULL in feed / feed_frames / finish) the accepted frames are kept in the
 * resolver, in ts order, and read here (valid until the next call on the resolver). */
size_t adsbdec_b200_resolver_num_frames(const adsbdec_b200_resolver_t *r);
const adsbdec_b200_frame_t *adsbdec_b200_resolver_frames(const adsbdec_b200_resolver_t *r);

/* Raw candidate access for multi-GPU time-chunking: scan one chunk of a longer
 * capture and return every CRC-valid candidate whose preamble starts in [own_begin, own_end)
 * (offsets LOCAL to the chunk: 0 = its first sample pair), WITHOUT applying the sequential
 * rules - the host merges chunk lists by offset and runs the resolver once.  The chunk is
 * device-resident and 16-byte aligned; start_sample (multiple of 4) is the absolute index of
 * its first sample in the capture, so that the FIR rotation and the fidx wrap (air.c:34, :65)
 * line up.  History before the chunk reads as zero, so own_begin must be >= 8 unless the chunk
 * starts the capture; own_end + 1200 envelope samples must lie inside the chunk.  out must
 * hold cap entries; returns the count (sorted by n), or <0.  tries_out: gate-passing offsets
 * per DF class in the owned range (no sequential exclusion in chunk mode). */
int64_t adsbdec_b200_chunk_candidates(adsbdec_b200_t *h, const void *dev_samples, size_t n_samples,
                                      uint64_t start_sample, uint64_t own_begin, uint64_t own_end,
                                      adsbdec_b200_candidate_t *out, size_t cap,
                                      uint32_t tries_out[3]);

/* EXTENSION (see reserved[3]): which bits does the repair table flip for this 24-bit CRC syndrome?
 * Returns (first << 8) | second with positions counted from the first bit of a 112-bit frame (a 56-bit
 * frame is positions 56..111), second = 0xFF for a single bit, or -1 if no pattern of <= 2 bits has this
 * syndrome.  Host-only (no GPU needed): lets the table be checked against brute force. */
int adsbdec_b200_repair_lookup(uint32_t syndrome);

const char *adsbdec_b200_last_error(void);
const char *adsbdec_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
