// kernels.h - launch interface between the host side (api.cu) and the CUDA kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"

namespace adsb {

// ---- generic path (kernels_generic.cu) -----------------------------------

cudaError_t launch_envelope_generic(const uint16_t *raw, long long u_raw0, unsigned long long origin,
                                    long long j0, long long n, float *amp, int iq, cudaStream_t st);

cudaError_t launch_fill_pw(const uint16_t *raw, long long u_raw0, unsigned long long origin, CandRec *cands,
                           const unsigned long long *n_cand_ptr, unsigned long long cap, cudaStream_t st);

// pw for the n accepted frames of a job (*n_ptr, read on the device)
cudaError_t launch_fill_pw_frames(const uint16_t *raw, long long u_raw0, unsigned long long origin, FrameRec *frames,
                                  const unsigned long long *n_ptr, unsigned long long cap, bool wide, cudaStream_t st);

cudaError_t launch_scan_generic(const float *amp, long long j_amp0, long long n0, long long count,
                                int accept_df18, int want_tries, CandRec *cands,
                                unsigned long long cand_cap, unsigned long long *tries,
                                unsigned long long try_cap, Counters *ctr, const unsigned long long *fix_tab,
                                int fix_level, const unsigned long long *first_seen, cudaStream_t st);

// Packed 12-bit samples (libairspy layout, 8 samples per 12 bytes) -> u16 samples.  n_samples % 8 == 0,
// `packed` 4-byte and `out` 8-byte aligned.
cudaError_t launch_unpack12(const void *packed, uint16_t *out, long long n_samples, cudaStream_t st);

// ---- on-GPU ordering and greedy resolve (kernels_resolve.cu) ----------------

struct ResolveState {        // per stream: where the sequential scan stands (demod.c:128, :134)
	unsigned long long carry_start;  // start of the last accepted frame
	unsigned long long carry_pos;    // first offset behind it
	unsigned long long drift;        // iterations saved by skips so far: sum(skip - 1) (demod.c:99)
};

struct SlotResolve {         // per job, copied back to the host
	unsigned long long n_sorted;     // candidates of the job
	unsigned long long n_accept;     // of which accepted
	unsigned long long span0_start;  // scan position before this job (for the statistics pass and redo)
	unsigned long long span0_end;
	unsigned long long drift0;       // ResolveState::drift before this job
	unsigned int ok[3];              // accepted DF11 / DF17 / DF18 frames of the job (valid.c:53, :75)
	unsigned int pad_;
};

struct ResolveBuffers {      // per job slot scratch
	CandRec *sorted;         // cand_cap entries
	unsigned char *accept;   // cand_cap entries
	unsigned int *keys;      // cand_cap entries: offsets (relative to the job) grouped by bucket
	unsigned int *cnt, *off, *bacc, *bacc_x;        // bucket arrays, nb_cap + 1 entries each (cnt: zero between jobs)
	unsigned int *bdrift, *bdrift_x;                // sum(skip - 1) of the accepted frames per bucket / scanned
	unsigned long long *part;                       // totals of the 4096-entry tiles of a bucket scan
};

int resolve_bucket_count(long long n0, long long n1);
// wide: the chain will have the GPU to itself (four times the blocks of the grid-stride kernels)
cudaError_t launch_sort(const ResolveBuffers &rb, const CandRec *cands, const Counters *ctr, unsigned long long cap,
                        long long n0, long long n1, bool wide, cudaStream_t st);
cudaError_t launch_resolve(const ResolveBuffers &rb, const CandRec *cands, const Counters *ctr, unsigned long long cap,
                           long long n0, long long n1, ResolveState *state, SlotResolve *slot, FrameRec *out,
                           unsigned long long *keys_out, unsigned long long out_cap, bool wide, cudaStream_t st);
cudaError_t launch_set_state(ResolveState *state, unsigned long long carry_start, unsigned long long carry_pos,
                             unsigned long long drift, cudaStream_t st);
cudaError_t launch_count_tries_frames(const unsigned long long *tries, const unsigned long long *n_tries_ptr,
                                      unsigned long long cap, const FrameRec *frames, const SlotResolve *slot,
                                      unsigned long long frames_cap, const ResolveBuffers &rb, long long n0,
                                      long long n1, unsigned long long scan_end, unsigned long long *pend_out,
                                      unsigned long long *n_pend_out, unsigned long long pend_cap, StreamStats *stt,
                                      cudaStream_t st);

// ---- fused path (kernel_fused.cu) -----------------------------------------
//
// Three kernels per range of offsets:
//   k_fused  samples -> three bit arrays in HBM, one bit per offset (0.375 B per offset, L2-resident):
//            P (first preamble pulse pair beats its gap), Q (the mirrored test), B (PPM decision)
//   k_gate   word-parallel preamble AND + DF gate over the bit arrays -> list of "tries"
//   k_slice  PPM slicer + CRC-24 syndrome over the tries -> CRC-valid candidates
// Bit arrays: word g, bit 31-i holds offset J0 + 32 g + i (MSB first), J0 = the first span's j_begin.

// One contiguous run of envelope samples handled by one warp of k_fused ("span").
// Word W = the 32-bit pair of local samples (2W, 2W+1) = the newest input of envelope sample W.
struct FusedSpan {
	unsigned long long raw_w0;  // device address at which word 0 would sit (16-byte aligned; only
	                            // words in [w_lo, w_hi) are ever dereferenced)
	long long w_lo, w_hi;       // words present in the window, multiples of 4; words outside read as
	                            // mid-scale (before the stream start the FIR ring is zero, air.c:33)
	long long j_begin;          // first envelope sample computed: J0 + multiple of 224 (FIR phase 0, bit word 0)
	long long word_base;        // index in the bit arrays of the span's first bit word
	int n_words;                // bit words [0, n_words) of the span are stored
	int n_blocks;               // sub-blocks to run (each fused_block_outputs() envelope samples)
};

struct FusedParams {
	const FusedSpan *spans;
	int n_spans;
	int *span_ticket;      // dynamic span scheduler (zeroed before launch)
	Counters *ctr;         // only `wide` is written here
	unsigned int *bits_p, *bits_q, *bits_b;
	// parity tap (null in production): envelope of span 0, indexed from its j_begin
	float *dbg_amp;
	long long dbg_n;       // capacity of dbg_amp in samples
};

struct GateParams {
	const unsigned int *bits_p, *bits_q, *bits_b;
	long long j0;          // offset of bit 31 of word 0
	long long n0, n1;      // offsets owned by this launch: tries are emitted for n0 <= n < n1
	int accept_df18;
	unsigned long long *tries;
	unsigned long long try_cap;
	Counters *ctr;
	// address/parity extension (launch_gate_ap): the list above is the one of DF4/5/20/21 tries then
};

// EXTENSION (off by default): DF4 / DF5 / DF20 / DF21 carry CRC ^ ICAO address in their parity field.
// `first_seen[address]` (2^24 entries) is the smallest offset at which a clean DF11/17/18 candidate announces the
// address (~0: never); a frame of the four classes is a candidate when its syndrome is an address announced at
// an earlier offset.
struct AddressParityParams {
	const unsigned int *bits_b;
	long long j0;
	const unsigned long long *ap_tries;   // entries [0, ctr->n_ap)
	unsigned long long ap_cap;
	CandRec *cands;
	unsigned long long cand_cap;
	Counters *ctr;
	unsigned long long *first_seen;
};

struct SliceParams {
	const unsigned int *bits_b;
	long long j0;
	unsigned long long *tries;         // entries [ctr->try_mark, ctr->n_try) are sliced and get their DF class
	unsigned long long try_cap;
	CandRec *cands;
	unsigned long long cand_cap;
	Counters *ctr;
	const unsigned long long *fix_tab;  // repair.h table, null unless fix_level > 0 (extension, off by default)
	int fix_level;
};

// Geometry the host needs to cut spans.
int fused_block_outputs();      // envelope samples per sub-block (896)
int fused_span_quantum();       // span lengths are multiples of this many offsets (224 = lcm(28, 32))
int fused_threads();
size_t fused_smem_bytes();
// Configures the kernels (max dynamic smem, syndrome table) and reports the workers (warps) per SM; <0 on error.
int fused_prepare(int device);
cudaError_t launch_fused(const FusedParams &p, int workers, cudaStream_t st);
cudaError_t launch_gate(const GateParams &p, cudaStream_t st);
// extension: notes the addresses of candidates [ctr->cand_mark, ctr->n_cand), lists the DF4/5/20/21 tries of the
// range (launch_gate_ap, GateParams::tries = that list) and slices them against the address table
cudaError_t launch_note_addresses(const CandRec *cands, unsigned long long cand_cap, const Counters *ctr,
                                  unsigned long long *first_seen, cudaStream_t st);
cudaError_t launch_gate_ap(const GateParams &p, cudaStream_t st);
cudaError_t launch_slice_ap(const AddressParityParams &p, cudaStream_t st);
cudaError_t launch_slice(const SliceParams &p, cudaStream_t st);

}  // namespace adsb
