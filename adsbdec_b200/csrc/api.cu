// api.cu - the C ABI (include/adsbdec_b200.h): per-stream decoder handle, host staging,
// kernel orchestration and the hand-off to the sequential resolver.
//
// Reference seam (TLeconte/adsbdec): fileInput -> decodeiq -> deqframe -> validShort/Long -> netout
// (air.c:219-246, :54-101; demod.c:84-144; valid.c:39-82; output.c:159).  The host keeps the
// stream position, carries the un-scanned tail between batches the way air.c:96-98 carries
// ampbuff, and calls the sink in timestamp order on the caller's thread.
//
// Execution model: the stream is cut into batches; every batch becomes one scan JOB (kernels +
// result copies, all asynchronous on the compute stream).  Two jobs can be in flight: while the
// GPU scans batch k and the copy engine uploads batch k+1, the host sorts / resolves / emits the
// results of batch k-1.  The Try/Ok statistics pass of a job runs on the GPU after the job has
// been resolved (it needs the accepted spans), without any host round trip.
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <memory>
#include <new>
#include <thread>
#include <vector>

#include "../../include/adsbdec_b200.h"
#include "common.h"
#include "kernels.h"
#include "readpool.h"
#include "repair.h"
#include "resolver.h"

using namespace adsb;

namespace {

thread_local char t_err[512] = "";

int fail(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(t_err, sizeof t_err, fmt, ap);
	va_end(ap);
	return code;
}

#define CU(call)                                                                               \
	do {                                                                                       \
		cudaError_t e_ = (call);                                                               \
		if (e_ != cudaSuccess)                                                                 \
			return fail(ADSB_B200_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_),      \
			            __FILE__, __LINE__);                                                   \
	} while (0)

#define RC(call)          \
	do {                  \
		int rc_ = (call); \
		if (rc_)          \
			return rc_;   \
	} while (0)

constexpr int64_t kCarryBack = 128;        // raw samples kept in front of the next owned offset
constexpr int64_t kCarryMax = 1 << 14;     // room in front of a batch for the carried tail
constexpr int64_t kDefaultBatch = 32ll << 20;
constexpr int64_t kDeviceJob = 512ll << 20;  // most samples per scan job when the data is already in HBM
constexpr int64_t kDeviceJobMin = 128ll << 20;  // a piece of twice this is split in two so that host and GPU overlap
constexpr int64_t kAmpChunk = 8ll << 20;   // generic path: offsets per envelope scratch pass
constexpr int64_t kMinFused = 1 << 16;     // ranges smaller than this use the generic kernels
constexpr int64_t kStageSlice = 4ll << 20; // pageable pushes are staged and uploaded in slices
constexpr int64_t kBulkRead = 16ll << 20;  // decode_file: samples fetched per parallel read
constexpr int64_t kPackSlice = 32ll << 20; // push_packed12: samples uploaded and unpacked per step
constexpr unsigned long long kPendCap = 1ull << 18;
constexpr int kSlots = 3;                  // scan jobs in flight: three keep the GPU busy while the host hands frames out
constexpr int kResolveSMs = 0;             // SMs k_fused leaves free: none - the resolve chain of the previous job runs in blocks of
                                           // 128 threads x 32 registers, which fit BESIDE k_fused on every SM (kernels_resolve.cu)

__global__ void k_reset_counters(Counters *c)
{
	c->n_cand = 0;
	c->n_try = 0;
	c->overflow = 0;
	c->wide = 0;
	c->try_mark = 0;
}

__global__ void k_zero_u64(unsigned long long *p) { *p = 0; }

// A new stream on an open handle: statistics, scan position and the per-job result blocks back to zero.
__global__ void k_reset_stream(StreamStats *stats, ResolveState *state, SlotResolve *s0, SlotResolve *s1, SlotResolve *s2)
{
	static_assert(kSlots == 3, "one pointer per job slot");
	const unsigned int t = threadIdx.x;
	unsigned int *blocks[5] = {reinterpret_cast<unsigned int *>(stats), reinterpret_cast<unsigned int *>(state),
	                           reinterpret_cast<unsigned int *>(s0), reinterpret_cast<unsigned int *>(s1),
	                           reinterpret_cast<unsigned int *>(s2)};
	const unsigned int words[5] = {sizeof(StreamStats) / 4, sizeof(ResolveState) / 4, sizeof(SlotResolve) / 4,
	                               sizeof(SlotResolve) / 4, sizeof(SlotResolve) / 4};
#pragma unroll
	for (int b = 0; b < 5; b++)
		if (t < words[b])
			blocks[b][t] = 0u;
}

__global__ void k_mark_cands(Counters *c) { c->cand_mark = c->n_cand; }

// Builds the span table of one fused launch on the device and marks where its tries will begin.
// Span i covers envelope samples [j0 + i*per, j0 + (i+1)*per) = bit words [i*per/32, (i+1)*per/32), cut off at
// `words` (per is a multiple of 224: FIR phase 0 and a bit-word boundary at every span start).
__global__ void k_make_spans(FusedSpan *sp, int *ticket, Counters *ctr, int n_sp, long long j0, long long per,
                             long long words, unsigned long long raw_w0, long long w_lo, long long w_hi)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i == 0) {
		*ticket = 0;
		ctr->try_mark = ctr->n_try;
		ctr->cand_mark = ctr->n_cand;
		ctr->n_ap = 0;
	}
	if (i >= n_sp)
		return;
	FusedSpan s;
	s.j_begin = j0 + (long long)i * per;
	s.word_base = (long long)i * (per / 32);
	const long long nw = min(per / 32, words - s.word_base);
	s.n_words = (int)nw;
	// stage 2 of sub-block k stores words 28k-7 .. 28k+20
	s.n_blocks = (int)((nw + 7 + 27) / 28);
	s.raw_w0 = raw_w0;
	s.w_lo = w_lo;
	s.w_hi = w_hi;
	sp[i] = s;
}

inline double now_ms()
{
	return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ADSB_B200_TRACE=1: host timeline of the job pipeline on stderr at finish() (development aid).
struct Tracer {
	bool on = getenv("ADSB_B200_TRACE") != nullptr;
	std::vector<std::pair<const char *, double>> ev;
	void mark(const char *what)
	{
		if (on)
			ev.emplace_back(what, now_ms());
	}
	void dump()
	{
		if (!on || ev.empty())
			return;
		const double t0 = ev.front().second;
		for (auto &e : ev)
			fprintf(stderr, "  [trace] %8.3f ms  %s\n", e.second - t0, e.first);
		ev.clear();
	}
};

inline int64_t align_down(int64_t v, int64_t a) { return v - (((v % a) + a) % a); }

// One scan of offsets [n0, n1) over a device window, possibly still running.
struct Job {
	bool active = false;
	const uint16_t *buf = nullptr;  // window: buf[0] is local sample u0, samples up to u1 present
	int64_t u0 = 0, u1 = 0;
	int64_t n0 = 0, n1 = 0;         // offsets scanned (n1 <= n0: nothing, only a resolver feed)
	int64_t total_out = 0;          // envelope samples that exist once this window is in
	unsigned long long origin = 0;
	bool used_fused = false;
	bool resolve = true;            // feed the stream resolver on completion
	bool gpu_idle_after = false;    // nothing is queued behind this job's scan: its ordering chain runs alone
};

}  // namespace

struct adsbdec_b200 {
	adsbdec_b200_config_t cfg;
	int device = 0;
	cudaStream_t st = nullptr;       // kernels + result copies
	cudaStream_t st_copy = nullptr;  // uploads of sample batches
	cudaStream_t st_d2h = nullptr;   // a finished job's keys and frames on their way to the host store
	// Ordering / greedy / compaction / statistics of a job's candidates: a chain of small kernels that runs
	// BESIDE the next job's scan, in blocks of 128 threads x 32 registers - what is left on an SM next to a
	// k_fused block (kernels_resolve.cu).
	cudaStream_t st_res = nullptr;
	cudaStream_t st_stat = nullptr;  // the Try statistics passes (they only read a finished job's lists)
	cudaEvent_t ev_scan[kSlots] = {};   // the slot's scan kernels are done (recorded on st)
	cudaEvent_t ev_stats[kSlots] = {};  // the slot's lists have been read by the statistics pass
	unsigned long long origin = 0;   // absolute index of local sample 0 (FIR phase / fidx wrap)

	// stream position (local sample / offset indices, 0 = first sample ever pushed)
	int64_t pushed = 0;      // samples received
	int64_t own_next = 0;    // first offset not yet handed to a scan job
	bool finished = false;

	// batch windows: dev[i] holds [carried tail | batch samples]
	int64_t batch = 0;       // capacity in samples
	uint16_t *pin[2] = {nullptr, nullptr};
	uint16_t *dev[2] = {nullptr, nullptr};
	cudaEvent_t ev_dev_free[2] = {nullptr, nullptr};  // last reader of dev[i] (job + carry copy) done
	cudaEvent_t ev_pin_free[2] = {nullptr, nullptr};  // last upload out of pin[i] done
	cudaEvent_t ev_h2d = nullptr;
	// a job's results on their way to the host (st_d2h): its keys have arrived / its frames have arrived
	cudaEvent_t ev_keys = nullptr, ev_frm = nullptr;
	int cur = 0;
	int64_t win_u0 = 0;      // local index of the first sample held in dev[cur] (carry start)
	int64_t win_off = 0;     // element offset in dev[cur] where local sample win_u0 sits
	int64_t dev_fill = 0;    // samples of the window whose upload has been issued (incl. carry)
	int64_t pin_pos = 0;     // next free element of pin[cur] in this window
	int64_t pin_fill = 0;    // samples staged at pin[cur][pin_pos - pin_fill ..), upload not issued

	// per-slot job state and result buffers
	Job job[kSlots];
	int next_slot = 0;
	Counters *d_ctr[kSlots] = {}, *h_ctr[kSlots] = {};
	CandRec *d_cands[kSlots] = {};
	unsigned long long cand_cap[kSlots] = {};
	unsigned long long *d_tries[kSlots] = {};
	unsigned long long try_cap[kSlots] = {};
	// on-GPU ordering + greedy resolve (kernels_resolve.cu)
	ResolveBuffers rb[kSlots] = {};
	int nb_cap[kSlots] = {};
	SlotResolve *d_slot[kSlots] = {}, *h_slot[kSlots] = {};
	FrameRec *d_acc[kSlots] = {};  // accepted frames of the job, in order, stamped
	unsigned long long *d_key[kSlots] = {};  // (start << 1 | long) of the same frames: all the call schedule needs
	ResolveState *d_state = nullptr;
	int last_slot = 0;
	unsigned int *d_bits[kSlots] = {};  // P | Q | B bit arrays of the slot's fused range
	long long bits_cap[kSlots] = {};                // words per array
	FusedSpan *d_fspans[kSlots] = {};
	int fspan_cap[kSlots] = {};
	int *d_ticket[kSlots] = {};
	cudaEvent_t ev_done[kSlots] = {};
	cudaEvent_t tr_ref = nullptr;    // ADSB_B200_TRACE: the event the GPU timeline is printed against
	cudaEvent_t ev_a[kSlots] = {}, ev_b[kSlots] = {};
	cudaEvent_t ev_fa[kSlots] = {}, ev_fb[kSlots] = {};
	cudaEvent_t ev_fc[kSlots] = {};     // behind k_gate + k_slice of the slot's last fused range

	// statistics
	StreamStats *d_stats = nullptr, *h_stats = nullptr;
	unsigned long long *d_pend[2] = {nullptr, nullptr};
	int pend_in = 0;

	float *d_amp = nullptr;
	int fused_ctas = 0;      // autonomous workers (warps) of the fused kernel on this device (0 = unavailable)

	Resolver *res = nullptr;
	// Accepted frames, page-locked so that every job's result lands at its final place straight from the GPU:
	// [0, n_vis) have been reached by the deqframe call schedule (what the API shows), [n_vis, n_total) are
	// accepted by the greedy scan but still in front of the schedule (air.c:94-99).
	adsbdec_b200_frame_t *frames = nullptr;
	unsigned long long *fkeys = nullptr;  // fkeys[i] = frames[i].n << 1 | (112-bit frame): what the schedule reads
	size_t frames_cap = 0, n_total = 0, n_vis = 0;
	size_t emitted = 0;
	size_t counted = 0;      // frames[0 .. counted) are in stat_ok already (the resolver counts while it copies)
	uint32_t stat_ok[3] = {0, 0, 0};
	adsbdec_b200_sink_fn sink = nullptr;
	void *sink_user = nullptr;

	double kernel_ms = 0, fused_ms = 0, gate_ms = 0;
	Tracer tr;
	uint64_t launches = 0, fused_launches = 0, fused_samples = 0;

	std::vector<adsbdec_b200_candidate_t> cand_tmp;
	adsbdec_b200_telemetry_t tel = {};
	std::unique_ptr<ReadPool> readers;  // decode_file helpers, started on first use
	unsigned long long *d_fix = nullptr;  // repair.h syndrome table (extension; null when off)
	// packed 12-bit input: device scratch for one slice, two pinned staging buffers for pageable callers
	unsigned char *d_pk = nullptr, *h_pk[2] = {nullptr, nullptr};
	cudaEvent_t ev_pk[2] = {nullptr, nullptr};
	int pk_i = 0;
	int fix_level = 0;
	// EXTENSION, off by default: DF4/5/20/21 accepted against the addresses announced so far
	bool ap_mode = false;
	bool next_job_is_last = false;  // set around the submit of a resident push's last job (Job::gpu_idle_after)
	bool iq_mode = false;  // EXTENSION: the stream is signed 16-bit I/Q pairs at 10 MS/s, not real samples
	unsigned long long *d_first_seen = nullptr;     // [2^24] by ICAO address: offset of its first announcement
	unsigned long long *d_aptries[kSlots] = {};     // DF4/5/20/21 tries of the fused range being worked on
	unsigned long long ap_cap[kSlots] = {};

	// parity taps of the fused kernel (adsbdec_b200_debug_fused)
	float *dbg_amp = nullptr;
	long long dbg_n = 0, dbg_jb = 0;
	int dbg_slot = 0;
};

namespace {

int sync_compute(adsbdec_b200 *h)
{
	CU(cudaStreamSynchronize(h->st));
	CU(cudaStreamSynchronize(h->st_res));
	CU(cudaStreamSynchronize(h->st_stat));
	return 0;
}

int alloc_slot_outputs(adsbdec_b200 *h, int s, unsigned long long cand_cap, unsigned long long try_cap)
{
	if (cand_cap != h->cand_cap[s]) {
		if (h->d_cands[s]) cudaFree(h->d_cands[s]);
		if (h->d_acc[s]) cudaFree(h->d_acc[s]);
		if (h->d_key[s]) cudaFree(h->d_key[s]);
		if (h->rb[s].sorted) cudaFree(h->rb[s].sorted);
		if (h->rb[s].accept) cudaFree(h->rb[s].accept);
		if (h->rb[s].keys) cudaFree(h->rb[s].keys);
		h->d_cands[s] = h->rb[s].sorted = nullptr;
		h->d_acc[s] = nullptr;
		h->d_key[s] = nullptr;
		h->rb[s].accept = nullptr;
		h->rb[s].keys = nullptr;
		h->cand_cap[s] = 0;  // a failed allocation below must not leave the old capacity behind
		CU(cudaMalloc(&h->d_cands[s], cand_cap * sizeof(CandRec)));
		CU(cudaMalloc(&h->d_acc[s], cand_cap * sizeof(FrameRec)));
		CU(cudaMalloc(&h->d_key[s], cand_cap * sizeof(unsigned long long)));
		CU(cudaMalloc(&h->rb[s].sorted, cand_cap * sizeof(CandRec)));
		CU(cudaMalloc(&h->rb[s].accept, cand_cap));
		CU(cudaMalloc(&h->rb[s].keys, cand_cap * sizeof(unsigned int)));
		h->cand_cap[s] = cand_cap;
	}
	if (try_cap != h->try_cap[s]) {
		if (h->d_tries[s]) cudaFree(h->d_tries[s]);
		h->d_tries[s] = nullptr;
		h->try_cap[s] = 0;
		CU(cudaMalloc(&h->d_tries[s], try_cap * sizeof(unsigned long long)));
		h->try_cap[s] = try_cap;
		if (h->ap_mode) {  // four bit patterns pass that gate where three pass the reference's
			if (h->d_aptries[s]) cudaFree(h->d_aptries[s]);
			h->d_aptries[s] = nullptr;
			h->ap_cap[s] = 0;
			CU(cudaMalloc(&h->d_aptries[s], 2 * try_cap * sizeof(unsigned long long)));
			h->ap_cap[s] = 2 * try_cap;
		}
	}
	return 0;
}

// Lists of an idle slot sized up front for a job of `offsets` offsets (dense traffic: ~0.1 % of the offsets
// are CRC-valid candidates, ~2.6 % pass the gate), so that a big job does not start with a redo or two.
int presize_lists(adsbdec_b200 *h, int s, int64_t offsets)
{
	const unsigned long long cc = std::max<unsigned long long>(h->cand_cap[s], (unsigned long long)offsets / 256);
	// (the try list is also the fused path's survivor list: needed with or without statistics)
	const unsigned long long tc = std::max<unsigned long long>(h->try_cap[s], (unsigned long long)offsets / 24);
	if (cc == h->cand_cap[s] && tc == h->try_cap[s])
		return 0;
	RC(sync_compute(h));
	return alloc_slot_outputs(h, s, cc, tc);
}

int ensure_buckets(adsbdec_b200 *h, int s, int nb)
{
	if (nb + 1 <= h->nb_cap[s])
		return 0;
	RC(sync_compute(h));
	ResolveBuffers &rb = h->rb[s];
	unsigned int **arr[6] = {&rb.cnt, &rb.off, &rb.bacc, &rb.bacc_x, &rb.bdrift, &rb.bdrift_x};
	const int cap = nb + 1 + 4096;
	for (unsigned int **a : arr) {
		if (*a) cudaFree(*a);
		*a = nullptr;
		CU(cudaMalloc(a, (size_t)cap * sizeof(unsigned int)));
	}
	if (rb.part) cudaFree(rb.part);
	rb.part = nullptr;
	CU(cudaMalloc(&rb.part, (size_t)(cap / 512 + 2) * sizeof(unsigned long long)));
	// the bucket counts are zero between jobs: the scatter counts them back down (kernels_resolve.cu)
	CU(cudaMemset(rb.cnt, 0, (size_t)cap * sizeof(unsigned int)));
	h->nb_cap[s] = cap;
	return 0;
}

int ensure_batch_buffers(adsbdec_b200 *h, bool need_pinned)
{
	for (int i = 0; i < 2; i++) {
		if (!h->dev[i])
			CU(cudaMalloc(&h->dev[i], (size_t)(kCarryMax + h->batch + 64) * sizeof(uint16_t)));
		if (need_pinned && !h->pin[i])
			CU(cudaMallocHost(&h->pin[i], (size_t)h->batch * sizeof(uint16_t)));
	}
	return 0;
}

// ---- kernel launches of one job ----------------------------------------------------------

int launch_generic(adsbdec_b200 *h, int s, const Job &j, int64_t n0, int64_t n1)
{
	if (!h->d_amp)
		CU(cudaMalloc(&h->d_amp, (size_t)(kAmpChunk + kLookahead) * sizeof(float)));
	for (int64_t c0 = n0; c0 < n1; c0 += kAmpChunk) {
		const int64_t cnt = std::min(kAmpChunk, n1 - c0);
		CU(launch_envelope_generic(j.buf, j.u0, j.origin, c0, cnt + kLookahead, h->d_amp, h->iq_mode ? 1 : 0, h->st));
		if (h->ap_mode) {
			k_mark_cands<<<1, 1, 0, h->st>>>(h->d_ctr[s]);
			CU(cudaGetLastError());
		}
		CU(launch_scan_generic(h->d_amp, c0, c0, cnt, h->cfg.accept_df18, h->cfg.collect_stats, h->d_cands[s],
		                       h->cand_cap[s], h->d_tries[s], h->try_cap[s], h->d_ctr[s], h->d_fix, h->fix_level,
		                       nullptr, h->st));
		h->launches += 2;
		if (h->ap_mode) {
			// extension: the chunk's clean DF11/17/18 candidates announce their addresses, then a second pass
			// takes the DF4/5/20/21 frames whose address has been announced at an earlier offset
			CU(launch_note_addresses(h->d_cands[s], h->cand_cap[s], h->d_ctr[s], h->d_first_seen, h->st));
			CU(launch_scan_generic(h->d_amp, c0, c0, cnt, h->cfg.accept_df18, 0, h->d_cands[s], h->cand_cap[s],
			                       h->d_tries[s], h->try_cap[s], h->d_ctr[s], nullptr, 0, h->d_first_seen, h->st));
			h->launches += 3;
		}
	}
	return 0;
}

// Does a 2^32 boundary of the reference's `fidx` (air.c:34) fall where it would disturb the FIR
// windows of outputs [j0, j1)?  Those outputs go through the generic kernel.
bool wrap_in_range(unsigned long long org, int64_t j0, int64_t j1, int64_t *wrap_out)
{
	const unsigned long long a = (unsigned long long)std::max<int64_t>(0, 2 * j0 - 32) + org;
	const unsigned long long b = (unsigned long long)(2 * j1 + 32) + org;
	if ((a >> 32) == (b >> 32))
		return false;
	const unsigned long long w = ((a >> 32) + 1) << 32;  // first absolute sample of the next epoch
	*wrap_out = (int64_t)((w - org) / 2);                // envelope sample whose pair starts there
	return true;
}

// Runs the fused path over [n0, n1): k_fused (samples -> bit arrays), k_gate (-> tries), k_slice (-> candidates).
// Requires no fidx wrap in range.
int launch_fused_range(adsbdec_b200 *h, int s, const Job &j, int64_t n0, int64_t n1)
{
	const int S = fused_block_outputs();
	const int64_t Q = fused_span_quantum();
	// FIR phase.  Within one 2^32 epoch of the reference's fidx, slot(u) = (u + shift) % 14 with
	// shift = (origin - epoch_base) mod 14.  A span must start on an output j with
	// (2j + shift) % 14 == 0 (summation order becomes a compile-time constant per thread) and
	// j % 4 == 0 (16-byte aligned bulk copies): j = r (mod 28).
	const unsigned long long org = j.origin;
	const unsigned long long epoch_base = ((unsigned long long)(2 * n0) + org) & ~0xFFFFFFFFull;
	const int shift = (int)(((org % 14ull) + 14ull - (epoch_base % 14ull)) % 14ull);
	int r = -1;
	for (int c = 0; c < 28; c += 4)
		if ((2 * c + shift) % 14 == 0) {
			r = c;
			break;
		}
	if (r < 0)
		return fail(ADSB_B200_EINVAL, "internal: no FIR phase for shift %d", shift);
	// Bit arrays: word 0 starts at offset j0 <= n0 (FIR phase 0); the slicer reads the B bits of the last
	// owned offset up to n1 - 1 + 1190, the gate up to n1 - 1 + 120.
	const int64_t j0 = align_down(n0 - r, 28) + r;
	const int64_t words = (n1 + kLookahead + 31 - j0) / 32 + 1;
	// A worker is one warp of k_fused, and every worker gets exactly ONE span: a span computes up to one
	// sub-block of envelope past its end (the 15 samples of look-ahead of its last bits), so long spans keep
	// that redundancy small (60 s capture in three jobs: 67 k offsets per span, < 1 %), and the work per
	// offset is uniform, so equal spans finish together.  Small ranges: spans of at least 16 sub-blocks on
	// fewer workers.
	const int64_t total = words * 32;
	int64_t per = (total + h->fused_ctas - 1) / h->fused_ctas;
	per = std::max<int64_t>(per, 16 * (int64_t)S);
	per = std::min<int64_t>(per, 1ll << 29);
	per = (per + Q - 1) / Q * Q;
	const int n_sp = (int)((total + per - 1) / per);
	if (n_sp > h->fspan_cap[s]) {
		CU(cudaStreamSynchronize(h->st));
		if (h->d_fspans[s]) cudaFree(h->d_fspans[s]);
		h->d_fspans[s] = nullptr;
		h->fspan_cap[s] = 0;
		CU(cudaMalloc(&h->d_fspans[s], (size_t)(n_sp + 64) * sizeof(FusedSpan)));
		h->fspan_cap[s] = n_sp + 64;
	}
	const int64_t words_alloc = words + 16;  // the slicer fetches ten words per lane whatever it uses of them
	if (words_alloc > h->bits_cap[s]) {
		CU(cudaStreamSynchronize(h->st));
		if (h->d_bits[s]) cudaFree(h->d_bits[s]);
		h->d_bits[s] = nullptr;
		h->bits_cap[s] = 0;
		const int64_t cap = (words_alloc + words_alloc / 8 + 3) & ~(int64_t)3;  // the three arrays stay 16-byte aligned
		CU(cudaMalloc(&h->d_bits[s], (size_t)cap * 3 * sizeof(unsigned int)));
		h->bits_cap[s] = cap;
	}
	if (!h->d_ticket[s])
		CU(cudaMalloc(&h->d_ticket[s], sizeof(int)));
	if (j.u0 % 8 || ((uintptr_t)j.buf % 16))
		return fail(ADSB_B200_EINVAL, "internal: fused window not 16-byte aligned");
	const int64_t w_lo = j.u0 / 2, w_hi = align_down(j.u1 / 2, 4);
	if (j0 - 8 < w_lo && j.u0 != 0)
		return fail(ADSB_B200_EINVAL, "internal: window lacks FIR history");
	k_make_spans<<<(n_sp + 127) / 128, 128, 0, h->st>>>(
	    h->d_fspans[s], h->d_ticket[s], h->d_ctr[s], n_sp, j0, per, words,
	    (unsigned long long)(uintptr_t)j.buf - 2ull * (unsigned long long)j.u0, w_lo, w_hi);
	CU(cudaGetLastError());
	unsigned int *bp = h->d_bits[s], *bq = bp + h->bits_cap[s], *bb = bq + h->bits_cap[s];
	FusedParams p;
	p.spans = h->d_fspans[s];
	p.n_spans = n_sp;
	p.span_ticket = h->d_ticket[s];
	p.ctr = h->d_ctr[s];
	p.bits_p = bp;
	p.bits_q = bq;
	p.bits_b = bb;
	p.dbg_amp = h->dbg_amp;
	p.dbg_n = h->dbg_n;
	if (h->dbg_amp) {
		h->dbg_jb = j0;
		h->dbg_slot = s;
	}
	CU(cudaEventRecord(h->ev_fa[s], h->st));
	CU(launch_fused(p, std::min(h->fused_ctas, n_sp), h->st));
	CU(cudaEventRecord(h->ev_fb[s], h->st));
	GateParams g;
	g.bits_p = bp;
	g.bits_q = bq;
	g.bits_b = bb;
	g.j0 = j0;
	g.n0 = n0;
	g.n1 = n1;
	g.accept_df18 = h->cfg.accept_df18;
	g.tries = h->d_tries[s];
	g.try_cap = h->try_cap[s];
	g.ctr = h->d_ctr[s];
	CU(launch_gate(g, h->st));
	SliceParams sl;
	sl.bits_b = bb;
	sl.j0 = j0;
	sl.tries = h->d_tries[s];
	sl.try_cap = h->try_cap[s];
	sl.cands = h->d_cands[s];
	sl.cand_cap = h->cand_cap[s];
	sl.ctr = h->d_ctr[s];
	sl.fix_tab = h->d_fix;
	sl.fix_level = h->fix_level;
	CU(launch_slice(sl, h->st));
	CU(cudaEventRecord(h->ev_fc[s], h->st));
	h->launches += 4;
	if (h->ap_mode) {
		// extension: addresses announced by this range's clean candidates, then the DF4/5/20/21 tries of the
		// range against the table (same bit arrays, a list of their own)
		CU(launch_note_addresses(h->d_cands[s], h->cand_cap[s], h->d_ctr[s], h->d_first_seen, h->st));
		GateParams ga = g;
		ga.tries = h->d_aptries[s];
		ga.try_cap = h->ap_cap[s];
		CU(launch_gate_ap(ga, h->st));
		AddressParityParams ap;
		ap.bits_b = bb;
		ap.j0 = j0;
		ap.ap_tries = h->d_aptries[s];
		ap.ap_cap = h->ap_cap[s];
		ap.cands = h->d_cands[s];
		ap.cand_cap = h->cand_cap[s];
		ap.ctr = h->d_ctr[s];
		ap.first_seen = h->d_first_seen;
		CU(launch_slice_ap(ap, h->st));
		h->launches += 3;
	}
	h->fused_launches += 1;
	h->fused_samples += (uint64_t)(2 * (n1 - n0));
	return 0;
}

// Enqueues the kernels and result copies of job slot s (asynchronous).
int enqueue_job(adsbdec_b200 *h, int s, bool force_generic)
{
	Job &j = h->job[s];
	j.used_fused = false;
	if (j.n1 <= j.n0) {
		CU(cudaEventRecord(h->ev_done[s], h->st_res));
		return 0;
	}
	// the statistics pass of the job that used this slot before reads its lists on a stream of its own
	CU(cudaStreamWaitEvent(h->st, h->ev_stats[s], 0));
	CU(cudaStreamWaitEvent(h->st_res, h->ev_stats[s], 0));
	force_generic = force_generic || h->cfg.kernel == 1 || h->fused_ctas <= 0 || h->iq_mode;  // (no FIR in front of I/Q)
	k_reset_counters<<<1, 1, 0, h->st>>>(h->d_ctr[s]);
	CU(cudaGetLastError());
	CU(cudaEventRecord(h->ev_a[s], h->st));
	// pieces: [n0, n1) split around fidx wraps; each piece fused if large enough
	int64_t a = j.n0;
	while (a < j.n1) {
		int64_t b = j.n1, w = 0;
		if (wrap_in_range(j.origin, a, j.n1 + kLookahead, &w)) {
			// generic kernels over the neighbourhood of the wrap, fused (if any) before it
			const int64_t g0 = std::max(a, w - kLookahead - 64), g1 = std::min(j.n1, w + 64);
			if (g0 > a) {
				b = g0;
			} else {
				RC(launch_generic(h, s, j, a, g1));
				a = g1;
				continue;
			}
		}
		if (!force_generic && b - a >= kMinFused) {
			RC(launch_fused_range(h, s, j, a, b));
			j.used_fused = true;
		} else {
			RC(launch_generic(h, s, j, a, b));
		}
		a = b;
	}
	if (j.used_fused && !j.resolve) {  // chunk mode hands candidates out: all of them get their signal level
		CU(launch_fill_pw(j.buf, j.u0, j.origin, h->d_cands[s], &h->d_ctr[s]->n_cand, h->cand_cap[s], h->st));
		h->launches += 1;
	}
	// order the candidates on the GPU; for a stream job also apply the greedy skip there, so that
	// only accepted frames come back.  All of it on the resolve stream, behind the scan.
	RC(ensure_buckets(h, s, resolve_bucket_count(j.n0, j.n1)));
	CU(cudaEventRecord(h->ev_scan[s], h->st));
	CU(cudaStreamWaitEvent(h->st_res, h->ev_scan[s], 0));
	if (j.resolve) {
		CU(launch_resolve(h->rb[s], h->d_cands[s], h->d_ctr[s], h->cand_cap[s], j.n0, j.n1, h->d_state, h->d_slot[s],
		                  h->d_acc[s], h->d_key[s], h->cand_cap[s], j.gpu_idle_after, h->st_res));
		h->launches += 11;
		if (j.used_fused) {  // demod.c:127: the signal level, for the frames that were accepted
			CU(launch_fill_pw_frames(j.buf, j.u0, j.origin, h->d_acc[s], &h->d_slot[s]->n_accept, h->cand_cap[s],
			                         j.gpu_idle_after, h->st_res));
			h->launches += 1;
		}
	} else {
		CU(launch_sort(h->rb[s], h->d_cands[s], h->d_ctr[s], h->cand_cap[s], j.n0, j.n1, false, h->st_res));
		h->launches += 5;
	}
	CU(cudaEventRecord(h->ev_b[s], h->st_res));
	CU(cudaMemcpyAsync(h->h_ctr[s], h->d_ctr[s], sizeof(Counters), cudaMemcpyDeviceToHost, h->st_res));
	if (j.resolve)
		CU(cudaMemcpyAsync(h->h_slot[s], h->d_slot[s], sizeof(SlotResolve), cudaMemcpyDeviceToHost, h->st_res));
	// (the results themselves are fetched by complete_job, when their number and their place are known)
	CU(cudaEventRecord(h->ev_done[s], h->st_res));
	return 0;
}

void emit_new_frames(adsbdec_b200 *h)
{
	if (!h->sink && h->counted == h->n_vis) {
		h->emitted = h->n_vis;  // nothing to call, nothing to count
		return;
	}
	for (; h->emitted < h->n_vis; h->emitted++) {
		const adsbdec_b200_frame_t &f = h->frames[h->emitted];
		if (h->emitted >= h->counted) {
			const int df = f.bytes[0] >> 3;
			if (ok_class(df) >= 0)
				h->stat_ok[ok_class(df)]++;  // valid.c:53, :75
			h->counted = h->emitted + 1;
		}
		if (h->sink)
			h->sink(f.bytes, f.len, f.ts, f.pw, h->sink_user);
	}
	// a consumer that takes its frames through the sink never has to call clear_frames: the store is emptied
	// for it once it holds a million delivered frames (a live stream would otherwise grow it for ever)
	if (h->sink && h->n_vis >= (1u << 20))
		adsbdec_b200_clear_frames(h);
}

// Room for `more` frames behind the n_total the store holds.
int reserve_frames(adsbdec_b200 *h, size_t more)
{
	if (h->n_total + more <= h->frames_cap)
		return 0;
	size_t cap = std::max<size_t>(h->frames_cap * 2, 1 << 16);
	while (cap < h->n_total + more)
		cap *= 2;
	adsbdec_b200_frame_t *p = nullptr;
	unsigned long long *k = nullptr;
	CU(cudaMallocHost(&p, cap * sizeof(adsbdec_b200_frame_t)));
	CU(cudaMallocHost(&k, cap * sizeof(unsigned long long)));
	if (h->n_total) {
		memcpy(p, h->frames, h->n_total * sizeof(adsbdec_b200_frame_t));
		memcpy(k, h->fkeys, h->n_total * sizeof(unsigned long long));
	}
	if (h->frames)
		cudaFreeHost(h->frames);
	if (h->fkeys)
		cudaFreeHost(h->fkeys);
	h->frames = p;
	h->fkeys = k;
	h->frames_cap = cap;
	return 0;
}

inline int next_of(int s) { return (s + 1) % kSlots; }

// Stats pass for the tries of slot s (and the pending list), entirely on the GPU: a try counts if
// it lies below the resolver's scan end and outside the accepted frames of the job (which are
// still in d_acc[s]); tries beyond the scan end go to the other pending list.
int stats_pass(adsbdec_b200 *h, int s, bool with_slot_list)
{
	if (!h->cfg.collect_stats)
		return 0;
	const int pin = h->pend_in, pout = 1 - h->pend_in;
	const unsigned long long scan_end = h->res->scan_end();
	if (with_slot_list) {
		CU(launch_count_tries_frames(h->d_tries[s], &h->d_ctr[s]->n_try, h->try_cap[s], h->d_acc[s], h->d_slot[s],
		                             h->cand_cap[s], h->rb[s], h->job[s].n0, h->job[s].n1, scan_end, h->d_pend[pout],
		                             &h->d_stats->n_pend[pout], kPendCap, h->d_stats, h->st_stat));
		h->launches += 1;
	}
	CU(launch_count_tries_frames(h->d_pend[pin], &h->d_stats->n_pend[pin], kPendCap, h->d_acc[s], h->d_slot[s],
	                             h->cand_cap[s], h->rb[s], h->job[s].n0, h->job[s].n1, scan_end, h->d_pend[pout],
	                             &h->d_stats->n_pend[pout], kPendCap, h->d_stats, h->st_stat));
	k_zero_u64<<<1, 1, 0, h->st_stat>>>(&h->d_stats->n_pend[pin]);
	CU(cudaGetLastError());
	CU(cudaEventRecord(h->ev_stats[s], h->st_stat));
	h->launches += 2;
	h->pend_in = pout;
	return 0;
}

// Waits for job slot s, handles overflow / out-of-range fall-backs, feeds the resolver and emits.
int complete_job(adsbdec_b200 *h, int s)
{
	Job &j = h->job[s];
	if (!j.active)
		return 0;
	h->cand_tmp.clear();
	const bool scanned = j.n1 > j.n0;
	size_t n_res = 0;
	if (scanned) {
		bool force_generic = false;
		for (int attempt = 0;; attempt++) {
			const double t0 = now_ms();
			h->tr.mark("complete_job: wait");
			CU(cudaEventSynchronize(h->ev_done[s]));
			h->tr.mark("complete_job: job done");
			h->tel.wait_job_ms += now_ms() - t0;
			float ms = 0;
			cudaEventElapsedTime(&ms, h->ev_a[s], h->ev_b[s]);
			h->kernel_ms += ms;
			if (j.used_fused) {
				cudaEventElapsedTime(&ms, h->ev_fa[s], h->ev_fb[s]);
				h->fused_ms += ms;
				cudaEventElapsedTime(&ms, h->ev_fb[s], h->ev_fc[s]);
				h->gate_ms += ms;
				if (h->tr.on) {  // GPU timeline of the job relative to the start of the first job in flight's scan
					float t_fa = 0, t_fb = 0, t_fc = 0, t_b = 0;
					cudaEvent_t ref = h->tr_ref ? h->tr_ref : h->ev_a[s];
					if (!h->tr_ref)
						h->tr_ref = h->ev_a[s];
					cudaEventElapsedTime(&t_fa, ref, h->ev_fa[s]);
					cudaEventElapsedTime(&t_fb, ref, h->ev_fb[s]);
					cudaEventElapsedTime(&t_fc, ref, h->ev_fc[s]);
					cudaEventElapsedTime(&t_b, ref, h->ev_b[s]);
					fprintf(stderr, "  [trace-gpu] slot %d: k_fused %.3f..%.3f, gate+slice ..%.3f, chain done %.3f ms\n", s, t_fa,
					        t_fb, t_fc, t_b);
				}
			}
			const Counters &c = *h->h_ctr[s];
			if (j.used_fused && c.wide) {
				// samples beyond 12 bits: exact generic kernels
				force_generic = true;
			} else if (!(c.overflow & 7u)) {
				break;
			}
			if (attempt >= 8)
				return fail(ADSB_B200_ERANGE, "candidate lists keep overflowing");
			// Redo.  Everything issued since this job started ran on a truncated list: drain the GPU,
			// put the scan position back to where it was before this job, grow the lists and re-issue
			// this job and the newer ones in flight.
			RC(sync_compute(h));
			if (j.resolve)
				CU(launch_set_state(h->d_state, h->h_slot[s]->span0_start, h->h_slot[s]->span0_end,
				                    h->h_slot[s]->drift0, h->st_res));
			if (c.overflow & 7u)  // bit 2: the list of address/parity tries, sized with the try list
				RC(alloc_slot_outputs(h, s, (c.overflow & 1u) ? h->cand_cap[s] * 4 : h->cand_cap[s],
				                      (c.overflow & 6u) ? h->try_cap[s] * 4 : h->try_cap[s]));
			RC(enqueue_job(h, s, force_generic));
			for (int t = next_of(s); t != s; t = next_of(t))  // the newer jobs, oldest first
				if (h->job[t].active)
					RC(enqueue_job(h, t, false));
		}
		h->tel.candidates += h->h_ctr[s]->n_cand;
		h->tel.tries += h->cfg.collect_stats ? h->h_ctr[s]->n_try : 0;
		n_res = j.resolve ? (size_t)h->h_slot[s]->n_accept : (size_t)std::min<unsigned long long>(h->h_ctr[s]->n_cand, h->cand_cap[s]);
	} else {
		CU(cudaEventSynchronize(h->ev_done[s]));
	}
	j.active = false;
	static_assert(sizeof(CandRec) == sizeof(adsbdec_b200_candidate_t), "record layouts must agree");
	static_assert(sizeof(FrameRec) == sizeof(adsbdec_b200_frame_t), "record layouts must agree");
	if (!j.resolve) {
		h->cand_tmp.resize(n_res);  // chunk mode: sorted candidates for the caller
		if (n_res)
			CU(cudaMemcpy(h->cand_tmp.data(), h->rb[s].sorted, n_res * sizeof(CandRec), cudaMemcpyDeviceToHost));
		return 0;
	}
	// The greedy skip and the timestamps (demod.c:99, :128, :134) have been applied on the GPU.  The job's
	// accepted frames are fetched to their final place behind the earlier ones (the job is done: the copy
	// has nothing to wait for and runs beside whatever the GPU does next); what is left for the host is
	// the deqframe call schedule (air.c:94-99), which decides how far the scan has come - a few binary
	// searches over the new frames, no copy.
	const double t2 = now_ms();
	if (n_res) {
		RC(reserve_frames(h, n_res));
		// keys first: the call schedule below reads nothing else, and runs while the 40-byte records follow
		CU(cudaMemcpyAsync(h->fkeys + h->n_total, h->d_key[s], n_res * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
		                   h->st_d2h));
		CU(cudaEventRecord(h->ev_keys, h->st_d2h));
		CU(cudaMemcpyAsync(h->frames + h->n_total, h->d_acc[s], n_res * sizeof(FrameRec), cudaMemcpyDeviceToHost,
		                   h->st_d2h));
		CU(cudaEventRecord(h->ev_frm, h->st_d2h));
		CU(cudaEventSynchronize(h->ev_keys));
		h->tr.mark("complete_job: keys fetched");
		h->n_total += n_res;
	}
	h->n_vis = h->res->advance_keys(h->fkeys, h->n_vis, h->n_total, (uint64_t)std::max<int64_t>(j.n1, 0),
	                                (uint64_t)j.total_out);
	h->counted = h->n_vis;
	// (Ok counters: the GPU counted the job's accepted frames per class; the few the call schedule has not
	// reached yet are taken off again wherever the counters are read)
	if (scanned)
		for (int c = 0; c < 3; c++)
			h->stat_ok[c] += h->h_slot[s]->ok[c];
	const double t3 = now_ms();
	h->tr.mark("complete_job: frames fed");
	if (scanned) {  // (needs the scan end the schedule has just reached, nothing of the frames on the host)
		RC(stats_pass(h, s, true));
		h->last_slot = s;
	}
	const double t4 = now_ms();
	if (n_res)
		CU(cudaEventSynchronize(h->ev_frm));  // the frames themselves have landed
	emit_new_frames(h);
	h->tel.resolve_ms += t3 - t2;
	h->tel.stats_issue_ms += t4 - t3;
	h->tel.sink_ms += now_ms() - t4;
	h->tel.jobs += 1;
	return 0;
}

int complete_all(adsbdec_b200 *h)
{
	// oldest first: slots are used round-robin, next_slot holds the oldest job
	for (int k = 0, s = h->next_slot; k < kSlots; k++, s = next_of(s))
		RC(complete_job(h, s));
	return 0;
}

// Submits a scan of every offset that has its full look-ahead once the window [u0, u1) is in.
int submit_window(adsbdec_b200 *h, const uint16_t *buf, int64_t u0, int64_t u1)
{
	const int s = h->next_slot;
	RC(complete_job(h, s));  // the slot's previous job (two submissions ago)
	Job &j = h->job[s];
	j.buf = buf;
	j.u0 = u0;
	j.u1 = u1;
	j.total_out = u1 / 2;
	j.n0 = h->own_next;
	j.n1 = std::max<int64_t>(h->own_next, j.total_out - kLookahead);
	j.origin = h->origin;
	j.resolve = true;
	j.gpu_idle_after = h->next_job_is_last;
	j.active = true;
	h->own_next = j.n1;
	h->tr.mark("submit: enqueue");
	RC(enqueue_job(h, s, false));
	h->tr.mark("submit: enqueued");
	h->next_slot = next_of(s);
	return 0;
}

// Issues the upload of whatever is staged in pinned memory into the device window.
int drain_pinned(adsbdec_b200 *h)
{
	if (h->pin_fill == 0)
		return 0;
	CU(cudaMemcpyAsync(h->dev[h->cur] + h->win_off + h->dev_fill, h->pin[h->cur] + (h->pin_pos - h->pin_fill),
	                   (size_t)h->pin_fill * sizeof(uint16_t), cudaMemcpyHostToDevice, h->st_copy));
	CU(cudaEventRecord(h->ev_pin_free[h->cur], h->st_copy));
	h->dev_fill += h->pin_fill;
	h->pin_fill = 0;
	return 0;
}

// Opens the next window in the other device buffer, starting with the tail the next scan still
// needs (air.c:96-98 keeps the unscanned envelope; we keep the raw samples it is computed from).
// The copy is issued on the compute stream, behind the job that was just submitted.
int carry_to_next(adsbdec_b200 *h, const uint16_t *buf, int64_t u0, int64_t u1)
{
	int64_t c0 = std::max<int64_t>(u0, align_down(2 * h->own_next - kCarryBack, 8));
	c0 = std::max<int64_t>(c0, 0);
	const int64_t clen = u1 - c0;
	if (clen > kCarryMax)
		return fail(ADSB_B200_EINVAL, "internal: carry too large (%lld)", (long long)clen);
	const int nxt = 1 - h->cur;
	// local sample c0 (a multiple of 8) must land on a 16-byte boundary
	const int64_t off = kCarryMax - align_down(clen + 7, 8);
	if (clen > 0)
		CU(cudaMemcpyAsync(h->dev[nxt] + off, buf + (c0 - u0), (size_t)clen * sizeof(uint16_t),
		                   cudaMemcpyDeviceToDevice, h->st));
	h->cur = nxt;
	h->win_u0 = c0;
	h->win_off = off;
	h->dev_fill = clen;
	h->pin_pos = 0;
	h->pin_fill = 0;
	return 0;
}

// Ends the current batch window: scan it (asynchronously) and open the next one.
int flush_batch(adsbdec_b200 *h)
{
	RC(drain_pinned(h));
	const int old = h->cur;
	const int64_t u0 = h->win_u0, u1 = h->win_u0 + h->dev_fill;
	const uint16_t *buf = h->dev[old] + h->win_off;
	// the kernels must see the uploads
	CU(cudaEventRecord(h->ev_h2d, h->st_copy));
	CU(cudaStreamWaitEvent(h->st, h->ev_h2d, 0));
	RC(submit_window(h, buf, u0, u1));
	// The other device buffer is about to be refilled (carry copy now, uploads next), and a job may
	// have to be re-run on its window (list overflow, out-of-range samples): finish the job that used
	// it first.  Its host work overlaps the scan that was just submitted.
	for (int k = 0, t = h->next_slot; k < kSlots - 1; k++, t = next_of(t))  // all but the one just submitted
		RC(complete_job(h, t));
	RC(carry_to_next(h, buf, u0, u1));
	// dev[old] may be overwritten once this job and the carry copy out of it are done
	CU(cudaEventRecord(h->ev_dev_free[old], h->st));
	// uploads into the new window wait for the last readers of that buffer (recorded one flush ago)
	CU(cudaStreamWaitEvent(h->st_copy, h->ev_dev_free[h->cur], 0));
	return 0;
}

int64_t batch_room(const adsbdec_b200 *h)
{
	// samples that still fit in the current batch (the carry lives in the kCarryMax head room)
	const int64_t used = h->win_off + h->dev_fill + h->pin_fill - kCarryMax;
	return h->batch - std::max<int64_t>(0, used);
}

bool is_pinned_host(const void *p)
{
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	return at.type == cudaMemoryTypeHost;
}

}  // namespace

extern "C" {

void adsbdec_b200_default_config(adsbdec_b200_config_t *cfg)
{
	if (!cfg)
		return;
	memset(cfg, 0, sizeof *cfg);
	cfg->abi_version = ADSB_B200_ABI_VERSION;
	cfg->device = 0;
	cfg->accept_df18 = 0;
	cfg->emulate_reference_tail = 1;
	cfg->collect_stats = 1;
	cfg->kernel = 0;
	cfg->batch_samples = 0;
	cfg->start_sample = 0;
}

int adsbdec_b200_open(const adsbdec_b200_config_t *cfg, adsbdec_b200_t **out)
{
	if (!cfg || !out || cfg->abi_version != ADSB_B200_ABI_VERSION)
		return fail(ADSB_B200_EINVAL, "bad config / abi version");
	if (cfg->start_sample % 4)
		return fail(ADSB_B200_EINVAL, "start_sample must be a multiple of 4");
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count <= 0 || cfg->device < 0 || cfg->device >= count)
		return fail(ADSB_B200_ENODEV, "no usable CUDA device %d (%s); this library has no CPU fallback",
		            cfg->device, e == cudaSuccess ? "device count is 0 or ordinal out of range"
		                                          : cudaGetErrorString(e));
	CU(cudaSetDevice(cfg->device));
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, cfg->device));
	if (prop.major != 10)
		return fail(ADSB_B200_ENODEV, "device %d is sm_%d%d; kernels are built for sm_100a only",
		            cfg->device, prop.major, prop.minor);
	adsbdec_b200 *h = new (std::nothrow) adsbdec_b200();
	if (!h)
		return fail(ADSB_B200_ENOMEM, "out of memory");
	h->cfg = *cfg;
	h->origin = cfg->start_sample;
	h->device = cfg->device;
	h->batch = cfg->batch_samples ? (int64_t)cfg->batch_samples : kDefaultBatch;
	h->batch = std::max<int64_t>(h->batch, 1 << 17);
	h->batch = align_down(h->batch + 7, 8);
	h->res = new Resolver(cfg->emulate_reference_tail != 0);
	h->win_off = kCarryMax;
	int rc = 0;
	do {
		bool ok = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) == cudaSuccess &&
		          cudaStreamCreateWithFlags(&h->st_copy, cudaStreamNonBlocking) == cudaSuccess &&
		          cudaStreamCreateWithFlags(&h->st_d2h, cudaStreamNonBlocking) == cudaSuccess &&
		          cudaStreamCreateWithFlags(&h->st_res, cudaStreamNonBlocking) == cudaSuccess &&
		          cudaStreamCreateWithFlags(&h->st_stat, cudaStreamNonBlocking) == cudaSuccess &&
		          cudaEventCreateWithFlags(&h->ev_h2d, cudaEventDisableTiming) == cudaSuccess &&
		          cudaEventCreateWithFlags(&h->ev_keys, cudaEventDisableTiming) == cudaSuccess &&
		          cudaEventCreateWithFlags(&h->ev_frm, cudaEventDisableTiming) == cudaSuccess &&
		          cudaMalloc(&h->d_stats, sizeof(StreamStats)) == cudaSuccess &&
		          cudaMalloc(&h->d_state, sizeof(ResolveState)) == cudaSuccess &&
		          cudaMemset(h->d_state, 0, sizeof(ResolveState)) == cudaSuccess &&
		          cudaMallocHost(&h->h_stats, sizeof(StreamStats)) == cudaSuccess &&
		          cudaMemset(h->d_stats, 0, sizeof(StreamStats)) == cudaSuccess;
		for (int i = 0; ok && i < 2; i++) {
			ok = cudaEventCreateWithFlags(&h->ev_dev_free[i], cudaEventDisableTiming) == cudaSuccess &&
			     cudaEventCreateWithFlags(&h->ev_pin_free[i], cudaEventDisableTiming) == cudaSuccess &&
			     cudaMalloc(&h->d_pend[i], kPendCap * sizeof(unsigned long long)) == cudaSuccess;
		}
		for (int s = 0; ok && s < kSlots; s++) {
			ok = cudaEventCreate(&h->ev_a[s]) == cudaSuccess && cudaEventCreate(&h->ev_b[s]) == cudaSuccess &&
			     cudaEventCreate(&h->ev_fa[s]) == cudaSuccess && cudaEventCreate(&h->ev_fb[s]) == cudaSuccess &&
			     cudaEventCreate(&h->ev_fc[s]) == cudaSuccess &&
			     cudaEventCreateWithFlags(&h->ev_done[s], cudaEventDisableTiming) == cudaSuccess &&
			     cudaEventCreateWithFlags(&h->ev_scan[s], cudaEventDisableTiming) == cudaSuccess &&
			     cudaEventCreateWithFlags(&h->ev_stats[s], cudaEventDisableTiming) == cudaSuccess &&
			     cudaMalloc(&h->d_ctr[s], sizeof(Counters)) == cudaSuccess &&
			     cudaMalloc(&h->d_slot[s], sizeof(SlotResolve)) == cudaSuccess &&
			     cudaMemset(h->d_slot[s], 0, sizeof(SlotResolve)) == cudaSuccess &&
			     cudaMallocHost(&h->h_slot[s], sizeof(SlotResolve)) == cudaSuccess &&
			     cudaMallocHost(&h->h_ctr[s], sizeof(Counters)) == cudaSuccess &&
			     cudaMemset(h->d_ctr[s], 0, sizeof(Counters)) == cudaSuccess;
			if (ok)
				memset(h->h_ctr[s], 0, sizeof(Counters));
		}
		if (!ok) {
			rc = fail(ADSB_B200_ECUDA, "stream/event/counter setup: %s", cudaGetErrorString(cudaGetLastError()));
			break;
		}
		// reserved[0] / reserved[1]: initial list capacities (tests use tiny ones to exercise regrowth)
		// default: sized for a dense capture at this batch size (they regrow on demand anyway), so that
		// hundreds of small-batch decoders - one per receiver stream - stay cheap
		const unsigned long long cc =
		    cfg->reserved[0] ? cfg->reserved[0] : std::min<unsigned long long>(1ull << 18, std::max<unsigned long long>(4096, (unsigned long long)h->batch / 128));
		const unsigned long long tc =
		    cfg->reserved[1] ? cfg->reserved[1] : std::min<unsigned long long>(1ull << 20, std::max<unsigned long long>(16384, (unsigned long long)h->batch / 32));
		h->ap_mode = (cfg->reserved[3] & 0x100u) != 0;  // (sizes a list of its own below)
		for (int s = 0; s < kSlots && !rc; s++)
			rc = alloc_slot_outputs(h, s, cc, tc);
		if (rc) break;
		if (cfg->kernel != 1) {
			int per_sm = fused_prepare(cfg->device);  // workers (warps) per SM
			if (per_sm <= 0) {
				rc = fail(ADSB_B200_ENODEV, "fused kernel cannot be configured on device %d", cfg->device);
				break;
			}
			// (development knob: SMs left free for the resolve chain of the previous job; none by default, the
			// chain's blocks fit beside k_fused - see kResolveSMs)
			int reserve = kResolveSMs;
			if (const char *e = getenv("ADSB_B200_RESOLVE_SMS"))  // development knob
				reserve = atoi(e);
			const int sms = prop.multiProcessorCount > 4 * reserve && reserve >= 0 ? prop.multiProcessorCount - reserve
			                                                                      : prop.multiProcessorCount;
			h->fused_ctas = per_sm * sms;
		}
		// reserved[3]: opt-in error repair (repair.h) - 0 = the reference's behaviour
		if ((cfg->reserved[3] & 0xFFu) > 2 || (cfg->reserved[3] & ~0x3FFull)) {
			rc = fail(ADSB_B200_EINVAL,
			          "reserved[3]: repair level 0, 1 or 2, plus 0x100 for address/parity frames, 0x200 for I/Q input");
			break;
		}
		h->fix_level = (int)(cfg->reserved[3] & 0xFFu);
		h->ap_mode = (cfg->reserved[3] & 0x100u) != 0;
		h->iq_mode = (cfg->reserved[3] & 0x200u) != 0;
		if (h->ap_mode) {
			if (cudaMalloc(&h->d_first_seen, sizeof(unsigned long long) << 24) != cudaSuccess ||
			    cudaMemset(h->d_first_seen, 0xFF, sizeof(unsigned long long) << 24) != cudaSuccess) {
				rc = fail(ADSB_B200_ECUDA, "address table: %s", cudaGetErrorString(cudaGetLastError()));
				break;
			}
		}
		if (h->fix_level > 0) {
			std::vector<unsigned long long> tab(kFixSlots);
			fix_build_table(tab.data());
			if (cudaMalloc(&h->d_fix, kFixSlots * sizeof(unsigned long long)) != cudaSuccess ||
			    cudaMemcpy(h->d_fix, tab.data(), kFixSlots * sizeof(unsigned long long), cudaMemcpyHostToDevice) !=
			        cudaSuccess) {
				rc = fail(ADSB_B200_ECUDA, "repair table: %s", cudaGetErrorString(cudaGetLastError()));
				break;
			}
		}
	} while (0);
	if (rc) {
		adsbdec_b200_close(h);
		return rc;
	}
	*out = h;
	return 0;
}

void adsbdec_b200_close(adsbdec_b200_t *h)
{
	if (!h)
		return;
	cudaSetDevice(h->device);
	if (h->st_copy) cudaStreamSynchronize(h->st_copy);
	if (h->st) cudaStreamSynchronize(h->st);
	if (h->st_res) cudaStreamSynchronize(h->st_res);
	if (h->st_stat) cudaStreamSynchronize(h->st_stat);
	for (int i = 0; i < 2; i++) {
		if (h->dev[i]) cudaFree(h->dev[i]);
		if (h->pin[i]) cudaFreeHost(h->pin[i]);
		if (h->d_pend[i]) cudaFree(h->d_pend[i]);
		if (h->ev_dev_free[i]) cudaEventDestroy(h->ev_dev_free[i]);
		if (h->ev_pin_free[i]) cudaEventDestroy(h->ev_pin_free[i]);
	}
	for (int s = 0; s < kSlots; s++) {
		if (h->d_tries[s]) cudaFree(h->d_tries[s]);
		if (h->d_cands[s]) cudaFree(h->d_cands[s]);
		if (h->d_ctr[s]) cudaFree(h->d_ctr[s]);
		if (h->h_ctr[s]) cudaFreeHost(h->h_ctr[s]);
		if (h->d_acc[s]) cudaFree(h->d_acc[s]);
		if (h->d_key[s]) cudaFree(h->d_key[s]);
		if (h->rb[s].sorted) cudaFree(h->rb[s].sorted);
		if (h->rb[s].accept) cudaFree(h->rb[s].accept);
		if (h->rb[s].cnt) cudaFree(h->rb[s].cnt);
		if (h->rb[s].keys) cudaFree(h->rb[s].keys);
		if (h->rb[s].part) cudaFree(h->rb[s].part);
		if (h->rb[s].off) cudaFree(h->rb[s].off);
		if (h->rb[s].bacc) cudaFree(h->rb[s].bacc);
		if (h->rb[s].bacc_x) cudaFree(h->rb[s].bacc_x);
		if (h->rb[s].bdrift) cudaFree(h->rb[s].bdrift);
		if (h->rb[s].bdrift_x) cudaFree(h->rb[s].bdrift_x);
		if (h->d_slot[s]) cudaFree(h->d_slot[s]);
		if (h->h_slot[s]) cudaFreeHost(h->h_slot[s]);
		if (h->d_fspans[s]) cudaFree(h->d_fspans[s]);
		if (h->d_ticket[s]) cudaFree(h->d_ticket[s]);
		if (h->ev_a[s]) cudaEventDestroy(h->ev_a[s]);
		if (h->ev_b[s]) cudaEventDestroy(h->ev_b[s]);
		if (h->ev_fa[s]) cudaEventDestroy(h->ev_fa[s]);
		if (h->ev_fb[s]) cudaEventDestroy(h->ev_fb[s]);
		if (h->ev_fc[s]) cudaEventDestroy(h->ev_fc[s]);
		if (h->d_bits[s]) cudaFree(h->d_bits[s]);
		if (h->ev_done[s]) cudaEventDestroy(h->ev_done[s]);
		if (h->ev_scan[s]) cudaEventDestroy(h->ev_scan[s]);
		if (h->ev_stats[s]) cudaEventDestroy(h->ev_stats[s]);
	}
	if (h->d_fix) cudaFree(h->d_fix);
	if (h->d_pk) cudaFree(h->d_pk);
	for (int i = 0; i < 2; i++) {
		if (h->h_pk[i]) cudaFreeHost(h->h_pk[i]);
		if (h->ev_pk[i]) cudaEventDestroy(h->ev_pk[i]);
	}
	if (h->d_stats) cudaFree(h->d_stats);
	if (h->d_state) cudaFree(h->d_state);
	if (h->h_stats) cudaFreeHost(h->h_stats);
	if (h->d_amp) cudaFree(h->d_amp);
	if (h->ev_h2d) cudaEventDestroy(h->ev_h2d);
	if (h->ev_keys) cudaEventDestroy(h->ev_keys);
	if (h->ev_frm) cudaEventDestroy(h->ev_frm);
	if (h->st) cudaStreamDestroy(h->st);
	if (h->st_copy) cudaStreamDestroy(h->st_copy);
	if (h->st_d2h) cudaStreamDestroy(h->st_d2h);
	if (h->st_res) cudaStreamDestroy(h->st_res);
	if (h->st_stat) cudaStreamDestroy(h->st_stat);
	if (h->d_first_seen) cudaFree(h->d_first_seen);
	for (int s = 0; s < kSlots; s++)
		if (h->d_aptries[s]) cudaFree(h->d_aptries[s]);
	if (h->frames) cudaFreeHost(h->frames);
	if (h->fkeys) cudaFreeHost(h->fkeys);
	delete h->res;
	delete h;
}

int adsbdec_b200_reset(adsbdec_b200_t *h, uint64_t start_sample)
{
	if (!h)
		return fail(ADSB_B200_EINVAL, "null handle");
	if (start_sample % 4)
		return fail(ADSB_B200_EINVAL, "start_sample must be a multiple of 4");
	CU(cudaSetDevice(h->device));
	CU(cudaStreamSynchronize(h->st_copy));
	RC(sync_compute(h));
	h->origin = start_sample;
	h->cfg.start_sample = start_sample;
	h->pushed = 0;
	h->own_next = 0;
	h->finished = false;
	h->cur = 0;
	h->win_u0 = 0;
	h->win_off = kCarryMax;
	h->dev_fill = 0;
	h->pin_pos = 0;
	h->pin_fill = 0;
	h->next_slot = 0;
	h->pend_in = 0;
	for (int s = 0; s < kSlots; s++)
		h->job[s].active = false;
	delete h->res;
	h->res = new Resolver(h->cfg.emulate_reference_tail != 0);
	h->tr_ref = nullptr;
	h->n_total = h->n_vis = 0;  // the store keeps its capacity
	h->emitted = 0;
	h->counted = 0;
	for (int i = 0; i < 3; i++)
		h->stat_ok[i] = 0;
	// one launch instead of five memsets: everything that uses these runs on, or waits for, this stream
	k_reset_stream<<<1, 64, 0, h->st>>>(h->d_stats, h->d_state, h->d_slot[0], h->d_slot[1], h->d_slot[2]);
	CU(cudaGetLastError());
	if (h->d_first_seen)  // a new stream: nothing has been announced
		CU(cudaMemset(h->d_first_seen, 0xFF, sizeof(unsigned long long) << 24));
	return 0;
}

int adsbdec_b200_set_sink(adsbdec_b200_t *h, adsbdec_b200_sink_fn fn, void *user)
{
	if (!h)
		return fail(ADSB_B200_EINVAL, "null handle");
	h->sink = fn;
	h->sink_user = user;
	return 0;
}

int adsbdec_b200_push(adsbdec_b200_t *h, const uint16_t *samples, size_t n)
{
	if (!h || (!samples && n))
		return fail(ADSB_B200_EINVAL, "null argument");
	if (n % 4)
		return fail(ADSB_B200_EINVAL, "push length must be a multiple of 4 samples (air.c:59-82)");
	if (h->finished)
		return fail(ADSB_B200_ESTATE, "push after finish");
	if (n == 0)
		return 0;
	CU(cudaSetDevice(h->device));
	// Page-locked caller memory is uploaded straight from where it lies; pageable memory goes
	// through our own pinned staging buffers, in slices so that the upload of one slice overlaps
	// the memcpy of the next.
	const bool direct = is_pinned_host(samples);
	RC(ensure_batch_buffers(h, !direct));
	size_t done = 0;
	while (done < n) {
		const int64_t room = batch_room(h);
		if (room <= 0) {
			RC(flush_batch(h));
			continue;
		}
		size_t take = std::min<size_t>((size_t)room, n - done);
		if (direct) {
			RC(drain_pinned(h));
			CU(cudaMemcpyAsync(h->dev[h->cur] + h->win_off + h->dev_fill, samples + done, take * sizeof(uint16_t),
			                   cudaMemcpyHostToDevice, h->st_copy));
			h->dev_fill += (int64_t)take;
		} else {
			take = std::min<size_t>(take, (size_t)kStageSlice);
			if (h->pin_pos == 0)
				CU(cudaEventSynchronize(h->ev_pin_free[h->cur]));  // uploads of two windows ago
			memcpy(h->pin[h->cur] + h->pin_pos, samples + done, take * sizeof(uint16_t));
			h->pin_pos += (int64_t)take;
			h->pin_fill += (int64_t)take;
			if (h->pin_fill >= kStageSlice)
				RC(drain_pinned(h));
		}
		h->pushed += (int64_t)take;
		done += take;
	}
	if (direct) {
		const double t0 = now_ms();
		CU(cudaStreamSynchronize(h->st_copy));  // the caller may reuse its buffer on return
		h->tel.wait_upload_ms += now_ms() - t0;
	}
	return 0;
}

int adsbdec_b200_push_packed12(adsbdec_b200_t *h, const void *packed, size_t n)
{
	if (!h || (!packed && n))
		return fail(ADSB_B200_EINVAL, "null argument");
	if (n % 8)
		return fail(ADSB_B200_EINVAL, "packed pushes come in groups of 8 samples (12 bytes)");
	if (((uintptr_t)packed) % 4)
		return fail(ADSB_B200_EINVAL, "packed samples must be 4-byte aligned");
	if (h->iq_mode)
		return fail(ADSB_B200_EINVAL, "packed 12-bit samples are real samples: not on a stream opened for I/Q input");
	if (h->finished)
		return fail(ADSB_B200_ESTATE, "push after finish");
	if (n == 0)
		return 0;
	CU(cudaSetDevice(h->device));
	const bool direct = is_pinned_host(packed);
	RC(ensure_batch_buffers(h, false));
	const int64_t slice = std::min<int64_t>(kPackSlice, align_down(h->batch, 8));
	if (!h->d_pk)
		CU(cudaMalloc(&h->d_pk, (size_t)slice / 2 * 3));
	for (int i = 0; i < 2 && !direct; i++) {
		if (!h->h_pk[i])
			CU(cudaMallocHost(&h->h_pk[i], (size_t)slice / 2 * 3));
		if (!h->ev_pk[i])
			CU(cudaEventCreateWithFlags(&h->ev_pk[i], cudaEventDisableTiming));
	}
	const unsigned char *src = static_cast<const unsigned char *>(packed);
	size_t done = 0;
	while (done < n) {
		const int64_t room = align_down(batch_room(h), 8);
		if (room <= 0) {
			RC(flush_batch(h));
			continue;
		}
		RC(drain_pinned(h));  // earlier 16-bit pushes of this window go first
		const size_t take = (size_t)std::min<int64_t>(std::min<int64_t>(room, slice), (int64_t)(n - done));
		const size_t bytes = take / 2 * 3;
		const void *from = src + done / 2 * 3;
		if (!direct) {
			const int i = h->pk_i;
			CU(cudaEventSynchronize(h->ev_pk[i]));  // the upload that last used this staging buffer
			memcpy(h->h_pk[i], from, bytes);
			from = h->h_pk[i];
		}
		// copy and unpack on the upload stream, in order: the scratch buffer is free again when the next
		// slice's copy starts, and the scan waits for this stream anyway
		CU(cudaMemcpyAsync(h->d_pk, from, bytes, cudaMemcpyHostToDevice, h->st_copy));
		if (!direct) {
			CU(cudaEventRecord(h->ev_pk[h->pk_i], h->st_copy));
			h->pk_i ^= 1;
		}
		CU(launch_unpack12(h->d_pk, h->dev[h->cur] + h->win_off + h->dev_fill, (long long)take, h->st_copy));
		h->launches += 1;
		h->dev_fill += (int64_t)take;
		h->pushed += (int64_t)take;
		done += take;
	}
	if (direct) {
		const double t0 = now_ms();
		CU(cudaStreamSynchronize(h->st_copy));  // the caller may reuse its buffer on return
		h->tel.wait_upload_ms += now_ms() - t0;
	}
	return 0;
}

int adsbdec_b200_push_device(adsbdec_b200_t *h, const void *dev_samples, size_t n)
{
	if (!h || (!dev_samples && n))
		return fail(ADSB_B200_EINVAL, "null argument");
	if (n % 4)
		return fail(ADSB_B200_EINVAL, "push length must be a multiple of 4 samples (air.c:59-82)");
	if (h->finished)
		return fail(ADSB_B200_ESTATE, "push after finish");
	if (n == 0)
		return 0;
	CU(cudaSetDevice(h->device));
	h->tr.mark("push_device");
	RC(ensure_batch_buffers(h, false));
	const uint16_t *src = (const uint16_t *)dev_samples;
	const int64_t seam = 2 * kLookahead + 2 * kCarryBack;  // samples of the piece the seam window needs
	const bool direct = (int64_t)n >= (int64_t)kMinFused * 4 && ((uintptr_t)src % 16) == 0 &&
	                    (h->pushed % 8) == 0 && (n % 8) == 0;
	if (!direct) {
		// small or unaligned piece: append to the batch window on the device
		size_t done = 0;
		while (done < n) {
			const int64_t room = batch_room(h);
			if (room <= 0) {
				RC(flush_batch(h));
				continue;
			}
			RC(drain_pinned(h));
			const size_t take = std::min<size_t>((size_t)room, n - done);
			CU(cudaMemcpyAsync(h->dev[h->cur] + h->win_off + h->dev_fill, src + done, take * sizeof(uint16_t),
			                   cudaMemcpyDeviceToDevice, h->st_copy));
			h->dev_fill += (int64_t)take;
			h->pushed += (int64_t)take;
			done += take;
		}
		CU(cudaStreamSynchronize(h->st_copy));  // the caller may free its buffer on return
		return 0;
	}
	// Large aligned piece: (1) a seam window [carried tail | first `seam` samples of the piece] in
	// our own buffer, (2) the body straight from the caller's memory, in batch-sized scan jobs.
	RC(drain_pinned(h));
	if (h->dev_fill > 0 && batch_room(h) < seam)
		RC(flush_batch(h));
	const int64_t piece_u0 = h->pushed, piece_u1 = h->pushed + (int64_t)n;
	if (h->dev_fill > 0) {
		CU(cudaMemcpyAsync(h->dev[h->cur] + h->win_off + h->dev_fill, src, (size_t)seam * sizeof(uint16_t),
		                   cudaMemcpyDeviceToDevice, h->st_copy));
		CU(cudaEventRecord(h->ev_h2d, h->st_copy));
		CU(cudaStreamWaitEvent(h->st, h->ev_h2d, 0));
		RC(submit_window(h, h->dev[h->cur] + h->win_off, h->win_u0, h->win_u0 + h->dev_fill + seam));
	}
	h->pushed = piece_u1;
	// Equal jobs of at most kDeviceJob samples, at least two for a piece worth overlapping: big jobs keep
	// the per-job overheads and the tails of the kernel launches small, while the host's work on the
	// results of one job hides behind the scan of the next (60 s capture, 1144 Mi samples: three jobs,
	// 1.93-1.95 ms; five jobs of 256 Mi 2.07 ms; one job 2.3 ms, on the same box).
	int64_t dev_job = kDeviceJob;
	if (const char *e = getenv("ADSB_B200_DEVICE_JOB"))  // development knob: samples per scan job
		dev_job = std::max<int64_t>(1 << 22, atoll(e));
	const int64_t total = piece_u1 - piece_u0, cap = std::max(h->batch, dev_job);
	int64_t n_jobs = (total + cap - 1) / cap;
	if (n_jobs < 2 && total >= 2 * std::max(h->batch, kDeviceJobMin))
		n_jobs = 2;
	const int64_t per = align_down((total + n_jobs - 1) / n_jobs + 7, 8);
	if (piece_u0 == 0 && !h->job[0].active && !h->job[1].active && !h->job[2].active)  // a fresh stream: nothing refers to the lists yet
		for (int s = 0; s < kSlots; s++)
			RC(presize_lists(h, s, per / 2));
	// development knob: the piece cut at these percentages instead of into equal jobs, e.g. "25,50,25"
	std::vector<int64_t> ends;
	if (const char *e = getenv("ADSB_B200_JOB_SPLIT")) {
		double acc = 0;
		for (const char *q = e; *q;) {
			char *nx = nullptr;
			acc += strtod(q, &nx);
			if (nx == q)
				break;
			ends.push_back(std::min(piece_u1, piece_u0 + align_down((int64_t)((double)total * acc / 100.0) + 7, 8)));
			q = *nx ? nx + 1 : nx;
		}
		if (ends.empty() || ends.back() != piece_u1)
			ends.push_back(piece_u1);
	} else {
		for (int64_t end = piece_u0; end < piece_u1;)
			ends.push_back(end = std::min(piece_u1, end + per));
	}
	for (const int64_t end : ends) {
		// window: from the FIR history of the next owned offset to `end`, inside the caller's memory
		const int64_t w0 = std::max<int64_t>(piece_u0, align_down(2 * h->own_next - kCarryBack, 8));
		h->next_job_is_last = end == piece_u1;  // (we wait for it below: nothing runs beside its chain)
		const int rc = end > w0 ? submit_window(h, src + (w0 - piece_u0), w0, end) : 0;
		h->next_job_is_last = false;
		RC(rc);
	}
	// The tail the next push needs is copied while the last jobs still run (behind their scan kernels on the same
	// stream, into the window buffer no job of this push reads), so that the wait below finds it done.
	h->dev_fill = 0;
	RC(carry_to_next(h, src, piece_u0, piece_u1));
	RC(complete_all(h));  // the caller's memory is no longer needed after this ...
	CU(cudaStreamSynchronize(h->st));  // ... and the copy of the tail we keep
	return 0;
}

int adsbdec_b200_repair_lookup(uint32_t syndrome)
{
	static const std::vector<unsigned long long> tab = [] {  // built once, thread-safe (C++11 static init)
		std::vector<unsigned long long> t(kFixSlots);
		fix_build_table(t.data());
		return t;
	}();
	const uint32_t e = fix_lookup(tab.data(), syndrome & 0xFFFFFFu);
	return e == kFixNone ? -1 : (int)e;
}

int adsbdec_b200_flush(adsbdec_b200_t *h)
{
	if (!h)
		return fail(ADSB_B200_EINVAL, "null handle");
	if (h->finished)
		return fail(ADSB_B200_ESTATE, "flush after finish");
	CU(cudaSetDevice(h->device));
	const size_t before = h->emitted;
	// only samples that arrived since the last scan justify a job (the carried tail alone does not)
	if (h->dev[0] && 2 * h->own_next + 2 * kLookahead < h->pushed)
		RC(flush_batch(h));
	return (int)std::min<size_t>(h->emitted - before, 0x7FFFFFFF);
}

int adsbdec_b200_poll(adsbdec_b200_t *h)
{
	if (!h)
		return fail(ADSB_B200_EINVAL, "null handle");
	CU(cudaSetDevice(h->device));
	const size_t before = h->emitted;
	for (int k = 0; k < kSlots; k++) {
		const int s = (h->next_slot + k) % kSlots;  // oldest first: frames leave in stream order
		if (!h->job[s].active)
			continue;
		const cudaError_t e = cudaEventQuery(h->ev_done[s]);
		if (e == cudaErrorNotReady)
			break;
		if (e != cudaSuccess)
			return fail(ADSB_B200_ECUDA, "poll: %s", cudaGetErrorString(e));
		RC(complete_job(h, s));
	}
	return (int)std::min<size_t>(h->emitted - before, 0x7FFFFFFF);
}

int adsbdec_b200_finish(adsbdec_b200_t *h)
{
	if (!h)
		return fail(ADSB_B200_EINVAL, "null handle");
	if (h->finished)
		return 0;
	CU(cudaSetDevice(h->device));
	if (h->dev[0] && (h->pin_fill > 0 || h->dev_fill > 0))
		RC(flush_batch(h));
	RC(complete_all(h));
	// the stream path only ever feeds stamped frames, counted on the GPU: what the call schedule never reaches
	// (air.c:243-244 drops it) is taken off the counters
	// the last deqframe calls; what they do not reach is dropped (air.c:243-244) and comes off the Ok counters,
	// which the GPU keeps for everything it accepted
	h->n_vis = h->res->finish_keys(h->fkeys, h->n_vis, h->n_total);
	for (size_t i = h->n_vis; i < h->n_total; i++) {
		const int df = h->frames[i].bytes[0] >> 3;
		if (ok_class(df) >= 0)
			h->stat_ok[ok_class(df)]--;
	}
	h->n_total = h->n_vis;
	h->counted = h->n_vis;
	RC(stats_pass(h, h->last_slot, false));  // pending tries against the final scan end
	emit_new_frames(h);
	RC(sync_compute(h));
	h->tr.mark("finish: done");
	h->tr.dump();
	h->finished = true;
	return 0;
}

int adsbdec_b200_decode_file(adsbdec_b200_t *h, const char *path)
{
	if (!h || !path)
		return fail(ADSB_B200_EINVAL, "null argument");
	if (h->finished)
		return fail(ADSB_B200_ESTATE, "decode after finish");
	int fd = open(path, O_RDONLY);
	if (fd < 0)
		return fail(ADSB_B200_EIO, "cannot open %s: %s", path, strerror(errno));
	CU(cudaSetDevice(h->device));
	struct stat sb;
	const bool regular = fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && h->batch >= 4 * (int64_t)kReadSamples;
	int rc = 0;
	if (regular) {
		// air.c:233-239 reads 2 MiB at a time into ONE reused buffer and calls decodeiq(buf, n/2), which
		// eats groups of 4 and therefore reads up to 3 stale entries behind a ragged tail: whatever
		// the previous read left at those positions (zero if there was none).  Here every read goes
		// straight into the page-locked staging buffer; the stale entries are fetched from the file.
		rc = ensure_batch_buffers(h, true);
		off_t pos = 0;
		bool had_full_read = false;
		// every read but the last returns 2 MiB whatever the order it is issued in: fetch those in
		// parallel (readpool.h); reserved[2] = reader threads, 0 = up to 16, 1 = this thread alone
		int64_t bulk_left = (int64_t)(sb.st_size / (off_t)(kReadSamples * sizeof(uint16_t))) * kReadSamples;
		if (!rc && bulk_left > 0 && !h->readers) {
			int nthr = (int)h->cfg.reserved[2];
			if (nthr <= 0)
				nthr = std::max(1, std::min(16, (int)std::thread::hardware_concurrency()));  // (16 threads: 22.7 G samples/s, 12: 19.9)
			h->readers.reset(new (std::nothrow) ReadPool(std::min(nthr, 64) - 1));
			if (!h->readers)
				rc = fail(ADSB_B200_ENOMEM, "out of memory");
		}
		while (!rc) {
			if (batch_room(h) < (int64_t)kReadSamples + 4) {
				rc = flush_batch(h);
				if (rc) break;
			}
			if (h->pin_pos == 0 && cudaEventSynchronize(h->ev_pin_free[h->cur]) != cudaSuccess) {
				rc = fail(ADSB_B200_ECUDA, "waiting for the staging buffer");
				break;
			}
			if (bulk_left > 0) {
				int64_t take = std::min(bulk_left, std::min(kBulkRead, batch_room(h) - 4));
				take -= take % kReadSamples;
				if (!h->readers->read(fd, h->pin[h->cur] + h->pin_pos, pos, (size_t)take * sizeof(uint16_t))) {
					rc = fail(ADSB_B200_EIO, "short read from %s (file changed while decoding?)", path);
					break;
				}
				had_full_read = true;
				bulk_left -= take;
				pos += (off_t)take * (off_t)sizeof(uint16_t);
				h->pin_pos += take;
				h->pin_fill += take;
				h->pushed += take;
				rc = drain_pinned(h);
				continue;
			}
			uint16_t *dst = h->pin[h->cur] + h->pin_pos;
			const ssize_t got = pread(fd, dst, kReadSamples * sizeof(uint16_t), pos);
			if (got <= 0)
				break;
			size_t ns = (size_t)got / 2;
			if (ns % 4) {
				const size_t need = 4 - ns % 4;
				uint16_t stale[4] = {0, 0, 0, 0};
				if (had_full_read)  // same buffer positions, one read earlier
					if (pread(fd, stale, (need + 1) * sizeof(uint16_t), pos - (off_t)(kReadSamples * 2) + (off_t)(ns * 2)) < 0)
						memset(stale, 0, sizeof stale);
				if (got & 1)  // the odd byte overwrote the low half of the first stale entry
					stale[0] = (uint16_t)((stale[0] & 0xFF00u) | ((const uint8_t *)dst)[got - 1]);
				for (size_t i = 0; i < need; i++)
					dst[ns + i] = stale[i];
				ns += need;
			}
			had_full_read = (size_t)got == kReadSamples * sizeof(uint16_t);
			pos += got;
			h->pin_pos += (int64_t)ns;
			h->pin_fill += (int64_t)ns;
			h->pushed += (int64_t)ns;
			if (h->pin_fill >= kStageSlice)
				rc = drain_pinned(h);
		}
	} else {
		// pipes and the like: the reference's loop as it is, through push()
		uint16_t *buf = (uint16_t *)calloc(kReadSamples + 4, sizeof(uint16_t));
		if (!buf) {
			close(fd);
			return fail(ADSB_B200_ENOMEM, "out of memory");
		}
		for (;;) {
			ssize_t got = read(fd, buf, kReadSamples * sizeof(uint16_t));
			if (got <= 0)
				break;
			size_t ns = (size_t)got / 2;
			ns = (ns + 3) & ~(size_t)3;
			rc = adsbdec_b200_push(h, buf, ns);
			if (rc)
				break;
		}
		free(buf);
	}
	close(fd);
	if (rc)
		return rc;
	return adsbdec_b200_finish(h);
}

int adsbdec_b200_stats(adsbdec_b200_t *h, uint32_t try_[3], uint32_t ok[3])
{
	if (!h)
		return fail(ADSB_B200_EINVAL, "null handle");
	CU(cudaSetDevice(h->device));
	RC(complete_all(h));
	CU(cudaMemcpyAsync(h->h_stats, h->d_stats, sizeof(StreamStats), cudaMemcpyDeviceToHost, h->st_stat));
	CU(cudaStreamSynchronize(h->st_stat));
	if (h->h_stats->pend_overflow)
		return fail(ADSB_B200_ERANGE, "pending-try list overflow");
	uint32_t pend[3] = {0, 0, 0};  // frames the GPU has accepted but the deqframe call schedule has not reached
	for (size_t k = h->n_vis; k < h->n_total; k++) {
		const int df = h->frames[k].bytes[0] >> 3;
		if (ok_class(df) >= 0)
			pend[ok_class(df)]++;
	}
	for (int i = 0; i < 3; i++) {
		if (try_) try_[i] = h->h_stats->stat_try[i];
		if (ok) ok[i] = h->stat_ok[i] - pend[i];
	}
	return 0;
}

size_t adsbdec_b200_num_frames(const adsbdec_b200_t *h) { return h ? h->n_vis : 0; }
const adsbdec_b200_frame_t *adsbdec_b200_frames(const adsbdec_b200_t *h)
{
	return h && h->n_vis ? h->frames : nullptr;
}
void adsbdec_b200_clear_frames(adsbdec_b200_t *h)
{
	if (!h || h->emitted != h->n_vis)
		return;
	// every frame shown so far goes; the ones still in front of the call schedule move up
	const size_t keep = h->n_total - h->n_vis;
	if (keep) {
		memmove(h->frames, h->frames + h->n_vis, keep * sizeof(adsbdec_b200_frame_t));
		memmove(h->fkeys, h->fkeys + h->n_vis, keep * sizeof(unsigned long long));
	}
	h->n_total = keep;
	h->n_vis = 0;
	h->emitted = 0;
	h->counted = 0;
}

int adsbdec_b200_kernel_time(adsbdec_b200_t *h, double *ms, uint64_t *launches, int reset)
{
	if (!h)
		return fail(ADSB_B200_EINVAL, "null handle");
	if (ms) *ms = h->kernel_ms;
	if (launches) *launches = h->launches;
	if (reset) {
		h->kernel_ms = 0;
		h->launches = 0;
	}
	return 0;
}

int adsbdec_b200_telemetry(adsbdec_b200_t *h, adsbdec_b200_telemetry_t *out, int reset)
{
	if (!h || !out)
		return fail(ADSB_B200_EINVAL, "null argument");
	*out = h->tel;
	if (reset)
		h->tel = adsbdec_b200_telemetry_t{};
	return 0;
}

int adsbdec_b200_fused_time(adsbdec_b200_t *h, double *ms, uint64_t *launches, uint64_t *samples, int reset)
{
	if (!h)
		return fail(ADSB_B200_EINVAL, "null handle");
	if (ms) *ms = h->fused_ms;
	if (launches) *launches = h->fused_launches;
	if (samples) *samples = h->fused_samples;
	if (reset) {
		h->fused_ms = 0;
		h->fused_launches = 0;
		h->fused_samples = 0;
	}
	return 0;
}

int adsbdec_b200_gate_time(adsbdec_b200_t *h, double *ms, int reset)
{
	if (!h)
		return fail(ADSB_B200_EINVAL, "null handle");
	if (ms) *ms = h->gate_ms;
	if (reset)
		h->gate_ms = 0;
	return 0;
}

int adsbdec_b200_envelope(int device, const void *samples, size_t n_samples, int is_device,
                          uint64_t start_sample, float *amp_out, size_t count)
{
	if (!samples || !amp_out)
		return fail(ADSB_B200_EINVAL, "null argument");
	if (count > n_samples / 2)
		return fail(ADSB_B200_EINVAL, "count exceeds n_samples/2");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
		return fail(ADSB_B200_ENODEV, "no usable CUDA device %d; this library has no CPU fallback", device);
	CU(cudaSetDevice(device));
	uint16_t *d_raw = nullptr;
	float *d_amp = nullptr;
	const uint16_t *raw = (const uint16_t *)samples;
	if (!is_device) {
		CU(cudaMalloc(&d_raw, std::max<size_t>(n_samples, 1) * sizeof(uint16_t)));
		CU(cudaMemcpy(d_raw, samples, n_samples * sizeof(uint16_t), cudaMemcpyHostToDevice));
		raw = d_raw;
	}
	cudaError_t e = cudaMalloc(&d_amp, std::max<size_t>(count, 1) * sizeof(float));
	if (e == cudaSuccess)
		e = launch_envelope_generic(raw, 0, start_sample, 0, (long long)count, d_amp, 0, 0);
	if (e == cudaSuccess)
		e = cudaMemcpy(amp_out, d_amp, count * sizeof(float), cudaMemcpyDeviceToHost);
	if (d_raw) cudaFree(d_raw);
	if (d_amp) cudaFree(d_amp);
	if (e != cudaSuccess)
		return fail(ADSB_B200_ECUDA, "envelope: %s", cudaGetErrorString(e));
	return 0;
}

int64_t adsbdec_b200_chunk_candidates(adsbdec_b200_t *h, const void *dev_samples, size_t n_samples,
                                      uint64_t start_sample, uint64_t own_begin, uint64_t own_end,
                                      adsbdec_b200_candidate_t *out, size_t cap, uint32_t tries_out[3])
{
	if (!h || !dev_samples || (!out && cap))
		return fail(ADSB_B200_EINVAL, "null argument");
	if (h->ap_mode)  // a chunk does not know which addresses earlier chunks announced
		return fail(ADSB_B200_EINVAL, "address/parity frames (extension) are a stream-mode feature");
	if (own_end < own_begin || (own_end + kLookahead) * 2 > n_samples)
		return fail(ADSB_B200_EINVAL, "owned range needs 1200 envelope samples of look-ahead in the chunk");
	if (((uintptr_t)dev_samples % 16) != 0 || start_sample % 4)
		return fail(ADSB_B200_EINVAL, "chunk must be 16-byte aligned and start on a multiple of 4 samples");
	CU(cudaSetDevice(h->device));
	RC(complete_all(h));
	const int s = h->next_slot;
	Job &j = h->job[s];
	j.buf = (const uint16_t *)dev_samples;
	j.u0 = 0;
	j.u1 = (int64_t)n_samples;
	j.total_out = (int64_t)n_samples / 2;
	j.n0 = (int64_t)own_begin;
	j.n1 = (int64_t)own_end;
	j.origin = start_sample;
	j.resolve = false;
	j.gpu_idle_after = false;
	j.active = true;
	RC(presize_lists(h, s, j.n1 - j.n0));
	RC(enqueue_job(h, s, false));
	RC(complete_job(h, s));
	if (h->cand_tmp.size() > cap)
		return fail(ADSB_B200_ERANGE, "output capacity %zu < %zu candidates", cap, h->cand_tmp.size());
	if (!h->cand_tmp.empty())
		memcpy(out, h->cand_tmp.data(), h->cand_tmp.size() * sizeof(adsbdec_b200_candidate_t));
	if (tries_out) {
		// raw gate-pass counts in the owned range (no sequential exclusion in chunk mode)
		tries_out[0] = tries_out[1] = tries_out[2] = 0;
		const unsigned long long nt = own_end > own_begin && h->cfg.collect_stats ? h->h_ctr[s]->n_try : 0;
		if (nt) {
			std::vector<unsigned long long> t(nt);
			CU(cudaMemcpy(t.data(), h->d_tries[s], nt * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
			for (unsigned long long e : t)
				tries_out[e >> kTryClsShift]++;
		}
	}
	return (int64_t)h->cand_tmp.size();
}

int adsbdec_b200_debug_fused(adsbdec_b200_t *h, const uint16_t *samples, size_t n_samples,
                             uint64_t start_sample, float *amp_out, uint32_t *p_out, uint32_t *q_out,
                             uint32_t *b_out, size_t count, int64_t *j_begin_out)
{
	if (!h || !samples || !amp_out || !p_out || !q_out || !b_out || count % 32)
		return fail(ADSB_B200_EINVAL, "bad argument");
	if (h->fused_ctas <= 0)
		return fail(ADSB_B200_ENODEV, "fused kernel not configured");
	CU(cudaSetDevice(h->device));
	RC(complete_all(h));
	const int64_t n1 = (int64_t)(n_samples / 2) - kLookahead;
	if (n1 <= 0)
		return fail(ADSB_B200_EINVAL, "capture too short");
	uint16_t *d_raw = nullptr;
	CU(cudaMalloc(&d_raw, (n_samples + 64) * sizeof(uint16_t)));
	CU(cudaMemcpy(d_raw, samples, n_samples * sizeof(uint16_t), cudaMemcpyHostToDevice));
	CU(cudaMalloc(&h->dbg_amp, count * sizeof(float)));
	CU(cudaMemset(h->dbg_amp, 0, count * sizeof(float)));
	h->dbg_n = (long long)count;
	const int saved_ctas = h->fused_ctas;
	h->fused_ctas = 1;  // one worker, so span 0 covers the head of the capture in order
	const int s = h->next_slot;
	Job j;
	j.buf = d_raw;
	j.u0 = 0;
	j.u1 = (int64_t)n_samples;
	j.origin = start_sample;
	k_reset_counters<<<1, 1, 0, h->st>>>(h->d_ctr[s]);
	int rc = launch_fused_range(h, s, j, 0, n1);
	h->fused_ctas = saved_ctas;
	cudaError_t e = cudaStreamSynchronize(h->st);
	if (rc == 0 && e == cudaSuccess) {
		cudaMemcpy(amp_out, h->dbg_amp, count * sizeof(float), cudaMemcpyDeviceToHost);
		// the bit arrays of the range, word 0 at the range's first span (j_begin_out)
		const long long cap = h->bits_cap[h->dbg_slot];
		const size_t nw = (size_t)std::min<long long>((long long)count / 32, cap);
		memset(p_out, 0, count / 8);
		memset(q_out, 0, count / 8);
		memset(b_out, 0, count / 8);
		cudaMemcpy(p_out, h->d_bits[h->dbg_slot], nw * 4, cudaMemcpyDeviceToHost);
		cudaMemcpy(q_out, h->d_bits[h->dbg_slot] + cap, nw * 4, cudaMemcpyDeviceToHost);
		cudaMemcpy(b_out, h->d_bits[h->dbg_slot] + 2 * cap, nw * 4, cudaMemcpyDeviceToHost);
		if (j_begin_out) *j_begin_out = h->dbg_jb;
	}
	cudaFree(h->dbg_amp);
	cudaFree(d_raw);
	h->dbg_amp = nullptr;
	h->dbg_n = 0;
	if (rc) return rc;
	if (e != cudaSuccess)
		return fail(ADSB_B200_ECUDA, "debug_fused: %s", cudaGetErrorString(e));
	return 0;
}

/* ---- host-only helpers ---------------------------------------------------- */

struct adsbdec_b200_resolver {
	Resolver r;
	std::vector<adsbdec_b200_frame_t> out;
	explicit adsbdec_b200_resolver(bool e) : r(e) {}
};

adsbdec_b200_resolver_t *adsbdec_b200_resolver_new(int emulate_reference_tail)
{
	return new (std::nothrow) adsbdec_b200_resolver(emulate_reference_tail != 0);
}

void adsbdec_b200_resolver_free(adsbdec_b200_resolver_t *r) { delete r; }

static void resolver_emit(adsbdec_b200_resolver_t *r, adsbdec_b200_sink_fn fn, void *user)
{
	if (!fn)
		return;  // no sink: the frames stay in the resolver (adsbdec_b200_resolver_frames)
	for (const adsbdec_b200_frame_t &f : r->out)
		fn(f.bytes, f.len, f.ts, f.pw, user);
	r->out.clear();
}

int adsbdec_b200_resolver_feed(adsbdec_b200_resolver_t *r, const adsbdec_b200_candidate_t *c, size_t count,
                               uint64_t scanned_upto, uint64_t total_out, adsbdec_b200_sink_fn fn,
                               void *user)
{
	if (!r || (!c && count))
		return fail(ADSB_B200_EINVAL, "null argument");
	r->r.feed(c, count, scanned_upto, total_out, r->out);
	resolver_emit(r, fn, user);
	return 0;
}

int adsbdec_b200_resolver_feed_frames(adsbdec_b200_resolver_t *r, const adsbdec_b200_frame_t *f, size_t count,
                                      uint64_t scanned_upto, uint64_t total_out, uint32_t ok[3],
                                      adsbdec_b200_sink_fn fn, void *user)
{
	if (!r || (!f && count))
		return fail(ADSB_B200_EINVAL, "null argument");
	r->r.feed_frames(f, count, scanned_upto, total_out, r->out, ok);
	resolver_emit(r, fn, user);
	return 0;
}

int64_t adsbdec_b200_resolver_advance_keys(adsbdec_b200_resolver_t *r, const uint64_t *keys, size_t n_vis,
                                           size_t n_total, uint64_t scanned_upto, uint64_t total_out, int finish)
{
	if (!r || (!keys && n_total) || n_vis > n_total)
		return fail(ADSB_B200_EINVAL, "bad argument");
	static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "key type");
	const unsigned long long *k = reinterpret_cast<const unsigned long long *>(keys);
	if (finish)
		return (int64_t)r->r.finish_keys(k, n_vis, n_total);
	return (int64_t)r->r.advance_keys(k, n_vis, n_total, scanned_upto, total_out);
}

int adsbdec_b200_resolver_finish(adsbdec_b200_resolver_t *r, adsbdec_b200_sink_fn fn, void *user)
{
	if (!r)
		return fail(ADSB_B200_EINVAL, "null argument");
	r->r.finish(r->out);
	resolver_emit(r, fn, user);
	return 0;
}

size_t adsbdec_b200_resolver_num_frames(const adsbdec_b200_resolver_t *r) { return r ? r->out.size() : 0; }
const adsbdec_b200_frame_t *adsbdec_b200_resolver_frames(const adsbdec_b200_resolver_t *r)
{
	return r && !r->out.empty() ? r->out.data() : nullptr;
}

uint64_t adsbdec_b200_resolver_scan_end(const adsbdec_b200_resolver_t *r) { return r ? r->r.scan_end() : 0; }

const char *adsbdec_b200_last_error(void) { return t_err; }
const char *adsbdec_b200_version(void) { return "adsbdec_b200 0.2 (sm_100a)"; }

}  // extern "C"
