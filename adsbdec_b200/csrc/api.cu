// api.cu - the C ABI (include/adsbdec_b200.h): per-stream decoder handle, host staging,
// kernel orchestration and the hand-off to the sequential resolver.
//
// Reference seam (TLeconte/adsbdec): fileInput -> decodeiq -> deqframe -> validShort/Long -> netout
// (air.c:219-246, :54-101; demod.c:84-144; valid.c:39-82; output.c:159).  The host keeps the
// stream position, carries the un-scanned tail between batches the way air.c:96-98 carries
// ampbuff, and calls the sink in timestamp order on the caller's thread.
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <vector>

#include "../../include/adsbdec_b200.h"
#include "common.h"
#include "kernels.h"
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

constexpr int64_t kCarryBack = 128;        // raw samples kept in front of the next owned offset
constexpr int64_t kCarryMax = 1 << 14;     // room in front of a batch for the carried tail
constexpr int64_t kDefaultBatch = 32ll << 20;
constexpr int64_t kAmpChunk = 8ll << 20;   // generic path: offsets per envelope scratch pass
constexpr int64_t kMinFused = 1 << 16;     // windows smaller than this use the generic kernels

__global__ void k_set_counters(Counters *c, unsigned long long n_cand, unsigned long long n_try,
                               unsigned long long n_try_next, int clear_flags)
{
	c->n_cand = n_cand;
	c->n_try = n_try;
	c->n_try_next = n_try_next;
	if (clear_flags) {
		c->overflow = 0;
		c->wide = 0;
	}
}

inline int64_t align_down(int64_t v, int64_t a) { return v - (((v % a) + a) % a); }

}  // namespace

struct adsbdec_b200 {
	adsbdec_b200_config_t cfg;
	int device = 0;
	cudaStream_t st = nullptr;       // kernels + result copies
	cudaStream_t st_copy = nullptr;  // H2D of sample batches
	cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_copy[2] = {nullptr, nullptr};
	cudaEvent_t ev_fa = nullptr, ev_fb = nullptr;

	// stream position (local sample / offset indices, 0 = first sample ever pushed)
	unsigned long long origin = 0;  // absolute index of local sample 0 (FIR phase / fidx wrap)
	int64_t pushed = 0;      // samples received
	int64_t own_next = 0;    // first offset the GPU has not scanned yet
	bool finished = false;

	// current batch under construction: dev[cur] holds [carry | batch samples]
	int64_t batch = 0;       // capacity in samples
	uint16_t *pin[2] = {nullptr, nullptr};
	uint16_t *dev[2] = {nullptr, nullptr};
	int cur = 0;
	int64_t win_u0 = 0;      // local index of the first sample held in dev[cur] (carry start)
	int64_t win_off = 0;     // element offset in dev[cur] where local sample win_u0 sits
	int64_t dev_fill = 0;    // samples of the window already on the device (incl. carry)
	int64_t pin_fill = 0;    // samples staged in pin[cur], not yet copied

	// outputs
	CandRec *d_cands = nullptr, *h_cands = nullptr;
	unsigned long long cand_cap = 0;
	unsigned long long *d_tries[2] = {nullptr, nullptr};
	unsigned long long try_cap = 0;
	int try_cur = 0;
	unsigned long long try_carried = 0;
	Counters *d_ctr = nullptr, *h_ctr = nullptr;
	float *d_amp = nullptr;
	unsigned long long *d_spans = nullptr, *h_spans = nullptr;
	size_t span_cap = 0;
	FusedSpan *d_fspans = nullptr, *h_fspans = nullptr;
	int *d_ticket = nullptr;
	int fspan_cap = 0;
	int fused_ctas = 0;      // grid size of the fused kernel (0 = unavailable)

	Resolver *res = nullptr;
	std::vector<adsbdec_b200_frame_t> frames;
	size_t emitted = 0;
	uint32_t stat_ok[3] = {0, 0, 0};
	uint32_t raw_tries[3] = {0, 0, 0};
	adsbdec_b200_sink_fn sink = nullptr;
	void *sink_user = nullptr;

	double kernel_ms = 0, fused_ms = 0;
	uint64_t launches = 0, fused_launches = 0, fused_samples = 0;

	std::vector<adsbdec_b200_candidate_t> cand_tmp;

	// parity taps of the fused kernel (adsbdec_b200_debug_fused)
	float *dbg_amp = nullptr;
	unsigned int *dbg_p = nullptr, *dbg_q = nullptr, *dbg_b = nullptr;
	long long dbg_n = 0, dbg_jb = 0;
};

namespace {

int alloc_outputs(adsbdec_b200 *h, unsigned long long cand_cap, unsigned long long try_cap)
{
	if (h->d_cands) cudaFree(h->d_cands);
	if (h->h_cands) cudaFreeHost(h->h_cands);
	h->d_cands = nullptr;
	h->h_cands = nullptr;
	CU(cudaMalloc(&h->d_cands, cand_cap * sizeof(CandRec)));
	CU(cudaMallocHost(&h->h_cands, cand_cap * sizeof(CandRec)));
	h->cand_cap = cand_cap;
	if (try_cap != h->try_cap) {
		for (int i = 0; i < 2; i++) {
			unsigned long long *nt = nullptr;
			CU(cudaMalloc(&nt, try_cap * sizeof(unsigned long long)));
			if (h->d_tries[i]) {
				if (i == h->try_cur && h->try_carried)
					CU(cudaMemcpy(nt, h->d_tries[i], h->try_carried * sizeof(unsigned long long),
					              cudaMemcpyDeviceToDevice));
				cudaFree(h->d_tries[i]);
			}
			h->d_tries[i] = nt;
		}
		h->try_cap = try_cap;
	}
	return 0;
}

int ensure_batch_buffers(adsbdec_b200 *h)
{
	if (h->dev[0])
		return 0;
	for (int i = 0; i < 2; i++) {
		CU(cudaMalloc(&h->dev[i], (size_t)(kCarryMax + h->batch + 64) * sizeof(uint16_t)));
		CU(cudaMallocHost(&h->pin[i], (size_t)h->batch * sizeof(uint16_t)));
	}
	return 0;
}

// Launches the scan of offsets [n0, n1) over a device window (buf[0] = local sample u0, samples up
// to u1 present) with the generic kernels.  Counters must have been set by the caller.
int launch_generic(adsbdec_b200 *h, const uint16_t *buf, int64_t u0, int64_t n0, int64_t n1)
{
	if (!h->d_amp)
		CU(cudaMalloc(&h->d_amp, (size_t)(kAmpChunk + kLookahead) * sizeof(float)));
	for (int64_t c0 = n0; c0 < n1; c0 += kAmpChunk) {
		const int64_t cnt = std::min(kAmpChunk, n1 - c0);
		CU(launch_envelope_generic(buf, u0, h->origin, c0, cnt + kLookahead, h->d_amp, h->st));
		CU(launch_scan_generic(h->d_amp, c0, c0, cnt, h->cfg.accept_df18, h->cfg.collect_stats,
		                       h->d_cands, h->cand_cap, h->d_tries[h->try_cur], h->try_cap, h->d_ctr,
		                       h->st));
		h->launches += 2;
	}
	return 0;
}

// Does a 2^32 boundary of the reference's `fidx` (air.c:34) fall where it would disturb the FIR
// windows of outputs [j0, j1)?  Those outputs go through the generic kernel.
bool wrap_in_range(const adsbdec_b200 *h, int64_t j0, int64_t j1, int64_t *wrap_out)
{
	const unsigned long long org = h->origin;
	const unsigned long long a = (unsigned long long)std::max<int64_t>(0, 2 * j0 - 32) + org;
	const unsigned long long b = (unsigned long long)(2 * j1 + 32) + org;
	if ((a >> 32) == (b >> 32))
		return false;
	const unsigned long long w = ((a >> 32) + 1) << 32;  // first absolute sample of the next epoch
	*wrap_out = (int64_t)((w - org) / 2);                // envelope sample whose pair starts there
	return true;
}

// Cuts [n0, n1) into spans for the fused kernel and launches it.  Requires no fidx wrap in range.
int launch_fused_range(adsbdec_b200 *h, const uint16_t *buf, int64_t u0, int64_t u1, int64_t n0, int64_t n1)
{
	const int S = fused_block_outputs();
	const int lookback = fused_lookback();
	// FIR phase.  Within one 2^32 epoch of the reference's fidx, slot(u) = (u + shift) % 14 with
	// shift = (origin - epoch_base) mod 14.  A span must start on an output j with
	// (2j + shift) % 14 == 0 (summation order becomes a compile-time constant per thread) and
	// j % 4 == 0 (16-byte aligned bulk copies): j = r (mod 28).
	const unsigned long long org = h->origin;
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
	const int64_t total = n1 - n0;
	const int want = h->fused_ctas * 4;
	int64_t per = (total + want - 1) / want;
	per = std::max<int64_t>(per, 8 * (int64_t)S);
	per = std::min<int64_t>(per, 1ll << 29);
	per = (per + S - 1) / S * S;
	const int n_sp = (int)((total + per - 1) / per);
	if (n_sp > h->fspan_cap) {
		if (h->d_fspans) cudaFree(h->d_fspans);
		if (h->h_fspans) cudaFreeHost(h->h_fspans);
		h->fspan_cap = n_sp + 64;
		CU(cudaMalloc(&h->d_fspans, h->fspan_cap * sizeof(FusedSpan)));
		CU(cudaMallocHost(&h->h_fspans, h->fspan_cap * sizeof(FusedSpan)));
	}
	if (!h->d_ticket)
		CU(cudaMalloc(&h->d_ticket, sizeof(int)));
	if (u0 % 8 || ((uintptr_t)buf % 16))
		return fail(ADSB_B200_EINVAL, "internal: fused window not 16-byte aligned");
	const int64_t w_lo = u0 / 2, w_hi = align_down(u1 / 2, 4);
	for (int i = 0; i < n_sp; i++) {
		FusedSpan &sp = h->h_fspans[i];
		sp.own_begin = n0 + i * per;
		sp.own_end = std::min<int64_t>(n1, sp.own_begin + per);
		sp.j_begin = align_down(sp.own_begin - r, 28) + r;  // at or before own_begin, == r (mod 28)
		if (sp.j_begin - 8 < w_lo && u0 != 0)
			return fail(ADSB_B200_EINVAL, "internal: window lacks FIR history");
		sp.raw_w0 = (unsigned long long)(uintptr_t)buf - 2ull * (unsigned long long)u0;
		sp.w_lo = w_lo;
		sp.w_hi = w_hi;
		sp.n_blocks = (int)((sp.own_end + lookback - sp.j_begin + S - 1) / S);
		sp.pad_ = 0;
	}
	CU(cudaMemcpyAsync(h->d_fspans, h->h_fspans, n_sp * sizeof(FusedSpan), cudaMemcpyHostToDevice, h->st));
	CU(cudaMemsetAsync(h->d_ticket, 0, sizeof(int), h->st));
	FusedParams p;
	p.spans = h->d_fspans;
	p.n_spans = n_sp;
	p.accept_df18 = h->cfg.accept_df18;
	p.want_tries = h->cfg.collect_stats;
	p.cands = h->d_cands;
	p.cand_cap = h->cand_cap;
	p.tries = h->d_tries[h->try_cur];
	p.try_cap = h->try_cap;
	p.ctr = h->d_ctr;
	p.span_ticket = h->d_ticket;
	p.dbg_amp = h->dbg_amp;
	p.dbg_p = h->dbg_p;
	p.dbg_q = h->dbg_q;
	p.dbg_b = h->dbg_b;
	p.dbg_n = h->dbg_n;
	if (h->dbg_amp)
		h->dbg_jb = h->h_fspans[0].j_begin;
	CU(cudaEventRecord(h->ev_fa, h->st));
	CU(launch_fused(p, std::min(h->fused_ctas, n_sp), h->st));
	CU(cudaEventRecord(h->ev_fb, h->st));
	h->launches += 1;
	h->fused_launches += 1;
	h->fused_samples += (uint64_t)(2 * total);
	return 0;
}

// Scans offsets [n0, n1) of a device window into the candidate / try lists and brings the
// candidates back, sorted by n, in h->cand_tmp.  Retries with larger lists on overflow and with
// the generic kernels when the fused kernel reports out-of-range samples.
int scan_range(adsbdec_b200 *h, const uint16_t *buf, int64_t u0, int64_t u1, int64_t n0, int64_t n1)
{
	h->cand_tmp.clear();
	if (n1 <= n0)
		return 0;
	bool force_generic = h->cfg.kernel == 1 || h->fused_ctas <= 0;
	for (int attempt = 0; attempt < 8; attempt++) {
		k_set_counters<<<1, 1, 0, h->st>>>(h->d_ctr, 0, h->try_carried, 0, 1);
		CU(cudaGetLastError());
		CU(cudaEventRecord(h->ev_a, h->st));
		bool used_fused = false;
		// pieces: [n0, n1) split around fidx wraps; each piece fused if large enough
		int64_t a = n0;
		while (a < n1) {
			int64_t b = n1, w = 0;
			bool has_wrap = wrap_in_range(h, a, n1 + kLookahead, &w);
			if (has_wrap) {
				// generic over the neighbourhood of the wrap, fused (if any) before it
				int64_t g0 = std::max(a, w - kLookahead - 64), g1 = std::min(n1, w + 64);
				if (g0 > a) {
					b = g0;
				} else {
					int rc = launch_generic(h, buf, u0, a, g1);
					if (rc) return rc;
					a = g1;
					continue;
				}
			}
			if (!force_generic && b - a >= kMinFused) {
				int rc = launch_fused_range(h, buf, u0, u1, a, b);
				if (rc) return rc;
				used_fused = true;
			} else {
				int rc = launch_generic(h, buf, u0, a, b);
				if (rc) return rc;
			}
			a = b;
		}
		CU(cudaEventRecord(h->ev_b, h->st));
		CU(cudaMemcpyAsync(h->h_ctr, h->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, h->st));
		CU(cudaStreamSynchronize(h->st));
		float ms = 0;
		cudaEventElapsedTime(&ms, h->ev_a, h->ev_b);
		h->kernel_ms += ms;
		if (used_fused) {
			cudaEventElapsedTime(&ms, h->ev_fa, h->ev_fb);
			h->fused_ms += ms;
		}
		if (used_fused && h->h_ctr->wide) {
			force_generic = true;  // samples beyond 12 bits: exact int path only
			continue;
		}
		if (h->h_ctr->overflow) {
			unsigned long long cc = h->cand_cap, tc = h->try_cap;
			if (h->h_ctr->overflow & 1) cc *= 4;
			if (h->h_ctr->overflow & 2) tc *= 4;
			int rc = alloc_outputs(h, cc, tc);
			if (rc) return rc;
			continue;
		}
		const size_t nc = (size_t)h->h_ctr->n_cand;
		if (nc) {
			CU(cudaMemcpyAsync(h->h_cands, h->d_cands, nc * sizeof(CandRec), cudaMemcpyDeviceToHost, h->st));
			CU(cudaStreamSynchronize(h->st));
			std::sort(h->h_cands, h->h_cands + nc,
			          [](const CandRec &x, const CandRec &y) { return x.n < y.n; });
			h->cand_tmp.resize(nc);
			for (size_t i = 0; i < nc; i++) {
				adsbdec_b200_candidate_t &c = h->cand_tmp[i];
				c.n = h->h_cands[i].n;
				c.pw = h->h_cands[i].pw;
				c.len = h->h_cands[i].len;
				memcpy(c.bytes, h->h_cands[i].bytes, 16);
			}
		}
		return 0;
	}
	return fail(ADSB_B200_ERANGE, "candidate lists keep overflowing");
}

void emit_new_frames(adsbdec_b200 *h)
{
	for (; h->emitted < h->frames.size(); h->emitted++) {
		const adsbdec_b200_frame_t &f = h->frames[h->emitted];
		const int df = f.bytes[0] >> 3;
		h->stat_ok[df == 11 ? 0 : df == 17 ? 1 : 2]++;  // valid.c:53, :75
		if (h->sink)
			h->sink(f.bytes, f.len, f.ts, f.pw, h->sink_user);
	}
}

// Stats pass over the try list: counts visited tries below the resolver's scan end, carries the rest.
int stats_pass(adsbdec_b200 *h, size_t first_new_frame)
{
	if (!h->cfg.collect_stats)
		return 0;
	const unsigned long long n_try = h->h_ctr->n_try;
	size_t f0 = first_new_frame > 0 ? first_new_frame - 1 : 0;
	size_t ns = h->frames.size() - f0;
	if (ns > h->span_cap) {
		if (h->d_spans) cudaFree(h->d_spans);
		if (h->h_spans) cudaFreeHost(h->h_spans);
		h->span_cap = ns * 2 + 1024;
		CU(cudaMalloc(&h->d_spans, h->span_cap * 2 * sizeof(unsigned long long)));
		CU(cudaMallocHost(&h->h_spans, h->span_cap * 2 * sizeof(unsigned long long)));
	}
	for (size_t i = 0; i < ns; i++) {
		const adsbdec_b200_frame_t &f = h->frames[f0 + i];
		h->h_spans[2 * i] = f.n;
		h->h_spans[2 * i + 1] = f.n + (f.len == 7 ? kSkipShort : kSkipLong);
	}
	if (ns)
		CU(cudaMemcpyAsync(h->d_spans, h->h_spans, ns * 2 * sizeof(unsigned long long),
		                   cudaMemcpyHostToDevice, h->st));
	k_set_counters<<<1, 1, 0, h->st>>>(h->d_ctr, 0, n_try, 0, 0);
	CU(launch_count_tries(h->d_tries[h->try_cur], n_try, h->d_spans, (int)ns, h->res->scan_end(),
	                      h->d_tries[1 - h->try_cur], h->try_cap, h->d_ctr, h->st));
	h->launches += 1;
	CU(cudaMemcpyAsync(h->h_ctr, h->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, h->st));
	CU(cudaStreamSynchronize(h->st));
	if (h->h_ctr->overflow)
		return fail(ADSB_B200_ERANGE, "try list overflow in stats pass");
	h->try_carried = h->h_ctr->n_try_next;
	h->try_cur = 1 - h->try_cur;
	return 0;
}

// Processes a device window holding local samples [u0, u1): scans every offset that now has its
// full look-ahead, resolves, emits.
int process_window(adsbdec_b200 *h, const uint16_t *buf, int64_t u0, int64_t u1)
{
	const int64_t total_out = u1 / 2;
	const int64_t n1 = total_out - kLookahead;
	if (n1 > h->own_next) {
		int rc = scan_range(h, buf, u0, u1, h->own_next, n1);
		if (rc) return rc;
		h->own_next = n1;
	} else {
		h->cand_tmp.clear();
		h->h_ctr->n_try = h->try_carried;
	}
	const size_t f0 = h->frames.size();
	h->res->feed(h->cand_tmp.data(), h->cand_tmp.size(), (uint64_t)h->own_next, (uint64_t)total_out,
	             h->frames);
	int rc = stats_pass(h, f0);
	if (rc) return rc;
	emit_new_frames(h);
	return 0;
}

// Copies whatever is staged in pinned memory to the device window.
int drain_pinned(adsbdec_b200 *h)
{
	if (h->pin_fill == 0)
		return 0;
	CU(cudaMemcpyAsync(h->dev[h->cur] + h->win_off + h->dev_fill, h->pin[h->cur],
	                   (size_t)h->pin_fill * sizeof(uint16_t), cudaMemcpyHostToDevice, h->st));
	CU(cudaStreamSynchronize(h->st));
	h->dev_fill += h->pin_fill;
	h->pin_fill = 0;
	return 0;
}

// Ends the current batch: process the window, then start the next one with the carried tail.
int flush_batch(adsbdec_b200 *h)
{
	int rc = drain_pinned(h);
	if (rc) return rc;
	const int64_t u0 = h->win_u0, u1 = h->win_u0 + h->dev_fill;
	const uint16_t *buf = h->dev[h->cur] + h->win_off;
	rc = process_window(h, buf, u0, u1);
	if (rc) return rc;
	// carry: everything from kCarryBack samples before the next owned offset (air.c:96-98 keeps
	// the unscanned envelope; we keep the raw samples it is computed from)
	int64_t c0 = std::max<int64_t>(u0, align_down(2 * h->own_next - kCarryBack, 8));
	if (c0 < 0) c0 = 0;
	c0 = std::max<int64_t>(c0, u0 - (u0 % 8));
	const int64_t clen = u1 - c0;
	if (clen > kCarryMax)
		return fail(ADSB_B200_EINVAL, "internal: carry too large (%lld)", (long long)clen);
	const int nxt = 1 - h->cur;
	// place the carry so that local sample c0 (a multiple of 8) lands on a 16-byte boundary
	const int64_t off = kCarryMax - align_down(clen + 7, 8);
	if (clen > 0)
		CU(cudaMemcpyAsync(h->dev[nxt] + off, buf + (c0 - u0), (size_t)clen * sizeof(uint16_t),
		                   cudaMemcpyDeviceToDevice, h->st));
	CU(cudaStreamSynchronize(h->st));
	h->cur = nxt;
	h->win_u0 = c0;
	h->win_off = off;
	h->dev_fill = clen;
	h->pin_fill = 0;
	return 0;
}

int64_t batch_room(const adsbdec_b200 *h)
{
	// samples that still fit in the current batch (the carry lives in the kCarryMax head room)
	const int64_t used = h->win_off + h->dev_fill + h->pin_fill - kCarryMax;
	return h->batch - std::max<int64_t>(0, used);
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
		if (cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) != cudaSuccess ||
		    cudaStreamCreateWithFlags(&h->st_copy, cudaStreamNonBlocking) != cudaSuccess ||
		    cudaEventCreate(&h->ev_a) != cudaSuccess || cudaEventCreate(&h->ev_b) != cudaSuccess ||
		    cudaEventCreate(&h->ev_fa) != cudaSuccess || cudaEventCreate(&h->ev_fb) != cudaSuccess ||
		    cudaMalloc(&h->d_ctr, sizeof(Counters)) != cudaSuccess ||
		    cudaMallocHost(&h->h_ctr, sizeof(Counters)) != cudaSuccess) {
			rc = fail(ADSB_B200_ECUDA, "stream/event/counter setup: %s", cudaGetErrorString(cudaGetLastError()));
			break;
		}
		memset(h->h_ctr, 0, sizeof(Counters));
		if (cudaMemset(h->d_ctr, 0, sizeof(Counters)) != cudaSuccess) {
			rc = fail(ADSB_B200_ECUDA, "memset");
			break;
		}
		rc = alloc_outputs(h, 1ull << 18, 1ull << 20);
		if (rc) break;
		if (cfg->kernel != 1) {
			int per_sm = fused_prepare(cfg->device);
			if (per_sm <= 0) {
				rc = fail(ADSB_B200_ENODEV, "fused kernel cannot be configured on device %d", cfg->device);
				break;
			}
			h->fused_ctas = per_sm * prop.multiProcessorCount;
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
	if (h->st) cudaStreamSynchronize(h->st);
	for (int i = 0; i < 2; i++) {
		if (h->dev[i]) cudaFree(h->dev[i]);
		if (h->pin[i]) cudaFreeHost(h->pin[i]);
		if (h->d_tries[i]) cudaFree(h->d_tries[i]);
	}
	if (h->d_cands) cudaFree(h->d_cands);
	if (h->h_cands) cudaFreeHost(h->h_cands);
	if (h->d_ctr) cudaFree(h->d_ctr);
	if (h->h_ctr) cudaFreeHost(h->h_ctr);
	if (h->d_amp) cudaFree(h->d_amp);
	if (h->d_spans) cudaFree(h->d_spans);
	if (h->h_spans) cudaFreeHost(h->h_spans);
	if (h->d_fspans) cudaFree(h->d_fspans);
	if (h->h_fspans) cudaFreeHost(h->h_fspans);
	if (h->d_ticket) cudaFree(h->d_ticket);
	if (h->ev_a) cudaEventDestroy(h->ev_a);
	if (h->ev_b) cudaEventDestroy(h->ev_b);
	if (h->ev_fa) cudaEventDestroy(h->ev_fa);
	if (h->ev_fb) cudaEventDestroy(h->ev_fb);
	if (h->st) cudaStreamDestroy(h->st);
	if (h->st_copy) cudaStreamDestroy(h->st_copy);
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
	CU(cudaStreamSynchronize(h->st));
	h->origin = start_sample;
	h->cfg.start_sample = start_sample;
	h->pushed = 0;
	h->own_next = 0;
	h->finished = false;
	h->cur = 0;
	h->win_u0 = 0;
	h->win_off = kCarryMax;
	h->dev_fill = 0;
	h->pin_fill = 0;
	h->try_cur = 0;
	h->try_carried = 0;
	delete h->res;
	h->res = new Resolver(h->cfg.emulate_reference_tail != 0);
	h->frames.clear();
	h->emitted = 0;
	for (int i = 0; i < 3; i++)
		h->stat_ok[i] = 0;
	CU(cudaMemset(h->d_ctr, 0, sizeof(Counters)));
	memset(h->h_ctr, 0, sizeof(Counters));
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
	CU(cudaSetDevice(h->device));
	int rc = ensure_batch_buffers(h);
	if (rc) return rc;
	size_t done = 0;
	while (done < n) {
		int64_t room = batch_room(h);
		if (room <= 0) {
			rc = flush_batch(h);
			if (rc) return rc;
			continue;
		}
		const size_t take = std::min<size_t>((size_t)room, n - done);
		memcpy(h->pin[h->cur] + h->pin_fill, samples + done, take * sizeof(uint16_t));
		h->pin_fill += (int64_t)take;
		h->pushed += (int64_t)take;
		done += take;
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
	CU(cudaSetDevice(h->device));
	int rc = ensure_batch_buffers(h);
	if (rc) return rc;
	const uint16_t *src = (const uint16_t *)dev_samples;
	const int64_t seam = 2 * kLookahead + 2 * kCarryBack;  // samples of the piece the seam window needs
	const bool direct = (int64_t)n >= (int64_t)kMinFused * 4 && ((uintptr_t)src % 16) == 0 &&
	                    (h->pushed % 8) == 0;
	if (!direct) {
		// small or unaligned piece: append to the batch window on the device
		size_t done = 0;
		while (done < n) {
			int64_t room = batch_room(h);
			if (room <= 0) {
				rc = flush_batch(h);
				if (rc) return rc;
				continue;
			}
			rc = drain_pinned(h);
			if (rc) return rc;
			const size_t take = std::min<size_t>((size_t)room, n - done);
			CU(cudaMemcpyAsync(h->dev[h->cur] + h->win_off + h->dev_fill, src + done,
			                   take * sizeof(uint16_t), cudaMemcpyDeviceToDevice, h->st));
			h->dev_fill += (int64_t)take;
			h->pushed += (int64_t)take;
			done += take;
		}
		CU(cudaStreamSynchronize(h->st));
		return 0;
	}
	// Large aligned piece: (1) seam window = [carry | first `seam` samples of the piece] in our own
	// buffer, (2) the body straight from the caller's memory.
	rc = drain_pinned(h);
	if (rc) return rc;
	const int64_t piece_u0 = h->pushed;
	if (h->dev_fill > 0) {
		CU(cudaMemcpyAsync(h->dev[h->cur] + h->win_off + h->dev_fill, src, (size_t)seam * sizeof(uint16_t),
		                   cudaMemcpyDeviceToDevice, h->st));
		const int64_t u0 = h->win_u0, u1 = h->win_u0 + h->dev_fill + seam;
		rc = process_window(h, h->dev[h->cur] + h->win_off, u0, u1);
		if (rc) return rc;
	}
	h->pushed += (int64_t)n;
	rc = process_window(h, src, piece_u0, piece_u0 + (int64_t)n);
	if (rc) return rc;
	// carry the tail of the piece into the next window
	const int64_t u1 = piece_u0 + (int64_t)n;
	int64_t c0 = std::max<int64_t>(piece_u0, align_down(2 * h->own_next - kCarryBack, 8));
	const int64_t clen = u1 - c0;
	if (clen > kCarryMax)
		return fail(ADSB_B200_EINVAL, "internal: carry too large (%lld)", (long long)clen);
	const int nxt = 1 - h->cur;
	const int64_t off = kCarryMax - align_down(clen + 7, 8);
	CU(cudaMemcpyAsync(h->dev[nxt] + off, src + (c0 - piece_u0), (size_t)clen * sizeof(uint16_t),
	                   cudaMemcpyDeviceToDevice, h->st));
	CU(cudaStreamSynchronize(h->st));
	h->cur = nxt;
	h->win_u0 = c0;
	h->win_off = off;
	h->dev_fill = clen;
	h->pin_fill = 0;
	return 0;
}

int adsbdec_b200_finish(adsbdec_b200_t *h)
{
	if (!h)
		return fail(ADSB_B200_EINVAL, "null handle");
	if (h->finished)
		return 0;
	CU(cudaSetDevice(h->device));
	if (h->dev[0] && (h->pin_fill > 0 || h->dev_fill > 0)) {
		int rc = flush_batch(h);
		if (rc) return rc;
	}
	const size_t f0 = h->frames.size();
	h->res->finish(h->frames);
	h->h_ctr->n_try = h->try_carried;
	if (h->d_tries[0]) {
		int rc = stats_pass(h, f0);
		if (rc) return rc;
	}
	emit_new_frames(h);
	h->finished = true;
	return 0;
}

int adsbdec_b200_decode_file(adsbdec_b200_t *h, const char *path)
{
	if (!h || !path)
		return fail(ADSB_B200_EINVAL, "null argument");
	int fd = open(path, O_RDONLY);
	if (fd < 0)
		return fail(ADSB_B200_EIO, "cannot open %s: %s", path, strerror(errno));
	// air.c:233-239: one reused IQBUFFSZ buffer, read() then decodeiq(buf, n/2).  decodeiq eats
	// groups of 4 and so reads up to 3 stale entries behind a ragged tail; reusing one
	// zero-initialised buffer reproduces exactly what it would see.
	uint16_t *buf = (uint16_t *)calloc(kReadSamples + 4, sizeof(uint16_t));
	if (!buf) {
		close(fd);
		return fail(ADSB_B200_ENOMEM, "out of memory");
	}
	int rc = 0;
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
	CU(cudaMemcpy(h->h_ctr, h->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost));
	for (int i = 0; i < 3; i++) {
		if (try_) try_[i] = h->h_ctr->stat_try[i];
		if (ok) ok[i] = h->stat_ok[i];
	}
	return 0;
}

size_t adsbdec_b200_num_frames(const adsbdec_b200_t *h) { return h ? h->frames.size() : 0; }
const adsbdec_b200_frame_t *adsbdec_b200_frames(const adsbdec_b200_t *h)
{
	return h && !h->frames.empty() ? h->frames.data() : nullptr;
}
void adsbdec_b200_clear_frames(adsbdec_b200_t *h)
{
	if (!h)
		return;
	// keep the last frame: its span may still cover tries of the next stats pass
	if (h->emitted == h->frames.size() && h->frames.size() > 1) {
		h->frames.erase(h->frames.begin(), h->frames.end() - 1);
		h->emitted = h->frames.size();
	}
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
		CU(cudaMalloc(&d_raw, n_samples * sizeof(uint16_t)));
		CU(cudaMemcpy(d_raw, samples, n_samples * sizeof(uint16_t), cudaMemcpyHostToDevice));
		raw = d_raw;
	}
	cudaError_t e = cudaMalloc(&d_amp, std::max<size_t>(count, 1) * sizeof(float));
	if (e == cudaSuccess)
		e = launch_envelope_generic(raw, 0, start_sample, 0, (long long)count, d_amp, 0);
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
	if (own_end < own_begin || (own_end + kLookahead) * 2 > n_samples)
		return fail(ADSB_B200_EINVAL, "owned range needs 1200 envelope samples of look-ahead in the chunk");
	if (((uintptr_t)dev_samples % 16) != 0 || start_sample % 4)
		return fail(ADSB_B200_EINVAL, "chunk must be 16-byte aligned and start on a multiple of 4 samples");
	CU(cudaSetDevice(h->device));
	const unsigned long long saved = h->try_carried, saved_origin = h->origin;
	h->try_carried = 0;
	h->origin = start_sample;
	int rc = scan_range(h, (const uint16_t *)dev_samples, 0, (int64_t)n_samples, (int64_t)own_begin,
	                    (int64_t)own_end);
	h->try_carried = saved;
	h->origin = saved_origin;
	if (rc) return rc;
	if (h->cand_tmp.size() > cap)
		return fail(ADSB_B200_ERANGE, "output capacity %zu < %zu candidates", cap, h->cand_tmp.size());
	if (!h->cand_tmp.empty())
		memcpy(out, h->cand_tmp.data(), h->cand_tmp.size() * sizeof(adsbdec_b200_candidate_t));
	if (tries_out) {
		// raw gate-pass counts in the owned range (no sequential exclusion in chunk mode)
		tries_out[0] = tries_out[1] = tries_out[2] = 0;
		const unsigned long long nt = h->h_ctr->n_try;
		if (nt) {
			std::vector<unsigned long long> t(nt);
			CU(cudaMemcpy(t.data(), h->d_tries[h->try_cur], nt * sizeof(unsigned long long),
			              cudaMemcpyDeviceToHost));
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
	const int64_t n1 = (int64_t)(n_samples / 2) - kLookahead;
	if (n1 <= 0)
		return fail(ADSB_B200_EINVAL, "capture too short");
	uint16_t *d_raw = nullptr;
	CU(cudaMalloc(&d_raw, (n_samples + 64) * sizeof(uint16_t)));
	CU(cudaMemcpy(d_raw, samples, n_samples * sizeof(uint16_t), cudaMemcpyHostToDevice));
	CU(cudaMalloc(&h->dbg_amp, count * sizeof(float)));
	CU(cudaMalloc(&h->dbg_p, count / 8));
	CU(cudaMalloc(&h->dbg_q, count / 8));
	CU(cudaMalloc(&h->dbg_b, count / 8));
	CU(cudaMemset(h->dbg_amp, 0, count * sizeof(float)));
	CU(cudaMemset(h->dbg_p, 0, count / 8));
	CU(cudaMemset(h->dbg_q, 0, count / 8));
	CU(cudaMemset(h->dbg_b, 0, count / 8));
	h->dbg_n = (long long)count;
	const unsigned long long saved_origin = h->origin;
	const int saved_ctas = h->fused_ctas;
	h->origin = start_sample;
	h->fused_ctas = 1;  // one CTA, so span 0 covers the head of the capture in order
	k_set_counters<<<1, 1, 0, h->st>>>(h->d_ctr, 0, h->try_carried, 0, 1);
	int rc = launch_fused_range(h, d_raw, 0, (int64_t)n_samples, 0, n1);
	h->origin = saved_origin;
	h->fused_ctas = saved_ctas;
	cudaError_t e = cudaStreamSynchronize(h->st);
	if (rc == 0 && e == cudaSuccess) {
		cudaMemcpy(amp_out, h->dbg_amp, count * sizeof(float), cudaMemcpyDeviceToHost);
		cudaMemcpy(p_out, h->dbg_p, count / 8, cudaMemcpyDeviceToHost);
		cudaMemcpy(q_out, h->dbg_q, count / 8, cudaMemcpyDeviceToHost);
		cudaMemcpy(b_out, h->dbg_b, count / 8, cudaMemcpyDeviceToHost);
		if (j_begin_out) *j_begin_out = h->dbg_jb;
	}
	cudaFree(h->dbg_amp);
	cudaFree(h->dbg_p);
	cudaFree(h->dbg_q);
	cudaFree(h->dbg_b);
	cudaFree(d_raw);
	h->dbg_amp = nullptr;
	h->dbg_p = h->dbg_q = h->dbg_b = nullptr;
	h->dbg_n = 0;
	k_set_counters<<<1, 1, 0, h->st>>>(h->d_ctr, 0, h->try_carried, 0, 1);
	cudaStreamSynchronize(h->st);
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
	if (fn)
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

int adsbdec_b200_resolver_finish(adsbdec_b200_resolver_t *r, adsbdec_b200_sink_fn fn, void *user)
{
	if (!r)
		return fail(ADSB_B200_EINVAL, "null argument");
	r->r.finish(r->out);
	resolver_emit(r, fn, user);
	return 0;
}

uint64_t adsbdec_b200_resolver_scan_end(const adsbdec_b200_resolver_t *r) { return r ? r->r.scan_end() : 0; }

const char *adsbdec_b200_last_error(void) { return t_err; }
const char *adsbdec_b200_version(void) { return "adsbdec_b200 0.1 (sm_100a)"; }

}  // extern "C"
