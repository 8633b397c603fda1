// kernel_fused.cu - the B200 hot path: one persistent kernel that takes raw 16-bit samples from
// HBM and emits CRC-valid Mode S candidates, with the envelope never leaving the SM.
//
// What it computes is exactly the per-offset part of the reference (TLeconte/adsbdec):
//   air.c:59-92    envelope: fs/4 mixer, 14-tap FIR in ring-slot order, |.|^2, decimate by 2
//   demod.c:102-107 preamble test on int-truncated pair sums
//   demod.c:46-81  DF11 / DF17 / DF18 gate on the first five bits
//   demod.c:31-44  PPM bit decisions
//   valid.c:39-82, crc.h:36-42  CRC-24 syndrome == 0
// evaluated at EVERY offset; the sequential rules (greedy skip, ts, tail) are applied afterwards on
// the sparse output (kernels_resolve.cu + resolver.cpp), and pw (demod.c:127) is filled in for the
// few CRC-valid candidates by a follow-up kernel (kernels_generic.cu: k_fill_pw).
//
// Organisation: every WARP is an autonomous worker.  It owns ONE long contiguous span of the capture
// per launch (the host cuts a job into exactly as many spans as there are workers, so the ~1.4 k
// offsets of look-ahead a span has to compute past its end are ~2 % of its work) and streams it
// through its own slice of shared memory in sub-blocks of S = 28*32 = 896 envelope samples.  There
// is no CTA barrier in the steady state - only __syncwarp and one mbarrier per warp for its TMA
// loads - so the warps of an SM sit in different stages at any time and the FMA-heavy and ALU-heavy
// stages overlap on the pipes.
//
//   * load:    raw samples of a sub-block arrive by a 1-D TMA bulk copy (cp.async.bulk + mbarrier,
//              3.6 KB), issued by lane 0 as soon as the previous sub-block has been read;
//   * stage 1 (FIR): each lane owns 28 consecutive outputs.  28 = lcm(4,7): the ring-slot rotation of
//              the reference's summation order has period 7 outputs, so with spans cut at phase 0
//              every FMA's tap and order is a compile-time constant.  I and Q accumulate as the two
//              halves of packed fma.rn.f32x2 (the even/odd samples of a 32-bit word ARE the I/Q
//              pair); u16 -> float by the 0x4B000000 magic-number trick (no I2F);
//   * stage 2 (bits): per offset three decisions, all as sign bits collected by funnel shifts,
//              with f[n] = (int)(a[n]+a[n+10])  (floor via add.rm against 2^23):
//                P[n] = f[n] > 2*f[n+5]    (pulse pair at n, n+10 against the gap pair at n+5, n+15)
//                Q[n] = f[n+5] > 2*f[n]    (the mirrored test; the preamble is P[n] & Q[n+30])
//                B[n] = a[n] > a[n+5]      (PPM bit decision)
//              transposed inside 8-lane groups (8 runs of 28 = 7 words of 32) into bit rings;
//   * stage 3 (gate): word-parallel, one lane per 32 offsets: preamble = P[n] & Q[n+30]; DF gate =
//              five bits of B at stride 10.  It only needs 120 offsets of look-ahead, so it runs
//              right behind the bit front (on the words of the previous sub-block: no barrier of
//              its own) and appends its survivors to a small ring with a ballot + popc rank;
//   * stage 4 (slice+CRC): four lanes per survivor, EIGHT survivors per pass, and a pass only runs
//              when eight are waiting (or the oldest is about to lose its bits), so its ~100
//              instructions always work for 32 lanes.  A lane loads the ten B words that hold its
//              28 frame bits once, shifts them so that its first bit leads, pulls 7 nibbles out at
//              constant positions (funnel shift + multiply-gather), looks their syndromes up in a
//              28x16 table and the four lanes XOR-reduce; syndrome 0 -> the frame bits go to the
//              candidate list.  (REPAIR instantiation only: a non-zero syndrome is looked up in
//              repair.h's table - an extension, off by default.)
// Three __syncwarp per sub-block: before the raw stage is refilled, after the envelope (and the
// gate), after the bit arrays.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"
#include "kernels.h"
#include "repair.h"

namespace adsb {

namespace {

// One CTA of 20 warps per SM measured 9 % faster than 5 CTAs of 4 warps (same 20 warps, 96
// registers; profiles/README.md), 21 warps (80 registers) and 16 warps (126 registers) slower.
#ifndef ADSB_WPC
#define ADSB_WPC 20  // warps per CTA
#endif
#ifndef ADSB_MINB
#define ADSB_MINB 1  // CTAs per SM
#endif
#ifndef ADSB_LIST_CAP
#define ADSB_LIST_CAP 256  // survivor ring of a warp (tests build a tiny one to reach the overflow path)
#endif

constexpr int R = 28;               // outputs per lane per sub-block
constexpr int WPC = ADSB_WPC;
constexpr int NT = 32 * WPC;
constexpr int S = R * 32;           // envelope samples per sub-block of one warp (896)
constexpr int HIST = 8;             // words of FIR history in front of a sub-block (6 needed)
constexpr int STAGE_WORDS = S + HIST;
constexpr int LAG2 = 8;             // stage 2 works 8 runs behind the FIR front
constexpr int RUNS = 32 + LAG2;     // envelope ring, in runs of 28 floats
constexpr int WPB = S / 32;         // bit words per sub-block (28)
constexpr int WPQ = 64;             // P and Q rings, in 32-bit words
constexpr int WRB = 256;            // B ring: long enough for survivors to wait for a full slicer pass
constexpr int MIRROR_Q = 4;         // words 0.. of a ring mirrored behind its end (readers never wrap)
constexpr int MIRROR_B = 12;
constexpr int LAG3W = 39;           // stage 3 of sub-block k handles words 28k-39 .. 28k-12
constexpr int LIST_CAP = ADSB_LIST_CAP;
constexpr int BATCH = 8;            // survivors per slicer pass (four lanes each)
constexpr int FORCE_AGE = 6;        // a survivor older than this many sub-blocks is sliced at once
constexpr int TRY_FLUSH = 64;       // tries go to the global list in pieces of at least this many
// A span must run until the B bits of its last owned offset exist: stage 2 of sub-block k writes
// words up to 28k+20, the last frame bit of offset n sits at n+1190.
constexpr int LOOKBACK = 1440;

static_assert(LAG3W - 4 >= WPB + 7, "stage 3 of a sub-block must only need words of earlier sub-blocks");
static_assert(WPQ >= LAG3W + 21 + 1 && (WPQ & (WPQ - 1)) == 0, "P/Q ring too small");
static_assert((WRB & (WRB - 1)) == 0 && (LIST_CAP & (LIST_CAP - 1)) == 0, "rings must be powers of two");
static_assert(LOOKBACK % 32 == 0 && LOOKBACK >= 1190 + 32 * 7 + 1, "span look-back");
// a survivor found in sub-block k' starts at word >= 28k'-39 and must be sliced before stage 2 of
// sub-block k'+FORCE_AGE+1 overwrites the word behind its preamble (28k'-37 + WRB)
static_assert(28 * (FORCE_AGE + 1) + 20 < WRB - 37, "B ring too short for the slicer's waiting time");

struct WarpSmem {
	uint32_t raw[STAGE_WORDS];          // 3616 B
	float amp[RUNS * R];                // 4480 B
	uint32_t pbits[WPQ + MIRROR_Q];     // P[n]: first pulse pair beats its gap
	uint32_t qbits[WPQ + MIRROR_Q];     // Q[n]: pulse at n+5.. beats the gap at n..
	uint32_t bbits[WRB + MIRROR_B];     // B[n]: PPM decision
	uint32_t cand[LIST_CAP];            // survivor ring (offset relative to the span, class in the top bits)
	unsigned long long full;            // mbarrier of the raw stage
	unsigned long long pad_;
};
static_assert(sizeof(WarpSmem) % 16 == 0, "per-warp shared memory must keep 16-byte alignment");

struct Smem {
	uint32_t lut[28 * 16];              // lut[g][v]: syndrome of nibble g of a 112-bit frame holding v
	WarpSmem w[WPC];
};

// The syndrome table, built once per device by fused_prepare().  Frame bit b contributes
// x^(111-b) mod G (crc.h's table is the byte-wise form of the same polynomial division; the last
// 24 bits are XORed in unreduced, crc.h:40-42).  Entry [g][v]: nibble g (frame bits 4g..4g+3, MSB
// first) holding value v.  A 56-bit frame is the tail of the table: its nibble g is row g+14.
__device__ uint32_t g_lut[28 * 16];

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
	             : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "WAIT_%=:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra DONE_%=;\n"
	    "bra WAIT_%=;\n"
	    "DONE_%=:\n"
	    "}\n" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}

// 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes,
                                            unsigned long long *bar)
{
	asm volatile(
	    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
	        smem_u32(dst)),
	    "l"(src), "r"(bytes), "r"(smem_u32(bar))
	    : "memory");
}

__device__ __forceinline__ uint32_t lanemask_lt()
{
	uint32_t m;
	asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
	return m;
}

constexpr float kMagic = 8388608.0f;           // 2^23: float(0x4B000000 | x) == 2^23 + x
constexpr float kMagicMid = 8388608.0f + 2048.0f;
constexpr uint32_t FULL = 0xFFFFFFFFu;

}  // namespace

// REPAIR: the opt-in syndrome repair of repair.h (a separate instantiation: the reference-parity
// kernel carries none of it).
template <bool DEBUG, bool REPAIR>
__global__ void __launch_bounds__(NT, ADSB_MINB) k_fused(const FusedParams p)
{
	//@region setup
	extern __shared__ __align__(128) unsigned char smem_raw[];
	Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
	const int tid = threadIdx.x;
	const int lane = tid & 31;
	WarpSmem &ws = sm.w[tid >> 5];

	// ---- one-time setup (the only CTA barrier) ----------------------------
	for (int e = tid; e < 28 * 16; e += NT)
		sm.lut[e] = g_lut[e];
	if (lane == 0) {
		mbar_init(&ws.full, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	uint32_t phase = 0;  // mbarrier phase parity

	// air.c:36-45 taps paired for a 32-bit word of age g (0 = newest): the even sample (I, low
	// half) has age 2g+1, the odd sample (Q, high half) age 2g; tap = h[13-age] = h[age].
	const float h[7] = {0.012627f, 0.025254f, 0.037881f, 0.050508f, 0.063135f, 0.075761f, 0.088388f};
	float2 tp[7];
#pragma unroll
	for (int g = 0; g < 7; g++) {
		const int ai = 2 * g + 1, aq = 2 * g;  // ages
		tp[g] = make_float2(h[ai < 7 ? ai : 13 - ai], h[aq < 7 ? aq : 13 - aq]);
	}
	// lane constants of the 8-lane bit transposition and of the slicer
	const int j8 = lane & 7;
	const int wl = 7 * (lane >> 3) + j8 - 7;   // word written by this lane in sub-block k: 28k + wl
	const int sub = lane & 3;                  // quarter of a frame this lane slices
	const uint32_t lt = lanemask_lt();
	const uint32_t lane_mask3 = lane < WPB ? 0xFFFFFFFFu : 0u;      // lanes that own a word in stage 3
	const uint32_t df18_mask = p.accept_df18 ? 0xFFFFFFFFu : 0u;

	for (;;) {
		//@region span
		// ---- fetch a span (dynamic scheduler; the host cuts at most one span per worker) --------
		int span_idx = 0;
		if (lane == 0)
			span_idx = atomicAdd(p.span_ticket, 1);
		span_idx = __shfl_sync(FULL, span_idx, 0);
		if (span_idx >= p.n_spans)
			break;
		const FusedSpan sp = p.spans[span_idx];
		const long long jb = sp.j_begin;
		const int nb = sp.n_blocks;
		// address of the first word of sub-block 0 (word jb - HIST); only valid words are touched
		const char *raw_s0 = reinterpret_cast<const char *>(sp.raw_w0) + 4 * (jb - HIST);
		const uint32_t own_lo = (uint32_t)(sp.own_begin - jb), own_hi = (uint32_t)(sp.own_end - jb);
		// valid words of the window, relative to word jb - HIST (clamped to int range)
		const long long lo64 = sp.w_lo - (jb - HIST), hi64 = sp.w_hi - (jb - HIST);
		const int vlo = (int)(lo64 < 0 ? 0 : lo64);
		const int vhi = (int)(hi64 > 0x7FFFFFF0ll ? 0x7FFFFFF0ll : (hi64 < 0 ? 0 : hi64));
		// sub-blocks kf_lo <= k <= kf_hi lie entirely inside the window: one full-size copy each
		const int kf_lo = (vlo + S - 1) / S;
		const int kf_hi = vhi >= STAGE_WORDS ? (vhi - STAGE_WORDS) / S : -1;
		const uint32_t kf_n = kf_hi >= kf_lo ? (uint32_t)(kf_hi - kf_lo) : 0u;
		const bool kf_any = kf_hi >= kf_lo;

		auto issue_partial = [&](int kb) {
			// lane 0: one bulk copy for the part of sub-block kb that the window holds
			const int start = kb * S;
			const int a = max(start, vlo), b = min(start + STAGE_WORDS, vhi);
			if (b > a) {
				const uint32_t bytes = (uint32_t)(b - a) * 4u;
				mbar_expect_tx(&ws.full, bytes);
				tma_load_1d(&ws.raw[a - start], raw_s0 + 4ll * a, bytes, &ws.full);
			} else {
				mbar_arrive(&ws.full);
			}
		};
		auto is_full = [&](int kb) { return kf_any && (uint32_t)(kb - kf_lo) <= kf_n; };
		if (lane == 0) {
			if (is_full(0)) {
				mbar_expect_tx(&ws.full, STAGE_WORDS * 4u);
				tma_load_1d(&ws.raw[0], raw_s0, STAGE_WORDS * 4u, &ws.full);
			} else {
				issue_partial(0);
			}
		}

		// Survivor ring: entries [done, head) wait for the slicer, [flushed, head) for the global try
		// list (valid.c:46, :68: every offset that reaches the CRC is a "try").  All warp-uniform.
		uint32_t head = 0, done = 0, flushed = 0;

		auto flush_tries = [&]() {
			const uint32_t nt = head - flushed;
			unsigned long long base = 0;
			if (lane == 0)
				base = atomicAdd(&p.ctr->n_try, (unsigned long long)nt);
			base = __shfl_sync(FULL, base, 0);
			for (uint32_t i = lane; i < nt; i += 32) {
				const uint32_t ent = ws.cand[(flushed + i) & (LIST_CAP - 1)];
				if (base + i < p.try_cap)
					p.tries[base + i] = ((unsigned long long)(ent >> 30) << kTryClsShift) |
					                    (unsigned long long)(jb + (long long)(ent & 0x3FFFFFFFu));
				else
					atomicOr(&p.ctr->overflow, 2u);
			}
			flushed = head;
		};

		// slice + CRC for ring entries [done, done + n) (n <= BATCH), four lanes per survivor
		auto slice_pass = [&](uint32_t n) {
			const uint32_t ci = lane >> 2;
			const bool have = ci < n;
			const uint32_t ent = ws.cand[(done + ci) & (LIST_CAP - 1)];
			const uint32_t n_rel = ent & 0x3FFFFFFFu;
			const bool is_long = (ent >> 30) != kClsDF11;
			// short frames are the last 56 bits of the syndrome table and only use sub 0, 1
			const bool active = have && (is_long || sub < 2);
			// frame bit b sits at n + 80 + 10 b; nibble g starts at bit 4g
			const uint32_t pb0 = n_rel + 80u + 280u * (uint32_t)sub;
			const char *lutp = reinterpret_cast<const char *>(sm.lut) + (7 * sub + (is_long ? 0 : 14)) * 64;
			uint32_t acc = 0, bits28 = 0;
			if (active) {
				// The lane's 28 frame bits sit 10 apart in 271 consecutive B bits: fetch the ten
				// words once (mirrored past the end of the ring) and shift them so that the
				// lane's first bit is bit 31 of W[0] - every later position is then a constant.
				const uint32_t *bb = &ws.bbits[(pb0 >> 5) & (WRB - 1)];
				const uint32_t sh = pb0 & 31u;
				uint32_t r[10], W[9];
#pragma unroll
				for (int i = 0; i < 10; i++)
					r[i] = bb[i];
#pragma unroll
				for (int i = 0; i < 9; i++)
					W[i] = __funnelshift_l(r[i + 1], r[i], sh);
#pragma unroll
				for (int s_ = 0; s_ < 7; s_++) {
					const int kk = 40 * s_, i = kk >> 5, o = kk & 31;
					const uint32_t x = o ? __funnelshift_l(W[i + 1], W[i], o) : W[i];  // bits 31, 21, 11, 1
					// the product's bits 27..22 are zero (no two partial products land there)
					const uint32_t prod = (x & 0x80200802u) * 0x08040201u;
					acc ^= *reinterpret_cast<const uint32_t *>(lutp + 64 * s_ + (prod >> 26));
					bits28 = __funnelshift_l(prod, bits28, 4);
				}
			}
			// valid.c:48-52 / :70-74: CRC over all but the last 3 bytes, XOR the last 3, == 0
			acc ^= __shfl_xor_sync(FULL, acc, 1);
			acc ^= __shfl_xor_sync(FULL, acc, 2);
			bool ok = have && acc == 0;
			uint32_t fix = 0xFFFFu;  // (first << 8) | second flipped bit, 0xFF = none (repair.h)
			if (REPAIR && have && acc != 0) {
				const uint32_t e = fix_lookup(p.fix_tab, acc);
				if (e != kFixNone && fix_allowed(e, is_long, p.fix_level)) {
					ok = true;
					fix = e;
				}
			}
			if (__any_sync(FULL, ok)) {
				unsigned long long slot = 0;
				if (ok && sub == 0)
					slot = atomicAdd(&p.ctr->n_cand, 1ull);
				slot = __shfl_sync(FULL, slot, lane & ~3);
				if (ok) {
					if (slot < p.cand_cap) {
						uint32_t *out = reinterpret_cast<uint32_t *>(&p.cands[slot]);
						if (sub == 0) {
							const unsigned long long n = (unsigned long long)(jb + (long long)n_rel);
							// pw is filled in by k_fill_pw (demod.c:127)
							*reinterpret_cast<uint4 *>(out) = make_uint4(
							    (uint32_t)n, (uint32_t)(n >> 32), 0u, (is_long ? 14u : 7u) | kCandNibbleWords);
						}
						uint32_t word = bits28;
						if (REPAIR && fix != 0xFFFFu) {
							const uint32_t base = is_long ? 0u : 56u;
							const uint32_t b1 = (fix >> 8) - base, b2 = (fix & 0xFFu) - base;
							if (b1 / 28u == (uint32_t)sub)
								word ^= 0x08000000u >> (b1 % 28u);
							if ((fix & 0xFFu) != 0xFFu && b2 / 28u == (uint32_t)sub)
								word ^= 0x08000000u >> (b2 % 28u);
							if (sub == 0)  // bits 28, 29 of word 0: how many bits were repaired
								word |= ((fix & 0xFFu) != 0xFFu ? 2u : 1u) << 28;
						}
						out[4 + sub] = word;
					} else if (sub == 0) {
						atomicOr(&p.ctr->overflow, 1u);
					}
				}
			}
			done += n;
		};

		// stage 3: word handled by this lane in sub-block k (28k - 39 + lane), fully owned words
		int wd3 = lane - LAG3W;
		const bool ow_any = (own_hi >> 5) > ((own_lo + 31u) >> 5);
		const int ow_lo = ow_any ? (int)((own_lo + 31u) >> 5) : 0x40000000;  // none: every word takes the masked path
		const uint32_t ow_n = ow_any ? (own_hi >> 5) - 1u - (uint32_t)ow_lo : 0u;
		int q_fir = lane;          // ring run of this lane's FIR output, (k*32 + lane) % RUNS
		int wbase = wl;            // 28k + wl
		const char *src_next = raw_s0 + 4ll * S;  // global address of sub-block k+1

		for (int k = 0; k < nb; k++) {
			//@region stage1
			// ================= stage 1: envelope ============================
			{
				const bool partial = !is_full(k);  // warp-uniform
				if (partial) {
					// words the window does not hold read as mid-scale: before the stream start the
					// reference's ring is zero (air.c:33), past the end they are never used
					const int start = k * S;
					const int va = max(start, vlo) - start, vb = min(start + STAGE_WORDS, vhi) - start;
					for (int i = lane; i < STAGE_WORDS; i += 32)
						if (i < va || i >= vb)
							ws.raw[i] = 0x08000800u;
				}
				mbar_wait(&ws.full, phase);
				phase ^= 1u;
				if (partial)
					__syncwarp();  // the fill above was written by other lanes

				uint32_t w[36];
				const uint4 *src = reinterpret_cast<const uint4 *>(&ws.raw[R * lane]);
#pragma unroll
				for (int i = 0; i < 9; i++) {
					const uint4 t = src[i];
					w[4 * i] = t.x;
					w[4 * i + 1] = t.y;
					w[4 * i + 2] = t.z;
					w[4 * i + 3] = t.w;
				}

				// 12-bit domain check (the float tricks below are exact for samples < 4096)
				uint32_t any = 0;
#pragma unroll
				for (int i = HIST; i < 36; i++)
					any |= w[i];
				if (any & 0xF000F000u)
					p.ctr->wide = 1;
				// The raw stage may be refilled only once every lane HOLDS its words: the test above
				// consumed all of them, so past this barrier the loads have returned (issuing the copy
				// right behind the load instructions lost words to the incoming data now and then).
				__syncwarp();
				if (lane == 0 && k + 1 < nb) {
					asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
					if (is_full(k + 1)) {
						mbar_expect_tx(&ws.full, STAGE_WORDS * 4u);
						tma_load_1d(&ws.raw[0], src_next, STAGE_WORDS * 4u, &ws.full);
					} else {
						issue_partial(k + 1);
					}
				}
				src_next += 4ll * S;

				// u16 pair -> (x-2048) * sign as float2.  Word parity == parity of its index here
				// (jb, S, R*lane are even): odd words are samples t%4 in {2,3}, negated (air.c:79-82)
				float2 v[36];
#pragma unroll
				for (int L = 2; L < 36; L++) {
					const float lo = __uint_as_float(__byte_perm(w[L], 0x4B000000u, 0x7610));
					const float hi = __uint_as_float(__byte_perm(w[L], 0x4B000000u, 0x7632));
					if (L & 1)
						v[L] = __ffma2_rn(make_float2(lo, hi), make_float2(-1.0f, -1.0f),
						                  make_float2(kMagicMid, kMagicMid));
					else
						v[L] = __fadd2_rn(make_float2(lo, hi), make_float2(-kMagicMid, -kMagicMid));
				}
				float4 *dst = reinterpret_cast<float4 *>(&ws.amp[q_fir * R]);
				float o[4];
#pragma unroll
				for (int r_ = 0; r_ < R; r_++) {
					// air.c:69-74: ascending ring slot.  Pair-slot kap holds the word of age
					// (phi - kap) mod 7, phi = r_ % 7 by construction of the span start.
					const int phi = r_ % 7;
					float2 acc = __fmul2_rn(tp[phi], v[HIST + r_ - phi]);  // kap = 0
#pragma unroll
					for (int kap = 1; kap < 7; kap++) {
						const int g = (phi - kap + 7) % 7;
						acc = __ffma2_rn(tp[g], v[HIST + r_ - g], acc);
					}
					// air.c:76 as compiled by the reference flags: fma(I, I, round(Q*Q))
					o[r_ & 3] = __fmaf_rn(acc.x, acc.x, __fmul_rn(acc.y, acc.y));
					if ((r_ & 3) == 3) {
						dst[r_ >> 2] = make_float4(o[0], o[1], o[2], o[3]);
						if (DEBUG && p.dbg_amp && span_idx == 0) {
							const long long at = (long long)k * S + R * lane + (r_ - 3);
							if (at + 4 <= p.dbg_n)
								*reinterpret_cast<float4 *>(&p.dbg_amp[at]) = make_float4(o[0], o[1], o[2], o[3]);
						}
					}
				}
			}

			//@region stage3
			// ================= stage 3: preamble AND + DF gate ===============
			// Runs BEFORE stage 2 of this sub-block: it looks at words 28k-39 .. 28k-8 of the bit rings,
			// all written by stage 2 of earlier sub-blocks (stage 2 of this one writes 28k-7 .. 28k+20), so
			// no barrier is needed in front of it.  Entries appended here are sliced from the next
			// sub-block on, when stage 2 has produced the last of their 1190 look-ahead bits.
			const uint32_t ready = head;
			{
				// every lane works (lanes 28..31 and words in front of the span are masked, not branched
				// around): the ring reads of a negative word index land on valid, unused words
				const uint32_t o0 = (uint32_t)wd3 * 32u;
				const int wq = wd3 & (WPQ - 1);  // wq+1 is readable thanks to the mirror
				// demod.c:107: p1 > 2*s1 at n, p2 > 2*s2 at n+30
				uint32_t pre = ws.pbits[wq] & __funnelshift_l(ws.qbits[wq + 1], ws.qbits[wq], 30) & lane_mask3;
				// ownership: only offsets in [own_lo, own_hi) (relative to j_begin) are emitted here; words
				// ow_lo .. ow_lo + ow_n lie entirely inside
				if ((uint32_t)(wd3 - ow_lo) > ow_n) {
					const int o0s = wd3 * 32;
					const int lo = max((int)own_lo - o0s, 0);           // first owned bit index
					const int hi = min((int)own_hi - o0s, 32);          // one past the last
					const uint32_t m_lo = lo >= 32 ? 0u : 0xFFFFFFFFu >> lo;
					const uint32_t m_hi = hi >= 32 ? 0xFFFFFFFFu : hi <= 0 ? 0u : ~(0xFFFFFFFFu >> hi);
					pre &= m_lo & m_hi;
				}
				const uint32_t *bw = &ws.bbits[wd3 & (WRB - 1)];  // +2 .. +4 readable thanks to the mirror
				const uint32_t b2 = bw[2], b3 = bw[3], b4 = bw[4];
				// first five data bits sit at n+80, +90, +100, +110, +120 (demod.c:110, :46-81)
				const uint32_t x7 = __funnelshift_l(b3, b2, 16);  // 80  = 64 + 16
				const uint32_t x6 = __funnelshift_l(b3, b2, 26);  // 90  = 64 + 26
				const uint32_t x5 = __funnelshift_l(b4, b3, 4);   // 100 = 96 + 4
				const uint32_t x4 = __funnelshift_l(b4, b3, 14);  // 110 = 96 + 14
				const uint32_t x3 = __funnelshift_l(b4, b3, 24);  // 120 = 96 + 24
				const uint32_t c0 = pre & ~x5;
				const uint32_t m17 = c0 & x7 & ~x6 & ~x4 & x3;                   // 10001
				const uint32_t m18 = c0 & x7 & ~x6 & x4 & ~x3 & df18_mask;       // 10010 (only with -a, demod.c:59)
				const uint32_t m11 = c0 & ~x7 & x6 & x4 & x3;                    // 01011
				uint32_t any = m17 | m18 | m11;
				// warp-cooperative append: one bit per lane and round, ranked by ballot + popc
				uint32_t bal = __ballot_sync(FULL, any != 0u);
				if (bal) {
					const uint32_t floor_ = p.want_tries ? min(done, flushed) : done;
					do {
						if (any) {
							const int i = __clz(any);
							const uint32_t bit = 0x80000000u >> i;
							any &= ~bit;
							// class in the top two bits: DF11 = 0, DF17 = 1, DF18 = 2
							const uint32_t cls = ((m17 << i) >> 1 & 0x40000000u) | ((m18 << i) & 0x80000000u);
							const uint32_t idx = head + __popc(bal & lt);
							if (idx - floor_ < (uint32_t)LIST_CAP)
								ws.cand[idx & (LIST_CAP - 1)] = (o0 + (uint32_t)i) | cls;
						}
						head += __popc(bal);
						bal = __ballot_sync(FULL, any != 0u);
					} while (bal);
					if (head - floor_ > (uint32_t)LIST_CAP) {  // absurd survivor density: the host re-runs the job generically
						if (lane == 0)
							atomicOr(&p.ctr->overflow, 4u);
						head = floor_ + (uint32_t)LIST_CAP;
					}
				}
				wd3 += WPB;
			}
			__syncwarp();  // envelope of sub-block k and the survivor ring visible to the warp

			//@region stage2
			// ================= stage 2: decision bits ========================
			{
				// run handled here: 32k + lane - LAG2 (uniform sign per 8-lane group).  In sub-block 0 the
				// first eight lanes work on whatever the ring holds: their words (index < 0) are never read.
				uint32_t P = 0, Q = 0, B = 0;
				{
					int q2 = q_fir - LAG2;
					if (q2 < 0) q2 += RUNS;
					int q2n = q2 + 1;
					if (q2n == RUNS) q2n = 0;
					float a[44];
					const float4 *s0 = reinterpret_cast<const float4 *>(&ws.amp[q2 * R]);
					const float4 *s1 = reinterpret_cast<const float4 *>(&ws.amp[q2n * R]);
#pragma unroll
					for (int i = 0; i < 7; i++) {
						const float4 t = s0[i];
						a[4 * i] = t.x; a[4 * i + 1] = t.y; a[4 * i + 2] = t.z; a[4 * i + 3] = t.w;
					}
#pragma unroll
					for (int i = 0; i < 4; i++) {
						const float4 t = s1[i];
						a[28 + 4 * i] = t.x; a[29 + 4 * i] = t.y; a[30 + 4 * i] = t.z; a[31 + 4 * i] = t.w;
					}
					// f[i] = floor(a[i] + a[i+10]) as a float: demod.c:102-105's (int) of a float sum.
					// add.rm against 2^23 leaves floor(c) in the mantissa (c >= 0, c < 2^22).
					float f[34];
#pragma unroll
					for (int i = 0; i < 34; i += 2) {
						const float2 c = __fadd2_rn(make_float2(a[i], a[i + 1]), make_float2(a[i + 10], a[i + 11]));
						const float2 m = __fadd2_rd(c, make_float2(kMagic, kMagic));
						const float2 ff = __fadd2_rn(m, make_float2(-kMagic, -kMagic));
						f[i] = ff.x;
						f[i + 1] = ff.y;
					}
#pragma unroll
					for (int i = 0; i < R; i++) {
						// demod.c:107, first half: p1 > 2*s1 with p1 = f[n], s1 = f[n+5]
						//   <=> 2*s1 - p1 < 0 (exact small integers, so the sign bit is the answer)
						const float e = __fmaf_rn(f[i + 5], 2.0f, -f[i]);
						P = __funnelshift_l(__float_as_uint(e), P, 1);
						// second half, as seen from m = n+30: p2 = f[m+5] > 2*s2, s2 = f[m]
						const float e2 = __fmaf_rn(f[i], 2.0f, -f[i + 5]);
						Q = __funnelshift_l(__float_as_uint(e2), Q, 1);
						// demod.c:34: a[n] > a[n+5]  <=>  a[n+5] - a[n] < 0  (no FTZ: x-y==0 iff x==y)
						const float d = __fsub_rn(a[i + 5], a[i]);
						B = __funnelshift_l(__float_as_uint(d), B, 1);
					}
				}
				// 8 runs x 28 bits -> 7 words x 32 bits inside each 8-lane group (MSB = first offset)
				const uint32_t Pn = __shfl_down_sync(FULL, P, 1, 8);
				const uint32_t Qn = __shfl_down_sync(FULL, Q, 1, 8);
				const uint32_t Bn = __shfl_down_sync(FULL, B, 1, 8);
				if (j8 < 7) {
					const int sl = 4 + 4 * j8, sr = 24 - 4 * j8;
					const uint32_t pw_ = (P << sl) | (Pn >> sr);
					const uint32_t qw_ = (Q << sl) | (Qn >> sr);
					const uint32_t bw_ = (B << sl) | (Bn >> sr);
					const int wq = wbase & (WPQ - 1), wb = wbase & (WRB - 1);
					ws.pbits[wq] = pw_;
					ws.qbits[wq] = qw_;
					ws.bbits[wb] = bw_;
					if (wq < MIRROR_Q)  // mirrors: stages 3/4 read runs of words without wrapping
						ws.qbits[WPQ + wq] = qw_;
					if (wb < MIRROR_B)
						ws.bbits[WRB + wb] = bw_;
					if (DEBUG && p.dbg_p && span_idx == 0) {
						if (wbase >= 0 && wbase < p.dbg_n / 32) {
							p.dbg_p[wbase] = pw_;
							p.dbg_q[wbase] = qw_;
							p.dbg_b[wbase] = bw_;
						}
					}
				}
			}
			__syncwarp();  // bit arrays of sub-block k visible to the warp

			//@region stage4
			// ================= stage 4: slice + CRC, eight survivors per pass ==
			// Entries appended before this sub-block have all their bits now.  Full passes only, unless
			// the oldest entry has waited FORCE_AGE sub-blocks (its B words are about to be overwritten).
			{
				const bool last = k + 1 == nb;  // end of the span: LOOKBACK guarantees every entry has its bits
				const uint32_t avail = (last ? head : ready) - done;
				uint32_t nrun = last ? avail : (avail & ~(uint32_t)(BATCH - 1));
				if (avail != 0u && nrun == 0u) {
					const uint32_t oldest = ws.cand[done & (LIST_CAP - 1)] & 0x3FFFFFFFu;
					if ((int)oldest < (k - FORCE_AGE) * S)
						nrun = avail;
				}
				while (nrun) {
					const uint32_t n = min(nrun, (uint32_t)BATCH);
					slice_pass(n);
					nrun -= n;
				}
				if (p.want_tries && (head - flushed >= (uint32_t)TRY_FLUSH || (last && head != flushed)))
					flush_tries();
			}
			// no barrier here: the next one every lane meets is behind the mbarrier wait of sub-block
			// k+1, and nothing shared is written before it except the raw stage, which is free
			q_fir += 32;
			if (q_fir >= RUNS) q_fir -= RUNS;
			wbase += WPB;
		}
		//@region tail
		__syncwarp();
	}
}

int fused_block_outputs() { return S; }
int fused_lookback() { return LOOKBACK; }
int fused_threads() { return NT; }
size_t fused_smem_bytes() { return sizeof(Smem); }

// Configures the kernel and returns the number of autonomous workers (warps) per SM.
int fused_prepare(int device)
{
	(void)device;
	uint32_t syn[112], lut[28 * 16];
	for (int b = 0; b < 112; b++) {
		uint32_t r = 1;
		for (int i = 0; i < 111 - b; i++)
			r = (r & 0x800000u) ? ((r << 1) ^ kCrcPoly) & 0xFFFFFFu : (r << 1);
		syn[b] = r;
	}
	for (int e = 0; e < 28 * 16; e++) {
		const int g = e >> 4, v = e & 15;
		lut[e] = ((v & 8) ? syn[4 * g] : 0u) ^ ((v & 4) ? syn[4 * g + 1] : 0u) ^ ((v & 2) ? syn[4 * g + 2] : 0u) ^
		         ((v & 1) ? syn[4 * g + 3] : 0u);
	}
	if (cudaMemcpyToSymbol(g_lut, lut, sizeof lut) != cudaSuccess)
		return -1;
	if (cudaFuncSetAttribute(k_fused<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                         (int)fused_smem_bytes()) != cudaSuccess ||
	    cudaFuncSetAttribute(k_fused<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                         (int)fused_smem_bytes()) != cudaSuccess ||
	    cudaFuncSetAttribute(k_fused<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                         (int)fused_smem_bytes()) != cudaSuccess)
		return -1;
	int per_sm = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fused<false, false>, NT, fused_smem_bytes()) !=
	    cudaSuccess)
		return -1;
	return per_sm * WPC;
}

// `workers` = number of warps wanted (at most one per span).
cudaError_t launch_fused(const FusedParams &p, int workers, cudaStream_t st)
{
	if (workers <= 0)
		return cudaSuccess;
	const int grid = (workers + WPC - 1) / WPC;
	if (p.dbg_amp)
		k_fused<true, false><<<grid, NT, fused_smem_bytes(), st>>>(p);
	else if (p.fix_level > 0)
		k_fused<false, true><<<grid, NT, fused_smem_bytes(), st>>>(p);
	else
		k_fused<false, false><<<grid, NT, fused_smem_bytes(), st>>>(p);
	return cudaGetLastError();
}

}  // namespace adsb
