// kernel_fused.cu - the B200 hot path: raw 16-bit samples in HBM -> CRC-valid Mode S candidates, with the
// envelope never leaving the SM.
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
// Three kernels, split where the arithmetic density changes:
//
// k_fused (dominant; streams the samples once): every WARP is an autonomous worker.  It owns ONE long
//   contiguous span of the capture per launch and streams it through its own slice of shared memory in
//   sub-blocks of S = 28*32 = 896 envelope samples.  No CTA barrier in the steady state - only
//   __syncwarp and one mbarrier per warp for its TMA loads - so the warps of an SM sit in different
//   stages at any time and the FMA-heavy and ALU-heavy stages overlap on the pipes.
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
//              transposed inside 8-lane groups (8 runs of 28 = 7 words of 32) and stored to three bit
//              arrays in HBM: 3 bits per offset against 32 bits of input, and they stay in the L2.
//   The gate and the slicer used to live in this kernel; they are short dependent chains around
//   shared-memory loads and atomics, and at the 20 warps per SM the FIR's registers allow they took
//   half of the warps' time for a third of the instructions (profiles/r2a).
//
// k_gate (word-parallel, one lane per 128 offsets): preamble = P[n] & Q[n+30]; DF gate = five bits of B
//   at stride 10; survivors ("tries", valid.c:46, :68) are compacted with a warp scan and one atomic
//   per few hundred tries into the try list.
//
// k_slice (four lanes per try): a lane loads the ten B words that hold its 28 frame bits once, shifts
//   them so that its first bit leads, pulls 7 nibbles out at constant positions (funnel shift +
//   multiply-gather), looks their syndromes up in a 28x16 table and the four lanes XOR-reduce;
//   syndrome 0 -> the frame bits go to the candidate list.  (REPAIR instantiation only: a non-zero
//   syndrome is looked up in repair.h's table - an extension, off by default.)
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
#ifndef ADSB_PQ_INT
#define ADSB_PQ_INT 0      // preamble comparisons on the integer pipe instead of FFMA
#endif
#ifndef ADSB_B_INT
#define ADSB_B_INT 0       // PPM comparison on the integer pipe instead of FADD
#endif
#ifndef ADSB_BITS_SPLIT
#define ADSB_BITS_SPLIT 0  // two funnel-shift chains per bit stream
#endif
#ifndef ADSB_MINB
#define ADSB_MINB 1  // CTAs per SM
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
constexpr int QUANTUM = 224;        // lcm(28, 32): spans start on FIR phase 0 AND on a bit-word boundary

struct WarpSmem {
	uint32_t raw[STAGE_WORDS];          // 3616 B
	float amp[RUNS * R];                // 4480 B
	unsigned long long full;            // mbarrier of the raw stage
	unsigned long long pad_;
};
static_assert(sizeof(WarpSmem) % 16 == 0, "per-warp shared memory must keep 16-byte alignment");

struct Smem {
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

constexpr float kMagic = 8388608.0f;           // 2^23: float(0x4B000000 | x) == 2^23 + x
constexpr float kMagicMid = 8388608.0f + 2048.0f;
constexpr uint32_t FULL = 0xFFFFFFFFu;

}  // namespace

// ============================================================================================
// k_fused: samples -> P / Q / B bit arrays
// ============================================================================================
template <bool DEBUG>
__global__ void __launch_bounds__(NT, ADSB_MINB) k_fused(const FusedParams p)
{
	//@region setup
	extern __shared__ __align__(128) unsigned char smem_raw[];
	Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
	const int tid = threadIdx.x;
	const int lane = tid & 31;
	WarpSmem &ws = sm.w[tid >> 5];

	if (lane == 0) {
		mbar_init(&ws.full, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();
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
	// lane constants of the 8-lane bit transposition
	const int j8 = lane & 7;
	const int wl = 7 * (lane >> 3) + j8 - 7;   // word written by this lane in sub-block k: 28k + wl

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
		const uint32_t n_words = (uint32_t)sp.n_words;
		unsigned int *const out_p = p.bits_p + sp.word_base;
		unsigned int *const out_q = p.bits_q + sp.word_base;
		unsigned int *const out_b = p.bits_b + sp.word_base;
		// address of the first word of sub-block 0 (word jb - HIST); only valid words are touched
		const char *raw_s0 = reinterpret_cast<const char *>(sp.raw_w0) + 4 * (jb - HIST);
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

				// 12-bit domain check (the float tricks below are exact for samples < 4096), over every word this
				// lane uses: the history words in front of a job's first span belong to nobody else's check
				uint32_t any = 0;
#pragma unroll
				for (int i = 2; i < 36; i++)
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

			__syncwarp();  // envelope of sub-block k visible to the warp

			//@region stage2
			// ================= stage 2: decision bits ========================
			{
				// run handled here: 32k + lane - LAG2 (uniform sign per 8-lane group).  In sub-block 0 the
				// first eight lanes work on whatever the ring holds: their words (index < 0) are not stored.
				uint32_t P = 0, Q = 0, B = 0;
#if ADSB_BITS_SPLIT
				uint32_t P2 = 0, Q2 = 0, B2 = 0;
#endif
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
						// second half, as seen from m = n+30: p2 = f[m+5] > 2*s2, s2 = f[m]
#if ADSB_PQ_INT
						// the same on the integer pipe: for integer-valued floats x > 2y <=> bits(x) > bits(y) + 2^23
						const uint32_t e = __float_as_uint(f[i + 5]) + 0x00800000u - __float_as_uint(f[i]);
						const uint32_t e2 = __float_as_uint(f[i]) + 0x00800000u - __float_as_uint(f[i + 5]);
#else
						const uint32_t e = __float_as_uint(__fmaf_rn(f[i + 5], 2.0f, -f[i]));
						const uint32_t e2 = __float_as_uint(__fmaf_rn(f[i], 2.0f, -f[i + 5]));
#endif
						// demod.c:34: a[n] > a[n+5]  <=>  a[n+5] - a[n] < 0  (no FTZ: x-y==0 iff x==y)
#if ADSB_B_INT
						// non-negative floats order like their bit patterns
						const uint32_t d = __float_as_uint(a[i + 5]) - __float_as_uint(a[i]);
#else
						const uint32_t d = __float_as_uint(__fsub_rn(a[i + 5], a[i]));
#endif
#if ADSB_BITS_SPLIT
						// two chains per bit stream: the dependent funnel shifts are the longest chain of the stage
						if (i < R / 2) {
							P = __funnelshift_l(e, P, 1);
							Q = __funnelshift_l(e2, Q, 1);
							B = __funnelshift_l(d, B, 1);
						} else {
							P2 = __funnelshift_l(e, P2, 1);
							Q2 = __funnelshift_l(e2, Q2, 1);
							B2 = __funnelshift_l(d, B2, 1);
						}
#else
						P = __funnelshift_l(e, P, 1);
						Q = __funnelshift_l(e2, Q, 1);
						B = __funnelshift_l(d, B, 1);
#endif
					}
#if ADSB_BITS_SPLIT
					P = (P << (R / 2)) | P2;
					Q = (Q << (R / 2)) | Q2;
					B = (B << (R / 2)) | B2;
#endif
				}
				// 8 runs x 28 bits -> 7 words x 32 bits inside each 8-lane group (MSB = first offset)
				const uint32_t Pn = __shfl_down_sync(FULL, P, 1, 8);
				const uint32_t Qn = __shfl_down_sync(FULL, Q, 1, 8);
				const uint32_t Bn = __shfl_down_sync(FULL, B, 1, 8);
				// words [0, n_words) of the span go to the bit arrays (coalesced: 28 consecutive words per array)
				if (j8 < 7 && (uint32_t)wbase < n_words) {
					const int sl = 4 + 4 * j8, sr = 24 - 4 * j8;
					out_p[wbase] = (P << sl) | (Pn >> sr);
					out_q[wbase] = (Q << sl) | (Qn >> sr);
					out_b[wbase] = (B << sl) | (Bn >> sr);
				}
			}
			// no barrier here: the next one every lane meets is behind the register loads of sub-block k+1,
			// in front of its first envelope store (which reuses the ring runs read above)
			q_fir += 32;
			if (q_fir >= RUNS) q_fir -= RUNS;
			wbase += WPB;
		}
		//@region tail
		__syncwarp();
	}
}

// ============================================================================================
// k_gate: bit arrays -> tries (demod.c:107 preamble test, demod.c:46-81 DF gate), one lane per 32 offsets
// ============================================================================================
// A single counter cannot take an atomic per 1024 offsets (same-address atomics retire at ~1.2 per ns in
// the L2: 0.5 ms per 60 s capture), so every warp collects its tries in shared memory and reserves room
// in the global list a few hundred at a time.
constexpr int GATE_STAGE = 512;  // tries staged per warp (2 KB), as 32-bit offsets relative to j0

// The try list gets the bare offsets here; k_slice, which reads the first frame bits anyway, adds the DF
// class (demod.c:46-81: 10001 = DF17, 10010 = DF18, 01011 = DF11) that the statistics pass wants.
#ifndef ADSB_GATE_MINB
#define ADSB_GATE_MINB 4
#endif
#ifndef ADSB_SLICE_MINB
#define ADSB_SLICE_MINB 5
#endif
// AP (address/parity extension, off by default): the same kernel lists the offsets whose first five bits read
// 0010x / 1010x (DF4, DF5, DF20, DF21) in a list of their own (counter Counters::n_ap, overflow bit 2).
template <bool AP>
__global__ void __launch_bounds__(256, ADSB_GATE_MINB) k_gate(const GateParams p)
{
	//@region gate
	__shared__ uint32_t stage_all[8][GATE_STAGE];
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t lt = (1u << lane) - 1u;
	uint32_t *const stage = stage_all[threadIdx.x >> 5];
	const uint32_t warp = blockIdx.x * 8u + (threadIdx.x >> 5);
	const uint32_t stride = gridDim.x * 8u * 128u;  // words per step of the whole grid
	// words that hold owned offsets (n0 < n1, j0 <= n0, fewer than 2^31 offsets per launch) and the masks of
	// the two words the range may own only in part
	const uint32_t r0 = (uint32_t)(p.n0 - p.j0), r1 = (uint32_t)(p.n1 - 1 - p.j0);
	const uint32_t g_lo = r0 >> 5, g_hi = r1 >> 5;
	const uint32_t m_first = 0xFFFFFFFFu >> (r0 & 31u);
	const uint32_t m_last = 0xFFFFFFFFu << (31u - (r1 & 31u));
	const uint32_t df18_mask = p.accept_df18 ? 0xFFFFFFFFu : 0u;
	// a lane takes FOUR consecutive words (16-byte loads: five per lane and step instead of 24 four-byte ones)
	const uint4 *__restrict__ const bp = reinterpret_cast<const uint4 *>(p.bits_p);
	const uint4 *__restrict__ const bq = reinterpret_cast<const uint4 *>(p.bits_q);
	const uint4 *__restrict__ const bb = reinterpret_cast<const uint4 *>(p.bits_b);
	const unsigned int *__restrict__ const bq1 = p.bits_q;
	unsigned long long *const tries = p.tries;
	unsigned long long *const n_try = AP ? &p.ctr->n_ap : &p.ctr->n_try;
	const unsigned long long try_cap = p.try_cap;
	const long long j0 = p.j0;
	uint32_t staged = 0;  // entries in `stage`, warp-uniform
	bool ovf = false;

	// `staged` entries go to the try list: one atomic per few hundred tries, coalesced stores
	auto flush = [&]() {
		__syncwarp();  // the staged entries were written by other lanes
		unsigned long long base = 0;
		if (lane == 0)
			base = atomicAdd(n_try, (unsigned long long)staged);
		base = __shfl_sync(FULL, base, 0);
		for (uint32_t i = lane; i < staged; i += 32u) {
			if (base + i < try_cap)
				tries[base + i] = (unsigned long long)(j0 + (long long)stage[i]);
			else
				ovf = true;
		}
		__syncwarp();  // the stage may be overwritten
		staged = 0;
	};

	struct Words {
		uint4 p, q, b0, b1;  // P[G..G+3], Q[G..G+3], B[G..G+3], B[G+4..G+7]
		uint32_t q4;         // Q[G+4]
	};
	// (words past g_hi are never owned; the arrays are allocated 16 words beyond the last one written)
	auto fetch = [&](uint32_t G) {
		Words w;
		w.p = w.q = w.b0 = w.b1 = make_uint4(0u, 0u, 0u, 0u);
		w.q4 = 0u;
		if (G <= g_hi) {
			w.p = bp[G >> 2];
			w.q = bq[G >> 2];
			w.q4 = bq1[G + 4];
			w.b0 = bb[G >> 2];
			w.b1 = bb[(G >> 2) + 1];
		}
		return w;
	};
	// The words of the NEXT step are requested before this step is worked on: the kernel is a chain of
	// DRAM loads -> a few logic ops -> shared-memory appends.
	uint32_t G = (g_lo & ~3u) + warp * 128u + 4u * lane;
	Words nxt = fetch(G);
	// (G - 4 lane <= g_hi: warp-uniform; the range holds fewer than 2^26 words, so G + stride cannot wrap)
	for (; G - 4u * lane <= g_hi; G += stride) {
		const Words w = nxt;
		nxt = fetch(G + stride);
		const uint32_t P[4] = {w.p.x, w.p.y, w.p.z, w.p.w};
		const uint32_t Q[5] = {w.q.x, w.q.y, w.q.z, w.q.w, w.q4};
		const uint32_t B[8] = {w.b0.x, w.b0.y, w.b0.z, w.b0.w, w.b1.x, w.b1.y, w.b1.z, w.b1.w};
		uint32_t any[4];
		uint32_t c = 0;
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const uint32_t g = G + (uint32_t)k;
			// demod.c:107: p1 > 2*s1 at n, p2 > 2*s2 at n+30
			uint32_t pre = P[k] & __funnelshift_l(Q[k + 1], Q[k], 30);
			if (g < g_lo || g > g_hi)
				pre = 0u;
			if (g == g_lo)
				pre &= m_first;
			if (g == g_hi)
				pre &= m_last;
			// first five data bits sit at n+80, +90, +100, +110, +120 (demod.c:110, :46-81)
			const uint32_t b2 = B[k + 2], b3 = B[k + 3], b4 = B[k + 4];
			const uint32_t x7 = __funnelshift_l(b3, b2, 16);  // 80  = 64 + 16
			const uint32_t x6 = __funnelshift_l(b3, b2, 26);  // 90  = 64 + 26
			const uint32_t x5 = __funnelshift_l(b4, b3, 4);   // 100 = 96 + 4
			const uint32_t x4 = __funnelshift_l(b4, b3, 14);  // 110 = 96 + 14
			const uint32_t x3 = __funnelshift_l(b4, b3, 24);  // 120 = 96 + 24
			if (AP) {
				any[k] = pre & x5 & ~x6 & ~x4;           // 0010x, 1010x (extension)
			} else {
				const uint32_t c0 = pre & ~x5;
				const uint32_t lng = c0 & x7 & ~x6;
				any[k] = (lng & ~x4 & x3)                  // 10001
				         | (lng & x4 & ~x3 & df18_mask)    // 10010 (only with -a, demod.c:59)
				         | (c0 & ~x7 & x6 & x4 & x3);      // 01011
			}
			c += (uint32_t)__popc(any[k]);
		}
		// Where each lane's tries go in the stage: counts stay below 8 almost always, so the exclusive prefix
		// over the lanes is three ballots and population counts (no shared-memory atomics, no scan).
		const uint32_t v0 = __ballot_sync(FULL, c & 1u), v1 = __ballot_sync(FULL, c & 2u), v2 = __ballot_sync(FULL, c & 4u);
		const uint32_t big = __ballot_sync(FULL, c > 7u);
		uint32_t total, at;
		if (big == 0u) {
			total = (uint32_t)__popc(v0) + 2u * (uint32_t)__popc(v1) + 4u * (uint32_t)__popc(v2);
			at = (uint32_t)__popc(v0 & lt) + 2u * (uint32_t)__popc(v1 & lt) + 4u * (uint32_t)__popc(v2 & lt);
		} else {
			uint32_t x = c;  // inclusive scan by shuffles
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t y = __shfl_up_sync(FULL, x, d);
				if ((int)lane >= d)
					x += y;
			}
			total = __shfl_sync(FULL, x, 31);
			at = x - c;
		}
		if (total == 0u)
			continue;
		const uint32_t rel = G << 5;  // offset of bit 31 of this lane's first word, relative to j0
		if (total > (uint32_t)GATE_STAGE) {
			// absurd density (more than an eighth of 4096 offsets pass the gate): straight to the list
			unsigned long long dst = 0;
			if (lane == 0)
				dst = atomicAdd(n_try, (unsigned long long)total);
			dst = __shfl_sync(FULL, dst, 0) + at;
#pragma unroll
			for (int k = 0; k < 4; k++) {
				uint32_t m = any[k];
				while (m) {
					const int i = __clz(m);
					m &= ~(0x80000000u >> i);
					if (dst < try_cap)
						tries[dst] = (unsigned long long)(j0 + (long long)(rel + 32u * (uint32_t)k + (uint32_t)i));
					else
						ovf = true;
					dst++;
				}
			}
			continue;
		}
		if (staged + total > (uint32_t)GATE_STAGE)
			flush();
		at += staged;
#pragma unroll
		for (int k = 0; k < 4; k++) {
			uint32_t m = any[k];
			while (m) {
				const int i = __clz(m);
				m &= ~(0x80000000u >> i);
				stage[at++] = rel + 32u * (uint32_t)k + (uint32_t)i;
			}
		}
		staged += total;
	}
	if (staged)
		flush();
	if (ovf)
		atomicOr(&p.ctr->overflow, AP ? 4u : 2u);
}

// ============================================================================================
// k_slice: tries -> CRC-valid candidates (demod.c:31-44 slicer, valid.c:39-82 syndrome), one lane per try
// ============================================================================================
// 28 frame bits starting at bit position pb of the B array (10 apart: 271 consecutive bits): ten words are
// fetched once and shifted so that the first bit is bit 31 of W[0] - every later position is then a
// compile-time constant (funnel shift + multiply-gather of bits 31, 21, 11, 1 into a nibble).  Returns
// the 28 bits and XORs the syndromes of their seven nibbles (rows of `lutp`) into acc.
__device__ __forceinline__ uint32_t slice28(const unsigned int *__restrict__ bits_b, uint32_t pb, const char *lutp,
                                            uint32_t &acc)
{
	const unsigned int *bb = bits_b + (pb >> 5);
	const uint32_t sh = pb & 31u;
	uint32_t r[10], W[9];
#pragma unroll
	for (int i = 0; i < 10; i++)
		r[i] = bb[i];
#pragma unroll
	for (int i = 0; i < 9; i++)
		W[i] = __funnelshift_l(r[i + 1], r[i], sh);
	uint32_t bits28 = 0;
#pragma unroll
	for (int s_ = 0; s_ < 7; s_++) {
		const int kk = 40 * s_, i = kk >> 5, o = kk & 31;
		const uint32_t x = o ? __funnelshift_l(W[i + 1], W[i], o) : W[i];
		// the product's bits 27..22 are zero (no two partial products land there)
		const uint32_t prod = (x & 0x80200802u) * 0x08040201u;
		acc ^= *reinterpret_cast<const uint32_t *>(lutp + 64 * s_ + (prod >> 26));
		bits28 = __funnelshift_l(prod, bits28, 4);
	}
	return bits28;
}

constexpr int SLICE_STAGE = 64;  // candidate records staged per warp (2 KB): a pass adds at most 32

template <bool REPAIR>
__global__ void __launch_bounds__(256, ADSB_SLICE_MINB) k_slice(const SliceParams p)
{
	//@region slice
	__shared__ uint32_t lut[28 * 16];
	__shared__ uint4 stage_all[8][SLICE_STAGE * 2];
	for (int e = threadIdx.x; e < 28 * 16; e += 256)
		lut[e] = g_lut[e];
	__syncthreads();
	const uint32_t lane = threadIdx.x & 31u;
	uint4 *const stage = stage_all[threadIdx.x >> 5];
	const unsigned long long first = p.ctr->try_mark;
	unsigned long long last = p.ctr->n_try;
	if (last > p.try_cap)
		last = p.try_cap;
	if (last <= first)
		return;
	const uint32_t count = (uint32_t)min(last - first, 0xFFFFFF00ull);
	unsigned long long *tries = p.tries + first;
	const uint32_t stride = gridDim.x * 256u;
	uint32_t staged = 0;  // records in `stage`, warp-uniform
	bool ovf = false;

	// one atomic for all staged records (see k_gate), then a coalesced copy
	auto flush = [&]() {
		unsigned long long base = 0;
		if (lane == 0)
			base = atomicAdd(&p.ctr->n_cand, (unsigned long long)staged);
		base = __shfl_sync(FULL, base, 0);
		uint4 *out = reinterpret_cast<uint4 *>(p.cands);
		for (uint32_t i = lane; i < 2u * staged; i += 32u) {
			if (base + (i >> 1) < p.cand_cap)
				out[2ull * base + i] = stage[i];
			else
				ovf = true;
		}
		__syncwarp();
		staged = 0;
	};

	// whole warps iterate together (the ballot below needs every lane)
	// (the entry of the NEXT pass is requested a pass ahead: list entry -> frame bits -> syndrome is a chain of
	// dependent trips to the L2, and this kernel waits for them ten cycles of eleven)
	unsigned long long e_next = 0;
	{
		const uint32_t t = blockIdx.x * 256u + threadIdx.x;
		if (t < count)
			e_next = tries[t];
	}
	for (uint32_t t0 = blockIdx.x * 256u + (threadIdx.x & ~31u); t0 < count; t0 += stride) {
		const uint32_t t = t0 + lane;
		const bool have = t < count;
		const long long n_abs = have ? (long long)(e_next & kTryMask) : p.j0;
		if (t + stride < count && t + stride > t)
			e_next = tries[t + stride];
		// frame bit b sits at n + 80 + 10 b; the first one tells 10001 / 10010 (112 bits) from 01011 (56 bits,
		// the tail of the syndrome table: rows 14..27)
		const uint32_t pb = (uint32_t)(n_abs - p.j0) + 80u;
		const bool is_long = ((p.bits_b[pb >> 5] << (pb & 31u)) >> 31) != 0u;
		const char *lutp = reinterpret_cast<const char *>(lut) + (is_long ? 0 : 14 * 64);
		uint32_t acc = 0, w0 = 0, w1 = 0, w2 = 0, w3 = 0;
		if (have) {
			w0 = slice28(p.bits_b, pb, lutp, acc);
			w1 = slice28(p.bits_b, pb + 280u, lutp + 7 * 64, acc);
			if (is_long) {
				w2 = slice28(p.bits_b, pb + 560u, lutp + 14 * 64, acc);
				w3 = slice28(p.bits_b, pb + 840u, lutp + 21 * 64, acc);
			}
		}
		if (have)  // the class for the statistics pass (valid.c:46, :68): frame bit 3 tells DF18 from DF17
			tries[t] = ((unsigned long long)(is_long ? ((w0 >> 24) & 1u ? kClsDF18 : kClsDF17) : kClsDF11) << kTryClsShift) |
			           (unsigned long long)n_abs;
		// valid.c:48-52 / :70-74: CRC over all but the last 3 bytes, XOR the last 3, == 0
		bool ok = have && acc == 0;
		if (REPAIR && have && acc != 0) {
			// extension (repair.h): the syndrome of one or two flipped bits
			const uint32_t e = fix_lookup(p.fix_tab, acc);
			if (e != kFixNone && fix_allowed(e, is_long, p.fix_level)) {
				ok = true;
				const uint32_t base = is_long ? 0u : 56u;
				uint32_t nfix = 0;
				for (uint32_t fb = e >> 8;; fb = e & 0xFFu) {
					const uint32_t b = fb - base, m = 0x08000000u >> (b % 28u);
					if (b < 28u) w0 ^= m;
					else if (b < 56u) w1 ^= m;
					else if (b < 84u) w2 ^= m;
					else w3 ^= m;
					nfix++;
					if (nfix == 2 || (e & 0xFFu) == 0xFFu)
						break;
				}
				w0 |= nfix << 28;  // bits 28, 29 of word 0: how many bits were repaired
			}
		}
		const uint32_t okm = __ballot_sync(FULL, ok);
		if (okm) {
			if (ok) {
				uint4 *rec = stage + 2u * (staged + (uint32_t)__popc(okm & ((1u << lane) - 1u)));
				// pw is filled in by k_fill_pw (demod.c:127)
				rec[0] = make_uint4((uint32_t)n_abs, (uint32_t)((unsigned long long)n_abs >> 32), 0u,
				                    (is_long ? 14u : 7u) | kCandNibbleWords);
				rec[1] = make_uint4(w0, w1, w2, w3);
			}
			staged += (uint32_t)__popc(okm);
			__syncwarp();
			if (staged > (uint32_t)(SLICE_STAGE - 32))
				flush();
		}
	}
	if (staged)
		flush();
	if (ovf)
		atomicOr(&p.ctr->overflow, 1u);
}

// ============================================================================================
// EXTENSION (off by default, not in the reference): address/parity frames
// ============================================================================================
// A clean DF11 / DF17 / DF18 candidate announces its AA field (frame bits 8..31) at its offset.
__global__ void k_note_addresses(const CandRec *__restrict__ cands, unsigned long long cand_cap, const Counters *ctr,
                                 unsigned long long *first_seen)
{
	unsigned long long last = ctr->n_cand;
	if (last > cand_cap)
		last = cand_cap;
	for (unsigned long long i = ctr->cand_mark + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < last;
	     i += (unsigned long long)gridDim.x * blockDim.x) {
		const CandRec c = cands[i];
		uint32_t aa;
		if (c.len & kCandNibbleWords) {
			uint32_t w[2];
			memcpy(w, c.bytes, 8);
			if (w[0] >> 28)  // repaired by the other extension: not a clean announcement
				continue;
			aa = ((w[0] & 0xFFFFFu) << 4) | ((w[1] >> 24) & 0xFu);
		} else {
			if (c.bytes[15])
				continue;
			aa = (uint32_t)c.bytes[1] << 16 | (uint32_t)c.bytes[2] << 8 | c.bytes[3];
		}
		atomicMin(&first_seen[aa], (unsigned long long)c.n);
	}
}

// One lane per DF4/5/20/21 try: the syndrome of the frame bits IS the address the sender overlaid on the parity
// field; the frame is a candidate when that address was announced at an earlier offset.
__global__ void __launch_bounds__(256) k_slice_ap(const AddressParityParams p)
{
	__shared__ uint32_t lut[28 * 16];
	for (int e = threadIdx.x; e < 28 * 16; e += 256)
		lut[e] = g_lut[e];
	__syncthreads();
	unsigned long long count = p.ctr->n_ap;
	if (count > p.ap_cap)
		count = p.ap_cap;
	bool ovf = false;
	for (unsigned long long t = blockIdx.x * 256ull + threadIdx.x; t < count; t += (unsigned long long)gridDim.x * 256ull) {
		const long long n_abs = (long long)p.ap_tries[t];
		const uint32_t pb = (uint32_t)(n_abs - p.j0) + 80u;
		const bool is_long = ((p.bits_b[pb >> 5] << (pb & 31u)) >> 31) != 0u;
		const char *lutp = reinterpret_cast<const char *>(lut) + (is_long ? 0 : 14 * 64);
		uint32_t acc = 0, w0, w1, w2 = 0, w3 = 0;
		w0 = slice28(p.bits_b, pb, lutp, acc);
		w1 = slice28(p.bits_b, pb + 280u, lutp + 7 * 64, acc);
		if (is_long) {
			w2 = slice28(p.bits_b, pb + 560u, lutp + 14 * 64, acc);
			w3 = slice28(p.bits_b, pb + 840u, lutp + 21 * 64, acc);
		}
		if (acc == 0u || p.first_seen[acc] >= (unsigned long long)n_abs)
			continue;
		const unsigned long long k = atomicAdd(&p.ctr->n_cand, 1ull);
		if (k >= p.cand_cap) {
			ovf = true;
			continue;
		}
		uint4 *rec = reinterpret_cast<uint4 *>(p.cands + k);
		rec[0] = make_uint4((uint32_t)n_abs, (uint32_t)((unsigned long long)n_abs >> 32), 0u,
		                    (is_long ? 14u : 7u) | kCandNibbleWords);
		rec[1] = make_uint4(w0, w1, w2, w3);
	}
	if (ovf)
		atomicOr(&p.ctr->overflow, 1u);
}

int fused_block_outputs() { return S; }
int fused_span_quantum() { return QUANTUM; }
int fused_threads() { return NT; }
size_t fused_smem_bytes() { return sizeof(Smem); }

// Configures the kernels and returns the number of autonomous workers (warps) per SM.
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
	if (cudaFuncSetAttribute(k_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem_bytes()) !=
	        cudaSuccess ||
	    cudaFuncSetAttribute(k_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem_bytes()) !=
	        cudaSuccess)
		return -1;
	int per_sm = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fused<false>, NT, fused_smem_bytes()) != cudaSuccess)
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
		k_fused<true><<<grid, NT, fused_smem_bytes(), st>>>(p);
	else
		k_fused<false><<<grid, NT, fused_smem_bytes(), st>>>(p);
	return cudaGetLastError();
}

// Grids of the two light kernels: enough resident warps to hide the L2 latency of their loads.
cudaError_t launch_gate(const GateParams &p, cudaStream_t st)
{
	if (p.n1 <= p.n0)
		return cudaSuccess;
	const long long words = ((p.n1 - 1 - p.j0) >> 5) - ((p.n0 - p.j0) >> 5) + 1;
	const long long want = (words + 1023) / 1024;  // one step of 8 warps: 128 words each
	const long long most = 148 * ADSB_GATE_MINB;
	k_gate<false><<<(unsigned)(want < most ? want : most), 256, 0, st>>>(p);
	return cudaGetLastError();
}

cudaError_t launch_gate_ap(const GateParams &p, cudaStream_t st)
{
	if (p.n1 <= p.n0)
		return cudaSuccess;
	const long long words = ((p.n1 - 1 - p.j0) >> 5) - ((p.n0 - p.j0) >> 5) + 1;
	const long long want = (words + 1023) / 1024;
	const long long most = 148 * ADSB_GATE_MINB;
	k_gate<true><<<(unsigned)(want < most ? want : most), 256, 0, st>>>(p);
	return cudaGetLastError();
}

cudaError_t launch_note_addresses(const CandRec *cands, unsigned long long cand_cap, const Counters *ctr,
                                  unsigned long long *first_seen, cudaStream_t st)
{
	k_note_addresses<<<148, 256, 0, st>>>(cands, cand_cap, ctr, first_seen);
	return cudaGetLastError();
}

cudaError_t launch_slice_ap(const AddressParityParams &p, cudaStream_t st)
{
	k_slice_ap<<<148 * 4, 256, 0, st>>>(p);
	return cudaGetLastError();
}

cudaError_t launch_slice(const SliceParams &p, cudaStream_t st)
{
	if (p.fix_level > 0)
		k_slice<true><<<148 * 8, 256, 0, st>>>(p);
	else
		k_slice<false><<<148 * 8, 256, 0, st>>>(p);
	return cudaGetLastError();
}

}  // namespace adsb
