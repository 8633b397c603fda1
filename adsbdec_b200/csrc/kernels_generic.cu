// kernels_generic.cu - the straightforward ("generic") CUDA formulation of the hot path.
//
// One thread per envelope sample / per offset, envelope written to HBM.  It is NOT the fast
// path (kernel_fused.cu is); it exists because it handles every case with no preconditions:
//   * the uint32 `fidx` wrap of the FIR ring (air.c:34, :65) - the few outputs whose window
//     straddles a 2^32-sample boundary,
//   * samples outside the 12-bit ADC range (the fused kernel's float tricks assume < 4096),
//   * tiny seams between pushed pieces,
// and because a second, independent formulation is what the fused kernel is differentially
// tested against on the GPU.
//
// Reference citations are file:line into TLeconte/adsbdec.
#include "common.h"
#include "kernels.h"
#include "repair.h"

namespace adsb {

__constant__ float c_taps[kTaps] = ADSB_TAPS_INIT;

static __constant__ CrcTable c_crc = make_crc_table();

// Envelope sample j (local index) of a stream whose local sample u has FIR-ring slot
// (uint32_t)(u + phase_origin) % 14.   air.c:59-76, SURVEY App. A.
//
//   raw      : device pointer to local sample u_raw0
//   u_limit  : local samples [0, u_limit) exist; u < 0 reads as "ring still zero" (air.c:33)
//   j0, n    : compute amp[j0 .. j0+n)
//   amp      : output, amp[0] = envelope sample j0
__device__ float envelope_at(const uint16_t *__restrict__ raw, long long u_raw0,
                             unsigned long long phase_origin, long long j)
{
	const long long T = 2 * j + 1;  // newest input sample of this output
	// ring_pos after pushing sample T is T+1 (air.c:65: fidx++ after each store)
	const uint32_t pos = (uint32_t)((unsigned long long)(T + 1) + phase_origin);
	const uint32_t rot = kTaps - pos % kTaps;  // air.c:68: o = DSFLTLEN - fidx%DSFLTLEN
	float si = 0.0f, sq = 0.0f;
	const uint32_t posT = pos - 1u;  // (uint32)(T + origin)
	if (posT >= 32u) {
		// No 2^32 wrap of fidx among the last 32 samples: slot k simply holds the sample that is
		// (slot(T) - k) mod 14 back.  All 14 loads are independent of the FMA chain.
		const uint32_t slotT = posT % kTaps;
		float v[kTaps];
#pragma unroll
		for (int k = 0; k < kTaps; k++) {
			const long long u = T - (long long)((slotT + kTaps - (uint32_t)k) % kTaps);
			float x = 0.0f;  // never written: the static ring is zero (air.c:33)
			if (u >= 0) {
				x = (float)raw[u - u_raw0] - 2048.0f;  // air.c:64
				x = (u & 2) ? -x : x;                  // air.c:79-82: 2nd pair of each 4 negated
			}
			v[k] = x;
		}
#pragma unroll
		for (int k = 0; k < kTaps; k++) {
			const float tap = c_taps[(k + rot) % kTaps];  // dsfilter[k+o], the table is one period doubled
			if (k & 1)
				sq = __fmaf_rn(tap, v[k], sq);  // air.c:71 (contracted by the reference's -O3 -march=native)
			else
				si = __fmaf_rn(tap, v[k], si);  // air.c:70
		}
		return __fmaf_rn(si, si, __fmul_rn(sq, sq));  // air.c:76: vmulss q,q ; vfmadd i,i
	}
	// Around a wrap the ring is briefly "scrambled" (SURVEY F12): slot k holds the newest sample
	// u <= T with (uint32)(u+origin)%14 == k, which can be up to 17 back.  Search for it.
#pragma unroll 1
	for (int k = 0; k < kTaps; k++) {
		float v = 0.0f;
		for (int back = 0; back < 32; back++) {
			long long u = T - back;
			if (u < 0)
				break;
			if ((uint32_t)((unsigned long long)u + phase_origin) % kTaps == (uint32_t)k) {
				float x = (float)raw[u - u_raw0] - 2048.0f;
				v = (u & 2) ? -x : x;
				break;
			}
		}
		const float tap = c_taps[(k + rot) % kTaps];
		if (k & 1)
			sq = __fmaf_rn(tap, v, sq);
		else
			si = __fmaf_rn(tap, v, si);
	}
	return __fmaf_rn(si, si, __fmul_rn(sq, sq));
}

__global__ void k_envelope_generic(const uint16_t *__restrict__ raw, long long u_raw0,
                                   unsigned long long phase_origin, long long j0, long long n,
                                   float *__restrict__ amp, int iq)
{
	long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	if (iq) {
		// EXTENSION: the stream is complex baseband, one signed 16-bit I/Q pair per envelope
		// sample; the envelope is |z|^2 in the form of air.c:76
		const long long u = 2 * (j0 + i) - u_raw0;
		const float I = (float)(int16_t)raw[u], Q = (float)(int16_t)raw[u + 1];
		amp[i] = __fmaf_rn(I, I, __fmul_rn(Q, Q));
		return;
	}
	amp[i] = envelope_at(raw, u_raw0, phase_origin, j0 + i);
}

// pw = (p1+p2)/4 (demod.c:102-105, :127) for the candidates the fused kernel found: it hands over
// the frame bits only, the four envelope samples are recomputed here from the raw window.
__global__ void k_fill_pw(const uint16_t *__restrict__ raw, long long u_raw0, unsigned long long phase_origin,
                          CandRec *cands, const unsigned long long *n_cand_ptr, unsigned long long cap)
{
	unsigned long long n_cand = *n_cand_ptr;
	if (n_cand > cap)
		n_cand = cap;
	// four lanes per candidate: a[n], a[n+10], a[n+35], a[n+45]
	const int sub = threadIdx.x & 3;
	const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x / 4;
	const unsigned long long rounds = (n_cand + stride - 1) / stride;
	unsigned long long i = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) / 4;
	for (unsigned long long r = 0; r < rounds; r++, i += stride) {
		const bool todo = i < n_cand && (cands[i].len & kCandNibbleWords);  // else a complete generic record
		float a = 0.0f;
		if (todo) {
			const long long n = (long long)cands[i].n;
			a = envelope_at(raw, u_raw0, phase_origin, n + (sub == 0 ? 0 : sub == 1 ? 10 : sub == 2 ? 35 : 45));
		}
		const float pair = __fadd_rn(a, __shfl_xor_sync(0xFFFFFFFFu, a, 1));
		const int pi = __float2int_rz(pair);
		const int sum = pi + __shfl_xor_sync(0xFFFFFFFFu, pi, 2);
		if (todo && sub == 0)
			cands[i].pw = (uint32_t)(sum / 4);
	}
}

// The same for ACCEPTED frames (the stream path): a third of the candidates survive the greedy rule, so the
// signal level is only worked out for those, behind the compaction.
// (128 threads x 32 registers: fits beside k_fused of the next job, see kernels_resolve.cu)
__global__ void __launch_bounds__(128, 16) k_fill_pw_frames(const uint16_t *__restrict__ raw, long long u_raw0, unsigned long long phase_origin,
                                 FrameRec *frames, const unsigned long long *n_ptr, unsigned long long cap)
{
	unsigned long long n_fr = *n_ptr;
	if (n_fr > cap)
		n_fr = cap;
	const int sub = threadIdx.x & 3;
	const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x / 4;
	const unsigned long long rounds = (n_fr + stride - 1) / stride;
	unsigned long long i = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) / 4;
	for (unsigned long long r = 0; r < rounds; r++, i += stride) {
		const bool todo = i < n_fr;
		float a = 0.0f;
		if (todo) {
			const long long n = (long long)frames[i].n;
			a = envelope_at(raw, u_raw0, phase_origin, n + (sub == 0 ? 0 : sub == 1 ? 10 : sub == 2 ? 35 : 45));
		}
		const float pair = __fadd_rn(a, __shfl_xor_sync(0xFFFFFFFFu, a, 1));
		const int pi = __float2int_rz(pair);
		const int sum = pi + __shfl_xor_sync(0xFFFFFFFFu, pi, 2);
		if (todo && sub == 0)
			frames[i].pw = (uint32_t)(sum / 4);
	}
}

__device__ __forceinline__ bool ppm(const float *a, int at)
{
	return a[at] > a[at + kPulse];  // demod.c:34
}

// One thread per offset.  amp[0] is envelope sample j_amp0; offsets [n0, n0+count) are owned.
// Every owned offset needs amp up to n+1195 (demod.c:89: idx < len-DECOFFSET).
__global__ void k_scan_generic(const float *__restrict__ amp, long long j_amp0, long long n0,
                               long long count, int accept_df18, int want_tries, CandRec *cands,
                               unsigned long long cand_cap, unsigned long long *tries,
                               unsigned long long try_cap, Counters *ctr, const unsigned long long *fix_tab,
                               int fix_level, const unsigned long long *first_seen)
{
	// first_seen != null: the pass of the address/parity extension - only DF4/5/20/21, no tries
	long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	const long long n = n0 + i;
	const float *a = amp + (n - j_amp0);
	// demod.c:102-105: float sums truncated to int
	int p1 = __float2int_rz(__fadd_rn(a[0], a[2 * kPulse]));
	int s1 = __float2int_rz(__fadd_rn(a[kPulse], a[3 * kPulse]));
	if (!(p1 > 2 * s1))
		return;
	int p2 = __float2int_rz(__fadd_rn(a[7 * kPulse], a[9 * kPulse]));
	int s2 = __float2int_rz(__fadd_rn(a[6 * kPulse], a[8 * kPulse]));
	if (!(p2 > 2 * s2))  // demod.c:107
		return;
	// demod.c:46-81: only 10001xxx (DF17), 10010xxx (DF18, -a) and 01011xxx (DF11) go on
	const int d = 16 * kPulse;
	bool b7 = ppm(a, d), b6 = ppm(a, d + 10), b5 = ppm(a, d + 20), b4 = ppm(a, d + 30),
	     b3 = ppm(a, d + 40);
	int len = 0;
	uint32_t cls = 0;
	if (first_seen) {
		if (b5 && !b6 && !b4)
			len = b7 ? 14 : 7;  // 1010x: DF20, DF21; 0010x: DF4, DF5
	} else if (b7 && !b6 && !b5) {
		if (b4 && !b3 && accept_df18) {
			len = 14;
			cls = kClsDF18;
		} else if (!b4 && b3) {
			len = 14;
			cls = kClsDF17;
		}
	} else if (!b7 && b6 && !b5 && b4 && b3) {
		len = 7;
		cls = kClsDF11;
	}
	if (!len)
		return;
	if (want_tries && !first_seen) {
		unsigned long long k = atomicAdd(&ctr->n_try, 1ull);
		if (k < try_cap)
			tries[k] = ((unsigned long long)cls << kTryClsShift) | (unsigned long long)n;
		else
			atomicOr(&ctr->overflow, 2u);
	}
	uint8_t msg[14];
	uint32_t crc = 0;
	for (int b = 0; b < len; b++) {
		uint32_t v = 0;
		for (int k = 0; k < 8; k++)
			v = (v << 1) | (ppm(a, d + 80 * b + 10 * k) ? 1u : 0u);  // demod.c:31-44
		msg[b] = (uint8_t)v;
		if (b < len - 3)
			crc = (crc << 8) ^ c_crc.v[v ^ ((crc >> 16) & 0xFFu)];  // crc.h:36-38
	}
	crc = (crc & 0xFFFFFFu) ^ ((uint32_t)msg[len - 3] << 16 | (uint32_t)msg[len - 2] << 8 |
	                           (uint32_t)msg[len - 1]);  // crc.h:40-42
	uint32_t nfix = 0;
	if (first_seen) {
		// the syndrome is the address overlaid on the parity field: announced at an earlier offset?
		if (crc == 0 || first_seen[crc] >= (unsigned long long)n)
			return;
	} else if (crc != 0) {  // valid.c:52, :74
		if (fix_level <= 0)
			return;
		// extension (repair.h): the syndrome of one or two flipped bits
		const uint32_t e = fix_lookup(fix_tab, crc);
		if (e == kFixNone || !fix_allowed(e, len == 14, fix_level))
			return;
		const uint32_t base = len == 14 ? 0u : 56u;
		const uint32_t b1 = (e >> 8) - base;
		msg[b1 >> 3] ^= (uint8_t)(0x80u >> (b1 & 7u));
		nfix = 1;
		if ((e & 0xFFu) != 0xFFu) {
			const uint32_t b2 = (e & 0xFFu) - base;
			msg[b2 >> 3] ^= (uint8_t)(0x80u >> (b2 & 7u));
			nfix = 2;
		}
	}
	unsigned long long k = atomicAdd(&ctr->n_cand, 1ull);
	if (k >= cand_cap) {
		atomicOr(&ctr->overflow, 1u);
		return;
	}
	CandRec r;
	r.n = (uint64_t)n;
	r.pw = (uint32_t)((p1 + p2) / 4);  // demod.c:127
	r.len = (uint32_t)len;
#pragma unroll
	for (int b = 0; b < 16; b++)
		r.bytes[b] = b < len ? msg[b] : 0;
	r.bytes[15] = (uint8_t)nfix;  // bits repaired (0 unless the extension is on)
	cands[k] = r;
}

// ---- packed 12-bit input -------------------------------------
// The airspy can deliver its 12-bit samples packed (the reference leaves that off, air.c:128-152):
// three 32-bit little-endian words hold eight samples, most significant bits first - the layout of
// libairspy's unpack_samples (airspy.c; not vendored in the reference).  One thread per group:
// 12 bytes in, 16 bytes out, straight into the sample window the scan kernels read.
__global__ void k_unpack12(const uint32_t *__restrict__ in, uint16_t *__restrict__ out, long long groups)
{
	for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < groups;
	     g += (long long)gridDim.x * blockDim.x) {
		const uint32_t w0 = in[3 * g], w1 = in[3 * g + 1], w2 = in[3 * g + 2];
		const uint32_t s0 = w0 >> 20, s1 = (w0 >> 8) & 0xFFFu, s2 = ((w0 & 0xFFu) << 4) | (w1 >> 28);
		const uint32_t s3 = (w1 >> 16) & 0xFFFu, s4 = (w1 >> 4) & 0xFFFu, s5 = ((w1 & 0xFu) << 8) | (w2 >> 24);
		const uint32_t s6 = (w2 >> 12) & 0xFFFu, s7 = w2 & 0xFFFu;
		uint2 *o = reinterpret_cast<uint2 *>(out + 8 * g);  // the window keeps pushes 8-byte aligned
		o[0] = make_uint2(s0 | (s1 << 16), s2 | (s3 << 16));
		o[1] = make_uint2(s4 | (s5 << 16), s6 | (s7 << 16));
	}
}

// ---- launch wrappers -----------------------------------------------------

cudaError_t launch_envelope_generic(const uint16_t *raw, long long u_raw0, unsigned long long origin,
                                    long long j0, long long n, float *amp, int iq, cudaStream_t st)
{
	if (n <= 0)
		return cudaSuccess;
	const int bs = 256;
	k_envelope_generic<<<(unsigned)((n + bs - 1) / bs), bs, 0, st>>>(raw, u_raw0, origin, j0, n, amp, iq);
	return cudaGetLastError();
}

cudaError_t launch_fill_pw(const uint16_t *raw, long long u_raw0, unsigned long long origin, CandRec *cands,
                           const unsigned long long *n_cand_ptr, unsigned long long cap, cudaStream_t st)
{
	k_fill_pw<<<592, 256, 0, st>>>(raw, u_raw0, origin, cands, n_cand_ptr, cap);
	return cudaGetLastError();
}

static int fill_blocks()
{
	static const int n = [] {
		const char *e = getenv("ADSB_B200_CHAIN_BLOCKS");  // development knob, see kernels_resolve.cu
		const int v = e ? atoi(e) : 0;
		return v > 0 ? v : 592;
	}();
	return n;
}

cudaError_t launch_fill_pw_frames(const uint16_t *raw, long long u_raw0, unsigned long long origin, FrameRec *frames,
                                  const unsigned long long *n_ptr, unsigned long long cap, bool wide, cudaStream_t st)
{
	k_fill_pw_frames<<<fill_blocks() * (wide ? 4 : 1), 128, 0, st>>>(raw, u_raw0, origin, frames, n_ptr, cap);
	return cudaGetLastError();
}

cudaError_t launch_scan_generic(const float *amp, long long j_amp0, long long n0, long long count,
                                int accept_df18, int want_tries, CandRec *cands,
                                unsigned long long cand_cap, unsigned long long *tries,
                                unsigned long long try_cap, Counters *ctr, const unsigned long long *fix_tab,
                                int fix_level, const unsigned long long *first_seen, cudaStream_t st)
{
	if (count <= 0)
		return cudaSuccess;
	const int bs = 256;
	k_scan_generic<<<(unsigned)((count + bs - 1) / bs), bs, 0, st>>>(
	    amp, j_amp0, n0, count, accept_df18, want_tries, cands, cand_cap, tries, try_cap, ctr, fix_tab, fix_level,
	    first_seen);
	return cudaGetLastError();
}

cudaError_t launch_unpack12(const void *packed, uint16_t *out, long long n_samples, cudaStream_t st)
{
	const long long groups = n_samples / 8;
	if (groups <= 0)
		return cudaSuccess;
	const int bs = 256;
	const long long want = (groups + bs - 1) / bs;
	k_unpack12<<<(unsigned)(want < 148 * 16 ? want : 148 * 16), bs, 0, st>>>(
	    static_cast<const uint32_t *>(packed), out, groups);
	return cudaGetLastError();
}

}  // namespace adsb
