// kernels_resolve.cu - the sequential half of demod.c:84-144, done on the GPU over the sparse
// candidate list of one scan job, so that only ACCEPTED frames travel back to the host.
//
//   1. order:   counting sort of the candidates by offset (4096-offset buckets: histogram, scan, the
//               bucket's offsets as keys, rank among them = final place); frame bits are repacked to bytes
//               on the way;
//   2. greedy:  demod.c:128/:134 - after an accepted frame the scan resumes behind it.  A candidate
//               that starts >= 1200 samples after its predecessor is accepted whatever happened
//               before (no frame is longer than 1200), so the list falls apart into independent
//               segments at such gaps; one thread walks each segment;
//   3. compact: a bucket-level prefix sum gives every accepted candidate its slot; they are written
//               out in offset order.
//
// The scan position that crosses jobs (end of the last accepted frame) lives in device memory and
// is advanced in stream order, so jobs can be pipelined without the host in the loop.  What stays
// on the host (resolver.cpp, fed with ACCEPTED frames only) is O(accepted): the iteration-count
// timestamp (demod.c:99) and the deqframe call schedule (air.c:94-99) that decides where the scan ends.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/adsbdec_b200.h"
#include "common.h"
#include "kernels.h"

namespace adsb {

namespace {

constexpr int kBShift = 12;  // bucket = 4096 offsets

// Footprint of the chain's kernels.  They run on their own stream beside k_fused of the NEXT job, which holds
// 640 threads x 96 registers on every SM: 4096 registers, 62 KB of shared memory and 44 warp slots are left
// there, i.e. room for exactly one block of 128 threads x 32 registers.  Built that way (ADSB_CHAIN_SMALL)
// the chain needs no SMs of its own; otherwise k_fused leaves it some (api.cu: kResolveSMs).
#ifndef ADSB_CHAIN_SMALL
#define ADSB_CHAIN_SMALL 1
#endif
#if ADSB_CHAIN_SMALL
#define ADSB_CHAIN_BOUNDS __launch_bounds__(128, 16)
constexpr int kChainThreads = 128;
constexpr int kScanThreads = 128;
#else
#define ADSB_CHAIN_BOUNDS
constexpr int kChainThreads = 256;
constexpr int kScanThreads = 1024;
#endif
// Blocks of the grid-stride kernels (development knob ADSB_B200_CHAIN_BLOCKS).
static int chain_blocks()
{
	static const int n = [] {
		const char *e = getenv("ADSB_B200_CHAIN_BLOCKS");
		const int v = e ? atoi(e) : 0;
		return v > 0 ? v : 296 * 256 / kChainThreads;
	}();
	return n;
}
// (a job whose chain will have the GPU to itself - the last one of a resident push - gets four times the blocks:
// the kernels wait for memory, and the one-block-per-SM slot beside k_fused is not what limits them then)
#define kChainBlocks (chain_blocks() * (wide ? 4 : 1))

__device__ __forceinline__ unsigned long long job_count(const Counters *ctr, unsigned long long cap)
{
	const unsigned long long n = ctr->n_cand;
	return n < cap ? n : cap;
}

__global__ void ADSB_CHAIN_BOUNDS k_hist(const CandRec *__restrict__ c, const Counters *ctr, unsigned long long cap, long long n0,
                       unsigned int *cnt)
{
	const unsigned long long m = job_count(ctr, cap);
	for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < m;
	     i += (unsigned long long)gridDim.x * blockDim.x)
		atomicAdd(&cnt[(unsigned long long)((long long)c[i].n - n0) >> kBShift], 1u);
}

// Exclusive scan of bucket counts (nb entries in, nb+1 out) in two launches: k_scan_sums leaves the total of
// every 4096-entry tile, k_scan_apply adds up the tiles in front of its own and scans it (one coalesced 16-byte
// load per thread, shuffles inside the warps, two block barriers).  A single block walking the whole array
// took 17-23 us of the 130 us a job's ordering chain needs; a thread-contiguous single-block variant was
// slower still (uncoalesced loads on one SM's L1).  DUAL scans two arrays at once (accepted frames and their
// timestamp drift per bucket) packed into one 64-bit value: both totals stay below 2^32 (a job holds
// < 2^29 offsets).
template <bool DUAL>
__device__ __forceinline__ void scan_load4(const unsigned int *__restrict__ in_a, const unsigned int *__restrict__ in_b,
                                           int i0, int nb, unsigned long long v[4])
{
	typedef unsigned long long T;
	if (i0 + 4 <= nb) {
		const uint4 a = *reinterpret_cast<const uint4 *>(in_a + i0);
		v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
		if (DUAL) {
			const uint4 b = *reinterpret_cast<const uint4 *>(in_b + i0);
			v[0] |= (T)b.x << 32; v[1] |= (T)b.y << 32; v[2] |= (T)b.z << 32; v[3] |= (T)b.w << 32;
		}
	} else {
#pragma unroll
		for (int k = 0; k < 4; k++)
			v[k] = i0 + k < nb ? (T)in_a[i0 + k] | (DUAL ? (T)in_b[i0 + k] << 32 : 0ull) : 0ull;
	}
}

constexpr int kScanTile = 4 * kScanThreads;

// Inclusive scan of one value per thread over the block; returns the block total in `total`.
__device__ __forceinline__ unsigned long long block_scan_incl(unsigned long long x, unsigned long long *wsum,
                                                              unsigned long long &total)
{
	typedef unsigned long long T;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const T y = __shfl_up_sync(0xFFFFFFFFu, x, d);
		if (lane >= d)
			x += y;
	}
	if (lane == 31)
		wsum[w] = x;
	__syncthreads();
	if (w == 0) {
		T sv = lane < kScanThreads / 32 ? wsum[lane] : 0ull;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const T y = __shfl_up_sync(0xFFFFFFFFu, sv, d);
			if (lane >= d)
				sv += y;
		}
		wsum[lane] = sv;
	}
	__syncthreads();
	total = wsum[31];
	return x + (w ? wsum[w - 1] : 0ull);
}

template <bool DUAL>
__global__ void __launch_bounds__(kScanThreads)
k_scan_sums(const unsigned int *__restrict__ in_a, const unsigned int *__restrict__ in_b, int nb,
            unsigned long long *__restrict__ part)
{
	__shared__ unsigned long long wsum[32];
	unsigned long long v[4], total;
	scan_load4<DUAL>(in_a, in_b, blockIdx.x * kScanTile + 4 * threadIdx.x, nb, v);
	block_scan_incl(v[0] + v[1] + v[2] + v[3], wsum, total);
	if (threadIdx.x == 0)
		part[blockIdx.x] = total;
}

template <bool DUAL>
__global__ void __launch_bounds__(kScanThreads)
k_scan_apply(const unsigned int *__restrict__ in_a, const unsigned int *__restrict__ in_b,
             unsigned int *__restrict__ out_a, unsigned int *__restrict__ out_b, int nb,
             const unsigned long long *__restrict__ part)
{
	typedef unsigned long long T;
	__shared__ T wsum[32];
	T base = 0;  // total of the tiles in front of this one (a few dozen at most)
	for (int k = 0; k < (int)blockIdx.x; k++)
		base += part[k];
	const int i0 = blockIdx.x * kScanTile + 4 * threadIdx.x;
	T v[4], total;
	scan_load4<DUAL>(in_a, in_b, i0, nb, v);
	const T mine = v[0] + v[1] + v[2] + v[3];
	T pre = base + block_scan_incl(mine, wsum, total) - mine;
	unsigned int oa[4], ob[4];
#pragma unroll
	for (int k = 0; k < 4; k++) {
		oa[k] = (unsigned int)pre;
		ob[k] = (unsigned int)(pre >> 32);
		pre += v[k];
	}
	if (i0 + 4 <= nb) {
		*reinterpret_cast<uint4 *>(out_a + i0) = make_uint4(oa[0], oa[1], oa[2], oa[3]);
		if (DUAL)
			*reinterpret_cast<uint4 *>(out_b + i0) = make_uint4(ob[0], ob[1], ob[2], ob[3]);
	} else {
#pragma unroll
		for (int k = 0; k < 4; k++)
			if (i0 + k < nb) {
				out_a[i0 + k] = oa[k];
				if (DUAL)
					out_b[i0 + k] = ob[k];
			}
	}
	if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
		out_a[nb] = (unsigned int)(base + total);
		if (DUAL)
			out_b[nb] = (unsigned int)((base + total) >> 32);
	}
}

template <bool DUAL>
void launch_scan(const unsigned int *in_a, const unsigned int *in_b, unsigned int *out_a, unsigned int *out_b, int nb,
                 unsigned long long *part, cudaStream_t st)
{
	const int tiles = (nb + kScanTile - 1) / kScanTile;
	k_scan_sums<DUAL><<<tiles, kScanThreads, 0, st>>>(in_a, in_b, nb, part);
	k_scan_apply<DUAL><<<tiles, kScanThreads, 0, st>>>(in_a, in_b, out_a, out_b, nb, part);
}

__device__ __forceinline__ void repack(CandRec &o, const CandRec &c)
{
	o = c;
	if (!(c.len & kCandNibbleWords))
		return;
	// four 28-bit words (MSB first) -> frame bytes
	o.len = c.len & ~kCandNibbleWords;
	uint32_t w[4];
	memcpy(w, c.bytes, 16);
	const uint32_t nfix = w[0] >> 28;  // bits repaired (repair.h; 0 in reference mode)
#pragma unroll
	for (int k = 0; k < 4; k++)
		w[k] &= 0x0FFFFFFFu;
	unsigned long long hi, lo;  // 112-bit frame as hi:lo (hi holds the top 48 bits) / 56-bit frame in lo
	int nbits;
	if (o.len == 14) {
		// f = w0<<84 | w1<<56 | w2<<28 | w3
		lo = ((unsigned long long)w[3]) | ((unsigned long long)w[2] << 28) | ((unsigned long long)w[1] << 56);
		hi = ((unsigned long long)w[1] >> 8) | ((unsigned long long)w[0] << 20);
		nbits = 112;
	} else {
		lo = ((unsigned long long)w[0] << 28) | w[1];
		hi = 0;
		nbits = 56;
	}
#pragma unroll
	for (int k = 0; k < 16; k++) {
		uint8_t v = 0;
		if (k < (int)o.len) {
			const int sh = nbits - 8 - 8 * k;  // bit position of this byte in the frame
			v = sh >= 64 ? (uint8_t)(hi >> (sh - 64))
			             : (sh > 56 ? (uint8_t)((lo >> sh) | (hi << (64 - sh))) : (uint8_t)(lo >> sh));
		}
		o.bytes[k] = v;
	}
	o.bytes[15] = (uint8_t)nfix;
}

// Ordering inside a bucket without moving records twice: the offsets of a bucket's candidates are dropped
// into its slice of `keys` in whatever order the atomics fall (the counts of k_hist are used up again as
// cursors and end at zero, ready for the next job), then every candidate counts the keys of its bucket
// that are smaller than its own - a handful - and goes straight to its final place.  No two candidates
// share an offset.
__global__ void ADSB_CHAIN_BOUNDS k_scatter_keys(const CandRec *__restrict__ c, const Counters *ctr, unsigned long long cap, long long n0,
                               const unsigned int *__restrict__ off, unsigned int *cnt, unsigned int *__restrict__ keys)
{
	const unsigned long long m = job_count(ctr, cap);
	for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < m;
	     i += (unsigned long long)gridDim.x * blockDim.x) {
		const unsigned long long rel = (unsigned long long)((long long)c[i].n - n0);
		const unsigned long long b = rel >> kBShift;
		keys[off[b] + atomicSub(&cnt[b], 1u) - 1u] = (unsigned int)rel;
	}
}

__global__ void ADSB_CHAIN_BOUNDS k_rank_move(const CandRec *__restrict__ c, const Counters *ctr, unsigned long long cap, long long n0,
                            const unsigned int *__restrict__ off, const unsigned int *__restrict__ keys,
                            CandRec *__restrict__ sorted)
{
	const unsigned long long m = job_count(ctr, cap);
	for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < m;
	     i += (unsigned long long)gridDim.x * blockDim.x) {
		const unsigned long long rel = (unsigned long long)((long long)c[i].n - n0);
		const unsigned long long b = rel >> kBShift;
		const unsigned int lo = off[b], hi = off[b + 1], mine = (unsigned int)rel;
		unsigned int rank = 0;
		for (unsigned int q = lo; q < hi; q++)
			rank += keys[q] < mine;
		CandRec r;
		repack(r, c[i]);
		sorted[lo + rank] = r;
	}
}

__device__ __forceinline__ unsigned long long skip_of(uint32_t len) { return len == 7 ? kSkipShort : kSkipLong; }

// Copies the scan position into the job's slot: the greedy kernel reads it from there while its
// last segment writes the new one, and the statistics pass needs the "before" value later.
__global__ void k_snapshot(const ResolveState *st, SlotResolve *slot)
{
	slot->span0_start = st->carry_start;
	slot->span0_end = st->carry_pos;
	slot->drift0 = st->drift;
	slot->ok[0] = slot->ok[1] = slot->ok[2] = 0;  // k_compact counts into them
}

// Greedy acceptance.  Thread i walks the segment that starts at candidate i (if one does): a few candidates in
// ordinary traffic.  A segment that is still open after kSoloSteps candidates - saturated traffic chains a whole
// job into ONE segment - is finished by its warp together: 32 candidates per coalesced load, the sequential rule
// applied to them out of registers (every lane runs the same few instructions on shuffled values), instead of
// one thread taking a trip to the L2 per candidate (30 ns instead of 300 per candidate).
constexpr int kSoloSteps = 48;

__global__ void ADSB_CHAIN_BOUNDS k_greedy(const CandRec *__restrict__ sorted, const unsigned int *__restrict__ off, int nb,
                         unsigned char *__restrict__ accept, ResolveState *st, const SlotResolve *slot)
{
	const unsigned long long m = off[nb];
	const unsigned int lane = threadIdx.x & 31u;
	// whole warps iterate together (the ballots below need every lane)
	for (unsigned long long i0 = blockIdx.x * (unsigned long long)blockDim.x + (threadIdx.x & ~31u); i0 < m;
	     i0 += (unsigned long long)gridDim.x * blockDim.x) {
		const unsigned long long i = i0 + lane;
		bool open = false;  // this lane's segment is not finished yet
		unsigned long long pos = 0, start = 0, j = i, prev = 0;
		if (i < m && (i == 0 || sorted[i].n - sorted[i - 1].n >= (unsigned long long)kLookahead)) {
			// demod.c:128/:134: idx = lidx.  The first segment of a job inherits the scan position; any
			// other segment starts >= 1200 behind a candidate of this job, i.e. behind every earlier frame.
			if (i == 0) {
				pos = slot->span0_end;
				start = slot->span0_start;
			}
			open = true;
			for (int steps = 0; steps < kSoloSteps; steps++) {
				const CandRec &c = sorted[j];
				const bool ok = c.n >= pos;
				accept[j] = ok ? 1 : 0;
				if (ok) {
					start = c.n;
					pos = c.n + skip_of(c.len);
				}
				prev = c.n;
				j++;
				if (j >= m || sorted[j].n - prev >= (unsigned long long)kLookahead) {
					open = false;
					break;
				}
			}
			if (!open && j >= m) {
				// the last segment leaves the scan position for the next job (ordered by the stream)
				st->carry_pos = pos;
				st->carry_start = start;
			}
		}
		// long segments, one after the other, by the whole warp
		for (unsigned int todo = __ballot_sync(0xFFFFFFFFu, open); todo; todo &= todo - 1u) {
			const int l = __ffs(todo) - 1;
			unsigned long long J = __shfl_sync(0xFFFFFFFFu, j, l), POS = __shfl_sync(0xFFFFFFFFu, pos, l);
			unsigned long long START = __shfl_sync(0xFFFFFFFFu, start, l), PREV = __shfl_sync(0xFFFFFFFFu, prev, l);
			bool done = false;
			while (!done) {
				const unsigned long long idx = J + lane;
				unsigned long long n_l = 0;
				unsigned int len_l = 0;
				if (idx < m) {
					n_l = sorted[idx].n;
					len_l = sorted[idx].len;
				}
				unsigned int acc = 0, cnt = 0;
				for (int t = 0; t < 32; t++) {  // (uniform: every lane sees the same shuffled values)
					if (J + (unsigned int)t >= m) {
						done = true;
						break;
					}
					const unsigned long long nt = __shfl_sync(0xFFFFFFFFu, n_l, t);
					const unsigned int lt = __shfl_sync(0xFFFFFFFFu, len_l, t);
					if (nt - PREV >= (unsigned long long)kLookahead) {
						done = true;
						break;
					}
					if (nt >= POS) {
						acc |= 1u << t;
						START = nt;
						POS = nt + skip_of(lt);
					}
					PREV = nt;
					cnt++;
				}
				if (lane < cnt)
					accept[J + lane] = (acc >> lane) & 1u;
				J += cnt;
			}
			if (J >= m && lane == 0) {
				st->carry_pos = POS;
				st->carry_start = START;
			}
		}
	}
}

__global__ void ADSB_CHAIN_BOUNDS k_bucket_stats(const CandRec *__restrict__ sorted, const unsigned char *__restrict__ accept,
                               const unsigned int *__restrict__ off, int nb, unsigned int *bacc, unsigned int *bdrift)
{
	for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
		unsigned int a = 0, d = 0;
		for (unsigned int i = off[b]; i < off[b + 1]; i++)
			if (accept[i]) {
				a++;
				d += (unsigned int)skip_of(sorted[i].len) - 1u;
			}
		bacc[b] = a;
		bdrift[b] = d;
	}
}

// Writes the accepted frames, in offset order, as the API hands them out.  ts counts the loop
// iterations of deqframe (demod.c:86, :99): every accepted frame before this one saved skip-1 of them.
__global__ void ADSB_CHAIN_BOUNDS k_compact(const CandRec *__restrict__ sorted, const unsigned char *__restrict__ accept,
                          const unsigned int *__restrict__ off, int nb, const unsigned int *__restrict__ bacc_x,
                          const unsigned int *__restrict__ bdrift_x, FrameRec *__restrict__ out,
                          unsigned long long *__restrict__ keys_out, unsigned long long out_cap, ResolveState *st,
                          SlotResolve *slot)
{
	const unsigned long long drift0 = slot->drift0;
	unsigned int ok[3] = {0, 0, 0};  // accepted frames per class (valid.c:53, :75), one atomic per warp at the end
	for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
		unsigned long long p = bacc_x[b];
		unsigned long long drift = drift0 + bdrift_x[b];
		for (unsigned int i = off[b]; i < off[b + 1]; i++)
			if (accept[i]) {
				const CandRec c = sorted[i];
				if (p < out_cap) {
					// 40 bytes as five 8-byte words
					unsigned long long w[5];
					w[0] = c.n + 1ull - drift;
					w[1] = c.n;
					w[2] = (unsigned long long)c.pw | ((unsigned long long)c.len << 32);
					memcpy(&w[3], c.bytes, 8);
					unsigned long long t = 0;
					memcpy(&t, c.bytes + 8, 6);
					if (c.len == 7) {  // a short frame leaves bytes 7..13 zero
						w[3] &= 0x00FFFFFFFFFFFFFFull;
						t = 0;
					}
					w[4] = t | ((unsigned long long)c.bytes[15] << 48);  // pad_[0] = bits repaired
					unsigned long long *o = reinterpret_cast<unsigned long long *>(&out[p]);
#pragma unroll
					for (int k = 0; k < 5; k++)
						o[k] = w[k];
					// what the host's call schedule (air.c:94-99) looks at, 8 bytes instead of 40: start and length
					keys_out[p] = (c.n << 1) | (c.len == 7 ? 0ull : 1ull);
				}
				drift += skip_of(c.len) - 1ull;
				p++;
				const int df = c.bytes[0] >> 3;
				if (ok_class(df) >= 0)
					ok[ok_class(df)]++;
			}
	}
#pragma unroll
	for (int c = 0; c < 3; c++) {
		const unsigned int w = __reduce_add_sync(0xFFFFFFFFu, ok[c]);
		if ((threadIdx.x & 31) == 0 && w)
			atomicAdd(&slot->ok[c], w);
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		slot->n_sorted = off[nb];
		slot->n_accept = bacc_x[nb];
		st->drift = drift0 + bdrift_x[nb];  // read by the next job's snapshot only
	}
}

// Try statistics (valid.c:46, :68 as the sequential scan sees them) against the accepted frames
// of this job, which are still on the device.  The bucket prefix counts of the compaction tell
// where the frames near a try sit, so no search is needed.
__global__ void ADSB_CHAIN_BOUNDS k_count_tries_frames(const unsigned long long *__restrict__ tries,
                                     const unsigned long long *__restrict__ n_tries_ptr, unsigned long long cap,
                                     const FrameRec *__restrict__ frames, const SlotResolve *slot,
                                     unsigned long long frames_cap, const unsigned int *__restrict__ bacc_x,
                                     long long n0, int nb, unsigned long long scan_end,
                                     unsigned long long *pend_out, unsigned long long *n_pend_out,
                                     unsigned long long pend_cap, StreamStats *st)
{
	unsigned long long n_tries = *n_tries_ptr;
	if (n_tries > cap)
		n_tries = cap;
	const unsigned long long nf = slot->n_accept < frames_cap ? slot->n_accept : frames_cap;
	const unsigned long long s0 = slot->span0_start, e0 = slot->span0_end;
	unsigned int mine[3] = {0, 0, 0};  // visited tries per class, added to the global counters once per warp
	for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n_tries;
	     i += (unsigned long long)gridDim.x * blockDim.x) {
		const unsigned long long e = tries[i];
		const unsigned long long n = e & kTryMask;
		const uint32_t cls = (uint32_t)(e >> kTryClsShift);
		// Is n strictly inside an accepted frame?  Only a frame that starts in n's bucket or the one
		// before can reach it (a frame spans <= 1200 < 4096 offsets); before this job's range only the
		// last frame accepted earlier can.
		bool inside = n > s0 && n < e0;
		const long long rel = (long long)n - n0;
		if (nb > 0 && rel >= 0) {
			long long b = rel >> kBShift;
			if (b > nb - 1) b = nb - 1;
			unsigned long long lo = bacc_x[b > 0 ? b - 1 : 0], hi = bacc_x[b + 1];
			if (hi > nf) hi = nf;
			for (unsigned long long q = lo; q < hi; q++) {
				const unsigned long long fs = frames[q].n;
				if (fs >= n)
					break;
				if (n < fs + skip_of((uint32_t)frames[q].len))
					inside = true;
			}
		}
		if (inside)
			continue;  // skipped by `idx = lidx` (demod.c:128, :134), whenever the scan gets there
		if (n >= scan_end) {
			// not reached yet by the deqframe schedule: decide in a later pass.  It has been cleared
			// against this job's frames already, and no later frame can start before it.
			const unsigned long long k = atomicAdd(n_pend_out, 1ull);
			if (k < pend_cap)
				pend_out[k] = e;
			else
				st->pend_overflow = 1;
			continue;
		}
		mine[0] += cls == 0;
		mine[1] += cls == 1;
		mine[2] += cls == 2;
	}
#pragma unroll
	for (int c = 0; c < 3; c++) {
		const unsigned int w = __reduce_add_sync(0xFFFFFFFFu, mine[c]);
		if ((threadIdx.x & 31) == 0 && w)
			atomicAdd(&st->stat_try[c], w);
	}
}

}  // namespace

int resolve_bucket_count(long long n0, long long n1) { return (int)(((n1 - n0) >> kBShift) + 2); }

// Sort only (chunk mode): rb.sorted = candidates in offset order, bytes repacked, count in off[nb].
cudaError_t launch_sort(const ResolveBuffers &rb, const CandRec *cands, const Counters *ctr, unsigned long long cap,
                        long long n0, long long n1, bool wide, cudaStream_t st)
{
	const int nb = resolve_bucket_count(n0, n1);
	// (rb.cnt is all zero here: zeroed when it was allocated, counted back down to zero by every job)
	k_hist<<<kChainBlocks, kChainThreads, 0, st>>>(cands, ctr, cap, n0, rb.cnt);
	launch_scan<false>(rb.cnt, nullptr, rb.off, nullptr, nb, rb.part, st);
	k_scatter_keys<<<kChainBlocks, kChainThreads, 0, st>>>(cands, ctr, cap, n0, rb.off, rb.cnt, rb.keys);
	k_rank_move<<<kChainBlocks, kChainThreads, 0, st>>>(cands, ctr, cap, n0, rb.off, rb.keys, rb.sorted);
	return cudaGetLastError();
}

// Sort + greedy + compaction: `out` receives the accepted candidates in order, slot->n_accept their number.
cudaError_t launch_resolve(const ResolveBuffers &rb, const CandRec *cands, const Counters *ctr, unsigned long long cap,
                           long long n0, long long n1, ResolveState *state, SlotResolve *slot, FrameRec *out,
                           unsigned long long *keys_out, unsigned long long out_cap, bool wide, cudaStream_t st)
{
	const int nb = resolve_bucket_count(n0, n1);
	cudaError_t e = launch_sort(rb, cands, ctr, cap, n0, n1, wide, st);
	if (e != cudaSuccess) return e;
	k_snapshot<<<1, 1, 0, st>>>(state, slot);
	k_greedy<<<kChainBlocks, kChainThreads, 0, st>>>(rb.sorted, rb.off, nb, rb.accept, state, slot);
	k_bucket_stats<<<(nb + 127) / 128, 128, 0, st>>>(rb.sorted, rb.accept, rb.off, nb, rb.bacc, rb.bdrift);
	launch_scan<true>(rb.bacc, rb.bdrift, rb.bacc_x, rb.bdrift_x, nb, rb.part, st);
	k_compact<<<(nb + 127) / 128, 128, 0, st>>>(rb.sorted, rb.accept, rb.off, nb, rb.bacc_x, rb.bdrift_x, out, keys_out,
	                                             out_cap, state, slot);
	return cudaGetLastError();
}

__global__ void k_set_state(ResolveState *st, unsigned long long carry_start, unsigned long long carry_pos,
                            unsigned long long drift)
{
	st->carry_start = carry_start;
	st->carry_pos = carry_pos;
	st->drift = drift;
}

cudaError_t launch_set_state(ResolveState *state, unsigned long long carry_start, unsigned long long carry_pos,
                             unsigned long long drift, cudaStream_t st)
{
	k_set_state<<<1, 1, 0, st>>>(state, carry_start, carry_pos, drift);
	return cudaGetLastError();
}

cudaError_t launch_count_tries_frames(const unsigned long long *tries, const unsigned long long *n_tries_ptr,
                                      unsigned long long cap, const FrameRec *frames, const SlotResolve *slot,
                                      unsigned long long frames_cap, const ResolveBuffers &rb, long long n0,
                                      long long n1, unsigned long long scan_end, unsigned long long *pend_out,
                                      unsigned long long *n_pend_out, unsigned long long pend_cap, StreamStats *stt,
                                      cudaStream_t st)
{
	const int nb = n1 > n0 ? resolve_bucket_count(n0, n1) : 0;
	const bool wide = false;
	k_count_tries_frames<<<4 * kChainBlocks, kChainThreads, 0, st>>>(tries, n_tries_ptr, cap, frames, slot, frames_cap, rb.bacc_x, n0, nb,
	                                         scan_end, pend_out, n_pend_out, pend_cap, stt);
	return cudaGetLastError();
}

}  // namespace adsb
