// repair.h - opt-in 1-/2-bit error repair by syndrome lookup.
//
// EXTENSION, off by default: the reference accepts a frame only if its CRC-24 syndrome is zero
// (valid.c:52, :74) and has no repair; with repair level 0 nothing here runs and
// the output is the reference's.  With level >= 1 a gate-passing offset whose syndrome is that of a
// single flipped bit is corrected and accepted; with level 2 also two flipped bits of a 112-bit
// frame.  The five DF bits are never touched (they chose the frame length and the gate), and a
// 56-bit frame is only ever repaired by one bit - the rules the dump1090 family uses.  The CRC is
// linear, so the syndrome of an error pattern is the XOR of its bits' syndromes, and with the Mode S
// generator every pattern of <= 2 bits in 112 has its own syndrome: one hash probe finds it.
//
// Bit positions count from the first bit of a 112-bit frame; a 56-bit frame is positions 56..111
// (its bits see the same powers of x as the tail of a long frame).
#pragma once
#include <stdint.h>

#include "common.h"

namespace adsb {

constexpr int kFixSlots = 1 << 14;          // 6328 patterns -> 39 % load
constexpr uint32_t kFixNone = 0xFFFFFFFFu;

// entry: bit 63 used | syndrome << 16 | first bit << 8 | second bit (0xFF = none)
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t fix_hash(uint32_t syndrome) { return (syndrome * 0x9E3779B1u) >> 18; }

// (first << 8) | second of the error pattern with this syndrome, or kFixNone.
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t fix_lookup(const unsigned long long *tab, uint32_t syndrome)
{
	uint32_t h = fix_hash(syndrome);
	for (;;) {
		const unsigned long long e = tab[h];
		if (!e)
			return kFixNone;
		if (((uint32_t)(e >> 16) & 0xFFFFFFu) == syndrome)
			return (uint32_t)e & 0xFFFFu;
		h = (h + 1) & (kFixSlots - 1);
	}
}

// May this pattern be applied?  level 1: one bit; level 2: two bits, long frames only.
#ifdef __CUDACC__
__host__ __device__
#endif
inline bool fix_allowed(uint32_t pattern, bool is_long, int level)
{
	const uint32_t p1 = pattern >> 8, p2 = pattern & 0xFFu;
	const uint32_t first = is_long ? 5u : 56u + 5u;  // never the DF field
	if (p1 < first)
		return false;
	if (p2 != 0xFFu && (!is_long || level < 2 || p2 < first))
		return false;
	return level >= 1;
}

// Host: fills tab[kFixSlots].  syn(b) = x^(111-b) mod G, the syndrome of bit b alone.
inline void fix_build_table(unsigned long long *tab)
{
	uint32_t syn[112];
	for (int b = 0; b < 112; b++) {
		uint32_t r = 1;
		for (int i = 0; i < 111 - b; i++)
			r = (r & 0x800000u) ? ((r << 1) ^ kCrcPoly) & 0xFFFFFFu : (r << 1);
		syn[b] = r;
	}
	for (int i = 0; i < kFixSlots; i++)
		tab[i] = 0;
	auto put = [&](uint32_t s, uint32_t p1, uint32_t p2) {
		uint32_t h = fix_hash(s);
		while (tab[h])
			h = (h + 1) & (kFixSlots - 1);
		tab[h] = (1ull << 63) | ((unsigned long long)s << 16) | (p1 << 8) | p2;
	};
	for (uint32_t a = 0; a < 112; a++) {
		put(syn[a], a, 0xFFu);
		for (uint32_t b = a + 1; b < 112; b++)
			put(syn[a] ^ syn[b], a, b);
	}
}

}  // namespace adsb
