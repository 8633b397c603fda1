// common.h - constants and device/host record layouts shared by the kernels and the host side.
//
// Reference citations are file:line into TLeconte/adsbdec.
#pragma once
#include <stdint.h>

namespace adsb {

constexpr int kPulse = 5;                 // adsbdec.h:1  PULSEW (0.5 us at 10 MS/s)
constexpr int kLookahead = 240 * kPulse;  // adsbdec.h:3  DECOFFSET = 1200 envelope samples
constexpr int kTaps = 14;                 // air.c:32     DSFLTLEN
constexpr int kAmpFlush = 8196 * kPulse;  // air.c:47     APBUFFSZ = 40980
constexpr int kReadSamples = 1 << 20;     // air.c:218    IQBUFFSZ
constexpr int kSkipShort = 80 + 80 * 7;   // demod.c:110,119-123: idx jumps to the end of the frame
constexpr int kSkipLong = 80 + 80 * 14;

// DF classes in print_stats order (valid.c:88-97)
enum : uint32_t { kClsDF11 = 0, kClsDF17 = 1, kClsDF18 = 2 };
// Index of a frame's DF in the Ok counters (valid.c:53, :75), -1 for the frames of the address/parity
// extension (DF4/5/20/21), which the reference's statistics do not know.
__host__ __device__ inline int ok_class(int df) { return df == 11 ? 0 : df == 17 ? 1 : df == 18 ? 2 : -1; }

// One CRC-valid candidate as written by the kernels (32 bytes).
struct alignas(16) CandRec {
	uint64_t n;        // local envelope-sample offset of the preamble start
	uint32_t pw;       // (p1+p2)/4, demod.c:127
	uint32_t len;      // 7 or 14
	uint8_t bytes[16]; // first len bytes
};
static_assert(sizeof(CandRec) == 32, "CandRec layout");

// One accepted frame as the host API hands it out: same layout as adsbdec_b200_frame_t.
struct FrameRec {
	uint64_t ts;       // demod.c:99 iteration counter
	uint64_t n;
	uint32_t pw;
	int32_t len;
	uint8_t bytes[14];
	uint8_t pad_[2];
};
static_assert(sizeof(FrameRec) == 40, "FrameRec must match adsbdec_b200_frame_t");
// Flag in CandRec::len: `bytes` holds four 28-bit words (frame bits 28k..28k+27, MSB first in bits
// 27..0 of word k) instead of the frame bytes; the host repacks them.
constexpr uint32_t kCandNibbleWords = 0x100u;

// A "try" (offset that reached validShort/validLong, valid.c:46,68): class in the top 2 bits.
constexpr int kTryClsShift = 62;
constexpr uint64_t kTryMask = (1ull << kTryClsShift) - 1;

// Device-side counters of one scan job (one 64-byte block, reset per job).
struct Counters {
	unsigned long long n_cand;      // entries in the job's candidate list
	unsigned long long n_try;       // entries in the job's try list
	unsigned int overflow;          // bit0: candidate list full, bit1: try list full
	unsigned int wide;              // a sample outside 0..4095 was seen by the fused kernel
	unsigned long long try_mark;    // n_try when the current fused range began: k_slice starts here
	unsigned long long cand_mark;   // n_cand when the current range began (address/parity extension)
	unsigned long long n_ap;        // entries in the range's list of address/parity tries (extension)
	unsigned int pad_[4];
};
static_assert(sizeof(Counters) == 64, "Counters layout");

// Device-side per-stream statistics (valid.c:30-31 as seen by the sequential scan).
struct StreamStats {
	unsigned int stat_try[3];       // visited tries per class (accumulates over the stream)
	unsigned int pend_overflow;
	unsigned long long n_pend[2];   // entries in the two pending-try lists (tries beyond the scan end)
	unsigned int pad_[8];
};
static_assert(sizeof(StreamStats) == 64, "StreamStats layout");

// Mode S CRC-24 generator (crc.h:1-34 is its byte-wise table).
constexpr uint32_t kCrcPoly = 0xFFF409u;

constexpr uint32_t crc24_table_entry(uint32_t b)
{
	uint32_t r = b << 16;
	for (int k = 0; k < 8; k++)
		r = (r & 0x800000u) ? ((r << 1) ^ kCrcPoly) : (r << 1);
	return r & 0xFFFFFFu;
}

struct CrcTable {
	uint32_t v[256];
};
constexpr CrcTable make_crc_table()
{
	CrcTable t{};
	for (uint32_t b = 0; b < 256; b++)
		t.v[b] = crc24_table_entry(b);
	return t;
}

// air.c:36-45 - one period of the triangular low-pass (decimal literals converted to float
// exactly as the reference's initialiser does).
#define ADSB_TAPS_INIT                                                                     \
	{ 0.012627f, 0.025254f, 0.037881f, 0.050508f, 0.063135f, 0.075761f, 0.088388f,        \
	  0.088388f, 0.075761f, 0.063135f, 0.050508f, 0.037881f, 0.025254f, 0.012627f }

}  // namespace adsb
