// resolver.h - the sequential half of demod.c:84-144 and air.c:94-99, applied to the sparse
// list of CRC-valid candidates the GPU returns.
//
// The GPU evaluates every offset independently.  What remains sequential in the reference is
//   * the greedy skip: after an accepted frame the scan resumes behind it (demod.c:128, :134),
//   * the timestamp: ts counts loop iterations, not samples (demod.c:86, :99),
//   * the buffer schedule: deqframe runs whenever >= 40980 envelope samples are queued and
//     leaves the last 1200 unscanned (air.c:94-99, demod.c:89), so the tail of a capture is
//     never decoded.
// All three are O(accepted frames).  The stream path does the first two on the GPU as well and
// feeds stamped frames (feed_frames); the chunk path feeds raw candidates (feed).
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/adsbdec_b200.h"

namespace adsb {

class Resolver {
public:
	explicit Resolver(bool emulate_tail) : emulate_tail_(emulate_tail) {}

	// Candidates in non-decreasing n, complete for all n < scanned_upto.
	void feed(const adsbdec_b200_candidate_t *c, size_t count, uint64_t scanned_upto, uint64_t total_out,
	          std::vector<adsbdec_b200_frame_t> &accepted);
	// The same for frames the GPU has already put through the greedy rule and stamped (kernels_resolve.cu):
	// only the deqframe call schedule is left, which decides how far the scan has come.
	// ok[0..2] += accepted DF11 / DF17 / DF18 frames (valid.c:53, :75), counted while they are copied.
	void feed_frames(const adsbdec_b200_frame_t *f, size_t count, uint64_t scanned_upto, uint64_t total_out,
	                 std::vector<adsbdec_b200_frame_t> &accepted, uint32_t ok[3]);
	// ok (may be null): counters for the stamped frames still accepted here.
	void finish(std::vector<adsbdec_b200_frame_t> &accepted, uint32_t ok[3] = nullptr);

	// In-place form of feed_frames for a caller that keeps ALL stamped frames in one array f[0 .. n_total):
	// f[0 .. n_vis) have been reached by the call schedule before, the rest is new or still in front of it.
	// Returns the new n_vis (the frames stay where they are; nothing is copied).
	size_t advance_frames(const adsbdec_b200_frame_t *f, size_t n_vis, size_t n_total, uint64_t scanned_upto,
	                      uint64_t total_out);
	size_t finish_frames(const adsbdec_b200_frame_t *f, size_t n_vis, size_t n_total);
	// The same on a compact copy of what the schedule looks at: k[i] = f[i].n << 1 | (f[i].len == 14).  The host
	// reads 8 bytes per frame instead of touching every cache line of the 40-byte records the GPU just wrote.
	size_t advance_keys(const unsigned long long *k, size_t n_vis, size_t n_total, uint64_t scanned_upto,
	                    uint64_t total_out);
	size_t finish_keys(const unsigned long long *k, size_t n_vis, size_t n_total);

	// DF11 / DF17 / DF18 frames fed as stamped frames that the call schedule has not reached yet.
	void count_pending(uint32_t out[3]) const;

	// Offsets below this have been visited (or skipped) by the emulated scan.
	uint64_t scan_end() const { return scan_end_; }
	// Where the next deqframe call would start (= absolute index of ampbuff[0]).
	uint64_t buffer_start() const { return base_; }

private:
	void run_calls(std::vector<adsbdec_b200_frame_t> &accepted);
	void scan_to(uint64_t cutoff, std::vector<adsbdec_b200_frame_t> &accepted);
	void scan_keys_to(uint64_t cutoff);
	void flush_stamped(std::vector<adsbdec_b200_frame_t> &accepted);

	bool emulate_tail_;
	bool inplace_ = false;    // advance_frames / finish_frames: consumed frames are not copied anywhere
	uint64_t base_ = 0;       // absolute index of ampbuff[0] at the next deqframe call
	uint64_t drift_ = 0;      // iterations saved by skips so far: sum(skip - 1)
	uint64_t scan_end_ = 0;
	uint64_t scanned_upto_ = 0;
	uint64_t total_out_ = 0;
	std::vector<adsbdec_b200_candidate_t> pending_;  // sorted by n; [head_, size) not yet consumed
	size_t head_ = 0;
	const adsbdec_b200_candidate_t *in_ = nullptr;  // array being fed, consumed in place
	size_t in_n_ = 0, in_i_ = 0;
	std::vector<adsbdec_b200_frame_t> fpending_;    // stamped frames the call schedule has not reached
	size_t fhead_ = 0, fdone_ = 0;                  // [fdone_, fhead_) consumed but not copied out yet
	const adsbdec_b200_frame_t *fin_ = nullptr;
	const unsigned long long *kin_ = nullptr;       // advance_keys / finish_keys: the compact form of fin_
	size_t fin_n_ = 0, fin_i_ = 0, fin_done_ = 0;
	uint32_t *ok_ = nullptr;  // where scan_to counts the stamped frames it accepts
};

}  // namespace adsb
