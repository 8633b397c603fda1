// resolver.cpp - see resolver.h.
#include "resolver.h"

#include <string.h>

#include <algorithm>

#include "common.h"

namespace adsb {

void Resolver::scan_to(uint64_t cutoff, std::vector<adsbdec_b200_frame_t> &accepted)
{
	// One deqframe(ampbuff, len) call whose loop bound len-DECOFFSET is `cutoff` in absolute
	// offsets (demod.c:89).  idx starts at ampbuff[0] = base_.  Candidates come from the few left
	// over from earlier feeds (pending_) and then, in place, from the array being fed.
	uint64_t pos = base_;
	// stamped frames: everything below the cutoff is accepted as it is
	for (int src = 0; src < 2; src++) {
		const adsbdec_b200_frame_t *f = src == 0 ? fpending_.data() : fin_;
		size_t &i = src == 0 ? fhead_ : fin_i_;
		const size_t n = src == 0 ? fpending_.size() : fin_n_, i0 = i;
		while (i < n && f[i].n < cutoff) {
			if (ok_) {
				const int df = f[i].bytes[0] >> 3;
				ok_[df == 11 ? 0 : df == 17 ? 1 : 2]++;
			}
			i++;
		}
		if (i > i0) {
			accepted.insert(accepted.end(), f + i0, f + i);
			pos = f[i - 1].n + (f[i - 1].len == 7 ? kSkipShort : kSkipLong);
		}
		if (i < n)
			break;  // the pending frames come before the fed ones
	}
	for (;;) {
		const adsbdec_b200_candidate_t *c;
		if (head_ < pending_.size())
			c = &pending_[head_];
		else if (in_i_ < in_n_)
			c = &in_[in_i_];
		else
			break;
		if (c->n >= cutoff)
			break;
		if (head_ < pending_.size())
			head_++;
		else
			in_i_++;
		if (c->n < pos)
			continue;  // inside the frame just accepted: never evaluated (demod.c:128, :134)
		const uint64_t skip = c->len == 7 ? kSkipShort : kSkipLong;
		accepted.emplace_back();
		adsbdec_b200_frame_t &f = accepted.back();
		f.ts = c->n + 1 - drift_;  // demod.c:99: one tick per loop iteration, not per sample
		f.n = c->n;
		f.pw = c->pw;
		f.len = (int32_t)c->len;
		memcpy(f.bytes, c->bytes, 14);
		if (c->len == 7)
			memset(f.bytes + 7, 0, 7);
		f.pad_[0] = c->bytes[15];  // bits repaired (extension; 0 in reference mode)
		f.pad_[1] = 0;
		drift_ += skip - 1;
		pos = c->n + skip;  // demod.c:110, :119-123: idx = lidx
	}
	base_ = std::max(cutoff, pos);  // deqframe returns idx; air.c:96-98 keeps the rest
	scan_end_ = cutoff;
}

void Resolver::run_calls(std::vector<adsbdec_b200_frame_t> &accepted)
{
	for (;;) {
		// air.c:94: the test runs after every 2 envelope samples; the call happens at the first
		// even sample count P with P - base >= APBUFFSZ.
		uint64_t call_at = base_ + kAmpFlush;
		call_at += call_at & 1;
		if (call_at > total_out_)
			return;
		const uint64_t cutoff = call_at - kLookahead;
		if (cutoff > scanned_upto_)
			return;  // the GPU has not covered this call's offsets yet
		scan_to(cutoff, accepted);
	}
}

void Resolver::feed(const adsbdec_b200_candidate_t *c, size_t count, uint64_t scanned_upto,
                    uint64_t total_out, std::vector<adsbdec_b200_frame_t> &accepted)
{
	if (head_ == pending_.size()) {
		pending_.clear();
		head_ = 0;
	}
	in_ = c;
	in_n_ = count;
	in_i_ = 0;
	scanned_upto_ = std::max(scanned_upto_, scanned_upto);
	total_out_ = std::max(total_out_, total_out);
	run_calls(accepted);
	// keep what the call schedule has not reached yet (the caller's array goes away)
	if (head_ == pending_.size()) {
		pending_.clear();
		head_ = 0;
	}
	pending_.insert(pending_.end(), c + in_i_, c + count);
	in_ = nullptr;
	in_n_ = in_i_ = 0;
}

void Resolver::feed_frames(const adsbdec_b200_frame_t *f, size_t count, uint64_t scanned_upto,
                           uint64_t total_out, std::vector<adsbdec_b200_frame_t> &accepted, uint32_t ok[3])
{
	ok_ = ok;
	if (fhead_ == fpending_.size()) {
		fpending_.clear();
		fhead_ = 0;
	}
	fin_ = f;
	fin_n_ = count;
	fin_i_ = 0;
	scanned_upto_ = std::max(scanned_upto_, scanned_upto);
	total_out_ = std::max(total_out_, total_out);
	run_calls(accepted);
	if (fhead_ == fpending_.size()) {
		fpending_.clear();
		fhead_ = 0;
	}
	fpending_.insert(fpending_.end(), f + fin_i_, f + count);
	fin_ = nullptr;
	fin_n_ = fin_i_ = 0;
	ok_ = nullptr;  // the caller's counters are only borrowed for the call
}

void Resolver::finish(std::vector<adsbdec_b200_frame_t> &accepted, uint32_t ok[3])
{
	ok_ = ok;
	run_calls(accepted);
	if (!emulate_tail_) {
		// Non-reference mode: scan every offset that has a complete look-ahead.
		uint64_t cutoff = std::min(scanned_upto_, total_out_ >= (uint64_t)kLookahead
		                                              ? total_out_ - kLookahead
		                                              : 0);
		if (cutoff > scan_end_)
			scan_to(cutoff, accepted);
	}
	pending_.clear();
	head_ = 0;  // the reference drops whatever is left in ampbuff at EOF (air.c:243-244)
	fpending_.clear();
	fhead_ = 0;
	ok_ = nullptr;
}

}  // namespace adsb
