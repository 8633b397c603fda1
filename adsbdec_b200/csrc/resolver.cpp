// resolver.cpp - see resolver.h.
#include "resolver.h"

#include <string.h>

#include <algorithm>

#include "common.h"

namespace adsb {

// scan_to for advance_keys / finish_keys: only stamped frames exist, as keys (start << 1 | long).
void Resolver::scan_keys_to(uint64_t cutoff)
{
	uint64_t pos = base_;
	const unsigned long long *k = kin_;
	const size_t n = fin_n_, i0 = fin_i_;
	size_t i = i0;
	// a call covers ~40 k offsets - a dozen frames of a dense capture: walk them (one predictable branch each);
	// search only what a long walk has not reached
	const unsigned long long kc = (unsigned long long)cutoff << 1;  // k < kc  <=>  start < cutoff
	const size_t w = std::min(n, i + 48);
	while (i + 4 <= w && k[i + 3] < kc)  // (four at a time: the walk is a chain of dependent loads otherwise)
		i += 4;
	while (i < w && k[i] < kc)
		i++;
	if (i == w && i < n && k[i] < kc)
		i = (size_t)(std::lower_bound(k + i, k + n, kc) - k);
	if (i > i0)
		pos = (k[i - 1] >> 1) + ((k[i - 1] & 1ull) ? kSkipLong : kSkipShort);
	fin_i_ = i;
	base_ = std::max(cutoff, pos);  // deqframe returns idx; air.c:96-98 keeps the rest
	scan_end_ = cutoff;
}

void Resolver::scan_to(uint64_t cutoff, std::vector<adsbdec_b200_frame_t> &accepted)
{
	if (kin_) {
		scan_keys_to(cutoff);
		return;
	}
	// One deqframe(ampbuff, len) call whose loop bound len-DECOFFSET is `cutoff` in absolute
	// offsets (demod.c:89).  idx starts at ampbuff[0] = base_.  Candidates come from the few left
	// over from earlier feeds (pending_) and then, in place, from the array being fed.
	uint64_t pos = base_;
	// stamped frames: everything below the cutoff is accepted as it is.  A call covers ~40 k offsets, i.e. a
	// dozen frames of a dense capture, so the end of its run is searched (first in a short window, then in the
	// rest) instead of walked, and the frames are copied once per feed (flush_stamped), not once per call:
	// 15 000 calls per minute of signal used to cost a millisecond of host time.
	for (int src = 0; src < 2; src++) {
		const adsbdec_b200_frame_t *f = src == 0 ? fpending_.data() : fin_;
		size_t &i = src == 0 ? fhead_ : fin_i_;
		const size_t n = src == 0 ? fpending_.size() : fin_n_, i0 = i;
		if (i < n && f[i].n < cutoff) {
			const auto below = [cutoff](const adsbdec_b200_frame_t &x, uint64_t) { return x.n < cutoff; };
			const size_t w = std::min(n, i + 64);
			size_t e = (size_t)(std::lower_bound(f + i, f + w, cutoff, below) - f);
			if (e == w && w < n)
				e = (size_t)(std::lower_bound(f + w, f + n, cutoff, below) - f);
			i = e;
		}
		if (i > i0) {
			if (ok_)
				for (size_t k = i0; k < i; k++) {
					const int df = f[k].bytes[0] >> 3;
					if (ok_class(df) >= 0)
						ok_[ok_class(df)]++;
				}
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

// Stamped frames consumed since the last flush: [fdone_, fhead_) of the pending vector, then
// [fin_done_, fin_i_) of the array being fed.
void Resolver::flush_stamped(std::vector<adsbdec_b200_frame_t> &accepted)
{
	if (inplace_)
		return;
	if (fhead_ > fdone_)
		accepted.insert(accepted.end(), fpending_.begin() + (ptrdiff_t)fdone_, fpending_.begin() + (ptrdiff_t)fhead_);
	fdone_ = fhead_;
	if (fin_ && fin_i_ > fin_done_)
		accepted.insert(accepted.end(), fin_ + fin_done_, fin_ + fin_i_);
	fin_done_ = fin_i_;
}

void Resolver::count_pending(uint32_t out[3]) const
{
	out[0] = out[1] = out[2] = 0;
	for (size_t k = fhead_; k < fpending_.size(); k++) {
		const int df = fpending_[k].bytes[0] >> 3;
		if (ok_class(df) >= 0)
			out[ok_class(df)]++;
	}
}

void Resolver::run_calls(std::vector<adsbdec_b200_frame_t> &accepted)
{
	struct Flush {  // whatever way the loop is left
		Resolver *r;
		std::vector<adsbdec_b200_frame_t> &a;
		~Flush() { r->flush_stamped(a); }
	} flush_on_exit{this, accepted};
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
		fhead_ = fdone_ = 0;
	}
	fin_ = f;
	fin_n_ = count;
	fin_i_ = fin_done_ = 0;
	scanned_upto_ = std::max(scanned_upto_, scanned_upto);
	total_out_ = std::max(total_out_, total_out);
	run_calls(accepted);
	if (fhead_ == fpending_.size()) {
		fpending_.clear();
		fhead_ = fdone_ = 0;
	}
	fpending_.insert(fpending_.end(), f + fin_i_, f + count);
	fin_ = nullptr;
	fin_n_ = fin_i_ = fin_done_ = 0;
	ok_ = nullptr;  // the caller's counters are only borrowed for the call
}

size_t Resolver::advance_frames(const adsbdec_b200_frame_t *f, size_t n_vis, size_t n_total, uint64_t scanned_upto,
                                uint64_t total_out)
{
	std::vector<adsbdec_b200_frame_t> none;
	inplace_ = true;
	ok_ = nullptr;
	fin_ = f + n_vis;
	fin_n_ = n_total - n_vis;
	fin_i_ = fin_done_ = 0;
	scanned_upto_ = std::max(scanned_upto_, scanned_upto);
	total_out_ = std::max(total_out_, total_out);
	run_calls(none);
	const size_t r = n_vis + fin_i_;
	fin_ = nullptr;
	fin_n_ = fin_i_ = fin_done_ = 0;
	inplace_ = false;
	return r;
}

size_t Resolver::advance_keys(const unsigned long long *k, size_t n_vis, size_t n_total, uint64_t scanned_upto,
                              uint64_t total_out)
{
	std::vector<adsbdec_b200_frame_t> none;
	inplace_ = true;
	ok_ = nullptr;
	kin_ = k + n_vis;
	fin_n_ = n_total - n_vis;
	fin_i_ = fin_done_ = 0;
	scanned_upto_ = std::max(scanned_upto_, scanned_upto);
	total_out_ = std::max(total_out_, total_out);
	run_calls(none);
	const size_t r = n_vis + fin_i_;
	kin_ = nullptr;
	fin_n_ = fin_i_ = fin_done_ = 0;
	inplace_ = false;
	return r;
}

size_t Resolver::finish_keys(const unsigned long long *k, size_t n_vis, size_t n_total)
{
	std::vector<adsbdec_b200_frame_t> none;
	inplace_ = true;
	ok_ = nullptr;
	kin_ = k + n_vis;
	fin_n_ = n_total - n_vis;
	fin_i_ = fin_done_ = 0;
	finish(none, nullptr);
	const size_t r = n_vis + fin_i_;
	kin_ = nullptr;
	fin_n_ = fin_i_ = fin_done_ = 0;
	inplace_ = false;
	return r;
}

size_t Resolver::finish_frames(const adsbdec_b200_frame_t *f, size_t n_vis, size_t n_total)
{
	std::vector<adsbdec_b200_frame_t> none;
	inplace_ = true;
	ok_ = nullptr;
	fin_ = f + n_vis;
	fin_n_ = n_total - n_vis;
	fin_i_ = fin_done_ = 0;
	finish(none, nullptr);
	const size_t r = n_vis + fin_i_;
	fin_ = nullptr;
	fin_n_ = fin_i_ = fin_done_ = 0;
	inplace_ = false;
	return r;
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
		if (cutoff > scan_end_) {
			scan_to(cutoff, accepted);
			flush_stamped(accepted);
		}
	}
	pending_.clear();
	head_ = 0;  // the reference drops whatever is left in ampbuff at EOF (air.c:243-244)
	fpending_.clear();
	fhead_ = fdone_ = 0;
	ok_ = nullptr;
}

}  // namespace adsb
