// Parallel file reader for adsbdec_b200_decode_file.
//
// The reference reads its capture with one thread, 2 MiB per read() (air.c:233-239); at ~170 M
// samples/s that is nowhere near a limit.  Here the decoder takes > 20 G samples/s from page-locked
// memory, and one thread copying out of the page cache (~3 G samples/s) would be the bottleneck of the
// whole `-f` path.  The bulk of a regular file (all the full 2 MiB reads, whose content does not
// depend on read order) is therefore fetched by a few helper threads with pread() straight into the
// page-locked staging buffer; only the ragged last read keeps the reference's sequential semantics.
#pragma once

#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace adsb {

class ReadPool {
public:
	explicit ReadPool(int helpers)
	{
		for (int i = 0; i < helpers; i++)
			th_.emplace_back([this] { loop(); });
	}
	~ReadPool()
	{
		{
			std::lock_guard<std::mutex> l(mu_);
			quit_ = true;
		}
		go_.notify_all();
		for (auto &t : th_)
			t.join();
	}
	ReadPool(const ReadPool &) = delete;
	ReadPool &operator=(const ReadPool &) = delete;

	// Reads bytes [off, off+bytes) of fd into dst with every thread (the caller's included).
	// False if any piece came back short or failed.
	bool read(int fd, void *dst, off_t off, size_t bytes)
	{
		{
			std::lock_guard<std::mutex> l(mu_);
			fd_ = fd;
			dst_ = static_cast<char *>(dst);
			off_ = off;
			bytes_ = bytes;
			next_.store(0, std::memory_order_relaxed);
			bad_.store(0, std::memory_order_relaxed);
			running_ = (int)th_.size();
			gen_++;
		}
		go_.notify_all();
		work();
		std::unique_lock<std::mutex> l(mu_);
		done_.wait(l, [this] { return running_ == 0; });
		return bad_.load(std::memory_order_relaxed) == 0;
	}

private:
	static constexpr size_t kPiece = 1u << 20;

	void work()
	{
		for (;;) {
			const size_t at = next_.fetch_add(kPiece, std::memory_order_relaxed);
			if (at >= bytes_)
				return;
			size_t left = bytes_ - at < kPiece ? bytes_ - at : kPiece, pos = at;
			while (left > 0) {
				const ssize_t got = pread(fd_, dst_ + pos, left, off_ + (off_t)pos);
				if (got <= 0) {
					bad_.store(1, std::memory_order_relaxed);
					return;
				}
				pos += (size_t)got;
				left -= (size_t)got;
			}
		}
	}

	void loop()
	{
		unsigned long long seen = 0;
		std::unique_lock<std::mutex> l(mu_);
		for (;;) {
			go_.wait(l, [&] { return quit_ || gen_ != seen; });
			if (quit_)
				return;
			seen = gen_;
			l.unlock();
			work();
			l.lock();
			if (--running_ == 0)
				done_.notify_one();
		}
	}

	std::vector<std::thread> th_;
	std::mutex mu_;
	std::condition_variable go_, done_;
	unsigned long long gen_ = 0;
	bool quit_ = false;
	int running_ = 0;
	int fd_ = -1;
	char *dst_ = nullptr;
	off_t off_ = 0;
	size_t bytes_ = 0;
	std::atomic<size_t> next_{0};
	std::atomic<int> bad_{0};
};

}  // namespace adsb
