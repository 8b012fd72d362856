// TEST INFRASTRUCTURE ONLY (see oracle.h).  Minimal persistent thread pool standing in for the
// reference's tbb::parallel_for over blocked ranges (CollapsedEMOptimizer.cpp:233,305,322) and for
// its std::thread-per-worker read loop (SailfishQuantify.cpp:903-935).
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace orc {

class Pool {
  public:
    explicit Pool(int n) : n_(n < 1 ? 1 : n) {
        for (int i = 1; i < n_; ++i) workers_.emplace_back([this, i] { loop(i); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int size() const { return n_; }

    // fn(begin, end) over [0, n) split into size() contiguous chunks (static partition, like a
    // blocked_range with one chunk per worker).
    void parallel_for(size_t n, const std::function<void(size_t, size_t)>& fn) {
        if (n_ == 1 || n < static_cast<size_t>(n_) * 4) { if (n) fn(0, n); return; }
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn; total_ = n; pending_ = n_ - 1; ++gen_;
        }
        cv_.notify_all();
        run_chunk(0);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }

    // fn(i) once on every worker i in [0, size())
    void run_each(const std::function<void(size_t)>& fn) {
        std::function<void(size_t, size_t)> w = [&](size_t b, size_t e) { for (size_t i = b; i < e; ++i) fn(i); };
        if (n_ == 1) { fn(0); return; }
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &w; total_ = static_cast<size_t>(n_); pending_ = n_ - 1; ++gen_;
        }
        cv_.notify_all();
        run_chunk(0);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }

  private:
    void run_chunk(int i) {
        const size_t per = (total_ + n_ - 1) / n_;
        const size_t b = per * i, e = std::min(total_, b + per);
        if (b < e) (*fn_)(b, e);
    }
    void loop(int i) {
        size_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            run_chunk(i);
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(size_t, size_t)>* fn_ = nullptr;
    size_t total_ = 0, gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

}  // namespace orc
