// b200pack_team.h -- the packer / reader thread team of a context (host-only C++, no CUDA).
#pragma once
#include <pthread.h>
#include <sched.h>

#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace b200h {

// A context's packer / reader threads: started once, parked on a condition variable between slots.  Filling a 256 MiB
// slot takes 5-8 ms; starting and joining a fresh team of 16 for it cost ~0.4 ms of that (build machine, 8 packers:
// 50.5 GiB/s with 256 MiB slots against 53.6 with 1 GiB slots, 39 with 64 MiB slots).
//   run(use, cpus, job)  runs job() on the calling thread and on up to use - 1 team threads and returns when every
//                        thread that picked the job up has come back.  job must be safe to run on any number of
//                        threads >= 1 (the packers pull grains from a shared counter): a team thread that wakes up
//                        after the caller has already finished the work simply does not take part.
// One caller at a time (a context's team is only used under the context mutex).
class PackTeam {
public:
    PackTeam() = default;
    PackTeam(const PackTeam&) = delete;
    PackTeam& operator=(const PackTeam&) = delete;
    ~PackTeam() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_go_.notify_all();
        for (auto& t : threads_) t.join();
    }

    void run(int use, const cpu_set_t* cpus, const std::function<void()>& job) {
        if (use <= 1) {
            job();
            return;
        }
        while ((int)threads_.size() < use - 1) {
            threads_.emplace_back([this] { loop(); });
            if (cpus) pthread_setaffinity_np(threads_.back().native_handle(), sizeof(cpu_set_t), cpus);  // best effort
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            job_ = &job;
            want_ = use - 1;
            taken_ = 0;
            running_ = 0;
            ++gen_;
        }
        cv_go_.notify_all();
        job();
        std::unique_lock<std::mutex> lk(m_);
        want_ = taken_;  // closed: nobody else may pick this job up (its captures die when run() returns)
        cv_done_.wait(lk, [&] { return running_ == 0; });
        job_ = nullptr;
    }

private:
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void()>* job = nullptr;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_go_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
                if (taken_ < want_) {
                    ++taken_;
                    ++running_;
                    job = job_;
                }
            }
            if (!job) continue;
            (*job)();
            std::lock_guard<std::mutex> lk(m_);
            if (--running_ == 0) cv_done_.notify_one();
        }
    }

    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_go_, cv_done_;
    const std::function<void()>* job_ = nullptr;
    uint64_t gen_ = 0;
    int want_ = 0, taken_ = 0, running_ = 0;
    bool stop_ = false;
};


}  // namespace b200h
