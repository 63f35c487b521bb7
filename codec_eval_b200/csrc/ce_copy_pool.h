// ce_copy_pool.h -- plain C++ (no CUDA): the host threads behind ce_evaluate_batch's pageable-memory path.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace ce {

// A few host threads that copy images from PAGEABLE caller memory into the pinned staging slot in parallel
// (ce_evaluate_batch).  The driver's own pageable path is one thread moving bytes through a small bounce buffer
// (~10 GB/s here, and it blocks the caller); four memcpy threads fill a pinned slot at several times that, after which
// ONE asynchronous copy moves the slot at PCIe rate.  Created on first use.
struct CopyPool {
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    const std::function<void(size_t)>* job = nullptr;
    size_t next = 0, count = 0, active = 0;
    uint64_t generation = 0;
    bool stop = false;
    void start(unsigned n);
    void run(size_t count, const std::function<void(size_t)>& fn);   // fn(i) for i in [0, count), returns when all done
    ~CopyPool();
};

inline void CopyPool::start(unsigned n) {
    if (!workers.empty()) return;
    for (unsigned t = 0; t < n; t++)
        workers.emplace_back([this] {
            uint64_t seen = 0;
            for (;;) {
                std::unique_lock<std::mutex> lk(m);
                cv_work.wait(lk, [&] { return stop || (generation != seen && next < count); });
                if (stop) return;
                const uint64_t gen = generation;
                while (generation == gen && next < count) {
                    const size_t i = next++;
                    active++;
                    lk.unlock();
                    (*job)(i);
                    lk.lock();
                    active--;
                }
                seen = gen;
                if (next >= count && active == 0) cv_done.notify_all();
            }
        });
}
inline void CopyPool::run(size_t n, const std::function<void(size_t)>& fn) {
    if (n == 0) return;
    if (workers.empty()) {
        for (size_t i = 0; i < n; i++) fn(i);
        return;
    }
    std::unique_lock<std::mutex> lk(m);
    job = &fn;
    next = 0;
    count = n;
    generation++;
    cv_work.notify_all();
    cv_done.wait(lk, [&] { return next >= count && active == 0; });
    job = nullptr;
    count = 0;
}
inline CopyPool::~CopyPool() {
    {
        std::lock_guard<std::mutex> lk(m);
        stop = true;
    }
    cv_work.notify_all();
    for (auto& t : workers) t.join();
}

}  // namespace ce
