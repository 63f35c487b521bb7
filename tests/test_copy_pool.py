"""The host thread pool behind ce_evaluate_batch's pageable-memory path (csrc/ce_copy_pool.h, plain C++): every item
of every job runs exactly once, jobs of any size (0, fewer than the workers, many) complete, and the pool can be
reused tens of thousands of times -- compiled and run here with g++, no GPU."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r'''
#include <atomic>
#include <cstdio>
#include "ce_copy_pool.h"
int main() {
    ce::CopyPool p;
    std::atomic<long> sum{0};
    p.run(10, [&](size_t i) { sum += (long)i; });          // no workers yet: runs inline
    if (sum != 45) return 2;
    p.start(4);
    p.start(4);                                             // idempotent
    for (int rep = 0; rep < 20000; rep++) {
        const size_t n = rep % 37;
        std::vector<int> hit(n, 0);
        const std::function<void(size_t)> f = [&](size_t i) { hit[i]++; sum += 1; };
        p.run(n, f);
        for (size_t i = 0; i < n; i++)
            if (hit[i] != 1) { printf("item %zu of job %d ran %d times\n", i, rep, hit[i]); return 1; }
    }
    printf("ok %ld\n", sum.load());
    return 0;
}
'''


def test_copy_pool_runs_every_item_exactly_once(tmp_path):
    src = tmp_path / "pool_test.cpp"
    src.write_text(HARNESS)
    exe = tmp_path / "pool_test"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "codec_eval_b200", "csrc"),
                           "-o", str(exe), str(src)])
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.startswith("ok ")
