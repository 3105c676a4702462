// Test-only host model of the barrier-free task queue of the latency mode (metis_search.cu: QueueWarp and the pop
// loop of het_search_kernel).  Threads stand for warps; the protocol is restated with std::atomic: per-slot
// sequence numbers (== ticket: free for that ticket, == ticket + 1: published), head / tail tickets, an `alive`
// counter for termination.  A task carries (chain id, steps left, checksum); running it releases the slot,
// pushes the successor if steps remain and retires.  The model checks that every step of every chain runs exactly
// once, that no payload is read torn or stale, and that all threads terminate - also with the tightest legal ring
// (ring == number of chains), where producers do wait for their slot.
//
//   g++ -O2 -std=c++17 -pthread queue_model.cpp -o queue_model && ./queue_model <threads> <chains> <seed> <slack>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

struct Slot { uint32_t chain, left, check; };

int main(int argc, char **argv) {
    const int threads = argc > 1 ? atoi(argv[1]) : 8;
    const uint32_t n = argc > 2 ? (uint32_t)atoi(argv[2]) : 1000;
    const uint32_t seed = argc > 3 ? (uint32_t)atoi(argv[3]) : 1;
    const uint32_t slack = argc > 4 ? (uint32_t)atoi(argv[4]) : 64;
    const uint32_t ring = n + slack;
    std::mt19937 rng(seed);
    std::vector<uint32_t> length(n);
    uint64_t total = 0;
    for (auto &l : length) { l = 1 + rng() % 19; total += l; }
    std::vector<Slot> slot(ring);
    std::vector<std::atomic<uint32_t>> seq(ring);
    std::vector<std::atomic<uint32_t>> ran(n);
    for (uint32_t i = 0; i < ring; ++i) seq[i].store(i < n ? i + 1 : i);
    for (uint32_t i = 0; i < n; ++i) { slot[i] = Slot{i, length[i], i * 2654435761u ^ length[i]}; ran[i].store(0); }
    std::atomic<uint32_t> head{0}, tail{n}, alive{n}, errors{0};
    std::atomic<uint64_t> executed{0};
    auto worker = [&]() {
        for (;;) {
            const uint32_t h = head.fetch_add(1);
            bool ready = false;
            for (;;) {
                if (seq[h % ring].load(std::memory_order_acquire) == h + 1) { ready = true; break; }
                if (alive.load(std::memory_order_acquire) == 0) break;
                std::this_thread::yield();
            }
            if (!ready) return;
            const Slot t = slot[h % ring];                                        // restore
            seq[h % ring].store(h + ring, std::memory_order_release);             // consumed()
            if (t.chain >= n || t.left == 0 || t.left > length[t.chain] || t.check != (t.chain * 2654435761u ^ t.left)) {
                errors.fetch_add(1);
                alive.fetch_sub(1);
                continue;
            }
            ran[t.chain].fetch_add(1);
            executed.fetch_add(1);
            if (t.left > 1) {                                                     // append() + publish()
                const uint32_t p = tail.fetch_add(1);
                alive.fetch_add(1);
                while (seq[p % ring].load(std::memory_order_acquire) != p) std::this_thread::yield();
                slot[p % ring] = Slot{t.chain, t.left - 1, t.chain * 2654435761u ^ (t.left - 1)};
                seq[p % ring].store(p + 1, std::memory_order_release);
            }
            alive.fetch_sub(1, std::memory_order_acq_rel);                        // retire after the successor is alive
        }
    };
    std::vector<std::thread> pool;
    for (int i = 0; i < threads; ++i) pool.emplace_back(worker);
    for (auto &t : pool) t.join();
    uint32_t wrong = 0;
    for (uint32_t i = 0; i < n; ++i) wrong += ran[i].load() != length[i];
    printf("{\"executed\": %llu, \"expected\": %llu, \"wrong_chains\": %u, \"errors\": %u, \"alive\": %u, \"pushed\": %u}\n",
           (unsigned long long)executed.load(), (unsigned long long)total, wrong, errors.load(), alive.load(),
           tail.load() - n);
    return (executed.load() == total && wrong == 0 && errors.load() == 0 && alive.load() == 0) ? 0 : 1;
}
