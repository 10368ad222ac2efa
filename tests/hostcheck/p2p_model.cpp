// Host-thread model of the peer-memory protocols of topomax_b200/csrc/tm_p2p.cuh (test
// infrastructure; CPU only).  One thread per rank; "windows" are plain arrays that the neighbours
// write into; arrival flags are std::atomic with release/acquire, exactly the roles of
// st.release.sys / ld.acquire.sys in the kernels.  The model checks the claim the kernels rely on:
// with TWO mailboxes per direction (epoch parity) and bidirectional exchanges, no acknowledgement
// is needed -- a mailbox is never overwritten before its owner has unpacked it -- and the same for
// the all-reduce slots.  Random delays shake the interleavings; `buffers = 1` is the negative
// control (single mailbox: overwrites before unpacking are then observed).
#include <atomic>
#include <chrono>
#include <cstdint>
#include <random>
#include <thread>
#include <vector>

namespace {

constexpr int ROWS = 8;  // payload words per message

struct Window {
    std::atomic<unsigned long long> halo_flag[2][2];  // [from below / above][parity]
    std::uint64_t mailbox[2][2][ROWS];
    std::atomic<unsigned long long> red_flag[16][2];
    double red_slot[2][16][4];
    Window() {
        for (auto& a : halo_flag)
            for (auto& f : a) f.store(0);
        for (auto& a : red_flag)
            for (auto& f : a) f.store(0);
    }
};

std::uint64_t payload(int rank, unsigned long long epoch, int dir, int i) {
    return (std::uint64_t)rank * 1000003ULL + epoch * 7919ULL + (std::uint64_t)dir * 101ULL + (std::uint64_t)i;
}

void jitter(std::mt19937& g) {
    const int k = (int)(g() % 16);
    if (k == 0) std::this_thread::sleep_for(std::chrono::microseconds(g() % 200));
    else if (k < 4) std::this_thread::yield();
}

}  // namespace

extern "C" int p2p_model_run(int ranks, int epochs, int seed, int buffers, int reduce_every) {
    std::vector<Window> win(ranks);
    std::atomic<int> violations{0};
    auto body = [&](int r) {
        std::mt19937 g(seed * 977 + r);
        unsigned long long halo_epoch = 0, red_epoch = 0;
        for (int e = 0; e < epochs; ++e) {
            // ---- halo exchange (p2p_halo_kernel): push, publish, wait on my own window, unpack
            const unsigned long long epoch = ++halo_epoch;
            const int par = buffers == 2 ? (int)(epoch & 1ULL) : 0;
            jitter(g);
            if (r + 1 < ranks) {
                for (int i = 0; i < ROWS; ++i) win[r + 1].mailbox[0][par][i] = payload(r, epoch, 0, i);
                jitter(g);
                win[r + 1].halo_flag[0][par].store(epoch, std::memory_order_release);
            }
            if (r > 0) {
                for (int i = 0; i < ROWS; ++i) win[r - 1].mailbox[1][par][i] = payload(r, epoch, 1, i);
                jitter(g);
                win[r - 1].halo_flag[1][par].store(epoch, std::memory_order_release);
            }
            if (r > 0) {
                while (win[r].halo_flag[0][par].load(std::memory_order_acquire) < epoch) std::this_thread::yield();
                jitter(g);
                for (int i = 0; i < ROWS; ++i)
                    if (win[r].mailbox[0][par][i] != payload(r - 1, epoch, 0, i)) violations++;
            }
            if (r + 1 < ranks) {
                while (win[r].halo_flag[1][par].load(std::memory_order_acquire) < epoch) std::this_thread::yield();
                jitter(g);
                for (int i = 0; i < ROWS; ++i)
                    if (win[r].mailbox[1][par][i] != payload(r + 1, epoch, 1, i)) violations++;
            }
            // ---- all-reduce (p2p_allreduce_kernel) every few exchanges, with its own epoch
            if (reduce_every > 0 && e % reduce_every == 0) {
                const unsigned long long re = ++red_epoch;
                const int rp = buffers == 2 ? (int)(re & 1ULL) : 0;
                const double mine = (double)(r + 1) * (double)re;
                for (int q = 0; q < ranks; ++q) win[q].red_slot[rp][r][0] = mine;
                jitter(g);
                for (int q = 0; q < ranks; ++q)
                    if (q != r) win[q].red_flag[r][rp].store(re, std::memory_order_release);
                for (int q = 0; q < ranks; ++q)
                    if (q != r)
                        while (win[r].red_flag[q][rp].load(std::memory_order_acquire) < re) std::this_thread::yield();
                jitter(g);
                double s = 0.0;
                for (int q = 0; q < ranks; ++q) s += win[r].red_slot[rp][q][0];
                const double want = (double)re * ranks * (ranks + 1) / 2.0;
                if (s != want) violations++;
            }
        }
    };
    std::vector<std::thread> th;
    for (int r = 0; r < ranks; ++r) th.emplace_back(body, r);
    for (auto& t : th) t.join();
    return violations.load();
}
