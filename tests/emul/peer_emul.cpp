// peer_emul.cpp -- TEST INFRASTRUCTURE.  Runs the sequencing of the peer-store halo exchange
// (chmy.jl_b200/csrc/peer_link.cuh: pl_exchange_dim, pl_slot, pl_grow_cap, the block layout -- the same source nvcc compiles
// into comm.cu) with one host thread per rank of a Cartesian topology.  A thread executes its operations in order, as a CUDA
// stream does; "mapped peer memory" is simply the neighbour's block in the shared address space; the payload is written and
// read with plain (non-atomic) stores and loads and the sequence flags with release / acquire atomics, mirroring
// st.release.sys / ld.acquire.sys of k_pl_flags.  Every unpack checks the complete message against what the neighbour must
// have sent for exactly this exchange; random delays shake the interleavings; built with -fsanitize=thread the run also
// proves the absence of data races on the slots.  tests/test_peer_protocol.py drives it.
//
//   peer_emul PX PY PZ ITERS SEED MAX_DELAY_US WORDS GROW_EVERY SLOW_UNPACK_US      (env: PEER_EMUL_TIMEOUT_S, PEER_EMUL_DIE=rank:iter)
//     message of dimension D at iteration it: WORDS * (D + 1) * (1 + it / GROW_EVERY) 64-bit words (GROW_EVERY = 0: fixed)
// Build: g++ -std=c++20 -O2 -pthread [-fsanitize=thread] [-DPL_SLOTS=1]
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

#include "../../chmy.jl_b200/csrc/peer_link.cuh"

using clk = std::chrono::steady_clock;

static std::atomic<bool>     g_abort{false};
static std::atomic<uint64_t> g_mismatch{0}, g_messages{0}, g_timeouts{0}, g_regrows{0}, g_handshake_errors{0};
static double                g_timeout_s = 20.0;

static uint64_t mix(uint64_t x) {       // splitmix64
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
// word i of the message rank `sender` packs on its side `side` of dimension D in iteration `it`
static uint64_t payload(int sender, int D, int side, uint64_t it, uint64_t i) {
    return mix(mix(((uint64_t)sender << 40) ^ ((uint64_t)D << 36) ^ ((uint64_t)side << 32) ^ it) + i);
}

// ---- pl_swap of comm.cu: a blocking pairwise exchange of a small record (there: NCCL send/recv + stream synchronise)
struct Hello { char* blk; size_t cap; };
struct Mailbox {
    std::mutex mu;
    std::condition_variable cv;
    std::map<std::pair<int, int>, std::vector<Hello>> box;      // (from, to) -> queue
    bool swap(int me, int peer, const Hello& out, Hello* in) {
        std::unique_lock<std::mutex> lk(mu);
        box[{me, peer}].push_back(out);
        cv.notify_all();
        const auto deadline = clk::now() + std::chrono::duration<double>(g_timeout_s);
        while (box[{peer, me}].empty()) {
            if (g_abort.load() || cv.wait_until(lk, deadline) == std::cv_status::timeout) {
                if (box[{peer, me}].empty()) return false;
            }
        }
        *in = box[{peer, me}].front();
        box[{peer, me}].erase(box[{peer, me}].begin());
        return true;
    }
};
static Mailbox g_mail;

struct Rank {
    int r = 0, nd = 0, dims[3] = {1, 1, 1}, coords[3] = {0, 0, 0}, nb[3][2];
    PlLink link[3][2];
    std::vector<char*> grave;
    std::mt19937_64 rng;
    int max_delay_us = 0, slow_unpack_us = 0;
    void delay() {
        if (max_delay_us <= 0) return;
        const uint64_t d = rng() % (uint64_t)(max_delay_us + 1);
        if (d && (rng() & 3) == 0) std::this_thread::sleep_for(std::chrono::microseconds(d));
        else if (rng() & 1) std::this_thread::yield();
    }
};

// the three stream operations of pl_exchange_dim, executed synchronously by the rank's thread
struct EmuOps {
    Rank*    me;
    int      D;
    uint64_t it;
    size_t   words;
    int push(int s, PlLink& l, int slot) {
        me->delay();
        uint64_t* dst = reinterpret_cast<uint64_t*>(l.remote + pl_off_slot(slot, l.cap));
        for (size_t i = 0; i < words; ++i) dst[i] = payload(me->r, D, s, it, i);          // the pack kernel's peer stores
        g_messages.fetch_add(1, std::memory_order_relaxed);
        return 0;
    }
    int post_and_wait(PlLink* const l[2], const uint64_t k[2]) {
        me->delay();
        for (int s = 0; s < 2; ++s)
            if (l[s]) std::atomic_ref<uint64_t>(*pl_flag(l[s]->remote, PL_OFF_DATA)).store(k[s], std::memory_order_release);
        for (int s = 0; s < 2; ++s) {
            if (!l[s]) continue;
            std::atomic_ref<uint64_t> f(*pl_flag(l[s]->local, PL_OFF_DATA));
            const auto t0 = clk::now();
            while (f.load(std::memory_order_acquire) < k[s]) {
                if (g_abort.load(std::memory_order_relaxed) || std::chrono::duration<double>(clk::now() - t0).count() > g_timeout_s) {
                    g_timeouts.fetch_add(1);
                    g_abort.store(true);
                    return 1;
                }
                std::this_thread::yield();
            }
        }
        return 0;
    }
    int unpack(int s, PlLink& l, int slot) {
        me->delay();
        const uint64_t* src = reinterpret_cast<const uint64_t*>(l.local + pl_off_slot(slot, l.cap));
        uint64_t bad = 0;
        const size_t half = words / 2;
        for (size_t i = 0; i < half; ++i) bad += src[i] != payload(l.peer, D, 1 - s, it, i);
        if (me->slow_unpack_us > 0) std::this_thread::sleep_for(std::chrono::microseconds(me->slow_unpack_us));
        for (size_t i = half; i < words; ++i) bad += src[i] != payload(l.peer, D, 1 - s, it, i);
        if (bad) g_mismatch.fetch_add(bad);
        return 0;
    }
};

// pl_link_ensure of comm.cu without the CUDA calls: (re)allocate when the message does not fit, pair the hand-shake
static int ensure(Rank& me, PlLink& l, size_t need) {
    if (l.mode == PL_MODE_PEER && need <= l.cap) return 0;
    const size_t cap = pl_grow_cap(l.mode == PL_MODE_PEER ? l.cap : 0, need);
    char* blk = static_cast<char*>(aligned_alloc(PL_ALIGN, pl_block_bytes(cap)));
    memset(blk, 0, pl_block_bytes(cap));
    Hello mine{blk, cap}, theirs{nullptr, 0};
    if (!g_mail.swap(me.r, l.peer, mine, &theirs) || theirs.cap != cap) {      // the peer did not get here in this exchange
        g_handshake_errors.fetch_add(1);
        g_abort.store(true);
        me.grave.push_back(blk);
        return 2;
    }
    if (l.mode == PL_MODE_PEER) g_regrows.fetch_add(1);
    if (l.local) me.grave.push_back(l.local);
    l.local = blk; l.remote = theirs.blk; l.cap = cap; l.seq = 0; l.mode = PL_MODE_PEER;
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 10) { fprintf(stderr, "usage: peer_emul PX PY PZ ITERS SEED MAX_DELAY_US WORDS GROW_EVERY SLOW_UNPACK_US\n"); return 2; }
    int dims[3] = {atoi(argv[1]), atoi(argv[2]), atoi(argv[3])};
    const uint64_t iters = strtoull(argv[4], nullptr, 10), seed = strtoull(argv[5], nullptr, 10);
    const int max_delay = atoi(argv[6]);
    const size_t words0 = strtoull(argv[7], nullptr, 10);
    const uint64_t grow_every = strtoull(argv[8], nullptr, 10);
    const int slow_unpack = atoi(argv[9]);
    if (getenv("PEER_EMUL_TIMEOUT_S")) g_timeout_s = atof(getenv("PEER_EMUL_TIMEOUT_S"));
    const int nd = dims[2] > 1 ? 3 : dims[1] > 1 ? 2 : 1;
    const int R = dims[0] * dims[1] * dims[2];
    std::vector<Rank> ranks(R);
    for (int r = 0; r < R; ++r) {           // row-major ranks, non-periodic neighbours (chmy_topo_create, comm.cu)
        Rank& k = ranks[r];
        k.r = r; k.nd = nd; k.rng.seed(mix(seed + r)); k.max_delay_us = max_delay;
        k.slow_unpack_us = (r & 1) ? slow_unpack : 0;
        int q = r;
        for (int a = nd - 1; a >= 0; --a) { k.dims[a] = dims[a]; k.coords[a] = q % dims[a]; q /= dims[a]; }
        for (int a = 0; a < 3; ++a)
            for (int s = 0; s < 2; ++s) {
                k.nb[a][s] = -1;
                memset(&k.link[a][s], 0, sizeof(PlLink));
                if (a < nd) {
                    int cc[3] = {k.coords[0], k.coords[1], k.coords[2]};
                    cc[a] += s == 0 ? -1 : 1;
                    if (cc[a] >= 0 && cc[a] < dims[a]) {
                        int nr = 0;
                        for (int b = 0; b < nd; ++b) nr = nr * dims[b] + cc[b];
                        k.nb[a][s] = nr;
                    }
                }
                k.link[a][s].peer = k.nb[a][s];
            }
    }
    int die_rank = -1;                      // PEER_EMUL_DIE=rank:iter -- that rank stops exchanging at that iteration
    uint64_t die_iter = 0;
    if (const char* d = getenv("PEER_EMUL_DIE")) { die_rank = atoi(d); if (const char* c = strchr(d, ':')) die_iter = strtoull(c + 1, nullptr, 10); }
    std::vector<std::thread> th;
    for (int r = 0; r < R; ++r)
        th.emplace_back([&, r] {
            Rank& me = ranks[r];
            for (uint64_t it = 0; it < iters && !g_abort.load(); ++it) {
                if (r == die_rank && it == die_iter) return;            // fault injection: this rank is lost
                for (int D = me.nd - 1; D >= 0; --D) {                 // bc!: D = N..1 (batch.jl:20-29)
                    const size_t words = words0 * (size_t)(D + 1) * (size_t)(1 + (grow_every ? it / grow_every : 0));
                    PlLink* pl[2] = {nullptr, nullptr};
                    bool failed = false;
                    for (int s = 0; s < 2; ++s) {
                        if (me.nb[D][s] < 0) continue;
                        if (ensure(me, me.link[D][s], words * 8)) { failed = true; break; }
                        pl[s] = &me.link[D][s];
                    }
                    if (failed) return;
                    EmuOps ops{&me, D, it, words};
                    if (pl_exchange_dim(ops, pl)) return;
                }
                me.delay();        // the inner-region kernel / the next launch
            }
        });
    for (auto& t : th) t.join();
    for (auto& k : ranks) {
        for (int a = 0; a < 3; ++a)
            for (int s = 0; s < 2; ++s) free(k.link[a][s].local);
        for (char* p : k.grave) free(p);
    }
    printf("{\"ranks\": %d, \"slots\": %d, \"iters\": %llu, \"messages\": %llu, \"mismatches\": %llu, \"timeouts\": %llu, "
           "\"regrows\": %llu, \"handshake_errors\": %llu}\n", R, (int)PL_SLOTS, (unsigned long long)iters,
           (unsigned long long)g_messages.load(), (unsigned long long)g_mismatch.load(), (unsigned long long)g_timeouts.load(),
           (unsigned long long)g_regrows.load(), (unsigned long long)g_handshake_errors.load());
    return (g_mismatch.load() || g_timeouts.load() || g_handshake_errors.load()) ? 1 : 0;
}
