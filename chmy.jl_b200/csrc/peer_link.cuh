// peer_link.cuh -- protocol of the peer-store halo exchange (comm.cu, CHMY_EXCHANGE_PEER): instead of pack -> ncclSend /
// ncclRecv -> unpack, the pack kernel of rank A stores the slabs of a (dim, side) straight into a receive slot that lives
// in the neighbour's HBM (mapped here with CUDA IPC, written over NVLink), and the two ranks hand-shake through one
// monotonic sequence flag per direction.  Replaces the Irecv / pack / Isend / poll / unpack loop of
// src/Distributed/exchange_halo.jl:13-61 and its per-field StackAllocator buffers (stack_allocator.jl:7-89).
//
// Everything that decides ORDER lives in this header as plain C++, so that tests/emul/peer_emul.cpp can run the very same
// sequencing with one host thread per rank (a thread executes its "stream" in order; the flags are the same 64-bit words)
// and tests/test_peer_protocol.py can check it for lost / overwritten / reordered messages, data races (ThreadSanitizer)
// and dead-locks without a GPU.
//
// One block per link (= one neighbour = one (dim, side)), allocated by the RECEIVER and mapped by the sender:
//      [ data_seq | pad | cookie | pad to 256 B ][ slot 0 : cap bytes ][ slot 1 : cap bytes ]   (rounded up to 2 MiB)
//   data_seq  written by the peer : "your slot (k & 1) holds my message k"           (release, system scope)
//   slots     written by the peer's pack kernel, read by the local unpack kernel
// Message k (1, 2, ...) of a link travels in slot k & 1.
//
// Stream order on every rank, per dimension (D = N..1 as bc! demands, batch.jl:20-29; both sides in one pass):
//      push           per side   : pack kernel, destination = the peer's slot k & 1
//      post_and_wait  both links : data_seq(peer) := k ; then data_seq(local) >= k     (acquire, system scope)
//      unpack         per side   : source = the local slot k & 1
//
// Why two slots and no acknowledgement: a link always carries one message in EACH direction per exchange (my side S talks
// to the neighbour's side 1-S, exchange_halo.jl:101-108).  Rank A overwrites slot k & 1 with message k + 2 only after its
// wait for B's message k + 1, which B posts -- in B's stream order -- after it has unpacked A's message k.  So the data
// flags alone order every overwrite after the read it would clobber; with ONE slot they would not (A may push k + 1 while
// B still unpacks k), which the emulation demonstrates (-DPL_SLOTS=1 must fail).
// Dead-lock freedom: pushes never wait; every wait of A is for a post that B issues after pushes only.
#pragma once

#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PL_HD __host__ __device__ __forceinline__
#else
#define PL_HD inline
#endif

#ifndef PL_SLOTS
#define PL_SLOTS 2
#endif
#define PL_FLAG_BYTES 256
#define PL_OFF_DATA   0
#define PL_OFF_COOKIE 128            // a nonce the owner writes and the mapper reads back: proves the mapping addresses THIS block
#define PL_ALIGN      256
#define PL_BLOCK_GRAIN (2u << 20)    // blocks are whole 2 MiB pages of their own (never sub-allocated next to other data)

PL_HD int    pl_slot(uint64_t k) { return (int)(k % PL_SLOTS); }
PL_HD size_t pl_round_cap(size_t bytes) { return (bytes + (PL_ALIGN - 1)) / PL_ALIGN * PL_ALIGN; }
PL_HD size_t pl_block_bytes(size_t cap) {
    const size_t b = (size_t)PL_FLAG_BYTES + (size_t)PL_SLOTS * cap;
    return (b + (PL_BLOCK_GRAIN - 1)) / PL_BLOCK_GRAIN * PL_BLOCK_GRAIN;
}
PL_HD size_t pl_off_slot(int slot, size_t cap) { return (size_t)PL_FLAG_BYTES + (size_t)slot * cap; }

// Slot capacity for a message of `need` bytes on a link whose slots hold `cap` bytes: unchanged while it fits, else the
// need plus a quarter (both ends of a link see the same message sizes, so both re-allocate in the same exchange).
PL_HD size_t pl_grow_cap(size_t cap, size_t need) {
    if (need <= cap) return cap;
    return pl_round_cap(need + need / 4);
}

enum { PL_MODE_UNSET = 0, PL_MODE_PEER = 1, PL_MODE_NCCL = 2 };

// host-side state of one link
struct PlLink {
    int      peer;         // neighbour's rank, -1 == none
    int      mode;         // PL_MODE_*: NCCL when the block could not be mapped on either end (agreed by both)
    char*    local;        // this rank's block (the peer writes it)
    char*    remote;       // the peer's block, mapped into this process
    size_t   cap;          // bytes per slot
    uint64_t seq;          // messages exchanged over this link since its block was (re)allocated
};

PL_HD uint64_t* pl_flag(char* block, int off) { return reinterpret_cast<uint64_t*>(block + off); }

// One dimension's exchange over the PEER links among link[0..1] (nullptr = side not exchanged or not a PEER link).
// OPS supplies the three stream operations; each returns 0 or an error code that aborts the exchange.
//   int push(int side, PlLink& l, int slot);
//   int post_and_wait(PlLink* const l[2], const uint64_t k[2]);
//   int unpack(int side, PlLink& l, int slot);
template <class OPS>
inline int pl_exchange_dim(OPS& ops, PlLink* const link[2]) {
    uint64_t k[2] = {0, 0};
    bool any = false;
    for (int s = 0; s < 2; ++s) {
        if (!link[s]) continue;
        k[s] = ++link[s]->seq;
        any  = true;
    }
    if (!any) return 0;
    int rc = 0;
    for (int s = 0; s < 2; ++s)
        if (link[s] && (rc = ops.push(s, *link[s], pl_slot(k[s])))) return rc;
    if ((rc = ops.post_and_wait(link, k))) return rc;
    for (int s = 0; s < 2; ++s)
        if (link[s] && (rc = ops.unpack(s, *link[s], pl_slot(k[s])))) return rc;
    return 0;
}
