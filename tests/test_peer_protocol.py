"""
CPU proof of the peer-store halo exchange's protocol (chmy.jl_b200/csrc/peer_link.cuh -> comm.cu, CHMY_EXCHANGE_PEER).

The header holds everything that decides order -- pl_exchange_dim (push both sides, post + wait the sequence flags, unpack),
the slot of a message, the block layout and the growth rule -- as plain C++.  tests/emul/peer_emul.cpp runs it with one
host thread per rank of a Cartesian topology (a thread = a stream: in-order), plain stores / loads for the payload and
release / acquire atomics for the flags, random delays, and checks EVERY word of EVERY received message against what the
neighbour must have packed for exactly that exchange (dimension, side, iteration).

  * two slots (the shipped protocol): no lost, overwritten or reordered message, no time-out (= no dead-lock), the
    re-allocation hand-shake pairs up on both ends of a link -- on chains, planes and the (2,2,2) box of the 8-GPU run;
  * the same runs under ThreadSanitizer: no data race on any slot, i.e. the flags alone order every overwrite after the
    read it would clobber (the argument in the header);
  * one slot (-DPL_SLOTS=1) must FAIL both ways: the test can see the failure it is there to exclude.

What this does not cover: CUDA IPC mapping, st.release.sys / ld.acquire.sys over NVLink and the kernels' launch order on
real streams -- the gated multi-GPU test (tests/test_z_b200_multigpu.py, CHMY_EXPERIMENTAL=1) is their first run.
"""
import json
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "peer_emul.cpp")
HDR = os.path.join(HERE, "..", "chmy.jl_b200", "csrc", "peer_link.cuh")

VARIANTS = {
    "peer_emul": ["-O2"],
    "peer_emul_tsan": ["-O1", "-g", "-fsanitize=thread"],
    "peer_emul_asan": ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"],
    "peer_emul_1slot": ["-O2", "-DPL_SLOTS=1"],
    "peer_emul_1slot_tsan": ["-O1", "-g", "-fsanitize=thread", "-DPL_SLOTS=1"],
}


def build(name):
    exe = os.path.join(HERE, "emul", name + ".bin")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-std=c++20", "-pthread", "-Wall"] + VARIANTS[name] + ["-o", exe, SRC])
    return exe


def run(name, dims, iters, seed, delay_us, words, grow_every=0, slow_unpack_us=0, timeout_s=20):
    exe = build(name)
    env = dict(os.environ, PEER_EMUL_TIMEOUT_S=str(timeout_s), TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0")
    px, py, pz = (list(dims) + [1, 1])[:3]
    r = subprocess.run([exe] + [str(x) for x in (px, py, pz, iters, seed, delay_us, words, grow_every, slow_unpack_us)],
                       capture_output=True, text=True, env=env, timeout=300)
    if "FATAL: ThreadSanitizer" in r.stderr:      # the sanitizer runtime cannot start here (address-space layout): not a verdict
        pytest.skip("ThreadSanitizer cannot run in this environment: " + r.stderr.strip().splitlines()[0][:200])
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    return r.returncode, json.loads(lines[-1]), r.stderr


def links(dims):
    """directed messages per iteration: every rank sends one message to each of its neighbours"""
    n = 1
    for d in dims:
        n *= d
    total = 0
    for a, d in enumerate(dims):
        total += 2 * (d - 1) * (n // d)
    return total


TOPOLOGIES = [(2,), (3,), (2, 2), (4, 2), (2, 2, 2), (3, 2, 2)]      # MPI.Dims_create shapes of 2 / 4 / 8 ranks and odd ones


@pytest.mark.parametrize("dims", TOPOLOGIES)
def test_no_message_is_lost_overwritten_or_reordered(dims):
    for seed, delay in ((1, 0), (2, 25), (3, 200)):
        iters = 1500 if delay < 100 else 150
        rc, out, err = run("peer_emul", dims, iters, seed, delay, 96)
        assert rc == 0 and out["mismatches"] == 0 and out["timeouts"] == 0 and out["handshake_errors"] == 0, (out, err[-500:])
        assert out["slots"] == 2 and out["messages"] == links(dims) * iters


@pytest.mark.parametrize("dims", [(2,), (2, 2), (2, 2, 2)])
def test_one_rank_far_slower_than_its_neighbours(dims):
    """odd ranks take 300 us per unpack: their neighbours run ahead as far as the protocol lets them (one exchange)"""
    rc, out, err = run("peer_emul", dims, 120, 7, 5, 512, slow_unpack_us=300)
    assert rc == 0 and out["mismatches"] == 0 and out["timeouts"] == 0, (out, err[-500:])


@pytest.mark.parametrize("dims", [(2,), (4, 2), (2, 2, 2)])
def test_growing_messages_pair_their_reallocation_handshakes(dims):
    """messages grow every 40 iterations: both ends of a link must re-allocate in the same exchange (pl_grow_cap is a
    function of the link's message sizes only), restart the link's sequence and never touch a retired block's slots"""
    rc, out, err = run("peer_emul", dims, 400, 11, 20, 64, grow_every=40)
    assert rc == 0 and out["mismatches"] == 0 and out["timeouts"] == 0 and out["handshake_errors"] == 0, (out, err[-500:])
    assert out["regrows"] > 0


@pytest.mark.parametrize("dims", [(2,), (2, 2), (2, 2, 2)])
def test_thread_sanitizer_sees_no_race_on_the_slots(dims):
    for kw in (dict(grow_every=0, slow_unpack_us=0), dict(grow_every=25, slow_unpack_us=0), dict(grow_every=0, slow_unpack_us=200)):
        rc, out, err = run("peer_emul_tsan", dims, 150, 5, 20, 128, **kw)
        assert "data race" not in err and "ThreadSanitizer" not in err, err[-3000:]
        assert rc == 0 and out["mismatches"] == 0 and out["timeouts"] == 0, out


def test_address_sanitizer_every_slot_access_inside_its_block():
    """slot offsets (pl_off_slot), capacities (pl_grow_cap) and block sizes (pl_block_bytes) agree: no access outside a
    block, none to a retired block, for growing messages of odd sizes"""
    for dims, words, grow in (((2, 2, 2), 96, 30), ((3, 2), 1, 7), ((2,), 33, 5)):
        rc, out, err = run("peer_emul_asan", dims, 300, 9, 10, words, grow_every=grow)
        assert "AddressSanitizer" not in err and "runtime error" not in err, err[-3000:]
        assert rc == 0 and out["mismatches"] == 0 and out["regrows"] > 0, out


def test_a_single_slot_fails_as_the_header_says():
    """Sensitivity: with ONE slot rank A may push message k + 1 while B still unpacks k.  The payload check and
    ThreadSanitizer must both see it -- otherwise the tests above prove nothing."""
    rc, out, err = run("peer_emul_1slot", (2,), 400, 4, 5, 512, slow_unpack_us=300)
    assert out["slots"] == 1 and out["mismatches"] > 0 and rc != 0, out
    rc, out, err = run("peer_emul_1slot_tsan", (2,), 100, 4, 5, 512, slow_unpack_us=300)
    assert "data race" in err, (out, err[-1000:])


def test_a_lost_neighbour_times_out_instead_of_hanging():
    """k_pl_flags gives up after CHMY_PEER_TIMEOUT_S and raises the context's error flag; the emulation's wait has the same
    rule.  Rank 1 of a (3,) chain stops exchanging at iteration 20: both neighbours must end with a time-out, not hang."""
    exe = build("peer_emul")
    env = dict(os.environ, PEER_EMUL_TIMEOUT_S="1", PEER_EMUL_DIE="1:20")
    r = subprocess.run([exe, "3", "1", "1", "50", "1", "0", "64", "0", "0"], capture_output=True, text=True, env=env, timeout=60)
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert r.returncode == 1 and out["timeouts"] >= 1 and out["mismatches"] == 0, out
    assert out["messages"] < links((3,)) * 50


def test_transport_knobs_refuse_bad_arguments_without_a_device():
    import ctypes as C
    import chmy_b200
    lib = chmy_b200.load_library()
    a, b = C.c_uint64(7), C.c_uint64(7)
    assert lib.chmy_set_exchange_mode(None, 1) != 0 and lib.chmy_exchange_stats(None, C.byref(a), C.byref(b)) != 0


def test_empty_messages_terminate():
    rc, out, err = run("peer_emul", (3,), 50, 1, 0, 0)
    assert rc == 0 and out["timeouts"] == 0, out
