"""Protocol-level model (CPU, threads) of the peer-memory exchanges in csrc/api.cu (fused all-reduce + solve) and
csrc/peer_graph.cu (device-side epochs, per-exchange slots, small all-reduce): G "ranks" run their in-order kernel sequences
with random delays against shared flag / slot arrays.  Every value written into a slot is tagged with (sweep, exchange), so a
reader that sees a stale or overwritten slot fails the test; a dead-lock fails by timeout.  This checks the slot-reuse
argument written in DESIGN.md section 8, not the CUDA code itself."""
import random
import threading
import time

import pytest


class Box:
    """one rank's exchange buffer: flags[0:G] for partial-M exchanges, flags[16:16+G] for the small all-reduce"""

    def __init__(self, nslots):
        self.flags = [0] * 32
        self.slots = [None] * nslots
        self.small = [None, None]


def run_ranks(G, body, timeout=60):
    errors = []

    def wrap(r):
        try:
            body(r)
        except Exception as ex:  # noqa: BLE001
            errors.append((r, repr(ex)))

    th = [threading.Thread(target=wrap, args=(r,), daemon=True) for r in range(G)]
    for t in th:
        t.start()
    t0 = time.time()
    for t in th:
        t.join(max(0.0, timeout - (time.time() - t0)))
    assert not any(t.is_alive() for t in th), "dead-lock: a rank never finished"
    assert not errors, errors


def jitter(rng):
    if rng.random() < 0.3:
        time.sleep(rng.random() * 2e-4)


def spin(cond):
    t0 = time.time()
    while not cond():
        if time.time() - t0 > 20:
            raise TimeoutError("peer never published")
        time.sleep(0)


@pytest.mark.parametrize("G,nexch", [(2, 2), (4, 2), (8, 2), (3, 3)])
def test_default_protocol_host_epoch_two_parity_slots(G, nexch):
    """api.cu mode_update_device: epoch = ++host counter, slot = epoch & 1, signal kernel, wait inside the solve kernel."""
    boxes = [Box(2) for _ in range(G)]
    sweeps = 40

    def body(r):
        rng = random.Random(1000 + r)
        epoch = 0
        for s in range(sweeps):
            for x in range(nexch):
                epoch += 1
                jitter(rng)
                boxes[r].slots[epoch & 1] = (s, x, r)              # second-level kernel writes my partial into my slot
                for q in range(G):                                 # peer_signal_kernel
                    boxes[q].flags[r] = epoch
                jitter(rng)
                spin(lambda: all(boxes[r].flags[q] >= epoch for q in range(G)))   # solve kernel: wait, then read every peer
                for q in range(G):
                    jitter(rng)
                    got = boxes[q].slots[epoch & 1]
                    assert got == (s, x, q), (r, s, x, q, got)

    run_ranks(G, body)


@pytest.mark.parametrize("G,order", [(2, 3), (4, 3), (8, 3), (4, 4)])
def test_peer_graph_protocol_device_epoch_slot_per_exchange_and_small_allreduce(G, order):
    """peer_graph.cu: slot = exchange index within the sweep (order-1 slots, no parity), device-side epochs, and two small
    all-reduces (column norms, Gram) per sweep, double buffered by the parity of their own device epoch."""
    nexch = order - 1
    boxes = [Box(max(2, nexch)) for _ in range(G)]
    sweeps = 40

    def body(r):
        rng = random.Random(2000 + r)
        epoch_dev, small_dev = 0, 0                                # live in device memory, advanced by the kernels themselves
        for s in range(sweeps):
            for x in range(nexch):
                jitter(rng)
                boxes[r].slots[x] = (s, x, r)
                epoch_dev += 1                                     # peer_signal_dev_kernel
                for q in range(G):
                    boxes[q].flags[r] = epoch_dev
                jitter(rng)
                e = epoch_dev                                      # peer_wait_dev_kernel
                spin(lambda: all(boxes[r].flags[q] >= e for q in range(G)))
                for q in range(G):                                 # row-solve kernel sums the peers' slots
                    jitter(rng)
                    got = boxes[q].slots[x]
                    assert got == (s, x, q), (r, s, x, q, got)
            for which in range(2):                                 # peer_allreduce_small_kernel (last mode): norms, Gram
                small_dev += 1
                e = small_dev
                jitter(rng)
                boxes[r].small[e & 1] = (s, which, r)
                for q in range(G):
                    boxes[q].flags[16 + r] = e
                spin(lambda: all(boxes[r].flags[16 + q] >= e for q in range(G)))
                for q in range(G):
                    jitter(rng)
                    got = boxes[q].small[e & 1]
                    assert got == (s, which, q), (r, s, which, q, got)

    run_ranks(G, body)


def test_single_slot_would_be_unsafe():
    """Negative control: with ONE slot and no other exchange in between, a fast rank overwrites what a slow rank still reads --
    the model must be able to see that (otherwise the two tests above prove nothing)."""
    G = 2
    boxes = [Box(1) for _ in range(G)]
    seen_stale = []

    def body(r):
        epoch = 0
        for s in range(200):
            epoch += 1
            boxes[r].slots[0] = (s, r)
            for q in range(G):
                boxes[q].flags[r] = epoch
            spin(lambda: all(boxes[r].flags[q] >= epoch for q in range(G)))
            if r == 1:
                time.sleep(2e-4)                                   # slow reader
            for q in range(G):
                if boxes[q].slots[0] != (s, q):
                    seen_stale.append((r, s, q))

    run_ranks(G, body)
    assert seen_stale, "the model failed to expose the single-slot race"
