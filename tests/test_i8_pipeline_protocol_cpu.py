"""Thread-level model of the mbarrier protocol of csrc/gemm_i8.cu (experimental, not yet run on hardware): the two TMA producer
rings (served in the kernel by ONE thread with non-blocking polls; modelled here as two free-running threads, a superset of
its interleavings), 256 converter threads (modelled as 8 warps arriving 32 times each), the single MMA thread and the 4 epilogue warps,
with the kernel's barrier counts, stage counts and wait parities transliterated.  Every stage buffer carries a
(tile, k-tile) tag: consuming a stale or overwritten buffer, a wrong parity or a dead-lock fails the test."""
import random
import threading
import time

import pytest

FST, DST = 3, 2   # I8_FSTAGES, I8_DSTAGES


class MBar:
    """mbarrier: `count` arrivals (+ optional transaction bytes) complete a phase; try_wait.parity(p) is true once the phase of
    parity p has completed, i.e. the barrier's current phase parity differs from p."""

    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0
        self.cv = threading.Condition()

    def _maybe_flip(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count
            self.cv.notify_all()

    def arrive(self, n=1, expect_tx=0):
        with self.cv:
            self.tx += expect_tx
            self.pending -= n
            assert self.pending >= 0, "more arrivals than the barrier was initialised for"
            self._maybe_flip()

    def complete_tx(self, nbytes):
        with self.cv:
            self.tx -= nbytes
            self._maybe_flip()

    def wait(self, parity):
        with self.cv:
            ok = self.cv.wait_for(lambda: (self.phase & 1) != parity, timeout=20)
            assert ok, "dead-lock"


def jitter(rng):
    if rng.random() < 0.25:
        time.sleep(rng.random() * 1e-4)


@pytest.mark.parametrize("tiles,ktc", [(1, 1), (1, 7), (3, 4), (4, 1), (2, 9)])
def test_i8_kernel_barrier_protocol(tiles, ktc):
    full_f = [MBar(1) for _ in range(FST)]
    empty_f = [MBar(256) for _ in range(FST)]
    full_d = [MBar(256 + 1) for _ in range(DST)]
    empty_d = [MBar(1) for _ in range(DST)]
    acc_full, acc_empty = MBar(1), MBar(128)
    F = [None] * FST          # FP64 stage contents
    A = [None] * DST          # A digit planes (tag written by the converters)
    B = [None] * DST          # B digit planes (tag written by the producer)
    TMEM = {"tile": None, "kts": []}
    errors = []
    consumed = []

    def guard(fn):
        def run():
            try:
                fn()
            except Exception as ex:  # noqa: BLE001
                errors.append(repr(ex))
        return run

    def producer_f():
        rng = random.Random(1)
        it = 0
        for w in range(tiles):
            for kt in range(ktc):
                sf = it % FST
                if it >= FST:
                    empty_f[sf].wait((it // FST - 1) & 1)
                full_f[sf].arrive(1, expect_tx=32768)
                jitter(rng)
                F[sf] = (w, kt)                       # TMA lands
                full_f[sf].complete_tx(32768)
                it += 1

    def producer_b():
        rng = random.Random(2)
        it = 0
        for w in range(tiles):
            for kt in range(ktc):
                sd = it % DST
                if it >= DST:
                    empty_d[sd].wait((it // DST - 1) & 1)
                full_d[sd].arrive(1, expect_tx=14336)
                jitter(rng)
                B[sd] = (w, kt)
                full_d[sd].complete_tx(14336)
                it += 1

    def mma():
        rng = random.Random(3)
        it = 0
        for w in range(tiles):
            if w > 0:
                acc_empty.wait((w - 1) & 1)
            TMEM["tile"], TMEM["kts"] = w, []
            for kt in range(ktc):
                sd = it % DST
                full_d[sd].wait((it // DST) & 1)
                assert A[sd] == (w, kt) and B[sd] == (w, kt), ("MMA read a wrong digit slot", w, kt, A[sd], B[sd])
                jitter(rng)
                TMEM["kts"].append(kt)
                consumed.append((w, kt))
                empty_d[sd].arrive()                  # tcgen05.commit
                it += 1
            acc_full.arrive()                         # tcgen05.commit after the last k-tile

    def converter(warp):
        def run():
            rng = random.Random(10 + warp)
            it = 0
            for w in range(tiles):
                for kt in range(ktc):
                    sf, sd = it % FST, it % DST
                    full_f[sf].wait((it // FST) & 1)
                    if it >= DST:
                        empty_d[sd].wait((it // DST - 1) & 1)
                    assert F[sf] == (w, kt), ("converter read a wrong FP64 stage", warp, w, kt, F[sf])
                    jitter(rng)
                    if A[sd] is not None and warp == 0:
                        pass
                    A[sd] = (w, kt)
                    full_d[sd].arrive(32)             # 32 lanes of this warp
                    empty_f[sf].arrive(32)
                    it += 1
                if warp < 4:
                    acc_full.wait(w & 1)
                    assert TMEM["tile"] == w and TMEM["kts"] == list(range(ktc)), ("epilogue saw an incomplete accumulator", w, TMEM)
                    jitter(rng)
                    acc_empty.arrive(32)
        return run

    threads = [threading.Thread(target=guard(f), daemon=True) for f in [producer_f, producer_b, mma] + [converter(wp) for wp in range(8)]]
    for t in threads:
        t.start()
    for t in threads:
        t.join(60)
    assert not any(t.is_alive() for t in threads), "dead-lock"
    assert not errors, errors
    assert consumed == [(w, kt) for w in range(tiles) for kt in range(ktc)]
