"""Randomised discrete-event model of the synchronisation protocol of attn_bwd_pipe_kernel (ops/csrc/attention.cu):
one MMA/TMA thread, T math threads, mbarriers with parities, in-order asynchronous tensor pipe, asynchronous TMA.
Checks, over many random schedules and latencies: no deadlock, no mbarrier phase overrun (a waiter can never be a
full phase behind), and no data hazard on S/dP (TMEM), dQ (TMEM), P~/dS (shared memory) and the Q/dO ring slots.

    python tools/sim_attn_bwd_pipe.py [--nqb 4] [--threads 4] [--trials 2000]

It models the *protocol* (who waits for what, in which order), not the arithmetic; it exists because the kernel was
written without GPU access.
"""
import argparse
import random


class Barrier:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0      # phase = number of completed phases

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "too many arrivals in one phase"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count


class Sim:
    def __init__(self, nqb, T, rng, mutate=0):
        self.nqb, self.T, self.rng, self.mutate = nqb, T, rng, mutate
        B = Barrier
        self.bar = dict(kv_full=B(1), sdp_ready=B(1), sdp_free=B(T), ds_ready=B(T), dq_ready=B(1), fin=B(1))
        for s in range(3):
            self.bar[f"qdo_full{s}"] = B(1)
            self.bar[f"qdo_empty{s}"] = B(1)
        # data state
        self.slot_block = [None, None, None]          # which query block each Q/dO ring slot holds
        self.slot_readers = [0, 0, 0]                 # MMA batches in flight that read the slot
        self.sdp_version = None                       # block whose S/dP sit in TMEM (None while being written)
        self.sdp_read = set()                         # math threads that have pulled the current S/dP
        self.dq_version = None
        self.dq_read = set()
        self.pds_version = None                       # block whose P~/dS sit in shared memory (complete)
        self.pds_written = set()
        self.pds_readers = 0
        # async engines
        self.pipe = []                                # in-order tensor pipe: list of dict(kind, block, remaining)
        self.tma = []                                 # independent loads: dict(slot, block, remaining)
        self.time = 0

    # ---- mbarrier wait with the kernel's parity semantics ----------------------------------------------
    def ready(self, name, use):
        """Can the `use`-th wait (0-based) on barrier `name` pass?  Also assert the waiter is not overrun."""
        ph = self.bar[name].phase
        assert ph <= use + 1, f"phase overrun on {name}: waiter at use {use}, barrier at phase {ph}"
        return ph >= use + 1

    # ---- asynchronous engines -----------------------------------------------------------------------------
    def issue(self, kind, block, slot=None, commits=()):
        self.pipe.append(dict(kind=kind, block=block, slot=slot, commits=list(commits), rem=self.rng.randint(1, 6), started=False))
        if slot is not None:
            self.slot_readers[slot] += 1

    def start_batch(self, b):
        b["started"] = True
        k, blk = b["kind"], b["block"]
        if k == "sdp":
            assert self.slot_block[b["slot"]] == blk, f"S/dP({blk}) reads slot {b['slot']} holding {self.slot_block[b['slot']]}"
            if self.sdp_version is not None or blk > 0:
                assert len(self.sdp_read) == self.T, f"S/dP({blk}) overwrites S/dP({self.sdp_version}) before all threads read it"
            self.sdp_version, self.sdp_read = None, set()
        elif k == "grad":
            assert self.slot_block[b["slot"]] == blk, f"dV/dK({blk}) reads slot {b['slot']} holding {self.slot_block[b['slot']]}"
            assert self.pds_version == blk and len(self.pds_written) == self.T, f"dV/dK/dQ({blk}) reads incomplete P~/dS ({self.pds_version})"
            self.pds_readers += 1
            if blk > 0:
                assert self.dq_version == blk - 1 and len(self.dq_read) == self.T, f"dQ({blk}) overwrites dQ({self.dq_version}) before it was drained"
            self.dq_version, self.dq_read = None, set()

    def finish_batch(self, b):
        k, blk = b["kind"], b["block"]
        if k == "sdp":
            self.sdp_version = blk
        elif k == "grad":
            self.dq_version = blk
            self.pds_readers -= 1
        if b["slot"] is not None:
            self.slot_readers[b["slot"]] -= 1
        for name in b["commits"]:
            self.bar[name].arrive()

    def advance_async(self):
        if self.pipe:
            b = self.pipe[0]
            if not b["started"]:
                self.start_batch(b)
            b["rem"] -= 1
            if b["rem"] <= 0:
                self.finish_batch(self.pipe.pop(0))
        for ld in list(self.tma):
            ld["rem"] -= 1
            if ld["rem"] <= 0:
                self.tma.remove(ld)
                self.slot_block[ld["slot"]] = ld["block"]
                self.bar[f"qdo_full{ld['slot']}"].arrive()

    def load(self, slot, block):
        assert self.slot_readers[slot] == 0, f"TMA overwrites ring slot {slot} (block {self.slot_block[slot]}) while MMAs still read it"
        self.slot_block[slot] = None
        self.tma.append(dict(slot=slot, block=block, rem=self.rng.randint(1, 12)))

    # ---- the two programs as generators; `yield cond` = spin until cond() --------------------------------
    def mma_thread(self):
        n = self.nqb
        self.bar["kv_full"].arrive()                  # K/V load (never reused) modelled as immediate
        for j in range(min(3, n)):
            self.load(j, j)

        def issue_sdp(j):
            yield lambda: self.ready(f"qdo_full{j % 3}", j // 3)
            self.issue("sdp", j, slot=j % 3, commits=["sdp_ready"])
        yield lambda: self.ready("kv_full", 0)
        yield from issue_sdp(0)
        for i in range(n):
            if i + 1 < n:
                if self.mutate != 2:                  # mutation 2: next S/dP without waiting for the readers
                    yield lambda: self.ready("sdp_free", i)
                yield from issue_sdp(i + 1)
            yield lambda: self.ready("ds_ready", i)
            self.issue("grad", i, slot=i % 3, commits=["dq_ready", f"qdo_empty{i % 3}"])
            if i >= 1 and i + 2 < n:
                sp = (i - 1) % 3
                if self.mutate != 3:                  # mutation 3: reload a ring slot without waiting for its readers
                    yield lambda: self.ready(f"qdo_empty{sp}", (i - 1) // 3)
                else:
                    sp = i % 3
                self.load(sp, i + 2)
        self.issue("fin", n, commits=["fin"])

    def math_thread(self, tid):
        n = self.nqb

        def drain(j):
            yield lambda: self.ready("dq_ready", j)
            assert self.dq_version == j, f"thread {tid} drains dQ({j}) but TMEM holds {self.dq_version}"
            self.dq_read.add(tid)
        for i in range(n):
            yield lambda: self.ready("sdp_ready", i)
            assert self.sdp_version == i, f"thread {tid} reads S/dP({i}) but TMEM holds {self.sdp_version}"
            yield lambda: True                        # the loads take a while
            self.sdp_read.add(tid)
            self.bar["sdp_free"].arrive()
            yield lambda: True                        # math in registers
            if i > 0 and self.mutate != 1:            # mutation 1: P~/dS stored without draining the previous dQ first
                yield from drain(i - 1)
            assert self.pds_readers == 0, f"thread {tid} overwrites P~/dS while MMAs of block {self.pds_version} read them"
            if self.pds_version != i:
                self.pds_version, self.pds_written = i, set()
            self.pds_written.add(tid)
            self.bar["ds_ready"].arrive()
        yield from drain(n - 1)
        yield lambda: self.ready("fin", 0)

    def run(self):
        progs = [self.mma_thread()] + [self.math_thread(t) for t in range(self.T)]
        conds = [None] * len(progs)
        alive = set(range(len(progs)))
        idle = 0
        while alive:
            self.advance_async()
            order = list(alive)
            self.rng.shuffle(order)
            moved = False
            for k in order:
                if self.rng.random() < 0.35:          # this thread is not scheduled in this tick
                    continue
                if conds[k] is not None and not conds[k]():
                    continue
                try:
                    conds[k] = next(progs[k])
                    moved = True
                except StopIteration:
                    alive.discard(k)
                    moved = True
            idle = 0 if (moved or self.pipe or self.tma) else idle + 1
            assert idle < 200, "deadlock: nothing can make progress"
        assert not self.pipe and not self.tma


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nqb", type=int, default=0, help="query blocks (0: sweep 1..6)")
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--trials", type=int, default=2000)
    ap.add_argument("--mutate", type=int, default=0,
                    help="self-test: break the protocol on purpose (1: no dQ drain before the P~/dS stores, 2: no sdp_free "
                         "wait, 3: early ring reload); the model must then report a hazard")
    a = ap.parse_args()
    for nqb in ([a.nqb] if a.nqb else range(1, 7)):
        caught = 0
        for t in range(a.trials):
            try:
                Sim(nqb, a.threads, random.Random(t * 7919 + nqb), a.mutate).run()
            except AssertionError as e:
                if not a.mutate:
                    raise
                caught += 1
                last = str(e)
        if a.mutate:
            print(f"nqb={nqb}: mutation {a.mutate} detected in {caught}/{a.trials} schedules" + (f" (e.g. {last})" if caught else ""))
        else:
            print(f"nqb={nqb}: {a.trials} random schedules, no deadlock / overrun / hazard")


if __name__ == "__main__":
    main()
