"""Scheduling logic of the continuous batcher (llava/_b2/batching.py) on the CPU: admit / step / retire, slot reuse, per-request
cancellation and error isolation, with a stand-in engine whose "tokens" encode (request, index) so every delivery can be checked.
The GPU behaviour of the same class is covered in tests/test_generate_gpu.py."""
import contextlib
import threading
import time

import pytest
import torch

from llava._b2.batching import ContinuousBatcher, RequestStream


class _FakeCuda:
    class Stream:
        def __init__(self, device=None):
            self.waited = 0

        def wait_event(self, ev):
            self.waited += 1

    class Event:
        def record(self):
            pass

    @staticmethod
    @contextlib.contextmanager
    def stream(s):
        yield


class _KV:
    closed = False

    def close(self):
        self.closed = True


class FakeEngine:
    """Token of request r at index i is 1000*r + i; r travels in embeds[0, 0, 0]."""
    device = "cpu"

    def __init__(self, slots, fail_request=None, step_delay=0.0):
        self.rows = [None] * slots          # slot -> [request id, next index]
        self.fail_request, self.step_delay = fail_request, step_delay
        self.prefills, self.steps = [], 0
        self.lock = threading.Lock()

    def new_kv(self, slots, max_seq):
        return _KV()

    def batch_begin(self, kv, slots):
        pass

    def prefill(self, kv, embeds, lens, mode, slot0=0):
        r = int(embeds[0, 0, 0])
        if r == self.fail_request:
            raise ValueError(f"request {r}: bad prompt")
        self.prefills.append((r, slot0, int(lens[0])))
        return torch.tensor([[float(r)]])

    def sample(self, logits, sampling, index=0):
        return torch.tensor([1000 * int(logits[0, 0]) + index])

    def check_async_error(self):
        pass

    def batch_set_row(self, kv, slot, active, sampling=None, first=None):
        with self.lock:
            self.rows[slot] = [first // 1000, 1] if active else None

    def stream_enqueue(self, kv, n):
        assert n == 1

    def stream_wait(self, kv, index, B, timeout_ms=0):
        assert index == self.steps, "ring entries must be consumed in order"
        time.sleep(self.step_delay)
        out = []
        with self.lock:
            for row in self.rows:
                if row is None:
                    out.append(-1)
                else:
                    out.append(1000 * row[0] + row[1])
                    row[1] += 1
        self.steps += 1
        return out


def _embeds(r):
    return torch.full((1, 4, 2), float(r))


def _drain(req, n):
    return [req.get(timeout=10) for _ in range(n)]


def test_requests_share_steps_and_slots_are_reused():
    eng = FakeEngine(slots=2)
    b = ContinuousBatcher(eng, slots=2, max_seq=64, _cuda=_FakeCuda)
    try:
        want = {1: 5, 2: 3, 3: 7, 4: 1, 5: 4}
        reqs = {r: b.submit(_embeds(r), 4, max_new_tokens=n) for r, n in want.items()}
        got = {r: _drain(reqs[r], want[r]) for r in want}
        for r, n in want.items():
            assert got[r] == [1000 * r + i for i in range(n)], (r, got[r])   # own tokens, in order, none from a neighbour
        assert b.stats["admitted"] == 5 and b.stats["max_active"] == 2
        assert sorted(p[0] for p in eng.prefills) == [1, 2, 3, 4, 5]
        assert {p[1] for p in eng.prefills} == {0, 1}                         # five requests went through two slots
        assert b.stats["rows_stepped"] >= sum(n - 1 for n in want.values())
        assert b.stream.waited == 5                                           # every admission waited for its producer's event
    finally:
        b.close()
    assert b.kv.closed


def test_cancel_frees_the_slot_and_a_bad_request_fails_alone():
    eng = FakeEngine(slots=2, fail_request=2, step_delay=0.002)
    b = ContinuousBatcher(eng, slots=2, max_seq=20_000, _cuda=_FakeCuda)
    try:
        long_ = b.submit(_embeds(1), 4, max_new_tokens=10_000)
        bad = b.submit(_embeds(2), 4, max_new_tokens=5)
        with pytest.raises(ValueError, match="request 2"):
            bad.get(timeout=10)
        assert _drain(long_, 3) == [1000, 1001, 1002]
        ok = b.submit(_embeds(3), 4, max_new_tokens=4)                          # takes the slot the bad request gave back
        assert _drain(ok, 4) == [3000, 3001, 3002, 3003]
        long_.cancel()                                                          # consumer saw eos / a stopping criterion
        nxt = b.submit(_embeds(4), 4, max_new_tokens=2)
        assert _drain(nxt, 2) == [4000, 4001]
        deadline = time.time() + 5
        while b.active and time.time() < deadline:
            time.sleep(0.01)
        assert not b.active and sorted(b.free) == [0, 1]
    finally:
        b.close()


def test_limits_and_request_stream_view():
    eng = FakeEngine(slots=2)
    with pytest.raises(ValueError):
        ContinuousBatcher(eng, slots=1, max_seq=64, _cuda=_FakeCuda)
    b = ContinuousBatcher(eng, slots=2, max_seq=16, _cuda=_FakeCuda)
    try:
        with pytest.raises(ValueError, match="exceeds"):
            b.submit(_embeds(1), 10, max_new_tokens=7)
        req = b.submit(_embeds(7), 4, max_new_tokens=3)
        view = RequestStream(req)                                               # what generate()'s host loop drives
        view.stream_begin(None, None, None)
        view.stream_enqueue(None, 5)
        assert [view.stream_wait(None, i, 1)[0] for i in range(3)] == [7000, 7001, 7002]
    finally:
        b.close()
    with pytest.raises(RuntimeError):
        b.submit(_embeds(8), 4)
