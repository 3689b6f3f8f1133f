"""CPU: host-side logic of generate()'s decode loop (`_stream_decode`: run-ahead scheduling over the streaming C-ABI, eos /
pad bookkeeping, streamer and stopping-criteria protocol) against a plain per-step restatement of the loop the reference's
callers get from HF generate: finished rows show the pad id, generation stops at the step where every row
has finished or a criterion fires, the result never exceeds max_new_tokens, the device is never asked for more steps than
max_new_tokens - 1 and never runs more than `run_ahead` tokens in front of the host."""
import torch

from llava.model.language_model.llava_llama import _stream_decode


def _next(tok):  # deterministic fake "model": next token is a function of the fed token only (rows independent)
    return (tok * 7 + 3) % 23


class FakeEngine:
    """stream_begin / stream_enqueue / stream_wait with the contract of include/b2llava.h, tokens computed eagerly."""

    def __init__(self, first):
        self.first = [int(x) for x in first]
        self.tokens = []          # tokens[t] = list of B ints, for every scheduled t
        self.enqueue_calls = []
        self.max_lead = 0

    def stream_begin(self, kv, logits, sampling):
        self.tokens = [list(self.first)]

    def stream_enqueue(self, kv, n):
        assert n >= 1
        self.enqueue_calls.append(n)
        for _ in range(n):
            self.tokens.append([int(_next(t)) for t in self.tokens[-1]])

    def stream_wait(self, kv, index, B, timeout_ms=0):
        assert index < len(self.tokens), "host waited for a token that was never scheduled"
        self.max_lead = max(self.max_lead, len(self.tokens) - 1 - index)
        return list(self.tokens[index])


def per_step_reference(first, max_new_tokens, eos_ids, pad, stop_at=None):
    B = len(first)
    finished = [False] * B
    cur, cols = list(first), []
    for step in range(max_new_tokens):
        shown = [pad if finished[b] else cur[b] for b in range(B)]
        cols.append(shown)
        for b in range(B):
            if shown[b] in eos_ids:
                finished[b] = True
        if (eos_ids and all(finished)) or (stop_at is not None and step + 1 >= stop_at):
            break
        cur = [int(_next(t)) for t in cur]  # rows never interact: what a finished row is fed does not matter
    return torch.tensor(cols, dtype=torch.long).t()


def run(first, max_new, eos, pad, run_ahead, streamer=None, criteria=None, prompt_len=3):
    eng = FakeEngine(first)
    prompt = torch.arange(len(first) * prompt_len).reshape(len(first), prompt_len)
    out = _stream_decode(eng, None, None, None, len(first), max_new, eos, pad, prompt, streamer, criteria, run_ahead=run_ahead)
    return out, eng


def test_stream_loop_equals_per_step_loop():
    for first in ([5], [0], [5, 11], [1, 2, 3, 4]):
        for eos in (set(), {2}, {15}, {9, 20}, {int(_next(first[0]))}, {first[0]}):
            for max_new in (1, 2, 7, 8, 9, 40):
                for ahead in (1, 4, 8):
                    pad = next(iter(eos)) if eos else 0
                    got, eng = run(first, max_new, eos, pad, ahead)
                    want = per_step_reference(first, max_new, eos, pad)
                    assert torch.equal(got, want), (first, eos, max_new, ahead, got, want)
                    assert got.shape[1] <= max_new and sum(eng.enqueue_calls) <= max(max_new - 1, 0)
                    assert eng.max_lead <= ahead + 1
                    if not eos:  # nothing stops early: every step the device ran was needed
                        assert sum(eng.enqueue_calls) == max_new - 1


def test_finished_rows_show_pad_and_stop_when_all_done():
    first = [5, 0]
    seq0, seq1 = [5], [0]
    for _ in range(30):
        seq0.append(int(_next(seq0[-1])))
        seq1.append(int(_next(seq1[-1])))
    eos = {seq0[2], seq1[6]} - set(seq1[:6]) - set(seq0[:2])
    assert len(eos) == 2
    out, _ = run(first, 30, eos, 99, 8)
    assert torch.equal(out, per_step_reference(first, 30, eos, 99))
    assert out.shape[1] == 7 and out[0, 3:].tolist() == [99] * 4


class RecordingStreamer:
    def __init__(self):
        self.got = []

    def put(self, value):
        assert value.dim() == 1 and value.dtype == torch.long
        self.got.append(value.tolist())


def test_streamer_and_stopping_criteria_protocol():
    seen = []

    def crit(ids, scores):  # reference signature: (output_ids so far incl. the prompt, scores) -> bool
        assert scores is None and ids.dim() == 2
        seen.append(ids.clone())
        return ids.shape[1] - 3 >= 5  # stop after 5 new tokens (prompt_len = 3)

    st = RecordingStreamer()
    out, eng = run([4], 50, set(), 0, 8, streamer=st, criteria=[crit])
    want = per_step_reference([4], 50, set(), 0, stop_at=5)
    assert torch.equal(out, want) and out.shape[1] == 5
    assert st.got == [[int(x)] for x in want[0]]
    # the criterion saw cat(prompt, new tokens) growing by one column per step
    assert [s.shape[1] for s in seen] == [4, 5, 6, 7, 8]
    assert torch.equal(seen[-1][:, :3], torch.arange(3).reshape(1, 3)) and torch.equal(seen[-1][:, 3:], want)
    # the device ran ahead, but never more than run_ahead tokens past what the host had consumed
    assert 4 <= sum(eng.enqueue_calls) <= 4 + 9


def test_tensor_valued_criteria_are_accepted():
    out, _ = run([4, 9], 20, set(), 0, 4, criteria=[lambda ids, s: torch.tensor([ids.shape[1] >= 6, True])])
    assert out.shape[1] == 3  # prompt_len 3 + 3 new tokens = 6 columns


def test_kv_pool_hands_out_exclusive_caches_and_blocks_at_its_cap():
    """_KVPool (one cache per generate() call): caches are created on demand up to `cap`, reused after release, and a
    third concurrent acquire waits for a release instead of sharing a cache (per-generate KV cache)."""
    import threading
    import time
    from llava.model.language_model.llava_llama import _KVPool

    class Eng:
        made = 0

        def new_kv(self, max_batch, max_seq):
            Eng.made += 1
            return object()

    pool = _KVPool(Eng(), max_batch=2, max_seq=64, cap=2)
    a, b = pool.acquire(), pool.acquire()
    assert a is not b and Eng.made == 2
    got = []
    t = threading.Thread(target=lambda: got.append(pool.acquire()))
    t.start()
    time.sleep(0.1)
    assert not got                     # cap reached: the third caller waits
    pool.release(a)
    t.join(timeout=5)
    assert got == [a] and Eng.made == 2  # ... and gets the released cache, not a new one
    pool.release(b)
    assert pool.acquire() is b

    class Failing:
        def new_kv(self, max_batch, max_seq):
            raise MemoryError("no room for another cache")

    bad = _KVPool(Failing(), 1, 8, cap=1)
    for _ in range(2):                 # a failed creation gives its place back (the second attempt must not dead-lock)
        try:
            bad.acquire()
        except MemoryError:
            pass
        else:
            raise AssertionError("expected MemoryError")
