"""CPU: host-side logic of generate()'s device-resident greedy loop (chunked decode_greedy + eos bookkeeping) against a
plain per-step restatement of the loop the reference's callers get from HF generate: finished rows show the
pad id, generation stops at the step where every row has finished, the result never exceeds max_new_tokens."""
import torch

from llava.model.language_model.llava_llama import _greedy_chunked


def _next(tok):  # deterministic fake "model": next token is a function of the fed token only (rows independent)
    return (tok * 7 + 3) % 23


class FakeEngine:
    def __init__(self):
        self.calls = []

    def decode_greedy(self, kv, first_tokens, n_steps):
        self.calls.append(int(n_steps))
        cur, rows = first_tokens.clone().to(torch.int64), []
        for _ in range(n_steps):
            cur = _next(cur)
            rows.append(cur.clone())
        return torch.stack(rows).to(torch.int32)  # [n_steps, B]


def per_step_reference(first, max_new_tokens, eos_ids, pad):
    B = first.numel()
    finished = torch.zeros(B, dtype=torch.bool)
    cur, cols = first.clone().to(torch.int64), []
    for step in range(max_new_tokens):
        shown = torch.where(finished, torch.full_like(cur, pad), cur)
        cols.append(shown)
        for b in range(B):
            if int(shown[b]) in eos_ids:
                finished[b] = True
        if (eos_ids and bool(finished.all())) or step == max_new_tokens - 1:
            break
        cur = _next(cur)  # rows never interact: what a finished row is fed does not matter for the others
    return torch.stack(cols, dim=1)


def test_chunked_greedy_equals_per_step_loop():
    for first in ([5], [0], [5, 11], [1, 2, 3, 4]):
        f = torch.tensor(first, dtype=torch.int32)
        for eos in (set(), {2}, {15}, {9, 20}, {int(_next(torch.tensor(first[0])))}, {first[0]}):
            for max_new in (1, 2, 15, 16, 17, 40):
                for chunk in (1, 4, 16):
                    eng = FakeEngine()
                    pad = next(iter(eos)) if eos else 0
                    got = _greedy_chunked(eng, None, f, max_new, eos, pad, chunk=chunk)
                    want = per_step_reference(f, max_new, eos, pad)
                    assert torch.equal(got, want), (first, eos, max_new, chunk, got, want)
                    assert got.shape[1] <= max_new and sum(eng.calls) <= max(max_new - 1, 0)
                    if not eos:
                        assert eng.calls in ([], [max_new - 1])  # no eos: one device-resident run
                    else:
                        assert all(c <= chunk for c in eng.calls)


def test_chunked_greedy_pads_finished_rows_and_stops_when_all_done():
    # row 0 hits eos early, row 1 later: row 0 shows pad afterwards, the result ends at row 1's eos step
    f = torch.tensor([5, 11], dtype=torch.int32)
    seq0, seq1 = [5], [11]
    for _ in range(30):
        seq0.append(int(_next(torch.tensor(seq0[-1]))))
        seq1.append(int(_next(torch.tensor(seq1[-1]))))
    eos = {seq0[2], seq1[6]} - set(seq1[:6])
    if seq0[2] in eos and seq1[6] in eos:
        out = _greedy_chunked(FakeEngine(), None, f, 30, eos, 99, chunk=4)
        want = per_step_reference(f, 30, eos, 99)
        assert torch.equal(out, want)
        assert int(out[0, -1]) == 99 or out.shape[1] <= 3
