"""GPU: token selection kernel (csrc/sampling.cu) against oracle/sampling_oracle.py, the streaming decode C-ABI
(b2_stream_begin / enqueue / wait) against the device-resident greedy loop, and generate() on the call pattern the
reference's serving code uses (llava/serve/model_worker.py:166-188): a worker Thread, a TextIteratorStreamer drained by
the caller, a keyword stopping criterion, several requests in flight on one model object."""
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import make_engine, make_model, synth_inputs  # noqa: E402
from llava import _b2  # noqa: E402
from oracle import llava_oracle as O  # noqa: E402
from oracle import sampling_oracle as S  # noqa: E402

DEV = "cuda"


@pytest.fixture(scope="module")
def tiny_engine():
    cfg = O.CONFIGS["tiny"]
    w = O.make_weights(cfg, seed=0)
    eng = make_engine(cfg, w, max_batch=16, max_seq=160, max_images=4)
    yield cfg, w, eng
    eng.close()


# ------------------------------------------------------------------------------------------------ selection kernel
def _check_draw(row, tok, T, k, p, seed, index, b):
    """the GPU token must be the one whose fixed-point CDF interval holds the Philox target (GPU expf and numpy exp may
    differ in the last ulp: a slack of 1e-6 of the total mass on the interval ends, and on the survivor boundary)."""
    want, info = S.sample_row(row, T, k, p, seed, index, b)
    if tok == want:
        return
    slack = 1e-6 * info["total"]
    assert info["lo"][tok] - slack <= info["target"] <= info["hi"][tok] + slack, (tok, want, T, k, p, index, b)


@pytest.mark.parametrize("V", [32000, 1000, 40])
def test_sample_kernel_matches_oracle(tiny_engine, V):
    _, _, eng = tiny_engine
    g = torch.Generator().manual_seed(V)
    B = 6
    logits = torch.randn(B, V, generator=g) * 2.5
    logits[1, :5] = float("-inf")
    logits[2, 7] = 40.0                      # one dominating token
    logits[3] = logits[3].round()            # many exact ties
    dev = logits.to(DEV)
    x = logits.numpy()
    # greedy == argmax with first-occurrence ties
    got = eng.sample(dev, _b2.make_sampling(False)).cpu().tolist()
    assert got == [S.greedy(x[b]) for b in range(B)]
    for (T, k, p) in [(1.0, 0, 1.0), (0.7, 50, 0.9), (0.2, 0, 0.7), (1.5, 5, 1.0), (1.0, 1, 0.5), (0.9, 0, 0.05), (2.0, V + 3, 0.999)]:
        for index in (0, 1, 77):
            sp = _b2.make_sampling(True, T, p, k, seed=0xC0FFEE + index)
            got = eng.sample(dev, sp, index=index).cpu().tolist()
            again = eng.sample(dev, sp, index=index).cpu().tolist()
            assert got == again                                   # a draw is a pure function of (logits, seed, index, row)
            for b in range(B):
                if b == 3 and p < 1.0:
                    continue  # exact ties at the nucleus boundary: HF keeps some of the tied tokens, this kernel all of them
                assert 0 <= got[b] < V
                _check_draw(x[b], got[b], T, k, p, 0xC0FFEE + index, index, b)  # tokens outside the survivor set have an empty interval


def test_sample_kernel_frequencies(tiny_engine):
    """4096 draws (64 rows with the same logits x 64 indices) follow the renormalised survivor distribution."""
    _, _, eng = tiny_engine
    g = torch.Generator().manual_seed(5)
    V, T, k, p = 1000, 0.8, 40, 0.95
    row = torch.randn(V, generator=g) * 3
    dev = row[None].repeat(64, 1).to(DEV)
    keep, e = S.kept_mask(row.numpy(), T, k, p)
    probs = e.astype(np.float64) / e.sum()
    counts = np.zeros(V)
    for index in range(64):
        toks = eng.sample(dev, _b2.make_sampling(True, T, p, k, seed=99), index=index).cpu().numpy()
        np.add.at(counts, toks, 1)
    assert counts[~keep].sum() == 0
    n = counts.sum()
    sel = probs * n > 5
    chi2 = (((counts - probs * n) ** 2)[sel] / (probs * n)[sel]).sum()
    assert chi2 < 3 * sel.sum(), (chi2, int(sel.sum()))
    assert len(set(counts.nonzero()[0])) > 5


# ------------------------------------------------------------------------------------------------ streaming C-ABI
def _prefill(eng, cfg, B, S_, seed):
    g = torch.Generator().manual_seed(seed)
    embeds = (torch.randn(B, S_, cfg["hidden"], generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    kv = eng.new_kv(B, 160)
    return kv, eng.prefill(kv, embeds, None, _b2.LOGITS_LAST)


@pytest.mark.parametrize("B", [1, 2, 4, 12])   # megakernel (<= 2), GEMV graph, stream-K GEMM graph
def test_streaming_greedy_equals_device_resident_greedy(tiny_engine, B):
    cfg, _, eng = tiny_engine
    n = 24
    kv, logits = _prefill(eng, cfg, B, 20, seed=B)
    first = eng.argmax(logits)
    ref = torch.cat([first[None].cpu(), eng.decode_greedy(kv, first, n - 1).cpu()])   # [n, B]
    kv.close()
    kv, logits2 = _prefill(eng, cfg, B, 20, seed=B)
    assert torch.equal(logits, logits2)
    eng.stream_begin(kv, logits2)
    eng.stream_enqueue(kv, 5)
    got = [eng.stream_wait(kv, t, B) for t in range(3)]          # host reads while the device is ahead
    eng.stream_enqueue(kv, n - 6)
    got += [eng.stream_wait(kv, t, B) for t in range(3, n)]
    assert got == ref.tolist()
    with pytest.raises(ValueError):
        eng.stream_wait(kv, n, B)                                # never scheduled
    # a second generation on the same cache object: new epoch, old ring entries are not mistaken for new tokens
    kv.reset()
    g = torch.Generator().manual_seed(100 + B)
    embeds = (torch.randn(B, 9, cfg["hidden"], generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    lg = eng.prefill(kv, embeds, None, _b2.LOGITS_LAST)
    eng.stream_begin(kv, lg)
    eng.stream_enqueue(kv, 3)
    second = [eng.stream_wait(kv, t, B) for t in range(4)]
    assert second[0] == eng.argmax(lg).cpu().tolist()
    torch.cuda.synchronize()
    kv.close()


@pytest.mark.parametrize("B", [1, 4])
def test_streaming_sampling_is_seeded_and_consistent_with_the_logits(tiny_engine, B):
    cfg, _, eng = tiny_engine
    runs = []
    for seed in (7, 7, 8):
        kv, logits = _prefill(eng, cfg, B, 16, seed=3)
        sp = _b2.make_sampling(True, 1.3, 0.95, 50, seed=seed)
        eng.stream_begin(kv, logits, sp)
        eng.stream_enqueue(kv, 15)
        toks = [eng.stream_wait(kv, t, B) for t in range(16)]
        # token 0 was drawn from the prefill logits with index 0: check it against the oracle
        for b in range(B):
            _check_draw(logits[b].cpu().numpy(), toks[0][b], 1.3, 50, 0.95, seed, 0, b)
        runs.append(toks)
        torch.cuda.synchronize()
        kv.close()
    assert runs[0] == runs[1] and runs[0] != runs[2]


# ------------------------------------------------------------------------------------------------ generate()
class CharTokenizer:
    """Minimal tokenizer surface used by TextIteratorStreamer / keyword stopping: id i <-> chr(97 + i % 26)."""
    bos_token_id = 1

    def decode(self, ids, **kw):
        return "".join(chr(97 + int(i) % 26) for i in (ids.tolist() if torch.is_tensor(ids) else ids))

    def batch_decode(self, ids, **kw):
        return [self.decode(r) for r in ids]


class KeywordStop:
    """Restatement of the reference's KeywordsStoppingCriteria protocol (llava/mm_utils.py:79-114: plain bool, looks at
    the decoded tail of cat(prompt ids, new tokens) beyond start_len); the reference class itself is exercised on the CPU
    in tests/test_dropin_cpu.py (the reference tree does not exist on the GPU box)."""

    def __init__(self, keyword, tokenizer, input_ids):
        self.keyword, self.tokenizer, self.start_len = keyword, tokenizer, input_ids.shape[1]

    def __call__(self, output_ids, scores, **kw):
        assert output_ids.shape[0] == 1
        tail = self.tokenizer.batch_decode(output_ids[:, self.start_len:][:, -len(self.keyword):])[0]
        return self.keyword in tail


def _tiny_model(max_batch=2):
    cfg = O.CONFIGS["tiny"]
    w = O.make_weights(cfg, seed=0)
    return cfg, make_model(cfg, w, max_batch=max_batch, max_seq=160)


def test_generate_from_a_worker_thread_with_iterator_streamer_and_keyword_stop():
    from transformers import TextIteratorStreamer

    cfg, model = _tiny_model()
    ids, images = synth_inputs(cfg, B=1, Lt=12, seed=9)
    ids_d, img_d = ids.to(DEV), images.to(DEV)
    ref = model.generate(ids_d, images=img_d, do_sample=False, max_new_tokens=40, eos_token_id=[])
    new = ref[0, ids.shape[1]:].cpu()
    tok = CharTokenizer()
    text = tok.decode(new)
    keyword = text[9:12]                                 # appears after 12 new tokens (or earlier, if it repeats)
    first_hit = text.find(keyword) + len(keyword)
    streamer = TextIteratorStreamer(tok, skip_prompt=True, skip_special_tokens=True, timeout=30)
    crit = KeywordStop(keyword, tok, ids_d)
    result = {}

    def work():  # exactly the call llava/serve/model_worker.py:174-185 makes
        result["out"] = model.generate(inputs=ids_d, images=img_d, do_sample=False, temperature=0.0, top_p=1.0,
                                       max_new_tokens=40, streamer=streamer, stopping_criteria=[crit], use_cache=True,
                                       eos_token_id=[])

    th = threading.Thread(target=work)
    th.start()
    pieces = "".join(streamer)                           # the caller drains the queue while generate() runs
    th.join()
    out = result["out"]
    assert out.shape[1] == ids.shape[1] + first_hit
    assert torch.equal(out.cpu(), ref[:, : out.shape[1]].cpu())
    assert pieces == text[:first_hit]
    model.invalidate_engine()


def test_concurrent_generate_threads_do_not_share_a_cache():
    """two requests in flight on one model must not see each other's context."""
    cfg, model = _tiny_model()
    prompts = [synth_inputs(cfg, B=1, Lt=10 + 3 * i, seed=20 + i) for i in range(4)]
    serial = [model.generate(p.to(DEV), images=im.to(DEV), do_sample=False, max_new_tokens=24, eos_token_id=[]).cpu()
              for p, im in prompts]
    for rounds in range(3):
        results, errors = [None] * 4, []

        def work(i):
            try:
                p, im = prompts[i]
                results[i] = model.generate(p.to(DEV), images=im.to(DEV), do_sample=False, max_new_tokens=24,
                                            eos_token_id=[]).cpu()
            except Exception as e:  # pragma: no cover
                errors.append(e)

        threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
        [t.start() for t in threads]
        [t.join() for t in threads]
        assert not errors, errors
        for i in range(4):
            assert torch.equal(results[i], serial[i]), (rounds, i)
    model.invalidate_engine()


def test_forward_caches_are_leases():
    cfg, model = _tiny_model()
    ids, images = synth_inputs(cfg, B=1, Lt=10, seed=1)
    outs = [model(input_ids=ids.to(DEV), images=images.to(DEV), use_cache=True) for _ in range(3)]
    nxt = outs[2].logits[:, -1].argmax(-1, keepdim=True)
    a = model(input_ids=nxt, past_key_values=outs[2].past_key_values, use_cache=True).logits
    b = model(input_ids=nxt, past_key_values=outs[1].past_key_values, use_cache=True).logits
    torch.testing.assert_close(a, b)                     # two live caches, same context
    with pytest.raises(RuntimeError, match="recycled"):  # the first one was recycled by the third forward
        model(input_ids=nxt, past_key_values=outs[0].past_key_values, use_cache=True)
    model.invalidate_engine()


def test_bad_ids_raise_instead_of_reading_out_of_bounds():
    """ids >= vocab / a leftover image placeholder must not fault the context."""
    cfg, model = _tiny_model()
    ids, images = synth_inputs(cfg, B=1, Lt=10, seed=1)
    bad = ids.clone()
    bad[0, 3] = cfg["vocab"] + 5
    with pytest.raises(ValueError, match="embedding table"):
        model.generate(bad, images=images.to(DEV), max_new_tokens=2)               # host ids: checked while the index is built
    with pytest.raises(ValueError, match="token id"):
        model.generate(bad.to(DEV), images=images.to(DEV), max_new_tokens=2)       # device ids: flagged by the gather kernel
    with pytest.raises(ValueError, match="token id"):
        model.generate(bad.to(DEV), images=None, max_new_tokens=2)                 # text path: flagged by the kernel
    with pytest.raises(ValueError, match="token id"):
        model.generate(ids.to(DEV), images=None, max_new_tokens=2)                 # -200 without images
    ok = model.generate(ids.to(DEV), images=images.to(DEV), max_new_tokens=2)      # the context survived all of it
    assert ok.shape[1] == ids.shape[1] + 2
    with pytest.raises(NotImplementedError):
        model.generate(ids.to(DEV), images=images.to(DEV), max_new_tokens=2, repetition_penalty=1.2)
    model.generate(ids.to(DEV), images=images.to(DEV), max_new_tokens=2, repetition_penalty=1.0)  # the "off" value is fine
    model.invalidate_engine()


# ------------------------------------------------------------------------------------------------ device-side splice
def test_device_splice_equals_host_index(tiny_engine):
    """b2_splice_ids (index built by a kernel from ids that stay on the device) == b2_splice over the host-built index, which
    tests/test_splice_host.py pins against the unmodified reference; a wrong placeholder count is flagged, never misread."""
    from llava.model.llava_arch import build_source_index

    cfg, _, eng = tiny_engine
    P, h = 16, cfg["hidden"]
    g = torch.Generator().manual_seed(11)
    for B, Lt, k, group in [(1, 12, 1, 1), (1, 40, 2, 1), (3, 9, 1, 1), (2, 17, 3, 1), (2, 11, 1, 2), (4, 130, 2, 1), (1, 700, 1, 1)]:
        n_img = B * k
        feat_rows = [P * group] * n_img
        feats = torch.randn(sum(feat_rows), h, generator=g).to(torch.bfloat16).to(DEV)
        ids = torch.randint(3, cfg["vocab"], (B, Lt), generator=g)
        for b in range(B):
            pos = torch.randperm(Lt - 1, generator=g)[:k] + 1
            ids[b, pos] = O.IMAGE_TOKEN_INDEX
        assert (Lt - k + k * P * group) * B <= 16 * 160  # this fixture's workspace
        got = eng.splice_ids(ids.to(DEV), k, feat_rows, feats)
        src, _, _, _, lens = build_source_index(ids.numpy(), np.ones((B, Lt), bool), np.full((B, Lt), -100), feats.shape[0],
                                                feat_rows, None, "right")
        want = eng.splice(torch.from_numpy(src.reshape(-1)).to(DEV), feats, B, src.shape[1])
        torch.cuda.synchronize()
        assert eng.take_async_error() == 0
        assert got.shape == want.shape and torch.equal(got, want), (B, Lt, k, group)
    # one placeholder too few / too many: flagged
    ids = torch.randint(3, cfg["vocab"], (1, 20), generator=g)
    ids[0, 4] = O.IMAGE_TOKEN_INDEX
    feats = torch.randn(2 * P, h, generator=g).to(torch.bfloat16).to(DEV)
    eng.splice_ids(ids.to(DEV), 2, [P, P], feats)
    torch.cuda.synchronize()
    assert eng.take_async_error() & _b2.ERR_SPLICE_SLOTS
    ids[0, 9] = ids[0, 12] = O.IMAGE_TOKEN_INDEX
    eng.splice_ids(ids.to(DEV), 2, [P, P], feats)
    torch.cuda.synchronize()
    assert eng.take_async_error() & _b2.ERR_SPLICE_SLOTS


def test_generate_with_device_ids_falls_back_to_reference_semantics_on_odd_inputs():
    """two images but ONE placeholder: the reference uses the first image and ignores the second (llava_arch.py:149-181).
    The device path assumes n_images / B placeholders per row, notices on its first sync that the assumption failed, and
    redoes the splice on the host path: same result as host-resident ids."""
    cfg, model = _tiny_model()
    ids, images = synth_inputs(cfg, B=1, Lt=12, seed=3)
    two = torch.cat([images, images.flip(-1)])
    a = model.generate(ids, images=two.to(DEV), do_sample=False, max_new_tokens=8, eos_token_id=[])          # host path
    b = model.generate(ids.to(DEV), images=two.to(DEV), do_sample=False, max_new_tokens=8, eos_token_id=[])  # device -> fallback
    c = model.generate(ids.to(DEV), images=images.to(DEV), do_sample=False, max_new_tokens=8, eos_token_id=[])  # device path
    assert torch.equal(a.cpu(), b.cpu()) and torch.equal(a.cpu(), c.cpu())
    model.invalidate_engine()


# ------------------------------------------------------------------------------------------------ continuous batching
def test_continuous_batching_of_concurrent_generate_threads():
    """six worker threads call generate() on ONE model (the reference worker's pattern, model_worker.py:174-185,
    230-243); with config.b2_continuous_batching = 4 their decode steps are shared. On the well-conditioned weight set every
    request must produce exactly the ids it produces alone on an unbatched model — requests join and leave the batch at
    different steps (different prompt lengths, max_new_tokens, one keyword stop, one sampled request)."""
    cfg = O.CONFIGS["tiny"]
    wc = O.condition_weights(O.make_weights(cfg, seed=0), cfg, seed=0)
    solo = make_model(cfg, wc, max_batch=1, max_seq=160)
    batched = make_model(cfg, wc, max_batch=4, max_seq=160, b2_continuous_batching=4)
    jobs = []
    for i in range(6):
        ids, images = synth_inputs(cfg, B=1, Lt=9 + 4 * i, seed=40 + i)
        jobs.append(dict(ids=ids.to(DEV), images=images.to(DEV), n=10 + 7 * i))
    tok = CharTokenizer()

    def run(model, j, i):
        kw = dict(do_sample=False, max_new_tokens=j["n"], eos_token_id=[])
        if i == 2:  # a keyword stop that fires mid-way (computed from the solo run below)
            kw["stopping_criteria"] = [KeywordStop(j["keyword"], tok, j["ids"])]
        if i == 4:
            torch.manual_seed(123)
            kw.update(do_sample=True, temperature=0.8, top_p=0.9)
        return model.generate(j["ids"], images=j["images"], **kw).cpu()

    free = solo.generate(jobs[2]["ids"], images=jobs[2]["images"], do_sample=False, max_new_tokens=jobs[2]["n"], eos_token_id=[])
    jobs[2]["keyword"] = tok.decode(free[0, jobs[2]["ids"].shape[1]:].cpu())[5:8]
    want = [run(solo, j, i) for i, j in enumerate(jobs)]
    got, errors = [None] * 6, []

    def work(i):
        try:
            got[i] = run(batched, jobs[i], i)
        except Exception as e:  # pragma: no cover
            errors.append((i, e))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    for i in range(6):
        if i == 4:  # sampled: same prompt echo and length; the draw stream differs from the solo run by construction (row index)
            assert got[i].shape == want[i].shape
            continue
        assert torch.equal(got[i], want[i]), (i, got[i], want[i])
    st = batched._batcher.stats
    assert st["admitted"] == 6 and st["max_active"] >= 2 and st["rows_stepped"] > st["steps"], st   # steps really were shared
    # a second wave reuses the freed slots (ring tags / slot resets)
    again = run(batched, jobs[1], 1)
    assert torch.equal(again, want[1])
    batched.invalidate_engine()
    solo.invalidate_engine()
