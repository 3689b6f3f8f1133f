"""ctypes binding of libb2llava.so (C ABI: include/b2llava.h).

PyTorch is used here only for device memory and streams: every call passes raw `tensor.data_ptr()` values
and `torch.cuda.current_stream().cuda_stream` across the ABI. There is no CPU or eager fallback: if the
library is missing or the device is not sm_100, calls raise.
"""
import ctypes
import os
import threading
import weakref

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get(
    "B2LLAVA_LIB", os.path.normpath(os.path.join(_HERE, "..", "..", "lib", "libb2llava.so"))
)

DT_BF16, DT_F16, DT_F32 = 0, 1, 2
ACT_NONE, ACT_QUICK_GELU, ACT_GELU_ERF, ACT_SWIGLU = 0, 1, 2, 3
LOGITS_NONE, LOGITS_LAST, LOGITS_ALL = 0, 1, 2
INT32_MIN = -(2**31)
ERR_TOKEN_RANGE, ERR_IMAGE_ROW_RANGE, ERR_SPLICE_SLOTS = 1, 2, 4

_c = ctypes
_vp, _i32, _i64, _f32 = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float


class ModelDesc(_c.Structure):
    _fields_ = [
        ("image_size", _c.c_int32), ("patch_size", _c.c_int32), ("vit_hidden", _c.c_int32),
        ("vit_inter", _c.c_int32), ("vit_layers", _c.c_int32), ("vit_heads", _c.c_int32),
        ("vit_select_layer", _c.c_int32), ("vit_ln_eps", _c.c_float),
        ("hidden", _c.c_int32), ("inter", _c.c_int32), ("layers", _c.c_int32), ("heads", _c.c_int32),
        ("vocab", _c.c_int32), ("rms_eps", _c.c_float), ("rope_theta", _c.c_float),
        ("max_batch", _c.c_int32), ("max_seq", _c.c_int32), ("max_images", _c.c_int32),
    ]


class Sampling(_c.Structure):
    """b2_sampling (include/b2llava.h): do_sample == 0 -> greedy argmax."""
    _fields_ = [("do_sample", _c.c_int32), ("temperature", _c.c_float), ("top_p", _c.c_float),
                ("top_k", _c.c_int32), ("seed", _c.c_ulonglong)]


class PreprocessPlan(_c.Structure):
    """b2_preprocess_plan (include/b2llava.h)."""
    _fields_ = [("img", _vp), ("H", _c.c_int32), ("W", _c.c_int32), ("pad_top", _c.c_int32), ("pad_left", _c.c_int32),
                ("bg", _c.c_uint8 * 4), ("h_bounds", _vp), ("h_kk", _vp), ("h_ksize", _c.c_int32), ("h_identity", _c.c_int32),
                ("v_bounds", _vp), ("v_kk", _vp), ("v_ksize", _c.c_int32), ("v_identity", _c.c_int32),
                ("y0", _c.c_int32), ("rows", _c.c_int32), ("x_lo", _c.c_int32), ("y_lo", _c.c_int32), ("out", _c.c_int32),
                ("tmp", _vp), ("mean", _c.c_float * 3), ("stdv", _c.c_float * 3), ("rescale", _c.c_float),
                ("pixels", _vp), ("u8_out", _vp)]


def make_sampling(do_sample=False, temperature=1.0, top_p=1.0, top_k=0, seed=0):
    return Sampling(int(bool(do_sample)), float(temperature), float(1.0 if top_p is None else top_p),
                    int(top_k or 0), int(seed) & (2**64 - 1))


# name -> (restype, argtypes); must list every symbol include/b2llava.h declares (tests check this)
SIGNATURES = {
    "b2_init": (_i32, [_i32]),
    "b2_last_error": (_c.c_char_p, []),
    "b2_version": (_i32, []),
    "b2_launch_count": (_c.c_ulonglong, []),
    "b2_model_create": (_i32, [_c.POINTER(ModelDesc), _c.POINTER(_vp)]),
    "b2_model_set_weight": (_i32, [_vp, _c.c_char_p, _vp, _c.POINTER(_i64), _i32, _i32]),
    "b2_model_finalize": (_i32, [_vp]),
    "b2_model_destroy": (_i32, [_vp]),
    "b2_model_enable_fp8_decode": (_i32, [_vp]),
    "b2_kv_create": (_i32, [_vp, _i32, _i32, _c.POINTER(_vp)]),
    "b2_kv_reset": (_i32, [_vp]),
    "b2_kv_destroy": (_i32, [_vp]),
    "b2_kv_lengths": (_i32, [_vp, _c.POINTER(_c.c_int32), _i32]),
    "b2_vit_encode": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "b2_project": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "b2_encode_images": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "b2_splice": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    "b2_splice_ids": (_i32, [_vp, _vp, _i32, _i32, _i32, _c.POINTER(_c.c_int32), _i32, _vp, _i32, _vp, _vp]),
    "b2_async_error": (_i32, [_vp, _c.POINTER(_c.c_int)]),
    "b2_stream_begin": (_i32, [_vp, _vp, _vp, _i32, _c.POINTER(Sampling), _vp]),
    "b2_stream_enqueue": (_i32, [_vp, _vp, _i32, _vp]),
    "b2_stream_wait": (_i32, [_vp, _i32, _c.POINTER(_c.c_int32), _i32]),
    "b2_op_preprocess_clip": (_i32, [_c.POINTER(PreprocessPlan), _vp]),
    "b2_op_sample": (_i32, [_vp, _i32, _i32, _c.POINTER(Sampling), _i32, _vp, _vp]),
    "b2_prefill": (_i32, [_vp, _vp, _vp, _c.POINTER(_c.c_int32), _i32, _i32, _vp, _i32, _vp]),
    "b2_prefill_slots": (_i32, [_vp, _vp, _vp, _c.POINTER(_c.c_int32), _i32, _i32, _i32, _vp, _i32, _vp]),
    "b2_batch_begin": (_i32, [_vp, _vp, _i32, _vp]),
    "b2_batch_set_row": (_i32, [_vp, _vp, _i32, _i32, _c.POINTER(Sampling), _i32, _vp]),
    "b2_decode_step": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "b2_decode_greedy": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    "b2_argmax": (_i32, [_vp, _i32, _i32, _vp, _vp]),
    "b2_op_gemm": (_i32, [_vp, _i32, _vp, _i32, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b2_op_gemv": (_i32, [_vp, _i64, _vp, _i32, _vp, _f32, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b2_op_gemm_skinny": (_i32, [_vp, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp, _vp]),
    "b2_op_gemm_skinny_fp8": (_i32, [_vp, _i32, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp, _vp]),
    "b2_op_quantize_rows_e4m3": (_i32, [_vp, _i64, _i32, _i32, _vp, _i64, _vp, _vp]),
    "b2_op_rmsnorm_quant_e4m3": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp]),
    "b2_op_gemm_skinny_workspace_bytes": (_i64, [_i32, _i32, _i32]),
    "b2_op_gemm_skinny_counter_bytes": (_i64, [_i32]),
    "b2_op_layernorm": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp]),
    "b2_op_rmsnorm": (_i32, [_vp, _vp, _vp, _i32, _i32, _f32, _vp]),
    "b2_op_flash_attn": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp]),
    "b2_op_rope_kv_write": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp]),
    "b2_op_decode_attn": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _f32, _vp]),
    "b2_op_decode_attn_scratch_bytes": (_i64, [_i32, _i32, _i32]),
    "b2_op_interleave_gate_up": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp]),
    "b2_op_im2col": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
}

_lib = None
_lib_lock = threading.Lock()
_inited_devices = set()


def load_library():
    """Load libb2llava.so (no GPU needed to load and resolve symbols)."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"libb2llava.so not found at {LIB_PATH}: build it with "
                    f"`python llava-plus-codebase_b200/build.py` (there is no fallback path)"
                )
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def last_error():
    return load_library().b2_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    """Error convention of the reference's callers: ValueError for bad arguments,
    RuntimeError for CUDA / state failures; never abort."""
    if rc == 0:
        return
    msg = f"{what}: {last_error()}" if what else last_error()
    if rc == -1:
        raise ValueError(msg)
    raise RuntimeError(msg)


def init(device_index):
    lib = load_library()
    if device_index not in _inited_devices:
        if not torch.cuda.is_available():
            raise RuntimeError("b2llava needs a CUDA device (sm_100a); no CPU fallback exists")
        check(lib.b2_init(int(device_index)), "b2_init")
        _inited_devices.add(device_index)
    return lib


def stream_ptr():
    return _vp(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return _vp(0)
    return _vp(t.data_ptr())


def launch_count():
    return int(load_library().b2_launch_count())


_TORCH_DT = {torch.bfloat16: DT_BF16, torch.float16: DT_F16, torch.float32: DT_F32}


class KVCache:
    """Device KV cache handle (b2_kv): [layers][B][heads][max_seq][128] bf16 for K and V."""

    def __init__(self, engine, max_batch, max_seq):
        self.engine = engine
        self.max_batch, self.max_seq = int(max_batch), int(max_seq)
        h = _vp()
        check(engine.lib.b2_kv_create(engine.handle, self.max_batch, self.max_seq, ctypes.byref(h)), "b2_kv_create")
        self.handle = h

    def reset(self):
        check(self.engine.lib.b2_kv_reset(self.handle), "b2_kv_reset")

    def lengths(self, n=None):
        n = self.max_batch if n is None else n
        arr = (_c.c_int32 * n)()
        check(self.engine.lib.b2_kv_lengths(self.handle, arr, n), "b2_kv_lengths")
        return list(arr)

    def get_seq_length(self, layer_idx=0):
        return self.lengths(1)[0]

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.engine.lib.b2_kv_destroy(self.handle)
            self.handle = _vp(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """Thin owner of a b2_model handle. All tensors are torch CUDA tensors used as raw device memory."""

    def __init__(self, desc: dict, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("b2llava Engine requires a CUDA device; there is no CPU path")
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.lib = init(self.index)
        self.desc = ModelDesc(**desc)
        h = _vp()
        with torch.cuda.device(self.index):
            check(self.lib.b2_model_create(ctypes.byref(self.desc), ctypes.byref(h)), "b2_model_create")
        self.handle = h
        self.hidden, self.vocab = desc["hidden"], desc["vocab"]
        self.vit_hidden = desc["vit_hidden"]
        self.num_patches = (desc["image_size"] // desc["patch_size"]) ** 2
        self.finalized = False
        self._kvs = weakref.WeakSet()

    # -- weights ------------------------------------------------------------------------------------
    def set_weight(self, key, tensor):
        t = tensor.detach()
        if t.dtype not in _TORCH_DT:
            t = t.float()
        t = t.contiguous()
        shape = (_i64 * max(t.dim(), 1))(*([int(s) for s in t.shape] or [1]))
        with torch.cuda.device(self.index):
            check(self.lib.b2_model_set_weight(self.handle, key.encode(), ptr(t), shape, max(t.dim(), 1),
                                               _TORCH_DT[t.dtype]), f"set_weight({key})")

    def finalize(self):
        with torch.cuda.device(self.index):
            check(self.lib.b2_model_finalize(self.handle), "b2_model_finalize")
        self.finalized = True

    def enable_fp8_decode(self):
        """BASELINE configs[4]: e4m3 weights for decode at batch >= 7 (see include/b2llava.h): W8A8 numerics, opt-in."""
        with torch.cuda.device(self.index):
            check(self.lib.b2_model_enable_fp8_decode(self.handle), "b2_model_enable_fp8_decode")

    def new_kv(self, max_batch, max_seq):
        with torch.cuda.device(self.index):
            kv = KVCache(self, max_batch, max_seq)
        self._kvs.add(kv)
        return kv

    # -- hot path -----------------------------------------------------------------------------------
    def _bf16(self, t):
        return t.to(device=self.device, dtype=torch.bfloat16).contiguous()

    def vit_encode(self, pixels):
        pixels = self._bf16(pixels)
        B = pixels.shape[0]
        out = torch.empty(B, self.num_patches, self.vit_hidden, dtype=torch.bfloat16, device=self.device)
        with torch.cuda.device(self.index):
            check(self.lib.b2_vit_encode(self.handle, ptr(pixels), B, ptr(out), stream_ptr()), "b2_vit_encode")
        return out

    def project(self, feats):
        feats = self._bf16(feats)
        rows = feats.numel() // self.vit_hidden
        out = torch.empty(*feats.shape[:-1], self.hidden, dtype=torch.bfloat16, device=self.device)
        with torch.cuda.device(self.index):
            check(self.lib.b2_project(self.handle, ptr(feats), rows, ptr(out), stream_ptr()), "b2_project")
        return out

    def encode_images(self, pixels):
        pixels = self._bf16(pixels)
        B = pixels.shape[0]
        out = torch.empty(B, self.num_patches, self.hidden, dtype=torch.bfloat16, device=self.device)
        with torch.cuda.device(self.index):
            check(self.lib.b2_encode_images(self.handle, ptr(pixels), B, ptr(out), stream_ptr()), "b2_encode_images")
        return out

    def splice(self, src_index, image_feats, B, S):
        """src_index: int32 device tensor [B*S]; image_feats: bf16 [n_rows, hidden] or None. Ids / rows out of range
        become zero rows and are reported by check_async_error() (never an out-of-bounds read)."""
        out = torch.empty(B, S, self.hidden, dtype=torch.bfloat16, device=self.device)
        n_rows = 0 if image_feats is None else int(image_feats.shape[0])
        with torch.cuda.device(self.index):
            check(self.lib.b2_splice(self.handle, ptr(src_index), ptr(image_feats), n_rows, B * S, ptr(out), stream_ptr()),
                  "b2_splice")
        return out

    def splice_ids(self, input_ids, k_per_row, feat_rows, image_feats):
        """Device-side splice for equal-length unpadded rows: input_ids int64 [B, Lt] ON THE DEVICE, k_per_row image
        placeholders per row, feat_rows[j] = feature rows of image slot j (row-major slot order). Returns embeds
        [B, S, hidden] with S = Lt - k + sum(feat rows of one row) and no host synchronisation. A wrong placeholder count is
        reported by take_async_error() (ERR_SPLICE_SLOTS) once the stream has run."""
        B, Lt = input_ids.shape
        n_img = len(feat_rows)
        assert n_img == B * k_per_row and input_ids.dtype == torch.int64 and input_ids.is_cuda and input_ids.is_contiguous()
        per_row = [sum(feat_rows[b * k_per_row:(b + 1) * k_per_row]) for b in range(B)]
        assert len(set(per_row)) == 1, "rows must receive the same number of feature rows"
        S = Lt - k_per_row + per_row[0]
        off = (_c.c_int32 * (n_img + 1))(*([0] + [sum(feat_rows[:j + 1]) for j in range(n_img)]))
        out = torch.empty(B, S, self.hidden, dtype=torch.bfloat16, device=self.device)
        with torch.cuda.device(self.index):
            check(self.lib.b2_splice_ids(self.handle, ptr(input_ids), B, Lt, int(k_per_row), off, n_img, ptr(image_feats), S,
                                         ptr(out), stream_ptr()), "b2_splice_ids")
        return out

    def take_async_error(self):
        """Bit mask of the input problems kernels have flagged since the last call (ERR_* below); clears it."""
        code = _c.c_int(0)
        check(self.lib.b2_async_error(self.handle, ctypes.byref(code)), "b2_async_error")
        return code.value

    def check_async_error(self):
        """Raise ValueError if a kernel flagged bad inputs (ids outside the embedding table, image placeholder without
        features). Call after a point where the stream has been synchronised (e.g. once the first token is on the host)."""
        if self.take_async_error():
            raise ValueError(last_error())

    def prefill(self, kv, embeds, seq_lens=None, logits_mode=LOGITS_LAST, slot0=0):
        """`slot0`: first cache slot to fill (continuous batching); the other slots keep their contents."""
        embeds = self._bf16(embeds)
        B, S = embeds.shape[0], embeds.shape[1]
        lens = None
        if seq_lens is not None:
            lens = (_c.c_int32 * B)(*[int(x) for x in seq_lens])
        logits = None
        if logits_mode == LOGITS_LAST:
            logits = torch.empty(B, self.vocab, dtype=torch.float32, device=self.device)
        elif logits_mode == LOGITS_ALL:
            logits = torch.empty(B, S, self.vocab, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.index):
            check(self.lib.b2_prefill_slots(self.handle, kv.handle, ptr(embeds), lens, B, S, int(slot0), ptr(logits), logits_mode,
                                            stream_ptr()), "b2_prefill")
        return logits

    def decode_step(self, kv, tokens, want_logits=True):
        """tokens: int32 tensor [B] (cpu or cuda). Returns fp32 logits [B, vocab] (device)."""
        tokens = tokens.to(torch.int32).contiguous()
        B = tokens.numel()
        logits = torch.empty(B, self.vocab, dtype=torch.float32, device=self.device) if want_logits else None
        with torch.cuda.device(self.index):
            check(self.lib.b2_decode_step(self.handle, kv.handle, ptr(tokens), B, ptr(logits), _vp(0), stream_ptr()),
                  "b2_decode_step")
        return logits

    def decode_greedy(self, kv, first_tokens, n_steps, out=None):
        """Runs n_steps greedy steps on the device (CUDA-graph replay). Returns int32 [n_steps, B];
        `out` may be a pinned CPU tensor to receive the tokens directly."""
        first_tokens = first_tokens.to(torch.int32).contiguous()
        B = first_tokens.numel()
        if out is None:
            out = torch.empty(n_steps, B, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.index):
            check(self.lib.b2_decode_greedy(self.handle, kv.handle, ptr(first_tokens), B, int(n_steps), ptr(out),
                                            stream_ptr()), "b2_decode_greedy")
        return out

    # -- streaming decode (device runs ahead, host reads tokens from mapped pinned memory) ---------------
    def stream_begin(self, kv, logits, sampling=None):
        """Token 0 is chosen from the prefill logits [B, vocab] on the device and published as ring index 0."""
        logits = logits.contiguous()
        sp = sampling if sampling is not None else make_sampling()
        with torch.cuda.device(self.index):
            check(self.lib.b2_stream_begin(self.handle, kv.handle, ptr(logits), int(logits.shape[0]), ctypes.byref(sp),
                                           stream_ptr()), "b2_stream_begin")

    def batch_begin(self, kv, B):
        with torch.cuda.device(self.index):
            check(self.lib.b2_batch_begin(self.handle, kv.handle, int(B), stream_ptr()), "b2_batch_begin")

    def batch_set_row(self, kv, slot, active, sampling=None, first_token=0):
        sp = sampling if sampling is not None else make_sampling()
        with torch.cuda.device(self.index):
            check(self.lib.b2_batch_set_row(self.handle, kv.handle, int(slot), int(bool(active)), ctypes.byref(sp), int(first_token),
                                            stream_ptr()), "b2_batch_set_row")

    def stream_enqueue(self, kv, n_steps):
        with torch.cuda.device(self.index):
            check(self.lib.b2_stream_enqueue(self.handle, kv.handle, int(n_steps), stream_ptr()), "b2_stream_enqueue")

    def stream_wait(self, kv, index, B, timeout_ms=60000):
        """Blocks (GIL released by ctypes) until token `index` is visible; returns a list of B ints."""
        out = (_c.c_int32 * B)()
        check(self.lib.b2_stream_wait(kv.handle, int(index), out, int(timeout_ms)), "b2_stream_wait")
        return list(out)

    def sample(self, logits, sampling, index=0):
        """One selection per row of fp32 logits [B, V] with csrc/sampling.cu (greedy or temperature/top-k/top-p)."""
        logits = logits.to(device=self.device, dtype=torch.float32).contiguous()
        B, V = logits.shape
        out = torch.empty(B, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.index):
            check(self.lib.b2_op_sample(ptr(logits), B, V, ctypes.byref(sampling), int(index), ptr(out), stream_ptr()),
                  "b2_op_sample")
        return out

    def argmax(self, logits):
        B, V = logits.shape
        out = torch.empty(B, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.index):
            check(self.lib.b2_argmax(ptr(logits), B, V, ptr(out), stream_ptr()), "b2_argmax")
        return out

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            for kv in list(self._kvs):  # caches point into the model: release them first
                kv.close()
            self.lib.b2_model_destroy(self.handle)
            self.handle = _vp(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
