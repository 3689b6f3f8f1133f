"""LLaVA multimodal glue on the B200 engine: same class surface as the reference's
llava/model/llava_arch.py (LlavaMetaModel :27-82, LlavaMetaForCausalLM :85-284), different execution:

* `encode_images` (ref :94-97) runs the CLIP tower + mlp2x_gelu projector inside libb2llava.so.
* `prepare_inputs_labels_for_multimodal` (ref :99-240) keeps the reference's return contract and edge-case
  semantics (global image index, zero-image rows consuming a slot, 5-D / list images, padding side,
  truncation, None-mirroring), but the embedding gather + image-feature splice is ONE device kernel driven by
  a host-built source-row index instead of per-row torch.cat chains.
"""
from abc import ABC, abstractmethod

import numpy as np
import torch
import torch.nn as nn

from ..constants import IGNORE_INDEX, IMAGE_TOKEN_INDEX, DEFAULT_IMAGE_PATCH_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN
from .multimodal_encoder.builder import build_vision_tower
from .multimodal_projector.builder import build_vision_projector

_PAD_ROW = -(2**31)


class LlavaMetaModel:
    """Holds `vision_tower` + `mm_projector` next to the decoder weights (ref llava_arch.py:27-40)."""

    def __init__(self, config):
        super(LlavaMetaModel, self).__init__(config)
        if hasattr(config, "mm_vision_tower"):
            self.vision_tower = build_vision_tower(config, delay_load=True)
            self.mm_projector = build_vision_projector(config)

    def get_vision_tower(self):
        vision_tower = getattr(self, "vision_tower", None)
        if type(vision_tower) is list:
            vision_tower = vision_tower[0]
        return vision_tower

    def initialize_vision_modules(self, model_args, fsdp=None):
        """Entry point of the reference's training script (ref llava_arch.py:42-82), kept because callers of the class
        surface may use it to attach a tower to a bare language model: make sure a loaded tower and a projector exist and
        record the multimodal settings in the config. Training-only details (FSDP list wrapping aside) are not reproduced."""
        wrap = fsdp is not None and len(fsdp) > 0
        tower = self.get_vision_tower()
        if tower is None:
            tower = build_vision_tower(model_args)
            self.vision_tower = [tower] if wrap else tower
        else:
            tower.load_model()
        cfg = self.config
        cfg.mm_vision_tower = model_args.vision_tower
        cfg.use_mm_proj = True
        cfg.mm_projector_type = getattr(model_args, "mm_projector_type", "linear")
        cfg.mm_hidden_size = tower.hidden_size
        cfg.mm_vision_select_layer = model_args.mm_vision_select_layer
        cfg.mm_vision_select_feature = model_args.mm_vision_select_feature
        if getattr(self, "mm_projector", None) is None:
            self.mm_projector = build_vision_projector(cfg)
        adapter = getattr(model_args, "pretrain_mm_mlp_adapter", None)
        if adapter is not None:
            state = torch.load(adapter, map_location="cpu")
            self.mm_projector.load_state_dict({k.split("mm_projector.", 1)[1]: v for k, v in state.items() if "mm_projector" in k})
        self._engine_dirty()

    def _engine_dirty(self):
        owner = getattr(self, "_owner", None)
        if owner is not None and owner() is not None:
            owner().invalidate_engine()


def build_source_index(input_ids, attention_mask, labels, num_image_rows, feats_per_image, max_length, padding_side,
                       vocab_size=None):
    """Host half of the splice (ref llava_arch.py:143-225), pure numpy.

    input_ids [B, Lt] int64 numpy (IMAGE_TOKEN_INDEX marks an image), attention_mask bool [B, Lt],
    labels int64 [B, Lt]; feats_per_image: list with the number of feature rows of each image slot (global,
    row-major order, ref :149-179). With `vocab_size`, ids outside the embedding table raise ValueError here instead of
    reaching the gather kernel. Returns (src_index int32 [B,S], new_labels int64 [B,S], mask bool [B,S],
    position_ids int64 [B,S], lens list[int]).
    """
    B = input_ids.shape[0]
    if vocab_size is not None:
        bad = ((input_ids >= vocab_size) | ((input_ids < 0) & (input_ids != IMAGE_TOKEN_INDEX))) & attention_mask
        if bad.any():
            b, i = np.argwhere(bad)[0]
            raise ValueError(f"input_ids[{b}, {i}] = {int(input_ids[b, i])} is outside the embedding table [0, {vocab_size}) "
                             f"(and is not IMAGE_TOKEN_INDEX = {IMAGE_TOKEN_INDEX})")
    rows_src, rows_lab = [], []
    cur_image_idx = 0
    feat_offset = np.concatenate([[0], np.cumsum(feats_per_image)]).astype(np.int64)
    for b in range(B):
        ids = input_ids[b][attention_mask[b]]
        lab = labels[b][attention_mask[b]]
        img_pos = np.nonzero(ids == IMAGE_TOKEN_INDEX)[0]
        if len(img_pos) == 0:
            # ref :152-159: a row without <image> still consumes one image slot (empty slice of its features)
            if cur_image_idx >= len(feats_per_image):
                raise IndexError("list index out of range")  # same failure as the reference (App. C.2)
            cur_image_idx += 1
            rows_src.append(ids.astype(np.int64))
            rows_lab.append(lab)
            continue
        src_parts, lab_parts = [], []
        prev = -1
        for p in list(img_pos) + [len(ids)]:
            src_parts.append(ids[prev + 1:p].astype(np.int64))
            lab_parts.append(lab[prev + 1:p])
            if p < len(ids):
                if cur_image_idx >= len(feats_per_image):
                    raise IndexError("list index out of range")
                n = int(feats_per_image[cur_image_idx])
                start = int(feat_offset[cur_image_idx])
                src_parts.append(-(np.arange(start, start + n, dtype=np.int64)) - 1)
                lab_parts.append(np.full((n,), IGNORE_INDEX, dtype=lab.dtype))
                cur_image_idx += 1
            prev = p
        rows_src.append(np.concatenate(src_parts))
        rows_lab.append(np.concatenate(lab_parts))
    if max_length is not None:  # ref :190-193
        rows_src = [r[:max_length] for r in rows_src]
        rows_lab = [r[:max_length] for r in rows_lab]
    lens = [int(r.shape[0]) for r in rows_src]
    S = max(lens)
    src = np.full((B, S), _PAD_ROW, dtype=np.int64)
    new_labels = np.full((B, S), IGNORE_INDEX, dtype=np.int64)
    mask = np.zeros((B, S), dtype=bool)
    pos = np.zeros((B, S), dtype=np.int64)
    for b in range(B):
        n = lens[b]
        if n == 0:
            continue
        sl = slice(S - n, S) if padding_side == "left" else slice(0, n)  # ref :206-225
        src[b, sl] = rows_src[b]
        new_labels[b, sl] = rows_lab[b]
        mask[b, sl] = True
        pos[b, sl] = np.arange(n)
    return src.astype(np.int32), new_labels, mask, pos, lens


class LlavaMetaForCausalLM(ABC):

    @abstractmethod
    def get_model(self):
        pass

    def get_vision_tower(self):
        return self.get_model().get_vision_tower()

    def encode_images(self, images):
        """ref llava_arch.py:94-97: vision tower then mm_projector — one engine call, [n,3,H,W] -> [n,P,h]."""
        engine = self._ensure_engine()
        return engine.encode_images(images)

    def _image_features(self, images):
        """ref :114-121. Returns (features [rows, h] bf16 on device, rows-per-image-slot list)."""
        P = self.get_vision_tower().num_patches
        if type(images) is list or images.ndim == 5:
            concat_images = torch.cat([image for image in images], dim=0)
            feats = self.encode_images(concat_images)
            split_sizes = [image.shape[0] for image in images]
            return feats.reshape(-1, feats.shape[-1]), [n * P for n in split_sizes]
        feats = self.encode_images(images)
        return feats.reshape(-1, feats.shape[-1]), [P] * feats.shape[0]

    def prepare_inputs_labels_for_multimodal(self, input_ids, position_ids, attention_mask, past_key_values, labels, images):
        return self._prepare_multimodal(input_ids, position_ids, attention_mask, past_key_values, labels, images)[0]

    def _spliced_embeds(self, input_ids, attention_mask, images, force_host=False):
        """generate()'s use of the splice: (inputs_embeds [B,S,h], valid length per row, speculative), or (None, None, False)
        when the multimodal branch passes the ids through (single-token prompts).

        Ids that already live on the device are spliced THERE (b2_splice_ids: prefix count over `ids == IMAGE_TOKEN_INDEX`, no
        D2H of the ids, no host loop; the reference syncs per row, llava_arch.py:143-187) for the layout every generation
        caller uses — equal-length rows, no padding mask, the same number of placeholders per row. The output length then
        follows from shapes alone, on the assumption that each row holds n_images / B placeholders: `speculative` tells the
        caller to confirm it (Engine.take_async_error() & ERR_SPLICE_SLOTS after its first sync) and to redo the splice with
        force_host=True otherwise, which reproduces the reference's exact semantics (extra images ignored, IndexError ...)."""
        tower = self.get_vision_tower()
        if (not force_host and tower is not None and images is not None and attention_mask is None and input_ids.is_cuda
                and input_ids.shape[1] > 1):
            B, Lt = input_ids.shape
            image_feats, feats_per_image = self._image_features(images)
            n_img = len(feats_per_image)
            k = n_img // B if n_img % B == 0 else 0
            per_row = {sum(feats_per_image[b * k:(b + 1) * k]) for b in range(B)} if k else set()
            max_len = getattr(self.config, "tokenizer_model_max_length", None)
            if k >= 1 and len(per_row) == 1 and n_img < 128 and (max_len is None or Lt - k + max(per_row) <= max_len):
                engine = self._ensure_engine()
                ids = input_ids.to(torch.int64).contiguous()
                embeds = engine.splice_ids(ids, k, feats_per_image, image_feats)
                return embeds, [embeds.shape[1]] * B, True
            out, lens = self._prepare_multimodal(input_ids, None, attention_mask, None, None, images,
                                                 encoded=(image_feats, feats_per_image))
            return out[4], lens, False
        out, lens = self._prepare_multimodal(input_ids, None, attention_mask, None, None, images)
        return out[4], lens, False

    def _prepare_multimodal(self, input_ids, position_ids, attention_mask, past_key_values, labels, images, encoded=None):
        """The reference's 6-tuple plus the per-row valid lengths of the spliced sequence (no state is kept on `self`:
        concurrent calls from several threads must not see each other's lengths)."""
        vision_tower = self.get_vision_tower()
        if vision_tower is None or images is None or input_ids.shape[1] == 1:
            # decode branch (ref :103-112): the engine keeps the cache length itself, so only the mask/position
            # bookkeeping the caller can observe is reproduced
            if past_key_values is not None and vision_tower is not None and images is not None and input_ids.shape[1] == 1:
                target_shape = past_key_values.get_seq_length() + 1
                if attention_mask is not None:
                    attention_mask = torch.cat((attention_mask, torch.ones(
                        (attention_mask.shape[0], target_shape - attention_mask.shape[1]),
                        dtype=attention_mask.dtype, device=attention_mask.device)), dim=1)
                    position_ids = torch.sum(attention_mask, dim=1).unsqueeze(-1) - 1
            return (input_ids, position_ids, attention_mask, past_key_values, None, labels), None

        image_feats, feats_per_image = encoded if encoded is not None else self._image_features(images)

        if getattr(self.config, "tune_mm_mlp_adapter", False) and getattr(self.config, "mm_use_im_start_end", False):
            raise NotImplementedError

        _labels, _position_ids, _attention_mask = labels, position_ids, attention_mask
        ids_np = input_ids.detach().cpu().numpy().astype(np.int64)  # one D2H of B*Lt ints (ref syncs per row)
        mask_np = np.ones_like(ids_np, dtype=bool) if attention_mask is None else attention_mask.detach().bool().cpu().numpy()
        lab_np = np.full_like(ids_np, IGNORE_INDEX) if labels is None else labels.detach().cpu().numpy().astype(np.int64)

        src, new_labels, mask, pos, lens = build_source_index(
            ids_np, mask_np, lab_np, image_feats.shape[0], feats_per_image,
            getattr(self.config, "tokenizer_model_max_length", None),
            getattr(self.config, "tokenizer_padding_side", "right"),
            vocab_size=self.get_model().embed_tokens.weight.shape[0])
        B, S = src.shape
        engine = self._ensure_engine()
        src_dev = torch.from_numpy(src.reshape(-1)).to(engine.device, non_blocking=True)
        new_input_embeds = engine.splice(src_dev, image_feats, B, S)

        dev = input_ids.device
        new_labels_t = None if _labels is None else torch.from_numpy(new_labels).to(device=dev, dtype=_labels.dtype)
        attention_mask_t = None if _attention_mask is None else torch.from_numpy(mask).to(device=dev, dtype=_attention_mask.dtype)
        position_ids_t = None if _position_ids is None else torch.from_numpy(pos).to(device=dev, dtype=_position_ids.dtype)
        return (None, position_ids_t, attention_mask_t, past_key_values, new_input_embeds, new_labels_t), lens

    def initialize_vision_tokenizer(self, model_args, tokenizer):
        """Token bookkeeping for the optional <im_patch>/<im_start>/<im_end> tokens (ref llava_arch.py:242-284).
        LLaVA-1.5 enables neither flag; training-only freezing/adapter-loading branches are out of scope."""
        new_tokens = []
        if getattr(model_args, "mm_use_im_patch_token", False):
            new_tokens.append(DEFAULT_IMAGE_PATCH_TOKEN)
        if getattr(model_args, "mm_use_im_start_end", False):
            new_tokens += [DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN]
        if not new_tokens:
            return
        n_old = self.get_input_embeddings().weight.shape[0]
        tokenizer.add_tokens(new_tokens, special_tokens=True)
        self.resize_token_embeddings(len(tokenizer))
        n_new = self.get_input_embeddings().weight.shape[0] - n_old
        if n_new > 0 and getattr(model_args, "mm_use_im_start_end", False):
            # new rows start at the mean of the existing rows, as the reference initialises them
            for emb in (self.get_input_embeddings().weight.data, self.get_output_embeddings().weight.data):
                emb[n_old:] = emb[:n_old].mean(dim=0, keepdim=True)
        if getattr(model_args, "pretrain_mm_mlp_adapter", None) and getattr(model_args, "mm_use_im_start_end", False):
            raise NotImplementedError("loading <im_start>/<im_end> rows from a pretraining adapter is a training-only path")
        self.invalidate_engine()
