"""Constants of the multimodal path.

Only the values the B200 hot path and its callers need are defined here; they are part of the reference's wire
format (token ids inside `input_ids`, label masking) and therefore must keep the reference's names and values
(reference: llava/constants.py:1-13):

* `IMAGE_TOKEN_INDEX` (-200) marks an `<image>` placeholder inside `input_ids` (written by
  `llava.mm_utils.tokenizer_image_token`); `prepare_inputs_labels_for_multimodal` replaces each marker by the 576
  projected patch features of one image. It is negative so it can never collide with a vocabulary id.
* `IGNORE_INDEX` (-100) is the label value the loss ignores (the image span gets it).
* the `DEFAULT_*` strings are the textual placeholders of the prompt templates.
* the two heart-beat periods and `LOGDIR` belong to the serving control plane (out of scope here) and are kept
  only so that `from llava.constants import *` in the reference's own `llava.serve` modules keeps working when
  those modules are layered on top of this package.
"""

# --- label / token-id conventions used by the splice -------------------------------------------------------
IGNORE_INDEX: int = -100
IMAGE_TOKEN_INDEX: int = -200

# --- prompt placeholders -----------------------------------------------------------------------------------
DEFAULT_IMAGE_TOKEN: str = "<image>"
IMAGE_PLACEHOLDER: str = "<image-placeholder>"
DEFAULT_IMAGE_PATCH_TOKEN: str = "<im_patch>"
DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN = "<im_start>", "<im_end>"

# --- serving control plane (seconds) and log directory: consumed by the reference's llava.serve only -------
WORKER_HEART_BEAT_INTERVAL: int = 15
CONTROLLER_HEART_BEAT_EXPIRATION: int = 2 * WORKER_HEART_BEAT_INTERVAL
LOGDIR: str = "."
