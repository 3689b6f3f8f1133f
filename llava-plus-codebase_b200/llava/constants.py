"""Wire-format constants of the multimodal path (names and values of the reference's llava/constants.py:1-13)."""

IGNORE_INDEX = -100          # label value the loss ignores (the image span gets it)
IMAGE_TOKEN_INDEX = -200     # `<image>` placeholder inside input_ids, replaced by 576 projected patch features

DEFAULT_IMAGE_TOKEN = "<image>"
IMAGE_PLACEHOLDER = "<image-placeholder>"
DEFAULT_IMAGE_PATCH_TOKEN = "<im_patch>"
DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN = "<im_start>", "<im_end>"

# used only by the reference's own llava.serve modules when they are layered on top of this package
WORKER_HEART_BEAT_INTERVAL = 15
CONTROLLER_HEART_BEAT_EXPIRATION = 30
LOGDIR = "."
