from .inference import BoxCoder, PostProcessor, make_roi_box_post_processor  # noqa: F401
