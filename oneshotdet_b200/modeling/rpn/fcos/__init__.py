from .inference import FCOSPostProcessor, make_fcos_postprocessor  # noqa: F401
