from . import loss, model, strategy  # noqa: F401
