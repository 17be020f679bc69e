from .tulip import TULIP, tulip_base, tulip_large  # noqa: F401
