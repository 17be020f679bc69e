"""Minimal stand-in for the `timm` names the reference imports (compat/README.md); timm is not installed in this image."""
__version__ = "0.0-tulip-b200-shim"
