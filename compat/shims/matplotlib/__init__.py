"""The three matplotlib names engine_upsampling.py:20-37 uses to colour range images for TensorBoard."""
