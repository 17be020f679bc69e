"""Drop-in seam: with this directory ahead of the reference's `tulip/` on sys.path, the unchanged
driver's `import model.tulip as tulip` (main_lidar_upsampling.py:29) resolves to the CUDA module."""
from tulip_b200.model.tulip import *  # noqa: F401,F403
from tulip_b200.model.tulip import tulip_base, tulip_large, TULIP  # noqa: F401
