"""trimesh.PointCloud(...).export(path) as engine_upsampling.py:318-325 uses it (only with --save_pcd): ASCII PLY writer."""
import numpy as np


class PointCloud:
    def __init__(self, vertices, colors=None):
        self.vertices = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)

    def export(self, path):
        with open(path, "w") as f:
            f.write(f"ply\nformat ascii 1.0\nelement vertex {len(self.vertices)}\nproperty float x\nproperty float y\nproperty float z\nend_header\n")
            for v in self.vertices:
                f.write(f"{v[0]:.6f} {v[1]:.6f} {v[2]:.6f}\n")
