"""Stand-in for the `pyquaternion` package the reference's networks/occ3d_proj.py imports (absent from this image):
only `Quaternion(q).rotation_matrix` for a unit quaternion (w, x, y, z) is used (occ3d_proj.py:71)."""
import numpy as np


class Quaternion:
    def __init__(self, q):
        self.q = np.asarray(q, dtype=np.float64)

    @property
    def rotation_matrix(self):
        w, x, y, z = self.q / np.linalg.norm(self.q)
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
