"""OccupancyRay — Occupancy Ray-shape Sampling projector on the GPU (reference: networks/occ3d_proj.py:10-113).

Same constructor arguments and `project(sample_token)` contract as the reference class (6 x h x w x sample_point class
ids, 17 = empty), with two differences: the camera table / Occ3D index can be passed in directly (the reference's
`camera.pkl` / `occ3d_idx.pkl` blobs are not part of the repository), and `project_rows` returns the normalised bf16
rows the foreground ControlNet branch consumes (dataset/utils.py:412-420) without materialising the (6, 320, h, w)
tensor.  Ray origins / directions are computed on the host with the reference's own formulae (occ3d_proj.py:26-42:
1400 pixels per camera, negligible), the 2.7 M nearest-voxel lookups run in `dd_ors_project`."""
import os
import pickle
from typing import Dict, Optional

import numpy as np
import torch

CAMS = ['CAM_FRONT_LEFT', 'CAM_FRONT', 'CAM_FRONT_RIGHT', 'CAM_BACK_RIGHT', 'CAM_BACK', 'CAM_BACK_LEFT']  # occ3d_proj.py:63


def quaternion_rotation_matrix(q) -> np.ndarray:
    """unit quaternion (w, x, y, z) -> 3x3 rotation (what pyquaternion.Quaternion(q).rotation_matrix returns)"""
    w, x, y, z = np.asarray(q, dtype=np.float64) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


class OccupancyRay:
    def __init__(self, image_shape=(900, 1600), sample_point=200, sample_step=0.2, compress_ratio=8,
                 dataroot='./data/nuscenes/', device='cuda', camera_data: Optional[Dict] = None,
                 occ3d_idx: Optional[Dict] = None, pkl_root='magicdrive/networks'):
        if camera_data is None:
            with open(os.path.join(pkl_root, 'camera.pkl'), 'rb') as f:
                camera_data = pickle.load(f)
        if occ3d_idx is None and os.path.exists(os.path.join(pkl_root, 'occ3d_idx.pkl')):
            with open(os.path.join(pkl_root, 'occ3d_idx.pkl'), 'rb') as f:
                occ3d_idx = pickle.load(f)
        self.camera_data, self.occ3d_idx = camera_data, occ3d_idx
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("dualdiff_b200 has no CPU path: OccupancyRay needs a CUDA device")
        self.dataroot = dataroot
        self.image_shape = image_shape
        self.sample_point = sample_point
        self.sample_step = sample_step
        self.compress_ratio = compress_ratio
        self.image_shape_compress = [int(image_shape[0] * compress_ratio), int(image_shape[1] * compress_ratio)]

    # occ3d_proj.py:26-42 + :64-87, for all six cameras
    def rays(self, cam_data):
        h, w = self.image_shape_compress
        origins, dirs = [], []
        for key in CAMS:
            T = np.eye(4)
            T[:3, :3] = quaternion_rotation_matrix(np.array(cam_data[key]['rotation']))
            T[:3, 3] = np.array(cam_data[key]['translation'])
            Rt = torch.from_numpy(T).to(torch.float32)
            K = torch.Tensor(cam_data[key]['intrinsic']).to(torch.float32)
            xx, yy = torch.meshgrid(torch.arange(0, w), torch.arange(0, h), indexing='ij')
            u = (xx.flatten() // self.compress_ratio).to(torch.float32)
            v = (yy.flatten() // self.compress_ratio).to(torch.float32)
            pix = torch.stack([u, v, torch.ones_like(u)], dim=1)
            d = torch.matmul(Rt[:3, :3], torch.matmul(torch.inverse(K), pix.T)).T
            d = d / torch.norm(d, dim=1, keepdim=True)
            o = Rt[:3, 3].expand_as(d)
            origins.append(o.view(w, h, 3).permute(1, 0, 2))
            dirs.append(d.view(w, h, 3).permute(1, 0, 2))
        o = torch.stack(origins).reshape(-1, 3).contiguous().to(self.device)
        d = torch.stack(dirs).reshape(-1, 3).contiguous().to(self.device)
        return o, d

    def _semantics(self, sample_token):
        f = os.path.join(self.dataroot, self.occ3d_idx[sample_token], 'labels.npz')
        return torch.from_numpy(np.load(f)['semantics'].astype(np.uint8))

    def project_arrays(self, semantics: torch.Tensor, cam_data) -> torch.Tensor:
        """semantics: integer labels [200, 200, 16] -> int64 [6, h, w, sample_point] (the reference's return value)"""
        from .. import ops
        h, w = self.image_shape_compress
        o, d = self.rays(cam_data)
        ids, _ = ops.ors_project(o, d, semantics.to(torch.uint8).contiguous().to(self.device), sample_point=self.sample_point,
                                 sample_step=self.sample_step)
        return ids.view(6, h, w, self.sample_point).long()

    def project(self, sample_token):
        return self.project_arrays(self._semantics(sample_token), self.camera_data[sample_token])

    def project_rows(self, semantics: torch.Tensor, cam_data, use_fg=True, use_bg=True) -> torch.Tensor:
        """bf16 [6*h*w, sample_point]: filter(ids) / 17 in the channels-last row layout of the fg branch"""
        from .. import ops
        o, d = self.rays(cam_data)
        _, rows = ops.ors_project(o, d, semantics.to(torch.uint8).contiguous().to(self.device), sample_point=self.sample_point,
                                  sample_step=self.sample_step, want_ids=False, want_rows=True, keep_fg=use_fg, keep_bg=use_bg)
        return rows
