"""Runs the REFERENCE's own OccupancyRay.project (networks/occ3d_proj.py, unmodified, imported from /root/reference) on
synthetic cameras and a synthetic Occ3D volume, and stores inputs + output as tests/golden/ors_small.pt.
The class reads two pickles and an .npz from disk: the script writes synthetic ones into a temp dir and chdirs there.
    python oracle/make_golden_ors.py"""
import os, pickle, sys, tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference/MD_txt_con_fusion")


def synthetic_scene(seed=0):
    rng = np.random.default_rng(seed)
    sem = np.full((200, 200, 16), 17, dtype=np.uint8)
    sem[:, :, :3] = 11                                   # ground layers: driveable surface
    for _ in range(60):                                  # boxes of random classes
        x, y, z = rng.integers(0, 180), rng.integers(0, 180), rng.integers(2, 10)
        sem[x:x + rng.integers(3, 20), y:y + rng.integers(3, 20), z:z + rng.integers(1, 6)] = rng.integers(0, 17)
    cams = {}
    yaw0 = {'CAM_FRONT_LEFT': 55, 'CAM_FRONT': 0, 'CAM_FRONT_RIGHT': -55, 'CAM_BACK_RIGHT': -110, 'CAM_BACK': 180, 'CAM_BACK_LEFT': 110}
    for name, yaw in yaw0.items():
        a = np.deg2rad(yaw + rng.uniform(-2, 2))
        # camera looking along +x of the ego frame rotated by yaw: columns = camera axes (x right, y down, z forward)
        fwd = np.array([np.cos(a), np.sin(a), 0.0]); right = np.array([np.sin(a), -np.cos(a), 0.0]); down = np.array([0, 0, -1.0])
        Rm = np.stack([right, down, fwd], axis=1)
        qw = np.sqrt(max(0.0, 1 + Rm[0, 0] + Rm[1, 1] + Rm[2, 2])) / 2
        if qw > 1e-6:
            q = [qw, (Rm[2, 1] - Rm[1, 2]) / (4 * qw), (Rm[0, 2] - Rm[2, 0]) / (4 * qw), (Rm[1, 0] - Rm[0, 1]) / (4 * qw)]
        else:
            q = [0.0, 1.0, 0.0, 0.0]
        cams[name] = dict(translation=[rng.uniform(-1, 1), rng.uniform(-0.5, 0.5), rng.uniform(1.4, 1.6)], rotation=q,
                          intrinsic=[[1266.4 + rng.uniform(-5, 5), 0, 816.3], [0, 1266.4 + rng.uniform(-5, 5), 491.5], [0, 0, 1]])
    return sem, cams


def main():
    from pyquaternion import Quaternion
    sem, cams = synthetic_scene(0)
    token = "synthetic0"
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "magicdrive", "networks"))
    os.makedirs(os.path.join(tmp, "data", "scene0"))
    pickle.dump({token: cams}, open(os.path.join(tmp, "magicdrive", "networks", "camera.pkl"), "wb"))
    pickle.dump({token: "scene0"}, open(os.path.join(tmp, "magicdrive", "networks", "occ3d_idx.pkl"), "wb"))
    np.savez(os.path.join(tmp, "data", "scene0", "labels.npz"), semantics=sem)
    os.chdir(tmp)
    from magicdrive.networks.occ3d_proj import OccupancyRay            # the reference's own class
    image_shape, ratio, S = (896, 1600), 400 / 8 / 1600, 64
    proj = OccupancyRay(image_shape=image_shape, sample_point=S, compress_ratio=ratio, dataroot=os.path.join(tmp, "data"))
    out = proj.project(token)                                          # [6, 28, 50, S] int64
    Ks = torch.stack([torch.tensor(cams[k]['intrinsic'], dtype=torch.float32) for k in
                      ['CAM_FRONT_LEFT', 'CAM_FRONT', 'CAM_FRONT_RIGHT', 'CAM_BACK_RIGHT', 'CAM_BACK', 'CAM_BACK_LEFT']])
    Rts = []
    for k in ['CAM_FRONT_LEFT', 'CAM_FRONT', 'CAM_FRONT_RIGHT', 'CAM_BACK_RIGHT', 'CAM_BACK', 'CAM_BACK_LEFT']:
        T = np.eye(4); T[:3, :3] = Quaternion(np.array(cams[k]['rotation'])).rotation_matrix; T[:3, 3] = np.array(cams[k]['translation'])
        Rts.append(torch.from_numpy(T).float())
    torch.save(dict(semantics=torch.from_numpy(sem), Ks=Ks, Rts=torch.stack(Rts), image_shape=image_shape, compress_ratio=ratio,
                    sample_point=S, sample_step=0.2, out=out.to(torch.uint8)), os.path.join(ROOT, "tests", "golden", "ors_small.pt"))
    print("golden written:", tuple(out.shape), "classes present:", sorted(set(out.flatten().tolist()))[:18],
          "non-empty fraction:", float((out != 17).float().mean()))


if __name__ == "__main__":
    main()
