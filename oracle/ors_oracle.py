"""Oracle (test infrastructure, CPU) for the Occupancy Ray-shape Sampling projector -- SURVEY.md §8f rank 2.

Restates networks/occ3d_proj.py:26-113 (`OccupancyRay.compute_rays` + `.project`): for each of the six cameras and
each pixel of the compressed image grid, `sample_point` points are taken along the pixel's ray every `sample_step`
metres, and the semantic class of the Occ3D voxel (200 x 200 x 16, 0.4 m, z from -1 m) nearest to each point is
returned; points outside the grid get class 17.  The reference does this with a one-hot (18-channel) volume and
`F.grid_sample(mode='nearest', padding_mode='zeros', align_corners=False)` followed by an argmax; the restatement
gathers the label directly, which is the same function.  Pinned by tests/golden/ors_*.pt, produced by running the
reference's own class on synthetic cameras / voxels (oracle/make_golden_ors.py)."""
import torch

CAMS = ['CAM_FRONT_LEFT', 'CAM_FRONT', 'CAM_FRONT_RIGHT', 'CAM_BACK_RIGHT', 'CAM_BACK', 'CAM_BACK_LEFT']  # occ3d_proj.py:63


def compute_rays(K, Rt, u, v):
    """occ3d_proj.py:26-42"""
    K_inv = torch.inverse(K.float())
    R, t = Rt[:3, :3].float(), Rt[:3, 3].float()
    pix = torch.stack([u.float(), v.float(), torch.ones_like(u, dtype=torch.float32)], dim=1)
    d = torch.matmul(R, torch.matmul(K_inv, pix.T)).T
    d = d / torch.norm(d, dim=1, keepdim=True)
    return t.expand_as(d), d


def nearest_index(g, size):
    """F.grid_sample(mode='nearest', align_corners=False): unnormalise, round half to even"""
    return torch.round(((g + 1.0) * size - 1.0) / 2.0)


def project(semantics, Ks, Rts, image_shape, compress_ratio, sample_point=320, sample_step=0.2):
    """semantics: int [200, 200, 16]; Ks [6, 3, 3]; Rts [6, 4, 4] -> int64 [6, h, w, sample_point] (occ3d_proj.py:50-113)"""
    h, w = int(image_shape[0] * compress_ratio), int(image_shape[1] * compress_ratio)
    sem = semantics.long()
    D, H, W = sem.shape          # grid_sample's (D, H, W) = the volume's three axes in storage order
    outs = []
    for c in range(6):
        xx, yy = torch.meshgrid(torch.arange(w), torch.arange(h), indexing='ij')
        gx, gy = xx.flatten() // compress_ratio, yy.flatten() // compress_ratio          # :80-81
        o, d = compute_rays(Ks[c], Rts[c], gx, gy)
        o = o.view(w, h, 3).permute(1, 0, 2).contiguous()
        d = d.view(w, h, 3).permute(1, 0, 2).contiguous()
        steps = torch.arange(sample_point).float() * sample_step
        pts = o.unsqueeze(2) + steps.view(1, 1, -1, 1) * d.unsqueeze(2)                   # [h, w, S, 3] metres
        grid = pts / 40
        grid[..., 2] = grid[..., 2] * 40 / 3.2 - 2.2 / 3.2                                # :92
        # :93-94 reorders to (x, y, z) <- (z', y, x): grid_sample's x indexes the LAST volume axis (16 levels),
        # y the middle axis, z the first axis
        ix = nearest_index(grid[..., 2], W)     # last axis  <- height
        iy = nearest_index(grid[..., 1], H)     # middle axis <- y
        iz = nearest_index(grid[..., 0], D)     # first axis <- x
        inb = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H) & (iz >= 0) & (iz < D)
        lab = sem[iz.clamp(0, D - 1).long(), iy.clamp(0, H - 1).long(), ix.clamp(0, W - 1).long()]
        outs.append(torch.where(inb, lab, torch.full_like(lab, 17)))
    return torch.stack(outs, 0)
