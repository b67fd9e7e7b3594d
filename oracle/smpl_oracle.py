"""TEST INFRASTRUCTURE ONLY — CPU restatement (fp32/fp64, plain PyTorch ops) of the SMPL forward the decoders call
(reference lib/models/ktd.py:100-114, lib/models/spin.py:92-104 -> lib/models/smpl.py:84-106 -> smplx.SMPL.forward).

**Parity unpinned.**  The arithmetic lives in `smplx==0.1.13` (reference requirements.txt:4), which is neither vendored
under /root/reference nor installed, and the model data (SMPL_NEUTRAL.pkl, J_regressor_extra.npy) is licensed and absent.
This file restates the PUBLISHED algorithm of `smplx.lbs.lbs` (Loper et al., SMPL, SIGGRAPH Asia 2015; smplx/lbs.py):

    v_shaped = v_template + shapedirs . betas
    J        = J_regressor @ v_shaped
    v_posed  = v_shaped + posedirs^T . vec(R_1..23 - I)
    G_j      = G_parent(j) [R_j | J_j - J_parent(j)]            (kinematic chain, root: [R_0 | J_0])
    A_j      = [G_j.R | G_j.t - G_j.R J_j]
    verts    = (sum_j W_vj A_j) [v_posed; 1]
    joints45 = [G_j.t (24) | verts[extra_vertex_ids] (21)]      (smplx VertexJointSelector)
    joints54 = [joints45 | J_regressor_extra @ verts (9)];  out = joints54[joint_map] (49)      (lib/models/smpl.py:96-99)

and is anchored by known-answer tests that hold for ANY asset pack (tests/test_smpl_oracle.py): identity pose + zero
betas -> v_template; a pure global rotation R -> R (v - J_0) + J_0; pose blend shapes vanish for a pure global rotation;
joints of the rest pose = J_regressor @ v_shaped.  The synthetic asset pack (`synthetic_assets`) has the true shapes.
"""
import numpy as np
import torch

PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]   # = last ancestor, ktd.py:10-35
# smplx.vertex_ids 'smplh' entries used by VertexJointSelector(use_hands=True, use_feet_keypoints=True), in its order:
# face (nose, reye, leye, rear, lear), feet (LBigToe, LSmallToe, LHeel, RBigToe, RSmallToe, RHeel), finger tips (l then r:
# thumb, index, middle, ring, pinky).  Quoted from memory of smplx 0.1.13 — part of the unpinned boundary.
EXTRA_VERTEX_IDS = [332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787,
                    2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133]
# lib/models/smpl.py:15-55: JOINT_MAP[name] for name in JOINT_NAMES
JOINT_MAP = [24, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
             8, 5, 45, 46, 4, 7, 21, 19, 17, 16, 18, 20, 47, 48, 49, 50, 51, 52, 53, 24, 26, 25, 28, 27]


def synthetic_assets(seed=0, dtype=torch.float32):
    """Seeded stand-in for SMPL_NEUTRAL.pkl + J_regressor_extra.npy with the true shapes (documented as synthetic)."""
    g = torch.Generator().manual_seed(1234 + seed)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)  # noqa: E731
    nv = 6890
    v_template = r(nv, 3) * torch.tensor([0.25, 0.6, 0.12], dtype=torch.float64)
    shapedirs = 0.02 * r(nv, 3, 10)
    posedirs = 0.005 * r(207, nv * 3)

    def sparse_rows(rows, nnz):
        m = torch.zeros(rows, nv, dtype=torch.float64)
        for i in range(rows):
            idx = torch.randperm(nv, generator=g)[:nnz]
            w = torch.rand(nnz, generator=g, dtype=torch.float64) + 0.1
            m[i, idx] = w / w.sum()
        return m

    J_regressor = sparse_rows(24, 40)
    J_regressor_extra = sparse_rows(9, 20)
    lbs_weights = torch.zeros(nv, 24, dtype=torch.float64)
    idx = torch.stack([torch.randperm(24, generator=g)[:4] for _ in range(nv)])
    w = torch.rand(nv, 4, generator=g, dtype=torch.float64) + 0.05
    lbs_weights.scatter_(1, idx, w / w.sum(1, keepdim=True))
    return {"v_template": v_template.to(dtype), "shapedirs": shapedirs.to(dtype), "posedirs": posedirs.to(dtype),
            "J_regressor": J_regressor.to(dtype), "lbs_weights": lbs_weights.to(dtype),
            "J_regressor_extra": J_regressor_extra.to(dtype), "J_regressor_h36m": sparse_rows(17, 30).to(dtype),
            "parents": torch.tensor(PARENTS, dtype=torch.int32),
            "extra_vertex_ids": torch.tensor(EXTRA_VERTEX_IDS, dtype=torch.int32),
            "joint_map": torch.tensor(JOINT_MAP, dtype=torch.int32)}


def lbs(betas, rotmats, a):
    """smplx.lbs.lbs with pose2rot=False.  betas [B,10], rotmats [B,24,3,3] -> verts [B,6890,3], joints [B,24,3]."""
    B = betas.shape[0]
    dt = betas.dtype
    v_shaped = a["v_template"].to(dt)[None] + torch.einsum("bl,vkl->bvk", betas, a["shapedirs"].to(dt))
    J = torch.einsum("jv,bvk->bjk", a["J_regressor"].to(dt), v_shaped)
    ident = torch.eye(3, dtype=dt)
    pose_feature = (rotmats[:, 1:] - ident).reshape(B, 207)
    v_posed = v_shaped + (pose_feature @ a["posedirs"].to(dt)).reshape(B, -1, 3)
    parents = [int(p) for p in a["parents"]]
    G = []
    for j in range(24):
        rel = J[:, j] - (J[:, parents[j]] if parents[j] >= 0 else 0)
        M = torch.cat([torch.cat([rotmats[:, j], rel[:, :, None]], dim=2),
                       torch.tensor([0, 0, 0, 1], dtype=dt).expand(B, 1, 4)], dim=1)
        G.append(M if parents[j] < 0 else G[parents[j]] @ M)
    G = torch.stack(G, dim=1)                                           # [B,24,4,4]
    posed_joints = G[:, :, :3, 3]
    A = G.clone()
    A[:, :, :3, 3] = G[:, :, :3, 3] - torch.einsum("bjrc,bjc->bjr", G[:, :, :3, :3], J)
    T = torch.einsum("vj,bjrc->bvrc", a["lbs_weights"].to(dt), A)
    verts = torch.einsum("bvrc,bvc->bvr", T[:, :, :3, :3], v_posed) + T[:, :, :3, 3]
    return verts, posed_joints


def smpl_forward(betas, rotmats, a, J_regressor=None):
    """lib/models/smpl.py:94-106 (+ ktd.py:110-112 when J_regressor is given): verts [B,6890,3], joints [B,49 | 17,3]."""
    verts, j24 = lbs(betas, rotmats, a)
    if J_regressor is not None:
        return verts, torch.einsum("jv,bvk->bjk", J_regressor.to(verts.dtype), verts)
    extra = verts[:, a["extra_vertex_ids"].long()]
    reg = torch.einsum("jv,bvk->bjk", a["J_regressor_extra"].to(verts.dtype), verts)
    j54 = torch.cat([j24, extra, reg], dim=1)
    return verts, j54[:, a["joint_map"].long()]
