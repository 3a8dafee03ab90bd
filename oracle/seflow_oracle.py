"""CPU restatement of the chamfer nearest-neighbour op and of seflowLoss.  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

* ``chamfer_forward`` -- NmDistanceKernel (OpenSceneFlow/assets/cuda/chamfer3D/chamfer3D.cu:33-74): squared distance
  d = (x1-x0)^2 + (y1-y0)^2 + (z1-z0)^2 in fp32 as nvcc contracts it (fma(dz,dz, fma(dx,dx, dy*dy)) -- read off the SASS of
  that expression -- emulated through float64), strict `<` scan = lowest index among equal distances, (1e20, -1) for an empty target.  Pinned on the GPU against
  the reference's own compiled extension (oracle/_ref/chamfer3D_ref.so, tests/test_gpu_chamfer.py).
* ``ChamferStandIn`` -- the object ``MyCUDAChamferDis`` of OSF/src/lossfuncs.py:14-16, differentiable on the CPU
  (gradient = NmDistanceGradKernel, chamfer3D.cu:92-114).
* ``seflow_loss`` -- seflowLoss (OSF/src/lossfuncs.py:22-100).  Pinned against the reference's own function (AST-extracted,
  run with ChamferStandIn) in tests/test_oracle_golden.py."""
import numpy as np
import torch

TRUNCATED_DIST = 4


def _nn(q, t, chunk=2048):
    q32, t32 = q.astype(np.float32), t.astype(np.float32)
    nq, nt = q32.shape[0], t32.shape[0]
    dist = np.full(nq, np.float32(1e20), np.float32)
    idx = np.full(nq, -1, np.int32)
    if nt == 0:
        return dist, idx
    for a in range(0, nq, chunk):
        qq = q32[a:a + chunk]
        dx = (t32[None, :, 0] - qq[:, None, 0]).astype(np.float32)
        dy = (t32[None, :, 1] - qq[:, None, 1]).astype(np.float32)
        dz = (t32[None, :, 2] - qq[:, None, 2]).astype(np.float32)
        d = (dy * dy).astype(np.float32)
        d = (dx.astype(np.float64) * dx.astype(np.float64) + d.astype(np.float64)).astype(np.float32)
        d = (dz.astype(np.float64) * dz.astype(np.float64) + d.astype(np.float64)).astype(np.float32)
        i = np.argmin(d, axis=1)                       # first minimum = lowest index
        dist[a:a + chunk] = d[np.arange(d.shape[0]), i]
        idx[a:a + chunk] = i
    return dist, idx


def chamfer_forward(pc0, pc1):
    """-> dist0 f32[N], dist1 f32[M], idx0 i32[N], idx1 i32[M] (numpy)."""
    d0, i0 = _nn(pc0, pc1)
    d1, i1 = _nn(pc1, pc0)
    return d0, d1, i0, i1


def chamfer_backward(pc0, pc1, idx0, idx1, g0, g1):
    gp0 = np.zeros(pc0.shape, np.float64)
    gp1 = np.zeros(pc1.shape, np.float64)
    if pc0.shape[0] and pc1.shape[0]:
        a = (2.0 * g0.astype(np.float64))[:, None] * (pc0.astype(np.float64) - pc1.astype(np.float64)[idx0])
        gp0 += a
        np.add.at(gp1, idx0, -a)
        b = (2.0 * g1.astype(np.float64))[:, None] * (pc1.astype(np.float64) - pc0.astype(np.float64)[idx1])
        gp1 += b
        np.add.at(gp0, idx1, -b)
    return gp0.astype(np.float32), gp1.astype(np.float32)


class _ChamferFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pc0, pc1):
        d0, d1, i0, i1 = chamfer_forward(pc0.detach().numpy(), pc1.detach().numpy())
        ctx.save_for_backward(pc0, pc1)
        ctx.idx = (i0, i1)
        out = (torch.from_numpy(d0), torch.from_numpy(d1), torch.from_numpy(i0), torch.from_numpy(i1))
        ctx.mark_non_differentiable(out[2], out[3])
        return out

    @staticmethod
    def backward(ctx, g0, g1, _a, _b):
        pc0, pc1 = ctx.saved_tensors
        g0 = g0 if g0 is not None else torch.zeros(pc0.shape[0])
        g1 = g1 if g1 is not None else torch.zeros(pc1.shape[0])
        a, b = chamfer_backward(pc0.detach().numpy(), pc1.detach().numpy(), ctx.idx[0], ctx.idx[1], g0.numpy(), g1.numpy())
        return torch.from_numpy(a), torch.from_numpy(b)


class ChamferStandIn:
    """nnChamferDis (assets/cuda/chamfer3D/__init__.py:54-92) on the CPU: what lossfuncs.py binds as MyCUDAChamferDis."""

    def __call__(self, input0, input1, truncate_dist=-1):
        dist0, dist1, _, _ = _ChamferFn.apply(input0.contiguous(), input1.contiguous())
        if truncate_dist <= 0:
            return torch.mean(dist0) + torch.mean(dist1)
        return torch.nanmean(dist0[dist0 <= truncate_dist]) + torch.nanmean(dist1[dist1 <= truncate_dist])

    def disid_res(self, input0, input1):
        return _ChamferFn.apply(input0.contiguous(), input1.contiguous())

    def dis_res(self, input0, input1):
        return _ChamferFn.apply(input0.contiguous(), input1.contiguous())[:2]


def seflow_loss(res_dict):
    """seflowLoss (OSF/src/lossfuncs.py:22-100)."""
    cham = ChamferStandIn()
    l0, l1 = res_dict["pc0_labels"], res_dict["pc1_labels"]
    pc0, pc1, est = res_dict["pc0"], res_dict["pc1"], res_dict["est_flow"]
    pseudo = pc0 + est                                                             # :32
    have_dyn = (int((l0 > 0).sum()) > 256) and (int((l1 > 0).sum()) > 256)         # :35-38
    e0, e1, _, _ = cham.disid_res(pseudo, pc1)                                     # :43
    r0, r1, ri0, _ = cham.disid_res(pc0, pc1)                                      # :44
    out = {"chamfer_dis": e0[e0 <= TRUNCATED_DIST].mean() + e1[e1 <= TRUNCATED_DIST].mean()}   # :45
    dyn = torch.tensor(0.0)
    if have_dyn:
        dyn = dyn + cham(pseudo[l0 > 0], pc1[l1 > 0], truncate_dist=TRUNCATED_DIST)            # :51-52
    out["dynamic_chamfer_dis"] = dyn
    static, norms = torch.tensor(0.0), []
    ri0 = ri0.long()
    for label in torch.unique(l0):                                                 # :63
        mask = l0 == label
        if label == 0:
            static = static + torch.linalg.vector_norm(est[mask], dim=-1).mean()  # :66-67
        elif label > 0 and have_dyn:
            nnd = r0[mask]
            order = torch.argsort(nnd, descending=True)                            # :75
            near = l1[ri0[mask][order]]                                            # :76
            nz = torch.nonzero(near > 0)
            if nz.shape[0] <= 0:
                continue
            mi = order[nz.squeeze(1)[0]]                                           # :80
            max_flow = pc1[ri0[mask][mi]] - pc0[mask][mi]                          # :83
            norms.append(torch.linalg.vector_norm(est[mask] - max_flow, dim=-1))   # :86
    moved = torch.tensor(0.0)
    if norms:
        moved = torch.cat(norms).mean()                                            # :88-89
    elif have_dyn:
        moved = r0[r0 <= TRUNCATED_DIST].mean() + r1[r1 <= TRUNCATED_DIST].mean()  # :90-91
    out["static_flow_loss"], out["cluster_based_pc0pc1"] = static, moved
    return out


def make_scene(n0, n1, seed, n_clusters=6):
    """Two frames with a static background (label 0) and a few moving clusters (labels 1..K), estimated flow near the truth."""
    g = torch.Generator().manual_seed(seed)
    pc0 = torch.cat([torch.rand(n0, 2, generator=g) * 40 - 20, torch.rand(n0, 1, generator=g) * 2], 1)
    l0 = torch.zeros(n0, dtype=torch.int16)
    flow = torch.zeros(n0, 3)
    per = max(n0 // (3 * n_clusters), 1)
    for k in range(n_clusters):
        sel = slice(k * per, (k + 1) * per)
        c = torch.rand(3, generator=g) * torch.tensor([30.0, 30.0, 1.0]) - torch.tensor([15.0, 15.0, 0.0])
        pc0[sel] = c + 0.5 * torch.randn(per, 3, generator=g)
        l0[sel] = k + 1
        flow[sel] = torch.randn(3, generator=g) * torch.tensor([0.8, 0.8, 0.0])
    keep = torch.randperm(n0, generator=g)[:min(n1, n0)]
    pc1 = (pc0 + flow)[keep] + 0.02 * torch.randn(len(keep), 3, generator=g)
    l1 = l0[keep].clone()
    if n1 > len(keep):
        extra = n1 - len(keep)
        pc1 = torch.cat([pc1, torch.cat([torch.rand(extra, 2, generator=g) * 40 - 20, torch.rand(extra, 1, generator=g) * 2], 1)])
        l1 = torch.cat([l1, torch.zeros(extra, dtype=torch.int16)])
    est = flow + 0.05 * torch.randn(n0, 3, generator=g)
    return {"pc0": pc0.contiguous(), "pc1": pc1.contiguous(), "pc0_labels": l0, "pc1_labels": l1, "est_flow": est}
