"""The optimisation iteration around the renderer boundary — host-array form and device-resident form.

One iteration of the reference's surface optimisation (exp_bunny/test.py:161-216) is

    transient, grad = inverseRendering(mesh, gt_transient, weight, opt)          # the hot path (this library)
    smoothing_val, smoothing_grad = renderStreamedNormalSmoothing(mesh)           # O(F) regulariser (this library)
    l2, original_l2 = evaluate_loss_with_normal_smoothness(...)                   # exp_bunny/rendering.py:360-367
    grad += smooth_weight * smoothing_grad
    Adam_Modified.step()                                                          # exp_bunny/adam_modified.py:60-107

`HostIteration` runs exactly that call sequence on NumPy arrays through the reference-signature facade (`rendering`), i.e. what
an unmodified driver does.  `DeviceIteration` keeps vertices, target, weight, transient, gradient, regulariser gradient, loss and
the Adam state in HBM as torch CUDA tensors that the C ABI uses in place: an iteration moves no [L,B] array over PCIe, runs no
NumPy code, and reads back 8 bytes (the loss) only when asked.  Both produce the same vertex trajectory (tests/test_gpu_loop.py).

Not here (out of scope): El Topo / CGAL remeshing between iterations — `rendering.removeTriangle` (the
renderer-boundary part of the remeshing step, exp_bunny/rendering.py:271-278) is provided, the remeshers themselves stay on the host
in the reference.
"""
import numpy as np

from . import rendering, renderer, scenes

__all__ = ['AdamModified', 'HostIteration', 'DeviceIteration', 'RenderOptions', 'Mesh']


class RenderOptions(object):
    """The reference's ad-hoc `opt` object (exp_bunny/test.py:16-46) with the attributes the renderer facade reads."""
    max_distance_bin = 1200
    distance_resolution = 1.2e-3
    normal = 'fn'
    smooth_weight = 0.0001
    gamma = 0
    bin_refine_resolution = 10
    sigma_bin = 1
    testing_flag = 1
    loss_flag = 0
    alpha_flag = False
    albedo_flag = False
    jitter = False

    def __init__(self, sample_num, lighting, lighting_normal, **kw):
        self.sample_num = int(sample_num)
        self.lighting = np.ascontiguousarray(lighting, dtype=np.float32)
        self.lighting_normal = np.ascontiguousarray(lighting_normal, dtype=np.float32)
        for k, v in kw.items():
            setattr(self, k, v)


class Mesh(object):
    """The reference's ad-hoc `mesh` object: v [V,3] f32, f [F,3] i32, f_affinity [F,3] i32 (exp_bunny/rendering.py:88-101)."""

    def __init__(self, v, f):
        self.v = np.ascontiguousarray(v, dtype=np.float32)
        self.f = np.ascontiguousarray(f, dtype=np.int32)
        self.f_affinity = scenes.face_affinity(self.f)


class AdamModified(object):
    """exp_bunny/adam_modified.py:60-107: Adam (betas 0.9 / 0.999, eps 1e-8, no weight decay, no amsgrad) whose denominator
    sqrt(v)+eps is AVERAGED over the xyz components of a vertex (`torch.mean(denom, dim=1, keepdim=True)`, :99), so a vertex moves
    along its gradient direction.  Works on NumPy float32 arrays or, in place, on torch tensors."""

    def __init__(self, lr, betas=(0.9, 0.999), eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = float(lr), float(betas[0]), float(betas[1]), float(eps)
        self.t = 0; self.m = None; self.v = None

    def _step_size(self):
        return self.lr * (1.0 - self.b2 ** self.t) ** 0.5 / (1.0 - self.b1 ** self.t)      # :102-104

    def step(self, p, g):
        """p <- p - step * m / mean_xyz(sqrt(v) + eps).  NumPy: returns the new float32 array; torch: updates p in place."""
        self.t += 1
        if isinstance(p, np.ndarray):
            g = np.asarray(g, dtype=np.float32)
            if self.m is None:
                self.m = np.zeros_like(p, dtype=np.float32); self.v = np.zeros_like(p, dtype=np.float32)
            self.m = (np.float32(self.b1) * self.m + np.float32(1 - self.b1) * g).astype(np.float32)                    # :89
            self.v = (np.float32(self.b2) * self.v + np.float32(1 - self.b2) * g * g).astype(np.float32)                # :90
            denom = (np.sqrt(self.v) + np.float32(self.eps)).mean(axis=1, keepdims=True, dtype=np.float32)           # :97-99
            return (p - np.float32(self._step_size()) * self.m / denom).astype(np.float32)                              # :105
        import torch
        g = g.to(torch.float32)
        if self.m is None:
            self.m = torch.zeros_like(p); self.v = torch.zeros_like(p)
        self.m.mul_(self.b1).add_(g, alpha=1 - self.b1)
        self.v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
        denom = (self.v.sqrt() + self.eps).mean(dim=1, keepdim=True)
        p.addcdiv_(self.m, denom, value=-self._step_size())
        return p


class HostIteration(object):
    """The reference's call sequence on host arrays (what an unmodified exp_*/test.py does each iteration)."""

    def __init__(self, mesh, gt_transient, weight, opt, lr, ctx=None):
        self.mesh, self.gt, self.weight, self.opt = mesh, gt_transient, weight, opt
        self.adam = AdamModified(lr)
        self.ctx = ctx

    def step(self):
        """-> (l2_with_regulariser, l2).  mesh.v is replaced by the updated float32 array."""
        mesh, opt = self.mesh, self.opt
        transient, grad, _ = rendering.inverseRendering(mesh, self.gt, self.weight, opt, ctx=self.ctx)
        smoothing_val, smoothing_grad = rendering.renderStreamedNormalSmoothing(mesh, ctx=self.ctx)
        loss, l2 = rendering.evaluate_loss_with_normal_smoothness(self.gt, self.weight, transient, smoothing_val, mesh, opt)
        grad = grad + opt.smooth_weight * smoothing_grad                                  # exp_bunny/test.py:181
        mesh.v = np.ascontiguousarray(self.adam.step(mesh.v, grad.astype(np.float32)))  # :211-215 (.float())
        self.transient = transient
        return float(loss), float(l2)


class DeviceIteration(object):
    """The same iteration with every array resident in HBM (torch CUDA tensors used in place by the C ABI).

    step(read_loss=True) returns (loss, l2) as floats (one 16-byte D2H); with read_loss=False it returns the two 0-d device tensors and
    the iteration has no host synchronisation at all."""

    def __init__(self, mesh, gt_transient, weight, opt, lr, ctx=None, device=0):
        import torch
        from . import _ffi
        self.torch = torch
        self.ctx = ctx or _ffi.default_context(device)
        self.dev = torch.device('cuda', self.ctx.device)
        self.opt = opt
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        self.o, self.n = to(opt.lighting), to(opt.lighting_normal)
        self.v, self.f, self.aff = to(mesh.v), to(mesh.f), to(mesh.f_affinity)
        self.gt = gt_transient if torch.is_tensor(gt_transient) else to(np.asarray(gt_transient, dtype=np.float64))
        self.weight = weight if torch.is_tensor(weight) else to(np.asarray(weight, dtype=np.float64))
        L, B = self.gt.shape
        self.L = L
        self.T = torch.zeros((L, B), dtype=torch.float64, device=self.dev)
        self.pl = torch.zeros(B, dtype=torch.float64, device=self.dev)
        self.G = torch.zeros((mesh.v.shape[0], 3), dtype=torch.float64, device=self.dev)
        self.S = torch.zeros((mesh.v.shape[0], 3), dtype=torch.float64, device=self.dev)
        self.sval = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.adam = AdamModified(lr)
        self.lo, self.hi, self.res = 0.0, opt.max_distance_bin * opt.distance_resolution, opt.distance_resolution

    def step(self, read_loss=True):
        torch, opt = self.torch, self.opt
        self.G.zero_()                                                                    # the reference allocates np.zeros per call; the ABI accumulates ('+=')
        renderer.renderStreamedGradient(self.o, self.n, self.v, self.f, opt.sample_num, self.lo, self.hi, self.res, self.T, self.pl, self.G,
                                        self.gt, self.weight, opt.bin_refine_resolution, opt.sigma_bin, opt.testing_flag, getattr(opt, 'loss_flag', 0), ctx=self.ctx)
        sval = renderer.renderStreamedNormalSmoothing(self.v, self.f, self.aff, self.S, ctx=self.ctx, value_out=self.sval)
        d = self.T - self.gt
        l2 = (d * d * self.weight).sum() / self.L                                         # ||diff * sqrt(w)||^2 / L   (rendering.py:360-364)
        loss = l2 + opt.smooth_weight * (sval if torch.is_tensor(sval) else float(sval))
        g = self.G + opt.smooth_weight * self.S
        self.adam.step(self.v, g)
        if read_loss:
            both = torch.stack([loss.reshape(()), l2.reshape(())]).cpu()
            return float(both[0]), float(both[1])
        return loss, l2

    def vertices(self):
        return self.v.cpu().numpy()
