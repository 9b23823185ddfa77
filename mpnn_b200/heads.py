"""Prediction head + loss of the reference drivers as one launch each way (csrc/head.cu).

    test_graph_norm.py:86-90:  nn.Sequential(GraphWrapper(BasicModel(...)), nn.BatchNorm1d(out), nn.Linear(out, targets))
                               criterion = nn.MSELoss()

`BNLinearMSE(bn, linear)` wraps the driver's OWN stock modules (same parameters, buffers and state_dict entries) and
returns `mse_loss(linear(bn(x)), target)`; the predictions of the last call are kept in `.prediction`.  Problems that
do not fit one CTA's shared memory (`mpnn_head_supported`) raise: the caller keeps its stock modules for those.
"""
import torch
from torch import nn
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import check, f32c, ptr, stream


class _BNLinearMSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target, gamma, beta, W, b, running_mean, running_var, nbt, training, momentum, eps):
        lib = _lib.load()
        if not x.is_cuda:
            raise RuntimeError("mpnn_b200.heads: CUDA tensors required (there is no CPU fallback)")
        x, target, W = f32c(x), f32c(target), f32c(W)
        B, C = x.shape
        T = W.shape[0]
        dev = x.device
        y = torch.empty(B, T, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        stats = torch.empty(2 * C, dtype=torch.float32, device=dev)
        check(lib.mpnn_head_bn_linear_mse_fwd(ptr(x), ptr(target), ptr(gamma), ptr(beta), ptr(running_mean),
                                              ptr(running_var), ptr(nbt), ptr(W), ptr(b), B, C, T, int(training),
                                              float(momentum), float(eps), ptr(y), ptr(loss), ptr(stats), stream()),
              "head_bn_linear_mse_fwd")
        ctx.save_for_backward(x, target, gamma, beta, W, y, stats)
        ctx.dims = (B, C, T, int(training), b is not None)
        ctx.mark_non_differentiable(y)
        ctx.set_materialize_grads(False)     # no zero-filled gradient for the non-differentiable predictions
        return loss, y

    @staticmethod
    @once_differentiable
    def backward(ctx, gloss, _gy):
        lib = _lib.load()
        x, target, gamma, beta, W, y, stats = ctx.saved_tensors
        B, C, T, training, has_b = ctx.dims
        dev = x.device
        if gloss is None:
            return (None,) * 12
        gloss = f32c(gloss.reshape(1))
        dx = torch.empty_like(x)
        dgamma = torch.empty_like(gamma) if gamma is not None else None
        dbeta = torch.empty_like(beta) if beta is not None else None
        dW = torch.empty_like(W)
        db = torch.empty(T, dtype=torch.float32, device=dev)
        check(lib.mpnn_head_bn_linear_mse_bwd(ptr(x), ptr(target), ptr(gamma), ptr(beta), ptr(W), ptr(y), ptr(stats),
                                              ptr(gloss), B, C, T, training, ptr(dx), ptr(dgamma), ptr(dbeta), ptr(dW),
                                              ptr(db), stream()), "head_bn_linear_mse_bwd")
        return dx, None, dgamma, dbeta, dW, (db if has_b else None), None, None, None, None, None, None


class BNLinearMSE(nn.Module):
    """loss = mse_loss(linear(bn(x)), target) for x [B, C], target [B, T]  (mean over all B*T elements)."""

    def __init__(self, bn, linear):
        super(BNLinearMSE, self).__init__()
        if not isinstance(bn, nn.BatchNorm1d) or not isinstance(linear, nn.Linear):
            raise TypeError("BNLinearMSE wraps an nn.BatchNorm1d and an nn.Linear")
        self.bn, self.linear = bn, linear
        self.prediction = None

    def forward(self, x, target):
        bn, lin = self.bn, self.linear
        B, C = x.shape
        T = lin.out_features
        training = bn.training or not bn.track_running_stats
        lib = _lib.load()
        if (not lib.mpnn_head_supported(B, C, T) or bn.momentum is None or target.shape != (B, T)
                or (training and B < 2)):
            # no silent dispatch to another implementation: the caller keeps its stock modules for such a head
            raise RuntimeError("mpnn_b200.heads.BNLinearMSE: B=%d, C=%d, T=%d%s is not served by the fused head kernel "
                               "(mpnn_head_supported); use the stock BatchNorm1d / Linear / mse_loss modules for it"
                               % (B, C, T, ", momentum=None" if bn.momentum is None else ""))
        tracked = bn.track_running_stats
        loss, y = _BNLinearMSEFn.apply(x, target, bn.weight if bn.affine else None, bn.bias if bn.affine else None,
                                       lin.weight, lin.bias, bn.running_mean if tracked else None,
                                       bn.running_var if tracked else None,
                                       bn.num_batches_tracked if (tracked and training) else None, training,
                                       bn.momentum, bn.eps)
        self.prediction = y
        return loss
