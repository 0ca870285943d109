"""Shared machinery of the drop-in modules: eval-only guard, plan cache, zero-copy chaining."""
import collections

import torch
import torch.nn as nn

from .. import _lib


def require_device(t):
    """The product path is CUDA-only; tests of the host logic replace this hook (tests/conftest.py)."""
    if not t.is_cuda:
        raise _lib.SvxError("swinvox_b200 modules run on CUDA tensors only (sm_100a); there is no CPU fallback. "
                            f"Got a tensor on {t.device}.")


def mark_owned(t, act_or_buf):
    """tag a tensor returned by one of our modules so the next module can bind to its memory directly"""
    t._svx_src = act_or_buf
    return t


def owned_src(t):
    return getattr(t, "_svx_src", None)


class PlannedModule(nn.Module):
    """nn.Module whose forward replays a recorded libswinvox_b200 plan (one plan per input signature).
    Parameters live in ordinary nn layers (the reference's state_dict layout); `forward` never calls them."""

    use_graph = False
    # plans kept per module (least recently used are dropped): each holds the activations of one input signature
    # (about 6.5 GB for the encoder at 64 x 3 views: 34 MB per image), so a long-lived process that varies B or V must not keep them all
    max_plans = 4

    def __init__(self):
        super().__init__()
        self._plans = collections.OrderedDict()
        self._param_version = None

    def _guard(self, *tensors):
        if self.training:
            raise RuntimeError(f"{type(self).__name__} (swinvox_b200) is inference-only: call .eval() first "
                               "(BatchNorm uses running statistics, Dropout is the identity; no backward exists).")
        if torch.is_grad_enabled():
            raise RuntimeError(f"{type(self).__name__} (swinvox_b200) has no backward: wrap the call in torch.no_grad().")
        for t in tensors:
            require_device(t)
            if t.dtype != torch.float32:
                raise TypeError(f"expected float32 input, got {t.dtype}")

    def _versions(self):
        return tuple(p._version for p in self.parameters()) + tuple(b._version for b in self.buffers())

    def _plan_for(self, key, builder):
        """plans bake prepared copies of the weights: rebuild when any parameter/buffer was modified in place
        (load_state_dict, .to(), manual edits) or moved"""
        ver = self._versions()
        if ver != self._param_version:
            self._plans.clear()
            self._param_version = ver
        if key not in self._plans:
            self._plans[key] = builder()
            while len(self._plans) > self.max_plans:
                self._plans.popitem(last=False)
        else:
            self._plans.move_to_end(key)
        return self._plans[key]

    def invalidate(self):
        self._plans.clear()

    def _apply(self, fn, *a, **k):
        self._plans = collections.OrderedDict()
        return super()._apply(fn, *a, **k)


def src_key(t):
    """plan-cache key component: the address of our own upstream buffer, or 0 for a foreign tensor.  A plan is
    bound either to that upstream buffer (zero copy) or to its own staging memory, never both, so feeding a
    foreign tensor never overwrites another module's output."""
    src = owned_src(t)
    return src.data_ptr() if isinstance(src, torch.Tensor) else 0


class PlanarInput:
    """A contiguous fp32 plan input: bound to one of our own upstream outputs (zero copy) or a private buffer."""

    def __init__(self, plan, x, shape):
        src = owned_src(x)
        self.bound = isinstance(src, torch.Tensor) and src.numel() == x.numel() and src.is_contiguous()
        self.buf = plan.hold(src).view(*shape) if self.bound else plan.empty(*shape)

    def feed(self, x):
        if not self.bound:
            self.buf.copy_(x.reshape(self.buf.shape))


class ChannelsLastInput:
    """A channels-last plan input [N*P, Cs]: bound to one of our own channels-last outputs (zero copy), or a
    private buffer filled from a planar tensor [N, C, P] by the library's transpose kernel."""

    def __init__(self, plan, x, N, C, P, Cs, round_in):
        self.N, self.C, self.P, self.Cs, self.round_in = N, C, P, Cs, round_in
        self.plan = plan
        src = owned_src(x)
        self.bound = isinstance(src, torch.Tensor) and tuple(src.shape) == (N * P, Cs)
        self.buf = plan.hold(src) if self.bound else plan.zeros(N * P, Cs)
        self.staging = None if self.bound else plan.empty(N, C, P)

    def feed(self, x):
        if self.bound:
            return
        from .. import engine as E
        self.staging.copy_(x.reshape(self.N, self.C, self.P))
        E.transpose_now(self.plan, self.staging, self.buf, self.N, self.C, self.P, self.Cs, True, self.round_in)
