"""Lazy construction of the native (libabopt_b200) handle behind an nn.Module."""
import copy
import weakref

import torch

from .. import _capi


def forbid_training_graph(module, what):
    """The native path computes values only.  The reference's train.py (AbDock/train.py:104-113) calls the model in
    training mode with autograd on and then loss.backward(): fail at the call, with a message, instead of later inside
    autograd with a generic one."""
    if module.training and torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters()):
        raise _capi.AboptError(f'{what} runs forward-only on the sm_100a kernels and records no autograd graph: call it under '
                               'torch.no_grad() / in eval() mode, or keep the reference class for training')


class NativeOwner:
    """Mixin for modules that can own an abopt_model handle.

    `_native_state()` returns {FullDPM-spelled key: tensor}; the handle is rebuilt whenever any
    of those tensors changes version, storage or device (load_state_dict, .to(), optimiser step).
    In-place writes through `.data` (`p.data.copy_(...)`, EMA / weight averaging) do not bump the
    version counter: call `invalidate_native()` after them.
    """
    _native_scope = _capi.SCOPE_FULL

    def _native_config(self):
        raise NotImplementedError

    def _native_state(self):
        raise NotImplementedError

    def _native_fingerprint(self, state):
        return tuple((k, t.data_ptr(), t._version, str(t.device)) for k, t in state.items())

    def invalidate_native(self):
        """Drop the packed device weights; the next call repacks them from the current parameters."""
        self.__dict__.pop('_native_cache', None)
        for child in self.children():
            if isinstance(child, NativeOwner):
                child.invalidate_native()

    def native(self):
        state = self._native_state()
        fp = self._native_fingerprint(state)
        cached = self.__dict__.get('_native_cache')
        if cached is not None and cached[0] == fp:
            return cached[1]
        devs = {t.device for t in state.values() if t.numel()}
        if len(devs) != 1:
            raise _capi.AboptError(f'parameters live on several devices: {devs}')
        dev = devs.pop()
        nm = _capi.NativeModel(self._native_config(), dev, state)
        self.__dict__['_native_cache'] = (fp, nm)
        return nm

    # ---- copy / pickle: the ctypes handle and the owner back-reference are per-object state
    def _rebind_children(self):
        """(Re)attach the children that execute on this module's handle (FullDPM -> eps_net, trans_*)."""

    def __getstate__(self):
        state = dict(super().__getstate__())
        state.pop('_native_cache', None)
        state.pop('_abopt_owner', None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._rebind_children()

    def __deepcopy__(self, memo):
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ('_native_cache', '_abopt_owner'):
                continue
            new.__dict__[k] = copy.deepcopy(v, memo)
        new._rebind_children()
        return new


class Owned:
    """Mixin (listed BEFORE nn.Module) for leaves that execute on their owner's handle: the back-reference is a weakref,
    which can neither be pickled nor meaningfully deep-copied; the owner re-attaches it (_rebind_children)."""

    def __getstate__(self):
        state = dict(super().__getstate__())
        state.pop('_abopt_owner', None)
        return state


def bind_owner(owner, children):
    ref = weakref.ref(owner)
    for child in children:
        child.__dict__['_abopt_owner'] = ref
