"""Lazy construction of the native (libabopt_b200) handle behind an nn.Module."""
import torch

from .. import _capi


class NativeOwner:
    """Mixin for modules that can own an abopt_model handle.

    `_native_state()` returns {FullDPM-spelled key: tensor}; the handle is rebuilt whenever any
    of those tensors changes version, storage or device (load_state_dict, .to(), optimiser step).
    """
    _native_scope = _capi.SCOPE_FULL

    def _native_config(self):
        raise NotImplementedError

    def _native_state(self):
        raise NotImplementedError

    def _native_fingerprint(self, state):
        return tuple((k, t.data_ptr(), t._version, str(t.device)) for k, t in state.items())

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
