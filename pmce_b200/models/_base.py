"""Shared plumbing of the drop-in modules: parameter containers + engine life-cycle."""
import torch
import torch.nn as nn

from ..engine import Engine, make_dims, PmceError


class EngineModule(nn.Module):
    """nn.Module whose parameters live in the reference schema and whose forward runs in libpmce_b200.

    The packed device copy of the weights is rebuilt lazily whenever the parameters may have changed:
    after construction, `load_state_dict`, `.cuda()/.to()` (`_apply`) or an explicit `refresh_weights()`.
    """

    _engine_prefix = ""

    def __init__(self):
        super().__init__()
        object.__setattr__(self, "_engine", None)
        object.__setattr__(self, "_wver", 1)
        object.__setattr__(self, "_packed_ver", 0)

    def _engine_dims(self):
        raise NotImplementedError

    def _engine_vj(self):
        return None

    def refresh_weights(self):
        """Mark the packed device copy stale (call after modifying parameters in place)."""
        object.__setattr__(self, "_wver", self._wver + 1)

    def _weights_version(self):
        return sum(m._wver for m in self.modules() if isinstance(m, EngineModule))

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.refresh_weights()
        return out

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.refresh_weights()
        return out

    def engine(self):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise PmceError("pmce_b200 modules run on CUDA only: call .cuda() first (there is no CPU path)")
        if self._engine is None:
            object.__setattr__(self, "_engine", Engine(self._engine_dims()))
        ver = self._weights_version()
        if ver != self._packed_ver or self._engine.weights is None or self._engine.weights.device != dev:
            named = [(self._engine_prefix + k, v) for k, v in self.state_dict().items()]
            self._engine.pack(named, dev, vj_relation=self._engine_vj())
            object.__setattr__(self, "_packed_ver", ver)
        return self._engine
