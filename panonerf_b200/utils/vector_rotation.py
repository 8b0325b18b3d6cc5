"""CUDA replacement for the batched rotation helper of utils/vector_rotation.py."""
from .. import ops


class RotToTarget:
    """utils/vector_rotation.py:50-89.  `rot2t(tvec [B,3]) -> [B,3,3]`: the Rodrigues rotation that takes the +y axis
    onto each (unit) target vector - a hemisphere of light directions defined around +y is carried onto a surface
    normal by `rot @ dirs`.  The antipodal target (0,-1,0) maps to the fixed matrix diag(1,-1,1) as upstream.
    Differentiable w.r.t. `tvec`."""

    def rot2t(self, tvec):
        t = tvec.reshape(-1, 3)
        t = t if t.requires_grad else ops._f32c(t)
        return ops.rot_to_target(t.contiguous())
