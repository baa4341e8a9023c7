"""smplx.body_models shim (see __init__.py)."""
import collections

ModelOutput = collections.namedtuple("ModelOutput", ["vertices", "joints", "full_pose", "betas", "global_orient", "body_pose"])
ModelOutput.__new__.__defaults__ = (None,) * 6
