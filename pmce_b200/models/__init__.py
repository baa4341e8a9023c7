"""Drop-in for the reference's `lib/models` package (reference lib/models/__init__.py:1-4): same
sub-module names, `get_model()` factories, `forward()` signatures and `state_dict` schema; the
arithmetic runs in libpmce_b200 (hand-written sm_100a CUDA)."""
from . import PMCE            # noqa: F401
from . import PoseEstimation  # noqa: F401
from . import CoevoDecoder    # noqa: F401
from . import project_net     # noqa: F401
