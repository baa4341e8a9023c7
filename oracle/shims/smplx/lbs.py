"""smplx.lbs shim (see __init__.py)."""
import torch


def vertices2joints(J_regressor, vertices):
    return torch.einsum("bik,ji->bjk", [vertices, J_regressor])
