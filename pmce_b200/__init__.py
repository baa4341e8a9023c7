"""pmce_b200 — B200-native (sm_100a) implementation of the PMCE per-clip forward hot path.

Public surface (mirrors the reference's interfaces for this path):
    pmce_b200.models.{PMCE,PoseEstimation,CoevoDecoder,project_net}.get_model   (reference lib/models)
    pmce_b200.smpl_layer.SMPL_Layer                                              (reference smplpytorch)
    pmce_b200.engine.JRegressor                                                  (reference lib/core/base.py:225)
    pmce_b200.dist.sharded_forward                                               (batch sharding + one all-gather)
All arithmetic runs in libpmce_b200.so (include/pmce_b200.h); there is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
