"""Stand-in for the (unpinned, absent) `timm` dependency of the reference.

TEST INFRASTRUCTURE ONLY. The reference imports `DropPath`, `Mlp`, `Attention`
(lib/models/PoseEstimation.py:9-10, lib/models/CoevoDecoder.py:6-7). timm is not pinned in
requirements.sh and not vendored, so the arithmetic at this boundary is restated from timm's
published ViT layers (0.4-0.6 era) and is "parity unpinned" there; parameter names
(qkv/proj/fc1/fc2) are pinned by the checkpoint schema. The in-repo copy of the same attention
math is lib/models/CoevoDecoder.py:107-131.
"""
