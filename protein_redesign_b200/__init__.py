"""prd-b200: the ProteinReDiff denoiser hot path on hand-written sm_100a CUDA kernels.

Public surface (mirrors the reference's Python module API):
    protein_redesign_b200.model.ProteinReDiffModel
    protein_redesign_b200.modules.{Denoiser, FoldingBlock, TriangleMultiplication, ...}
    protein_redesign_b200.models.AF2_modules.{SPAttention, OuterProductUpdate}
"""
__version__ = "0.1.0"
