"""OVERLAY for the reference tree: copy to mlagg/nnunetv2/training/nnUNetTrainer/variants/mamba/MambaSkip.py.

Same module path and the same public names as the reference file (SS2D_skip :266-543, DWConv :545-556,
ConvolutionalGLU :559-577, VSS_Conv_Block :669-753, VSS_Conv_Layer :756-804), so
`from nnunetv2.training.nnUNetTrainer.variants.mamba.MambaSkip import VSS_Conv_Layer` in the trainer file keeps working
and reference checkpoints load with strict=True.  The Multi-Scale Mamba Module then runs on the sm_100a kernels of
libmlagg_b200.so (fused 4-direction multi-scale selective scan, walk pack / unpack, depthwise stencils, tcgen05
projections); mamba-ssm and causal-conv1d are no longer imported.
"""
from mlagg_unet_b200.mamba_skip import (SS2D_skip, DWConv, ConvolutionalGLU, VSS_Conv_Block,  # noqa: F401
                                        VSS_Conv_Layer)
from mlagg_unet_b200.selective_scan_interface import selective_scan_fn  # noqa: F401  (mamba-ssm signature)

__all__ = ["SS2D_skip", "DWConv", "ConvolutionalGLU", "VSS_Conv_Block", "VSS_Conv_Layer", "selective_scan_fn"]
