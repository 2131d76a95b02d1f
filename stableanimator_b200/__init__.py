"""B200-native (sm_100a) implementation of StableAnimator's denoising hot path.

Public surface (mirrors the reference's operator / module interface for this path):
  * :class:`stableanimator_b200.unet.UNetB200`            -- UNetSpatioTemporalConditionModel.forward
  * :mod:`stableanimator_b200.processors`                 -- AnimationAttnProcessor,
        AnimationIDAttnNormalizedProcessor, XFormersAttnProcessor (attention-processor slots)
  * :class:`stableanimator_b200.pose_net.PoseNetB200`     -- PoseNet.forward
  * :class:`stableanimator_b200.sampler.AnimationSampler` -- the Euler/EDM loop with frame tiling + CFG
All arithmetic runs in ``libsa_b200.so`` (include/sa_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
