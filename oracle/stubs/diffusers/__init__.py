"""Minimal stand-in for the `diffusers` names imported by the reference's hot-path modules
(animation/modules/unet.py:8-10,19,25; pose_net.py:4; id_encoder.py).  TEST INFRASTRUCTURE ONLY:
used by oracle/ref_loader.py in the build container, where /root/reference exists, to import the
reference unmodified and generate golden vectors.  Never imported by the product."""
