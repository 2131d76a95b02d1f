"""GPU parity of the face-token encoder (FusionFaceIdB200) against the oracle restatement of
FusionFaceId (id_encoder.py:102-148) and the reference golden; and its hand-off to the UNet context."""
from pathlib import Path

import pytest
import torch

from util import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("batch", [1, 2])
def test_face_encoder_matches_oracle(batch):
    from oracle import face_oracle
    from stableanimator_b200.face_encoder import FusionFaceIdB200
    from stableanimator_b200.synth import synth_face_inputs, synth_face_state_dict
    sd = synth_face_state_dict()
    ide, clip = synth_face_inputs(batch)
    enc = FusionFaceIdB200(sd, "cuda")
    out = enc(ide.cuda(), clip.cuda())
    ref = face_oracle.fusion_face_id(sd, ide, clip)
    assert out.dtype == torch.float32 and out.shape == (batch, 4, 1024)
    assert_close(out.cpu(), ref, what="face tokens")  # fp32 output: plain max-norm / rel-L2 <= 2e-3


def test_face_encoder_matches_reference_golden():
    from stableanimator_b200.face_encoder import FusionFaceIdB200
    from stableanimator_b200.synth import synth_face_inputs, synth_face_state_dict
    g = torch.load(Path(__file__).resolve().parent / "golden" / "face_fusion.pt", map_location="cpu", weights_only=False)
    enc = FusionFaceIdB200(synth_face_state_dict(g["seed"]), "cuda")
    ide, clip = synth_face_inputs(g["batch"], g["seed"])
    assert_close(enc(ide.cuda(), clip.cuda()).cpu(), g["out"], what="face tokens vs reference")


def test_face_encoder_strict_state_dict_and_dtype():
    from stableanimator_b200.face_encoder import FusionFaceIdB200
    from stableanimator_b200.synth import synth_face_inputs, synth_face_state_dict
    sd = synth_face_state_dict()
    bad = dict(sd)
    bad.pop("norm.weight")
    with pytest.raises(KeyError):
        FusionFaceIdB200(bad, "cuda")
    enc = FusionFaceIdB200(sd, "cuda")
    ide, clip = synth_face_inputs(1)
    out = enc(ide.cuda().bfloat16(), clip.cuda().bfloat16())  # reference runs the module in the pipeline dtype
    assert out.dtype == torch.bfloat16
    with pytest.raises(AssertionError):
        enc(ide.cuda(), clip.cuda(), shortcut=True)
