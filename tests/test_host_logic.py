"""CPU: host-side logic of the drop-in modules — state-dict compatibility with the reference key names, packing
bookkeeping, window/shard arithmetic.  No kernels are launched."""
import numpy as np
import pytest
import torch

from oracle import sais_oracle as O


def test_vit_state_dict_keys_match_reference_names():
    import sais_b200.vision_transformer as vits
    m = vits.vit_small(patch_size=16)
    ref = O.make_vit_weights(0, "init")
    assert set(m.state_dict().keys()) == set(ref.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(ref[k].shape), k
    m.load_state_dict(ref, strict=True)
    assert sum(p.numel() for p in m.parameters()) == 21_665_664


def test_vit_rejects_other_architectures():
    import sais_b200.vision_transformer as vits
    with pytest.raises(NotImplementedError):
        vits.vit_base()
    with pytest.raises(NotImplementedError):
        vits.VisionTransformer(embed_dim=768)
    with pytest.raises(NotImplementedError):
        vits.vit_small(patch_size=8)


def test_full_model_state_dict_round_trip_with_reference_keys():
    from sais_b200.prepare_model import fullModel
    m = fullModel(data_type='reps', nclasses=2, domain='NH_02', rep_dim=384, encoder_type='ViT')
    sd = m.state_dict()
    # reference key names (prepare_model.py:62-101): ParameterDict entries, both encoders, heads
    assert "frame_pos_embeddings.0" in sd and "frame_pos_embeddings.1999" in sd and "clip_pos_embeddings.7" in sd
    assert sd["frame_pos_embeddings.5"].shape == (1, 384)
    assert "frame_pos_table" not in sd
    for k in ("frame_cls", "clip_cls", "linear.weight", "linear2.bias", "attentionA.weight",
              "attentionModules.2.bias", "finalModules.0.weight",
              "transEncoderFrame.layers.3.self_attn.in_proj_weight", "transEncoderFrame.layers.0.linear1.weight",
              "transEncoderClip.layers.2.norm2.bias", "transEncoderFrame.layers.1.self_attn.out_proj.bias"):
        assert k in sd, k
    assert sd["transEncoderFrame.layers.0.self_attn.in_proj_weight"].shape == (1152, 384)
    assert sd["transEncoderFrame.layers.0.linear1.weight"].shape == (2048, 384)
    # a reference checkpoint also carries the unused timm encoder and DDP's 'module.' prefix is stripped by loadModel
    sd2 = {k: v.clone() + 1 for k, v in sd.items()}
    sd2["encoder.cls_token"] = torch.zeros(1, 1, 768)
    m2 = fullModel(data_type='reps', nclasses=2, domain='NH_02', rep_dim=384, encoder_type='ViT')
    m2.load_state_dict(sd2, strict=True)
    assert torch.equal(m2.frame_pos_table[17], sd["frame_pos_embeddings.17"][0] + 1)
    assert torch.equal(m2.linear.weight, sd["linear.weight"] + 1)


def test_full_model_rejects_out_of_scope_configs():
    from sais_b200.prepare_model import fullModel
    with pytest.raises(NotImplementedError):
        fullModel(data_type='raw', encoder_type='R3D', rep_dim=512)
    with pytest.raises(NotImplementedError):
        fullModel(data_type='reps', encoder_type='ViT', rep_dim=384, importance_loss=True)


def test_head_param_count_matches_survey():
    from sais_b200.prepare_model import fullModel
    m = fullModel(data_type='reps', nclasses=2, domain='NH_02', rep_dim=384, encoder_type='ViT')
    n = sum(p.numel() for p in m.parameters())
    assert n == 19_180_681  # whole fullModel without the timm encoder
