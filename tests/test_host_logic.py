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


def _expand_ref_keys(golden_dir):
    import json
    d = json.loads((golden_dir / "fullmodel_state_keys.json").read_text())
    out = {}
    for k, v in d.items():
        if k.endswith(".*"):
            cnt, shape = v
            for i in range(cnt):
                out[f"{k[:-2]}.{i}"] = tuple(shape)
        else:
            out[k] = tuple(v)
    return out


def test_full_model_key_set_equals_executed_reference(golden_dir):
    """Exact key-set + shape equality with the EXECUTED reference's ``fullModel.state_dict()`` (fixture written by
    oracle/make_golden_keys.py from /root/reference; ``encoder.*`` — the unused timm ViT-B — is dropped on load)."""
    from sais_b200.prepare_model import fullModel
    ref = {k: s for k, s in _expand_ref_keys(golden_dir).items() if not k.startswith("encoder.")}
    own = {k: tuple(v.shape) for k, v in fullModel(data_type='reps', nclasses=2, domain='NH_02', rep_dim=384,
                                                   encoder_type='ViT').state_dict().items()}
    assert set(own) == set(ref), (sorted(set(own) - set(ref))[:5], sorted(set(ref) - set(own))[:5])
    assert own == ref


def test_loadmodel_checkpoint_round_trip(tmp_path, golden_dir):
    """``params.zip`` / ``prototypes.zip`` exactly as train.py:85-87,105-112 writes them — the deep-copied state dict of the
    DDP-wrapped model ('module.' prefix) and a deep-copied nn.ParameterDict — load through the product's loadModel
    helpers: strict key match, values intact, prototypes usable.  When /root/reference is mounted (build container) the
    checkpoint is produced by the reference's own fullModel."""
    import copy
    import torch.nn as nn
    from oracle import ref_import as R
    from sais_b200 import scoring
    from sais_b200.prepare_model import fullModel, load_checkpoint_params, load_prototypes

    if R.available():
        ref_model = R.build_full_model(R.load_prepare_model(), "RGB-Flow", nclasses=2)
        ref_sd = ref_model.state_dict()
    else:  # GPU box / no mount: same key set from the committed fixture, seeded values
        g = torch.Generator().manual_seed(5)
        ref_sd = {k: torch.rand(s, generator=g) for k, s in _expand_ref_keys(golden_dir).items()}
    ref_sd = dict(ref_sd)
    ref_sd["encoder.cls_token"] = torch.zeros(1, 1, 768)  # the real checkpoint also carries the timm ViT-B
    best_params_dict = copy.deepcopy({"module." + k: v for k, v in ref_sd.items()})
    protos = nn.ParameterDict({str(i): nn.Parameter(torch.rand(1, 256)) for i in range(2)})
    torch.save(best_params_dict, tmp_path / "params.zip")
    torch.save(copy.deepcopy(protos), tmp_path / "prototypes.zip")

    params = load_checkpoint_params(str(tmp_path))
    assert set(params) == set(ref_sd)
    m = fullModel(data_type='reps', nclasses=2, domain='NH_02', rep_dim=384, encoder_type='ViT')
    m.load_state_dict(params)  # strict
    back = m.state_dict()
    assert set(back) == {k for k in ref_sd if not k.startswith("encoder.")}
    for k in ("frame_cls", "linear.weight", "frame_pos_embeddings.1999", "clip_pos_embeddings.3",
              "transEncoderFrame.layers.2.self_attn.in_proj_weight", "transEncoderClip.layers.0.linear2.bias",
              "attentionModules.1.weight", "finalModules.2.bias"):
        assert torch.equal(back[k], ref_sd[k]), k
    loaded = load_prototypes(str(tmp_path), "cpu")
    assert isinstance(loaded, nn.ParameterDict) and sorted(loaded.keys()) == ["0", "1"]
    assert torch.equal(scoring.stack_prototypes(loaded), torch.vstack([protos["0"].detach(), protos["1"].detach()]))
    # a missing key is an error, as with the reference's strict load_state_dict (prepare_model.py:529)
    bad = {k: v for k, v in params.items() if k != "linear.bias"}
    with pytest.raises(RuntimeError):
        fullModel(data_type='reps', nclasses=2, domain='NH_02', rep_dim=384, encoder_type='ViT').load_state_dict(bad)
