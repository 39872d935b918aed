"""SAIS temporal head — drop-in for ``SAIS/scripts/prepare_model.py`` on the ``main.sh:27`` inference path
(``-t Prototypes -mod RGB-Flow -dim 384 -dt reps -sa``).

``fullModel`` keeps the reference constructor signature (:20), the ``state_dict`` key names (``frame_cls``,
``clip_cls``, ``frame_pos_embeddings.{0..1999}``, ``clip_pos_embeddings.*``, ``linear``, ``linear2``,
``transEncoderFrame.layers.*``, ``transEncoderClip.*``, ``attentionA/B``, ``attentionModules.*``,
``finalModules.*``), the forward signature ``model(x, f, xlens, flens, task, xpad, fpad, domains)`` (:246) and
the ``(output, attn)`` return convention (:444-448), including the list-of-3 TTA form.  All arithmetic runs in
``libsais_b200.so``: the RGB and flow streams of all TTA views are PACKED into one variable-length batch (they
share ``transEncoderFrame``, :350-351) and pushed through ``sais_temporal_forward`` once; the clip head and
prototype scoring are fused fp32 kernels.

Differences from the reference, all deliberate: inputs are NOT mutated in place (:192 does ``x += pos``); the
unused timm ViT-B ``encoder`` (:40) is not instantiated (``encoder.*`` keys in a checkpoint are ignored); only
``data_type='reps'``, ``encoder_type='ViT'``, ``self_attention=True`` are implemented, with ``task='Prototypes'`` (the
``main.sh`` path) and ``task='MIL'`` (:359-363, 451-488).  ``task='ClassificationHead'`` needs ``self.cls_head``, which the
reference creates only for ``data_type='raw'`` (:52-53) — unreachable with representations as input, so not provided.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib, ops
from .transformer import TransformerEncoder

N_POS = 2000  # prepare_model.py:67


class fullModel(nn.Module):  # noqa: N801 — reference class name
    def __init__(self, data_type='reps', nclasses=2, domain='NH_02', rep_dim=384, encoder_type='ViT',
                 modalities='RGB-Flow', encoder_depth=18, load_pretrained_params=True, freeze_encoder_params=True,
                 self_attention=True, importance_loss=False):
        super().__init__()
        if data_type != 'reps' or encoder_type != 'ViT' or rep_dim != 384 or not self_attention:
            raise NotImplementedError(
                "sais_b200.fullModel covers the SAIS inference hot path only: data_type='reps', "
                "encoder_type='ViT', rep_dim=384, self_attention=True")
        if importance_loss or '+' in domain:
            raise NotImplementedError("importance_loss / multi-task linearB are not on the main.sh inference path")
        if modalities not in ('RGB', 'Flow', 'RGB-Flow'):
            raise ValueError(f"unknown modalities {modalities!r}")
        self.data_type, self.encoder_type = data_type, encoder_type
        self.nclasses, self.domain, self.rep_dim = nclasses, domain, rep_dim
        self.modalities, self.self_attention, self.importance_loss = modalities, self_attention, importance_loss

        self.linear = nn.Linear(rep_dim, 256)
        self.linear2 = nn.Linear(256, 3)
        # torch.rand init as in the reference (:62-71); the 2000 [1,384] ParameterDict entries are held as one table
        self.frame_cls = nn.Parameter(torch.rand(1, rep_dim))
        self.clip_cls = nn.Parameter(torch.rand(1, rep_dim))
        self.frame_pos_table = nn.Parameter(torch.rand(N_POS, rep_dim))
        self.clip_pos_table = nn.Parameter(torch.rand(N_POS, rep_dim))
        self.transEncoderFrame = TransformerEncoder()
        self.transEncoderClip = TransformerEncoder()
        self.attentionA = nn.Linear(rep_dim, 256)
        self.attentionB = nn.Linear(rep_dim, 256)
        self.attentionModules = nn.ModuleDict({str(c): nn.Linear(256, 1) for c in range(3)})
        self.finalModules = nn.ModuleDict({str(c): nn.Linear(rep_dim, 1) for c in range(3)})
        self._register_state_dict_hook(_split_pos_tables)
        self._register_load_state_dict_pre_hook(_merge_pos_tables)
        self.eval()

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x, f, xlens, flens, task, xpad, fpad, domains):
        """x, f: fp32 [B,nsnip,T,384] (or lists of 3 TTA views); xpad, fpad: bool [B,nsnip,T+1], True = padded
        key, index 0 (CLS) never padded.  Returns ``(snip_sequence [B,256] or list, snip_attn [B*nsnip,T+1,T+1])``
        where ``snip_attn`` belongs to view 0 of the RGB stream (flow stream for modalities='Flow')."""
        if task == 'MIL':
            return self._forward_mil(x, xpad)
        if task != 'Prototypes':
            raise NotImplementedError("task must be 'Prototypes' or 'MIL' (ClassificationHead needs data_type='raw')")
        if self.training:
            raise _lib.SaisError("sais_b200.fullModel is inference-only; call .eval()")
        is_list = isinstance(x, list) if self.modalities != 'Flow' else isinstance(f, list)
        use_rgb = self.modalities in ('RGB', 'RGB-Flow')
        use_flow = self.modalities in ('Flow', 'RGB-Flow')
        xs = (x if is_list else [x]) if use_rgb else []
        fs = (f if is_list else [f]) if use_flow else []
        xps = (xpad if is_list else [xpad]) if use_rgb else []
        fps = (fpad if is_list else [fpad]) if use_flow else []
        nviews = len(xs) if use_rgb else len(fs)

        # pack every (view, modality) stream into one variable-length batch
        streams = [(t, p) for t, p in zip(xs, xps)] + [(t, p) for t, p in zip(fs, fps)]
        frames, pads, seq_lens, emit, spans = [], [], [], [], []
        cursor = 0
        for si, (t, p) in enumerate(streams):
            _lib.require_cuda(t, "x/f")
            B, nsnip, T, E = t.shape
            n = B * nsnip
            frames.append(t.reshape(n * T, E).float())
            if p is None:
                pads.append(torch.zeros(n * (T + 1), dtype=torch.uint8, device=t.device))
            else:
                p = p.reshape(n * (T + 1)).to(t.device)
                # bool storage is one 0/1 byte per element: reinterpret instead of launching a conversion kernel
                pads.append(p.view(torch.uint8) if p.dtype == torch.bool else p.to(torch.uint8))
            seq_lens += [T + 1] * n
            emit += [si == 0] * n
            spans.append((cursor, n, B, nsnip, T + 1))
            cursor += n
        x_frames = frames[0] if len(frames) == 1 else torch.cat(frames, 0)
        key_pad = pads[0] if len(pads) == 1 else torch.cat(pads, 0)
        cls, _, attn, _ = self.transEncoderFrame.run_packed(x_frames, seq_lens, key_pad, emit, self.frame_cls,
                                                            self.frame_pos_table)
        _, n0, _, _, S0 = spans[0]
        snip_attn = attn.view(n0, S0, S0)

        lin_w = self.linear.weight.detach().float().contiguous()
        lin_b = self.linear.bias.detach().float().contiguous()
        outs = []
        for v in range(nviews):
            a0, an, B, nsnip, _ = spans[v]
            cls_a = cls[a0:a0 + an]
            cls_b = None
            if use_rgb and use_flow:
                b0, bn, Bf, nsnip_f, _ = spans[nviews + v]
                if (Bf, nsnip_f) != (B, nsnip):
                    raise _lib.SaisError("RGB and flow streams must agree on [B, nsnippets]")
                cls_b = cls[b0:b0 + bn]
            outs.append(ops.clip_head(cls_a, cls_b, B, nsnip, lin_w, lin_b))
        return (outs if is_list else outs[0]), snip_attn

    @torch.no_grad()
    def _forward_mil(self, x, xpad):
        """task='MIL' (:286-295, 344-351, 359-363, 442-443): frame-level encoder per snippet -> clip-level encoder over the
        snippets' CLS vectors (+ clip positional embeddings, no CLS token, no mask) -> ReLU -> attention-based MIL head.
        Returns ``(snip_sequence [nsnip,B,384], snip_reps [B,nsnip,384], output_logits [B,nclasses], attention_dict)`` like
        the reference; the flow stream does not enter these outputs there either (``MIL_Head(snip_reps, flow_reps=None)``)."""
        if self.training:
            raise _lib.SaisError("sais_b200.fullModel is inference-only; call .eval()")
        if isinstance(x, list):
            raise NotImplementedError("task='MIL' takes a single view (the reference's list form never reaches getClipReps)")
        if self.nclasses > 3:
            raise NotImplementedError("attentionModules / finalModules hold three heads (prepare_model.py:84-101)")
        _lib.require_cuda(x, "x")
        B, nsnip, T, E = x.shape
        n = B * nsnip
        if xpad is None:
            key_pad = torch.zeros(n * (T + 1), dtype=torch.uint8, device=x.device)
        else:
            p = xpad.reshape(n * (T + 1)).to(x.device)
            key_pad = p.view(torch.uint8) if p.dtype == torch.bool else p.to(torch.uint8)
        cls, _, _, _ = self.transEncoderFrame.run_packed(x.reshape(n * T, E).float(), [T + 1] * n, key_pad, [False] * n,
                                                         self.frame_cls, self.frame_pos_table)
        tokens = ops.add_pos_rows(cls.view(B, nsnip, E), self.clip_pos_table[:nsnip].detach())  # getClipReps :455-457
        enc, _ = self.transEncoderClip._forward_tokens(tokens, None)                            # [nsnip,B,E], no ReLU yet
        att_c_w = torch.cat([self.attentionModules[str(c)].weight for c in range(self.nclasses)], 0)
        att_c_b = torch.cat([self.attentionModules[str(c)].bias for c in range(self.nclasses)], 0)
        fin_w = torch.cat([self.finalModules[str(c)].weight for c in range(self.nclasses)], 0)
        fin_b = torch.cat([self.finalModules[str(c)].bias for c in range(self.nclasses)], 0)
        reps, logits, attn = ops.mil_head(enc.permute(1, 0, 2).contiguous(),
                                          (self.attentionA.weight, self.attentionA.bias),
                                          (self.attentionB.weight, self.attentionB.bias), att_c_w, att_c_b, fin_w, fin_b)
        return tokens.permute(1, 0, 2).contiguous(), reps, logits, {c: attn[c] for c in range(self.nclasses)}

    def train(self, mode=True):
        if mode:
            raise _lib.SaisError("sais_b200.fullModel is inference-only")
        return super().train(False)


# ---------------------------------------------------------------------- state-dict compatibility
_TABLES = (("frame_pos_table", "frame_pos_embeddings."), ("clip_pos_table", "clip_pos_embeddings."))


def _split_pos_tables(module, state_dict, prefix, local_metadata):
    """state_dict(): expose the [2000,384] tables as the reference's 2000 ``[1,384]`` ParameterDict entries."""
    for table, ref_prefix in _TABLES:
        t = state_dict.pop(prefix + table)
        for i in range(t.shape[0]):
            state_dict[f"{prefix}{ref_prefix}{i}"] = t[i:i + 1]
    return state_dict


def _merge_pos_tables(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
    """load_state_dict(): fold reference ParameterDict entries into the tables; drop the unused ``encoder.*``."""
    for table, ref_prefix in _TABLES:
        keys = [k for k in state_dict if k.startswith(prefix + ref_prefix)]
        if not keys:
            continue
        rows = sorted(keys, key=lambda k: int(k.rsplit('.', 1)[1]))
        state_dict[prefix + table] = torch.cat([state_dict.pop(k).reshape(1, -1) for k in rows], 0)
    for k in [k for k in state_dict if k.startswith(prefix + "encoder.")]:
        state_dict.pop(k)


def load_checkpoint_params(savepath):
    """``params.zip`` as train.py:105-112 writes it (``copy.deepcopy(model['model'].state_dict())`` of the DDP-wrapped
    model: every key carries a ``module.`` prefix, stripped like prepare_model.py:523-527 does; unprefixed keys are
    accepted as well).  Returns the reference-named state dict on CPU."""
    params = torch.load(os.path.join(savepath, 'params.zip'), map_location='cpu', weights_only=True)
    return {(k.split('module.', 1)[1] if k.startswith('module.') else k): v for k, v in params.items()}


def load_prototypes(savepath, device):
    """``prototypes.zip``: a deep-copied ``nn.ParameterDict`` ``'0'..'P-1' -> [1,256]`` (train.py:87,110) — a pickled
    module, which the default ``weights_only=True`` unpickler of torch >= 2.6 rejects.  The file is the user's own
    checkpoint (trusted, exactly as in the reference, prepare_model.py:562)."""
    prototypes = torch.load(os.path.join(savepath, 'prototypes.zip'), map_location=device, weights_only=False)
    if not isinstance(prototypes, (dict, nn.ParameterDict)):
        raise _lib.SaisError(f"prototypes.zip holds a {type(prototypes).__name__}, expected a (Parameter)dict")
    return prototypes


def loadModel(rank, world_size, savepath, data_type, nclasses, domain, rep_dim, encoder_type, task, fold, lr=0.001,
              modalities='RGB-Flow', freeze_encoder_params=True, self_attention=True, importance_loss=False,
              inference=False):
    """Mirror of prepare_model.py:517-570: returns ``({'model': m, 'prototypes': dict}, optimizer, device)``.
    The model is placed on ``cuda:{rank}`` (the reference hard-wires CPU, :544)."""
    model = fullModel(data_type, nclasses, domain, rep_dim, encoder_type, modalities=modalities,
                      freeze_encoder_params=freeze_encoder_params, self_attention=self_attention,
                      importance_loss=importance_loss)
    if inference:
        model.load_state_dict(load_checkpoint_params(savepath))  # strict: a key mismatch raises, as in the reference
    device = torch.device(f'cuda:{rank}' if isinstance(rank, int) else rank)
    if device.type != 'cuda':
        raise _lib.SaisError("sais_b200.loadModel places the model on a CUDA device (there is no CPU path)")
    model.to(device)
    if not inference:
        prototypes = nn.ParameterDict({str(i): nn.Parameter(torch.rand(1, 256, device=device))
                                       for i in range(nclasses)})
    else:
        prototypes = load_prototypes(savepath, device)
    params = list(model.parameters()) + list(prototypes.values())
    optimizer = torch.optim.SGD(params, lr=lr)
    return {'model': model, 'prototypes': prototypes}, optimizer, device
