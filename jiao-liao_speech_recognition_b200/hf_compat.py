"""Checkpoint-name compatibility with the reference's pinned Hugging Face stack.

The reference loads its backbone and per-dialect adapters through ``transformers`` (/root/reference/requirements.txt:81),
so real checkpoints arrive with HF parameter names.  This module maps them onto ``JLForCTC``'s names — it moves
tensors between dictionaries on the host and touches no kernel:

* **wav2vec2 / XLS-R / MMS transformer stack** (``Wav2Vec2ForCTC.state_dict()``, SP/transformers/models/wav2vec2/
  modeling_wav2vec2.py:612-655, :1630): ``wav2vec2.encoder.layers.N.{layer_norm, attention.{q,k,v,out}_proj,
  final_layer_norm, feed_forward.{intermediate,output}_dense}``, ``wav2vec2.encoder.layer_norm``, ``lm_head`` — same
  leaf names as ours, prefix ``wav2vec2.`` dropped.
* **HF's per-language bottleneck adapter** (``…layers.N.adapter_layer.{norm, linear_1, linear_2}``, :931-953, files
  ``adapter.<lang>.safetensors`` / ``adapter.<lang>.bin``, :1046-1060, :1152-1168) is *exactly* a WFAdapter with
  bottleneck = rank = ``adapter_attn_dim`` whose inner factors are identities: W_down = I·linear_1.weight,
  W_up = linear_2.weight·I; it is loaded into the ``adapter_ffn`` slot (HF's site, :647-648).
* **Speech2Text encoder** (``Speech2TextEncoder.state_dict()``, SP/transformers/models/speech_to_text/
  modeling_speech_to_text.py:68-100, :561-608): ``conv.conv_layers.N`` → ``encoder.conv.N``, ``self_attn`` →
  ``attention``, ``self_attn_layer_norm`` → ``layer_norm``, ``fc1`` / ``fc2`` → ``feed_forward.{intermediate,output}_dense``.

The raw-waveform front end of true wav2vec2 (``feature_extractor``, ``feature_projection``, ``pos_conv_embed``) maps onto a
model built with ``front_end="wav2vec2"`` (``encoder.w2v.*``, SURVEY §8 f3: "layer" feature-extractor norm with conv bias);
a mel model reports those keys back as skipped.  ``masked_spec_embed`` (SpecAugment, training-time augmentation) is always
skipped.
"""
from __future__ import annotations

import os
import re
from typing import Dict, List, Tuple

import torch

ADAPTER_SAFE_FILE = "adapter.{}.safetensors"     # modeling_wav2vec2.py: WAV2VEC2_ADAPTER_SAFE_FILE
ADAPTER_PT_FILE = "adapter.{}.bin"               # modeling_wav2vec2.py: WAV2VEC2_ADAPTER_PT_FILE

_FRONT_END = ("feature_extractor.", "feature_projection.", "encoder.pos_conv_embed.", "masked_spec_embed", "encoder.embed_positions.")
_S2T_LEAF = [
    (re.compile(r"^encoder\.conv\.conv_layers\.(\d+)\."), r"encoder.conv.\1."),
    (re.compile(r"^(encoder\.layers\.\d+)\.self_attn_layer_norm\."), r"\1.layer_norm."),
    (re.compile(r"^(encoder\.layers\.\d+)\.self_attn\."), r"\1.attention."),
    (re.compile(r"^(encoder\.layers\.\d+)\.fc1\."), r"\1.feed_forward.intermediate_dense."),
    (re.compile(r"^(encoder\.layers\.\d+)\.fc2\."), r"\1.feed_forward.output_dense."),
]
_HF_ADAPTER = re.compile(r"^(encoder\.layers\.\d+)\.adapter_layer\.(norm|linear_1|linear_2)\.(weight|bias)$")


def _strip_prefix(k: str) -> str:
    for pre in ("wav2vec2.", "model."):
        if k.startswith(pre):
            k = k[len(pre):]
    if k.startswith(("conv.", "layers.", "layer_norm.")):          # bare Speech2TextEncoder / encoder state dict
        k = "encoder." + k
    return k


_W2V_FRONT = [
    (re.compile(r"^feature_extractor\.conv_layers\.(\d+)\.conv\.(weight|bias)$"), r"encoder.w2v.conv.\1.\2"),
    (re.compile(r"^feature_extractor\.conv_layers\.(\d+)\.layer_norm\.(weight|bias)$"), r"encoder.w2v.conv_norm.\1.\2"),
    (re.compile(r"^feature_projection\.layer_norm\.(weight|bias)$"), r"encoder.w2v.proj_norm.\1"),
    (re.compile(r"^feature_projection\.projection\.(weight|bias)$"), r"encoder.w2v.proj.\1"),
    (re.compile(r"^encoder\.pos_conv_embed\.conv\.(?:parametrizations\.weight\.original0|weight_g)$"), r"encoder.w2v.pos_conv.weight_g"),
    (re.compile(r"^encoder\.pos_conv_embed\.conv\.(?:parametrizations\.weight\.original1|weight_v)$"), r"encoder.w2v.pos_conv.weight_v"),
    (re.compile(r"^encoder\.pos_conv_embed\.conv\.bias$"), r"encoder.w2v.pos_conv.bias"),
]


def convert_hf_state_dict(sd: Dict[str, torch.Tensor], num_dialects: int = 1, dialect: int = 0,
                          with_front_end: bool = False) -> Tuple[Dict[str, torch.Tensor], List[str]]:
    """HF-named tensors → (``JLForCTC``-named tensors, list of skipped HF keys).  HF bottleneck-adapter weights become the
    factor set ``dialect`` of a WFAdapter in the ``adapter_ffn`` slot; with ``num_dialects`` > 1 the returned factor
    tensors hold only that set (shape [1, …]) under the key suffix ``@<dialect>`` for ``load_hf_state_dict`` to place."""
    out: Dict[str, torch.Tensor] = {}
    skipped: List[str] = []
    for key, val in sd.items():
        k = _strip_prefix(key)
        if with_front_end:          # raw-waveform front end of a model built with front_end = "wav2vec2" 
            mapped = None
            for pat, rep in _W2V_FRONT:
                if pat.match(k):
                    mapped = pat.sub(rep, k)
                    break
            if mapped is not None:
                out[mapped] = val.detach()
                continue
        if any(k.startswith(p) for p in _FRONT_END):
            skipped.append(key)
            continue
        m = _HF_ADAPTER.match(k)
        if m is not None:
            layer, part, leaf = m.groups()
            base = f"{layer}.adapter_ffn."
            v = val.detach().to(torch.float32)
            tag = "" if num_dialects == 1 else f"@{dialect}"
            if part == "norm":
                out[base + "norm." + leaf] = v
            elif part == "linear_1" and leaf == "weight":            # [a, d]: W_down = I · linear_1.weight
                a = v.shape[0]
                out[base + "down_B" + tag] = v.unsqueeze(0).clone()
                out[base + "down_A" + tag] = torch.eye(a).unsqueeze(0)
            elif part == "linear_1":
                out[base + "down_bias" + tag] = v.unsqueeze(0).clone()
            elif part == "linear_2" and leaf == "weight":            # [d, a]: W_up = linear_2.weight · I
                a = v.shape[1]
                out[base + "up_A" + tag] = v.unsqueeze(0).clone()
                out[base + "up_B" + tag] = torch.eye(a).unsqueeze(0)
            else:
                out[base + "up_bias" + tag] = v.unsqueeze(0).clone()
            continue
        for pat, rep in _S2T_LEAF:
            k = pat.sub(rep, k)
        if k.startswith(("encoder.", "lm_head.")):
            out[k] = val.detach()
        else:
            skipped.append(key)
    return out, skipped


def check_hf_architecture(sd: Dict[str, torch.Tensor], hf_config=None) -> None:
    """``JLEncoderLayer`` is the *stable-layer-norm* (pre-LN) layer of XLS-R / MMS / wav2vec2-large-lv60
    (``Wav2Vec2EncoderLayerStableLayerNorm``, modeling_wav2vec2.py:612-655) and the raw-waveform front end is the "layer"-norm
    feature extractor with conv bias.  Post-LN wav2vec2-base checkpoints (``do_stable_layer_norm=False``,
    ``feat_extract_norm="group"``, no conv bias) carry the same transformer leaf names and would load silently into the wrong
    arithmetic — refuse them.  Detection: the HF config when given, else the feature-extractor keys (a multi-layer "group"
    extractor has a norm on conv layer 0 only and no conv biases — the layout every post-LN wav2vec2-base checkpoint has)."""
    if hf_config is not None:
        get = (lambda k, d=None: hf_config.get(k, d)) if isinstance(hf_config, dict) else (lambda k, d=None: getattr(hf_config, k, d))
        if get("do_stable_layer_norm", True) is False:
            raise ValueError("checkpoint has do_stable_layer_norm=False (post-LN wav2vec2-base layers): only the stable-layer-norm "
                             "(pre-LN) architecture of XLS-R / MMS / wav2vec2-large is implemented")
        if get("feat_extract_norm", "layer") == "group":
            raise ValueError("checkpoint has feat_extract_norm='group': only the 'layer'-norm feature extractor (with conv bias) is implemented")
    keys = {_strip_prefix(k) for k in sd}
    if "feature_extractor.conv_layers.1.conv.weight" in keys:          # a real multi-layer feature extractor
        if "feature_extractor.conv_layers.1.layer_norm.weight" not in keys and "feature_extractor.conv_layers.1.conv.bias" not in keys:
            raise ValueError("checkpoint has a 'group'-norm feature extractor without conv bias (wav2vec2-base, do_stable_layer_norm=False): "
                             "its post-LN transformer layers share these parameter names but not this model's arithmetic — not supported")


def load_hf_state_dict(model, sd: Dict[str, torch.Tensor], strict: bool = False, dialect: int = 0, hf_config=None) -> Tuple[List[str], List[str]]:
    """Copy an HF-named state dict into ``model`` (a ``JLForCTC``).  Returns (missing model keys, skipped HF keys);
    ``strict`` raises if a model parameter outside the adapters stays unset or a shape differs.  Raises for post-LN /
    group-norm wav2vec2-base checkpoints (``check_hf_architecture``; pass the HF config as ``hf_config`` when there is one)."""
    check_hf_architecture(sd, hf_config)
    k_dialects = getattr(model.config, "num_dialects", 1)
    conv, skipped = convert_hf_state_dict(sd, num_dialects=k_dialects, dialect=dialect,
                                          with_front_end=getattr(model.config, "front_end", "mel") == "wav2vec2")
    own = model.state_dict()
    loaded = set()
    with torch.no_grad():
        for k, v in conv.items():
            name, _, tag = k.partition("@")
            if name not in own:
                skipped.append(k)
                continue
            dst = own[name]
            if tag:
                dst = dst[int(tag): int(tag) + 1]
            if tuple(dst.shape) != tuple(v.shape):
                raise ValueError(f"{k}: checkpoint shape {tuple(v.shape)} does not match the model's {tuple(dst.shape)}")
            dst.copy_(v.to(dst.device, dst.dtype))
            loaded.add(name)
    missing = [k for k in own if k not in loaded]
    if strict:
        hard = [k for k in missing if ".adapter_" not in k]
        if hard:
            raise ValueError(f"missing keys in the checkpoint: {hard[:8]}{' …' if len(hard) > 8 else ''}")
    return missing, skipped


def to_hf_state_dict(model, style: str = "wav2vec2") -> Dict[str, torch.Tensor]:
    """``JLForCTC`` → HF-named backbone + head tensors (adapters keep this package's names: HF has no module for them)."""
    if style not in ("wav2vec2", "speech_to_text"):
        raise ValueError("style must be 'wav2vec2' or 'speech_to_text'")
    out = {}
    for k, v in model.state_dict().items():
        if style == "wav2vec2":
            out[("wav2vec2." + k) if k.startswith("encoder.") else k] = v.detach().cpu()
            continue
        k2 = re.sub(r"^encoder\.conv\.(\d+)\.", r"encoder.conv.conv_layers.\1.", k)
        k2 = re.sub(r"^(encoder\.layers\.\d+)\.layer_norm\.", r"\1.self_attn_layer_norm.", k2)
        k2 = re.sub(r"^(encoder\.layers\.\d+)\.attention\.", r"\1.self_attn.", k2)
        k2 = re.sub(r"^(encoder\.layers\.\d+)\.feed_forward\.intermediate_dense\.", r"\1.fc1.", k2)
        k2 = re.sub(r"^(encoder\.layers\.\d+)\.feed_forward\.output_dense\.", r"\1.fc2.", k2)
        out[("model." + k2) if k2.startswith("encoder.") else k2] = v.detach().cpu()
    return out


# ------------------------------------------------------------------------------------------------ adapter files
def adapter_file(model_dir: str, lang: str) -> str:
    """Path of the adapter file of ``lang`` inside a local checkpoint directory — safetensors first, then ``.bin``
    (the lookup order of modeling_wav2vec2.py:1152-1225, local files only: there is no hub access on this path)."""
    for pat in (ADAPTER_SAFE_FILE, ADAPTER_PT_FILE):
        p = os.path.join(model_dir, pat.format(lang))
        if os.path.isfile(p):
            return p
    raise EnvironmentError(f"no {ADAPTER_SAFE_FILE.format(lang)} or {ADAPTER_PT_FILE.format(lang)} in {model_dir} (local files only)")


def read_tensor_file(path: str) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path, device="cpu")
    return torch.load(path, map_location="cpu", weights_only=True)


def write_tensor_file(tensors: Dict[str, torch.Tensor], path: str) -> None:
    tensors = {k: v.detach().cpu().contiguous() for k, v in tensors.items()}
    if path.endswith(".safetensors"):
        from safetensors.torch import save_file
        save_file(tensors, path, metadata={"format": "pt"})
    else:
        torch.save(tensors, path)
