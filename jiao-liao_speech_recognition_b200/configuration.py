"""Configuration of the Jiao-Liao ASR path.  Field names follow HF ``Wav2Vec2Config``
(``SP/transformers/models/wav2vec2/configuration_wav2vec2.py:165-219``) and ``Speech2TextConfig``
(conv_channels, input_feat_per_channel) so that a reference recipe's config maps one to one."""
from __future__ import annotations

from dataclasses import asdict, dataclass
from typing import Optional


@dataclass
class JLConfig:
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    conv_channels: int = 1024
    input_feat_per_channel: int = 80
    vocab_size: int = 5000
    pad_token_id: int = 0                   # = CTC blank (configuration_wav2vec2.py:211)
    ctc_loss_reduction: str = "sum"         # configuration_wav2vec2.py:203
    ctc_zero_infinity: bool = False         # :204
    layer_norm_eps: float = 1e-5            # :179
    initializer_range: float = 0.02
    # adapter kinds: "wf" = WFAdapter (factor set chosen by dialect id), "att" = AttAdapter (attention over the utterance's frames),
    # "fuse" = FusionAdapter (AdapterFusion-style attention over the outputs of the num_dialects source-dialect WFAdapter sets)
    adapter_attn: Optional[str] = None      # None | "wf" | "att" | "fuse"  — slot after the self-attention residual
    adapter_ffn: Optional[str] = None       # None | "wf" | "att" | "fuse"  — slot after the FFN residual (HF's adapter_layer site)
    wf_bottleneck: int = 256
    wf_rank: int = 32
    att_dim: int = 64
    num_dialects: int = 1
    logits_dtype: str = "float32"           # "float32" | "bfloat16"
    # front end: "mel" = 80-bin log-mel + Conv1d(k5, s2)+GLU x2 (Speech2Text); "wav2vec2" = raw-waveform conv stack with layer
    # norm, feature projection and grouped positional convolution (XLS-R / MMS / wav2vec2-large, configuration_wav2vec2.py:
    # 182-190, feat_extract_norm = "layer", conv_bias = True, do_stable_layer_norm = True)
    front_end: str = "mel"
    conv_dim: int = 512
    conv_kernel: tuple = (10, 3, 3, 3, 3, 2, 2)
    conv_stride: tuple = (5, 2, 2, 2, 2, 2, 2)
    num_conv_pos_embeddings: int = 128
    num_conv_pos_embedding_groups: int = 16

    def __post_init__(self):
        if self.hidden_size % 64 or self.hidden_size // self.num_attention_heads != 64:
            raise ValueError("hidden_size / num_attention_heads must be 64 (the attention kernel's head_dim)")
        for slot in (self.adapter_attn, self.adapter_ffn):
            if slot not in (None, "wf", "att", "fuse"):
                raise ValueError(f"unknown adapter kind {slot!r}")
            if slot == "fuse" and not 1 <= self.num_dialects <= 8:
                raise ValueError("the fusion adapter attends over 1..8 source-dialect adapters (num_dialects)")
        if self.att_dim != 64:
            raise ValueError("att_dim must be 64 (one attention head of head_dim 64)")
        if self.wf_rank % 8 or self.wf_bottleneck % 8:
            raise ValueError("wf_rank and wf_bottleneck must be multiples of 8")
        if self.ctc_loss_reduction not in ("sum", "mean"):
            raise ValueError("ctc_loss_reduction must be 'sum' or 'mean'")
        if self.logits_dtype not in ("float32", "bfloat16"):
            raise ValueError("logits_dtype must be 'float32' or 'bfloat16'")
        if self.front_end not in ("mel", "wav2vec2"):
            raise ValueError("front_end must be 'mel' or 'wav2vec2'")
        self.conv_kernel, self.conv_stride = tuple(self.conv_kernel), tuple(self.conv_stride)
        if self.front_end == "wav2vec2":
            if len(self.conv_kernel) != len(self.conv_stride) or not self.conv_kernel:
                raise ValueError("conv_kernel and conv_stride must have the same, non-zero length")
            if self.conv_kernel[0] > 16 or self.conv_dim % 8:
                raise ValueError("wav2vec2 front end: first conv kernel must be <= 16 and conv_dim a multiple of 8")
            g = self.num_conv_pos_embedding_groups
            if self.hidden_size % g or (self.hidden_size // g) % 8:
                raise ValueError("hidden_size / num_conv_pos_embedding_groups must be a multiple of 8")

    def to_dict(self):
        return asdict(self)

    @classmethod
    def base(cls, **kw):
        """12-layer d=768 encoder (BASELINE.json configs 1, 2)."""
        return cls(**kw)

    @classmethod
    def xlsr(cls, **kw):
        """XLS-R-300M / MMS-style model: the large transformer stack behind the raw-waveform front end."""
        d = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096, front_end="wav2vec2")
        d.update(kw)
        return cls(**d)

    @classmethod
    def large(cls, **kw):
        """24-layer d=1024 XLS-R / wav2vec2-large-style transformer stack (BASELINE.json config 3)."""
        d = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096)
        d.update(kw)
        return cls(**d)
