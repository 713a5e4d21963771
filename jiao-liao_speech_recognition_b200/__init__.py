"""jl-b200: B200-native (sm_100a) ASR forward + adapter fine-tune hot path of Jiao-Liao multi-dialect knowledge
transfer.  Public surface (HF-style, drop-in for the reference's pinned stack on this path):

    JLFeatureExtractor   80-bin Kaldi log-mel + utterance CMVN (fused CUDA kernels)
    JLWaveformFeatureExtractor   padded raw waveforms for the wav2vec2 / XLS-R front end (JLConfig.front_end = "wav2vec2")
    JLConfig             configuration (HF Wav2Vec2Config / Speech2TextConfig field names)
    JLEncoder            conv subsampler + pre-LN transformer with WFAdapter / AttAdapter slots
    WFAdapter, AttAdapter, FusionAdapter (AdapterFusion-style attention over the K source-dialect WFAdapter sets)
    JLForCTC             encoder + CTC head: forward(input_features, attention_mask, labels) → (loss, logits)
    AdapterTrainer       flat-bucket adapter fine-tuning step (CUDA graph + one NCCL all-reduce + fused AdamW)
    Transcriber          waveform → token ids inference step (CUDA graph)
    JLComm               C-ABI NCCL communicator for the one gradient all-reduce of the data-parallel step

All computation goes through ``libjl_b200.so`` (C ABI in ``include/jl_b200.h``); there is no CPU fallback.
The directory name contains a hyphen: import it with ``importlib.import_module("jiao-liao_speech_recognition_b200")``
or through the ``jl_b200`` alias module at the repository root.
"""
from . import _lib, hf_compat, ops, scoring  # noqa: F401
from .comm import JLComm  # noqa: F401
from .configuration import JLConfig  # noqa: F401
from .feature_extraction import JLFeatureExtractor, JLWaveformFeatureExtractor  # noqa: F401
from .modeling import AttAdapter, FusionAdapter, GradSink, JLEncoder, JLEngine, JLForCTC, PackedLayout, WFAdapter  # noqa: F401
from .training import AdapterTrainer, BucketLayout, FlatAdapterParams, Transcriber, ordered_trainables, shard_utterances  # noqa: F401

__version__ = "0.1.0"
