"""Generates tests/golden/golden.npz from the *installed dependency functions the reference's stack runs on this path*
(the reference itself ships no code or vectors — /root/reference/README.md:3):

  * HF ``Speech2TextFeatureExtractor`` (→ torchaudio ``kaldi.fbank`` + numpy utterance CMVN) on two short synthetic
    utterances;
  * ``torch.nn.functional.ctc_loss`` (+ autograd gradient w.r.t. the logits through fp32 log_softmax, as
    ``Wav2Vec2ForCTC.forward`` calls it) on small cases incl. repeated labels, an empty target and an infeasible one;
  * the greedy collapse rule of ``Wav2Vec2CTCTokenizer`` (groupby + drop pad) on hand-written frame-id strings.

Run in the build container:  python tests/golden/make_golden.py
Versions at generation time are stored in the file.
"""
import itertools
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import synth_wave  # noqa: E402


def main():
    import torchaudio
    import transformers
    from transformers import Speech2TextFeatureExtractor

    out = {}
    waves = [synth_wave(8000, 101), synth_wave(5237, 102)]
    fe = Speech2TextFeatureExtractor(feature_size=80, sampling_rate=16000, num_mel_bins=80)
    enc = fe([w.numpy() for w in waves], sampling_rate=16000, padding=True, return_tensors="np", return_attention_mask=True)
    out["wave0"], out["wave1"] = waves[0].numpy(), waves[1].numpy()
    out["hf_input_features"] = enc["input_features"].astype(np.float32)
    out["hf_attention_mask"] = enc["attention_mask"].astype(np.int32)
    import torchaudio.compliance.kaldi as kaldi
    out["kaldi_fbank0"] = kaldi.fbank(waves[0].unsqueeze(0) * 32768, num_mel_bins=80, sample_frequency=16000).numpy()

    # CTC cases: (T, V, labels per utterance, input lengths)
    g = torch.Generator().manual_seed(5)
    cases = [
        (12, 7, [[1, 2, 2, 3], [4, 4], []], [12, 9, 5]),          # repeats, shorter inputs, empty target
        (6, 5, [[1, 1, 1, 1], [2]], [6, 6]),                      # first is infeasible (needs 7 frames)
        (1, 4, [[3], []], [1, 1]),                                # single frame
    ]
    for ci, (t, v, labs, ilens) in enumerate(cases):
        b = len(labs)
        smax = max(1, max(len(x) for x in labs))
        labels = torch.full((b, smax), -100, dtype=torch.long)
        for i, x in enumerate(labs):
            labels[i, : len(x)] = torch.tensor(x, dtype=torch.long)
        logits = torch.randn(b, t, v, generator=g) * 2.0
        out[f"ctc{ci}_logits"] = logits.numpy()
        out[f"ctc{ci}_labels"] = labels.numpy().astype(np.int32)
        out[f"ctc{ci}_input_lengths"] = np.asarray(ilens, dtype=np.int32)
        for red, zi in itertools.product(("sum", "mean"), (False, True)):
            lg = logits.clone().requires_grad_(True)
            lp = F.log_softmax(lg, dim=-1, dtype=torch.float32).transpose(0, 1)
            tl = (labels >= 0).sum(-1)
            flat = labels.masked_select(labels >= 0)
            with torch.backends.cudnn.flags(enabled=False):
                loss = F.ctc_loss(lp, flat, torch.tensor(ilens), tl, blank=0, reduction=red, zero_infinity=zi)
            loss.backward()
            out[f"ctc{ci}_{red}_{int(zi)}_loss"] = np.asarray(loss.item(), dtype=np.float32)
            out[f"ctc{ci}_{red}_{int(zi)}_grad"] = lg.grad.numpy()

    # greedy collapse: frame ids → collapsed ids via the tokenizer's rule (groupby, then drop pad id 0)
    frames = [[0, 0, 0, 0], [3, 3, 0, 3, 3, 2, 2, 0, 0, 1], [1, 2, 3, 4], [5, 5, 5, 5], [0, 7, 7, 0, 7, 0, 0, 7, 7, 7]]
    for i, fr in enumerate(frames):
        col = [k for k, _ in itertools.groupby(fr)]
        col = [k for k in col if k != 0]
        out[f"greedy{i}_frames"] = np.asarray(fr, dtype=np.int32)
        out[f"greedy{i}_ids"] = np.asarray(col, dtype=np.int32)
    out["versions"] = np.asarray([f"torch {torch.__version__}", f"torchaudio {torchaudio.__version__}",
                                  f"transformers {transformers.__version__}"])
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    print("wrote golden.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
