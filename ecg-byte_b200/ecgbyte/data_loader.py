"""Batch form of the tokenisation call site of the reference's ECGTokenDataset
(ecg_byte/data_loader.py:52-132): what `__getitem__` does per sample in Python --
normalize_all -> ''.join -> encode_text -> 'signal_k' ids -> truncate / pad / labels / mask /
position ids -- done for a whole batch on the GPU by three kernels (fused quantise+encode,
then pack)."""
import numpy as np
import torch

from .api import Quantizer, Vocab, pack_training


def signal_token_lut(vocab_keys, tokenizer):
    """LLM id of every 'signal_{k}' token (main.py:144-146 adds them; data_loader.py:80 looks them up)."""
    keys = list(vocab_keys)
    lut = np.zeros(max(keys) + 1, np.int64)
    lut[keys] = tokenizer.convert_tokens_to_ids(["signal_%d" % k for k in keys])
    return torch.from_numpy(lut)


class ECGTokenBatcher:
    """Holds the device-resident vocabulary / quantiser / LUT and turns a batch of raw records plus
    tokenised question / answer ids into the training tensors of `_prepare_training`."""

    def __init__(self, merges, percentiles, lut, pad_to_max, pad_id, bos_id, eos_id, sig_start_id, sig_end_id,
                 dtype=torch.float64, device=None, out_stride=None):
        self.vocab = Vocab(merges=merges, device=device)
        self.quant = Quantizer(percentiles, dtype=dtype, device=device)
        self.dev = torch.device("cuda", self.vocab.device)
        self.lut = lut.to(self.dev)
        self.cfg = dict(pad_to_max=pad_to_max, pad_id=pad_id, bos_id=bos_id, eos_id=eos_id,
                        sig_start_id=sig_start_id, sig_end_id=sig_end_id)
        self.out_stride = out_stride

    def __call__(self, signals, questions, answers):
        """signals: [n, C, L] array / tensor of the quantiser's dtype; questions / answers: lists of
        id lists.  Returns dict with the reference's keys (batched)."""
        x = signals if isinstance(signals, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(signals))
        x = x.to(self.dev)
        # tokens beyond pad_to_max are never used (data_loader.py:106-107): truncate in the encoder output
        stride = self.out_stride or (self.cfg["pad_to_max"] + 8)
        tokens, lens = self.vocab.encode_batch(self.quant, x, out_stride=stride)
        flat, off, ql = [], [0], []
        for q, a in zip(questions, answers):
            flat.extend(q)
            flat.extend(a)
            off.append(len(flat))
            ql.append(len(q))
        ids, attn, labels, pos, status = pack_training(
            tokens, lens, self.lut, torch.tensor(flat, dtype=torch.int64), torch.tensor(off, dtype=torch.int64),
            torch.tensor(ql, dtype=torch.int32), **self.cfg)
        return {"tokenized_signal": ids, "attn_mask": attn, "quantized_signal_ids_input": labels,
                "position_ids": pos, "status": status}
