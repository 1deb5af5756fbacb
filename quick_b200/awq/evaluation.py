"""Perplexity of a (quantized or fp16) model over local text — the windowing and loss of the reference's
``evaluate_perplexity`` (quick/awq/evaluation/eval_utils.py:20-66: non-overlapping windows of ``seqlen`` tokens, mean
next-token cross entropy per window × seqlen, exp of the total over n_windows · seqlen).  The reference downloads
wikitext-2; here the text (or the token ids) comes from the caller.  lm-eval / MMLU / HumanEval / LibriSpeech
harnesses of that module are evaluation tooling outside the W4A16 hot path and are not rebuilt."""
from __future__ import annotations

from typing import Optional, Union

import torch
import torch.nn.functional as F


@torch.no_grad()
def evaluate_perplexity(model, tokenizer=None, data: Union[str, torch.Tensor, None] = None, seqlen: int = 2048,
                        max_windows: Optional[int] = None) -> float:
    """model: anything callable as ``model(ids)`` returning an object with ``.logits`` of every position (an
    AutoAWQForCausalLM wrapper — fused or not — or a HF model).  data: a string (needs tokenizer) or token ids (1, T)."""
    if isinstance(data, str):
        if tokenizer is None:
            raise ValueError("text data needs a tokenizer")
        ids = tokenizer(data, return_tensors="pt").input_ids
    elif isinstance(data, torch.Tensor):
        ids = data.reshape(1, -1).long()
    else:
        raise ValueError("data must be a string or a tensor of token ids (no dataset download in this library)")
    n = ids.numel() // seqlen
    if max_windows is not None:
        n = min(n, max_windows)
    if n == 0:
        raise ValueError(f"{ids.numel()} tokens are fewer than one window of {seqlen}")
    inner = getattr(model, "model", model)
    try:
        dev = next(inner.parameters()).device
    except StopIteration:
        dev = torch.device("cpu")
    total = torch.zeros((), dtype=torch.float64)
    for i in range(n):
        batch = ids[:, i * seqlen:(i + 1) * seqlen].to(dev)
        logits = model(batch).logits
        loss = F.cross_entropy(logits[:, :-1, :].float().reshape(-1, logits.shape[-1]), batch[:, 1:].reshape(-1))
        total += loss.double().cpu() * seqlen
    return float(torch.exp(total / (n * seqlen)))
