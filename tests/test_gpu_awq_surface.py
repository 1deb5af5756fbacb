"""GPU: the model-level plugin surface end to end on a tiny random-init Llama with grouped-query
attention — quantize on the GPU → save_quantized → from_quantized (HF module tree with WQLinear_QUICK linears, and the
fused runner) → forward / generate, every linear through the tcgen05 kernel, checked against the same network with
each packed linear replaced by its dequantised weight W16 = fp16(q − z)·s and torch.matmul in fp32.
Tolerance: |Δlogit| ≤ 3e-2·rms(logits) + 1e-3 (fp16 activations through two decoder layers; the GEMM itself is held to
1e-2 in test_gpu_parity.py)."""
import copy
import json
import os
import shutil

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _tiny_hf(tmp_path):
    import transformers
    cfg = transformers.LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                                   num_key_value_heads=2, vocab_size=512, max_position_embeddings=128)
    torch.manual_seed(0)
    path = str(tmp_path / "fp16")
    transformers.LlamaForCausalLM(cfg).half().save_pretrained(path)
    return path


def _dense_copy(hf_model):
    from quick_b200 import layout
    from quick_b200.awq.modules.linear.quick import WQLinear_QUICK
    from quick_b200.awq.utils.module import set_op_by_name
    m = copy.deepcopy(hf_model)
    for name, mod in list(m.named_modules()):
        if isinstance(mod, WQLinear_QUICK):
            q, z, s = layout.unpack_quick(mod.qweight, mod.qzeros, mod.scales)
            G = mod.group_size
            w16 = (q - z.repeat_interleave(G, 0)).half() * s.repeat_interleave(G, 0)
            lin = nn.Linear(mod.in_features, mod.out_features, bias=False, device=w16.device)
            lin.weight.data = w16.t().contiguous()
            set_op_by_name(m, name, lin)
    return m.float()


def _close(a, b, what):
    rms = b.float().pow(2).mean().sqrt().item()
    err = (a.float() - b.float()).abs().max().item()
    assert err <= 3e-2 * rms + 1e-3, f"{what}: max|err| {err:.4g} vs rms {rms:.4g}"
    return rms


def test_quantize_save_load_forward_generate(tmp_path, built):
    from oracle import quick_oracle as qo
    from quick_b200 import layout
    from quick_b200.awq import AutoAWQForCausalLM
    from quick_b200.awq.modules.linear.quick import WQLinear_QUICK

    fp_path = _tiny_hf(tmp_path)
    model = AutoAWQForCausalLM.from_pretrained(fp_path, device_map="cuda")
    assert next(model.model.parameters()).is_cuda and next(model.model.parameters()).dtype == torch.float16
    g = torch.Generator().manual_seed(1)
    calib = torch.randint(0, 512, (4, 32), generator=g)
    probe = torch.randint(0, 512, (2, 16), generator=g).cuda()
    model.quantize(None, quant_config={"zero_point": True, "q_group_size": 128, "w_bit": 4, "version": "QUICK"}, calib_data=calib)
    assert all(isinstance(m, WQLinear_QUICK) and m.qweight.is_cuda for m in
               (model.model.model.layers[1].self_attn.k_proj, model.model.model.layers[0].mlp.down_proj))

    # (1) the HF module tree with packed linears: every linear through the kernel
    dense = _dense_copy(model.model)
    with torch.no_grad():
        ref = dense(probe).logits
        got = model(probe).logits
    rms = _close(got, ref, "HF tree with WQLinear_QUICK linears")

    # (2) checkpoint round trip, un-fused: same integers, same logits
    out = str(tmp_path / "quick")
    model.save_quantized(out)
    loaded = AutoAWQForCausalLM.from_quantized(out, fuse_layers=False, max_new_tokens=64)
    sd0, sd1 = model.model.state_dict(), loaded.model.state_dict()
    assert sd0.keys() == sd1.keys() and all(torch.equal(sd0[k], sd1[k]) for k in sd0)
    with torch.no_grad():
        _close(loaded(probe).logits, got, "reloaded checkpoint")
    hf_gen = loaded.generate(probe, max_new_tokens=4, do_sample=False, attention_mask=torch.ones_like(probe), pad_token_id=0)
    assert hf_gen.shape == (2, 20) and torch.equal(hf_gen[:, :16], probe)

    # (3) the fused runner built from the same checkpoint
    fused = AutoAWQForCausalLM.from_quantized(out, fuse_layers=True, max_new_tokens=64, batch_size=2)
    runner = fused.model.model
    assert fused.model.qb200_fused and runner.cfg.num_kv_heads == 2 and runner.cfg.max_seq_len == 64
    res = fused(probe)
    _close(res.logits, ref, "fused runner, all positions")
    # the reference's stateful convention (examples/benchmark.py:47-60): single-token calls append to the cache
    seq = probe
    for _ in range(3):
        tok = res[0][:, -1].max(1)[1].unsqueeze(1)
        seq = torch.cat([seq, tok], 1)
        res = fused(tok, use_cache=True)
        assert res.logits.shape == (2, 1, 512) and runner.start_pos == seq.shape[1]
        with torch.no_grad():
            _close(res.logits[:, -1], dense(seq).logits[:, -1], f"stateful decode at position {seq.shape[1] - 1}")
    res = fused(probe)                      # a multi-token call starts over at position 0
    assert runner.start_pos == 16
    _close(res.logits, ref, "fused runner, second prefill")

    def consistent(seq, n_new, what):
        """every generated token must be (within tolerance) the arg-max of the dense model on the same prefix"""
        assert seq.shape == (2, 16 + n_new) and torch.equal(seq[:, :16], probe), what
        with torch.no_grad():
            lg = dense(seq[:, :-1]).logits[:, 15:, :]
        chosen = lg.gather(-1, seq[:, 16:, None]).squeeze(-1)
        assert (lg.max(-1).values - chosen).max().item() <= 6e-2 * rms + 2e-3, what

    gen = fused.generate(probe, max_new_tokens=8)
    consistent(gen, 8, "greedy generate (CUDA-graph decode)")
    consistent(fused.generate(probe, max_new_tokens=8, use_graph=False), 8, "greedy generate (eager decode)")
    consistent(fused.generate(probe, max_new_tokens=5), 5, "second generate on the same cache / graph")
    sampled = fused.generate(probe, max_new_tokens=6, do_sample=True, top_k=8, top_p=0.9, temperature=0.7,
                             eos_token_id=int(gen[0, 17]), generator=torch.Generator(device="cuda").manual_seed(3))
    assert sampled.shape[0] == 2 and 17 <= sampled.shape[1] <= 22 and torch.equal(sampled[:, :16], probe)
    with pytest.raises(ValueError, match="exceed the cache length"):
        fused.generate(probe, max_new_tokens=60)

    # (4) the same weights shipped in the AWQ "GEMM" layout convert at load (GPU converter kernels) to the same integers
    from safetensors.torch import load_file, save_file
    sd = load_file(os.path.join(out, "model.safetensors"))
    gemm_sd = {}
    for k, v in sd.items():
        if k.endswith(".qweight"):
            base = k[: -len(".qweight")]
            q, z, s = layout.unpack_quick(sd[base + ".qweight"], sd[base + ".qzeros"], sd[base + ".scales"])
            gq, gz = qo.pack_awq_gemm(q.numpy(), z.numpy())
            gemm_sd[base + ".qweight"], gemm_sd[base + ".qzeros"], gemm_sd[base + ".scales"] = torch.from_numpy(gq), torch.from_numpy(gz), s
        elif not (k.endswith(".qzeros") or k.endswith(".scales")):
            gemm_sd[k] = v
    gdir = str(tmp_path / "gemm")
    shutil.copytree(out, gdir)
    save_file(gemm_sd, os.path.join(gdir, "model.safetensors"), metadata={"format": "pt"})
    qc = json.load(open(os.path.join(gdir, "quant_config.json")))
    qc["version"] = "GEMM"
    json.dump(qc, open(os.path.join(gdir, "quant_config.json"), "w"))
    from_gemm = AutoAWQForCausalLM.from_quantized(gdir, fuse_layers=False)
    sd2 = from_gemm.model.state_dict()
    assert all(torch.equal(sd0[k], sd2[k]) for k in sd0)


def _save_tokenizer(path, vocab_size=512):
    """A local word-level tokenizer (no hub access): what AutoTokenizer.from_pretrained(model_path) finds."""
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    tok = Tokenizer(models.WordLevel({f"t{i}": i for i in range(vocab_size)}, unk_token="t0"))
    tok.pre_tokenizer = pre_tokenizers.Whitespace()
    PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="t0", pad_token="t1").save_pretrained(path)


def test_reference_benchmark_flow_through_the_quick_namespace(tmp_path, built):
    """The reference's own benchmark (examples/benchmark.py) against a random-init AWQ-QUICK checkpoint, through the
    reference's import path.  When the reference tree is present the UNMODIFIED script is executed (runpy, its own
    argument parser); on the GPU box, where /root/reference does not exist, the same call sequence is replayed line by
    line (benchmark.py:38-67 generate_torch, :91-150 run_round): `from quick.awq import AutoAWQForCausalLM`,
    `from_quantized(model_path, quant_file, max_new_tokens=…, batch_size=…, safetensors=…)`, `warmup(model)`,
    `model(inputs, use_cache=True)` with the cache position carried by the model, `out[0][:, -1].max(1)[1]`,
    `model.quant_config.version`."""
    import runpy
    import sys

    import numpy as np
    from quick.awq import AutoAWQForCausalLM                      # the reference's import lines (benchmark.py:6-7)
    from quick.awq.models.base import BaseAWQForCausalLM
    from transformers import AutoTokenizer

    fp_path = _tiny_hf(tmp_path)
    model = AutoAWQForCausalLM.from_pretrained(fp_path, device_map="cuda")
    calib = torch.randint(0, 512, (4, 32), generator=torch.Generator().manual_seed(1))
    model.quantize(None, quant_config={"zero_point": True, "q_group_size": 128, "w_bit": 4, "version": "QUICK"}, calib_data=calib)
    quant_path = str(tmp_path / "quick")
    model.save_quantized(quant_path)
    _save_tokenizer(quant_path)
    del model

    ref_script = "/root/reference/examples/benchmark.py"
    if os.path.exists(ref_script):
        argv = sys.argv
        sys.argv = [ref_script, "--model_path", quant_path, "--batch_size", "2"]
        try:
            runpy.run_path(ref_script, run_name="__main__")       # rounds beyond the cache length end the script's loop
        except (ValueError, RuntimeError) as e:                   # (the reference sweeps contexts up to 4096)
            assert "exceed" in str(e) or "position" in str(e), e
        finally:
            sys.argv = argv

    tokenizer = AutoTokenizer.from_pretrained(quant_path, trust_remote_code=True)
    batch_size, context, n_generate = 2, 16, 32      # cache = max_new_tokens = 32: decoding runs past it and rolls, like the reference
    input_ids = torch.randint(0, tokenizer.vocab_size, (batch_size, context)).cuda()
    model = AutoAWQForCausalLM.from_quantized(quant_path, "", max_new_tokens=n_generate, batch_size=batch_size, safetensors=True)
    assert isinstance(model, BaseAWQForCausalLM)
    warm_up = torch.randn((512, 512)).to(next(model.parameters()).device)       # warmup(model), benchmark.py:34-36
    torch.mm(warm_up, warm_up)
    context_time, generate_time = 0, []
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tokens = []
    with torch.inference_mode():
        for i in range(n_generate):
            start.record()
            inputs = torch.as_tensor(input_ids if i == 0 else token, device=next(model.parameters()).device)
            out = model(inputs, use_cache=True)
            end.record()
            torch.cuda.synchronize()
            token = out[0][:, -1].max(1)[1].unsqueeze(1)
            tokens.append(token)
            if i == 0:
                context_time += start.elapsed_time(end) * 1e-3
            else:
                generate_time.append(start.elapsed_time(end) * 1e-3)
    prefill_tps = input_ids.shape[1] / context_time * batch_size
    decode_tps = 1 / np.median(generate_time) * batch_size
    assert prefill_tps > 0 and decode_tps > 0 and model.quant_config.version == "QUICK"
    assert model.model.model.cfg.max_seq_len == 32 and model.model.model.start_pos <= 32     # the cache rolled (fused_utils.py:26-28)
    # the stateful loop produced a real greedy continuation: same tokens as generate() from the same prompt
    gen = model.generate(input_ids, max_new_tokens=8)
    assert torch.equal(gen[:, context:context + 8], torch.cat(tokens[:8], dim=1))
