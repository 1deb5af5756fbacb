"""CPU: the model-level plugin surface — AutoAWQForCausalLM.from_pretrained → quantize →
save_quantized → from_quantized on a tiny random-init Llama / Mistral, the AWQ search arithmetic against golden
vectors produced by the reference's own AwqQuantizer methods (tests/golden/make_golden_awq_search.py), the checkpoint
format, the AWQ-GEMM checkpoint conversion at load, and the fuser's structure.  Nothing here runs a GEMM: on a
CPU tensor the quantized linears raise (there is no fallback)."""
import glob
import json
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from quick_b200 import layout
from quick_b200.awq import AutoAWQForCausalLM
from quick_b200.awq.models._config import AwqConfig
from quick_b200.awq.models.base import shard_state_dict
from quick_b200.awq.modules.linear.quick import WQLinear_QUICK
from quick_b200.awq.quantize import quantizer as qz

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SEARCH_FILES = sorted(glob.glob(os.path.join(GOLD, "awqsearch_*.npz")))


def _tiny_hf(tmp_path, family="llama", layers=2, tie=False):
    import transformers
    kw = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=layers, num_attention_heads=4,
              num_key_value_heads=2, vocab_size=512, max_position_embeddings=128, tie_word_embeddings=tie)
    cfg_cls, cls = {"llama": (transformers.LlamaConfig, transformers.LlamaForCausalLM),
                    "mistral": (transformers.MistralConfig, transformers.MistralForCausalLM),
                    "qwen2": (transformers.Qwen2Config, transformers.Qwen2ForCausalLM)}[family]
    cfg = cfg_cls(sliding_window=64, **kw) if family == "mistral" else cfg_cls(**kw)
    torch.manual_seed(0)
    model = cls(cfg).half()
    path = str(tmp_path / f"fp16_{family}")
    model.save_pretrained(path)
    return path


def _dequant_linear(m: WQLinear_QUICK) -> nn.Linear:
    """fp32 nn.Linear holding what the packed module represents: W16 = fp16(q − z)·s."""
    q, z, s = layout.unpack_quick(m.qweight, m.qzeros, m.scales)
    G = m.group_size
    w16 = ((q - z.repeat_interleave(G, 0)).half() * s.repeat_interleave(G, 0)).float()       # (K, N)
    lin = nn.Linear(m.in_features, m.out_features, bias=m.bias is not None)
    lin.weight.data = w16.t().contiguous()
    if m.bias is not None:
        lin.bias.data = m.bias.float()
    return lin


def _dequantized_copy(hf_model):
    import copy
    from quick_b200.awq.utils.module import set_op_by_name
    m = copy.deepcopy(hf_model)
    for name, mod in list(m.named_modules()):
        if isinstance(mod, WQLinear_QUICK):
            set_op_by_name(m, name, _dequant_linear(mod))
    return m.float()


# ---------------------------------------------------------------------------------------------- search arithmetic
def test_group_quantiser_definition():
    torch.manual_seed(1)
    w = torch.randn(8, 256)
    dq, s, z = qz.pseudo_quantize_tensor(w, 4, 128, get_scale_zp=True)
    assert s.shape == z.shape == (8, 2) and dq.shape == w.shape
    wg = w.view(-1, 128)
    s_ref = (wg.amax(1) - wg.amin(1)).clamp(min=1e-5) / 15
    assert torch.equal(s.view(-1), s_ref)
    assert torch.equal(z.view(-1), (-torch.round(wg.amin(1) / s_ref)).clamp(0, 15))
    q = torch.round(dq.view(-1, 128) / s.view(-1, 1) + z.view(-1, 1))
    assert q.min() >= 0 and q.max() <= 15 and ((dq - w).abs().view(-1, 128) <= s.view(-1, 1) * 0.5001 + 1e-6).all()
    with pytest.raises(ValueError):
        qz.pseudo_quantize_tensor(torch.randn(4, 100), 4, 128)


class _GoldMLP(nn.Module):
    """The module the golden generator inspects: norm-free SwiGLU MLP (gate, up, down)."""

    def __init__(self, H, I):
        super().__init__()
        self.gate_proj, self.up_proj, self.down_proj = nn.Linear(H, I, bias=False), nn.Linear(H, I, bias=False), nn.Linear(I, H, bias=False)

    def forward(self, x):
        return self.down_proj(nn.functional.silu(self.gate_proj(x)) * self.up_proj(x))


@pytest.mark.parametrize("path", SEARCH_FILES, ids=os.path.basename)
def test_search_matches_reference_quantizer(path):
    d = np.load(path)
    G = int(d["G"])
    mlp = _GoldMLP(int(d["H"]), int(d["I"]))
    mlp.gate_proj.weight.data = torch.from_numpy(d["gate"]); mlp.up_proj.weight.data = torch.from_numpy(d["up"])
    mlp.down_proj.weight.data = torch.from_numpy(d["down"])
    x = torch.from_numpy(d["x"])
    before = [p.clone() for p in mlp.parameters()]
    # group quantiser: dequantised values, scales and zeros bit-identical
    dq, s, z = qz.pseudo_quantize_tensor(mlp.down_proj.weight.data, 4, G, get_scale_zp=True)
    assert np.array_equal(dq.numpy(), d["pq_dq"]) and np.array_equal(s.numpy(), d["pq_scales"]) and np.array_equal(z.numpy(), d["pq_zeros"])
    # scale search, several linears sharing one input (norm -> gate, up; inspected through the whole MLP)
    s1 = qz.search_best_scale([mlp.gate_proj, mlp.up_proj], x, module2inspect=mlp, w_bit=4, group_size=G, duo_scaling=bool(d["duo"]))
    np.testing.assert_allclose(s1.numpy(), d["scales_gate_up"], rtol=1e-6, atol=0)
    # scale search, one linear (up -> down)
    h = (nn.functional.silu(mlp.gate_proj(x)) * mlp.up_proj(x)).detach()
    s2 = qz.search_best_scale([mlp.down_proj], h, w_bit=4, group_size=G, duo_scaling=bool(d["duo"]))
    np.testing.assert_allclose(s2.numpy(), d["scales_down"], rtol=1e-6, atol=0)
    assert all(torch.equal(a, b) for a, b in zip(before, mlp.parameters())), "search must leave the weights unchanged"
    # clip search: the chosen threshold of every (channel, group)
    c = qz.search_best_clip(mlp.down_proj.weight.data, h, 4, G)
    assert c.shape == d["clip_down"].shape
    same = np.isclose(c.numpy(), d["clip_down"], rtol=1e-6, atol=0)
    assert same.mean() > 0.995, f"clip thresholds differ in {(~same).sum()} of {same.size} groups"
    # folding: the reference's scale_ln_fcs / scale_fc_fc / apply_clip (scale.py:16-26, 63-101) on the same layer
    class Holder(nn.Module):
        def __init__(self):
            super().__init__()
            self.norm = nn.LayerNorm(int(d["H"]))
            self.mlp = mlp
    layer = Holder()
    layer.norm.weight.data = torch.from_numpy(d["norm_w"].copy()); layer.norm.bias.data = torch.from_numpy(d["norm_b"].copy())
    feats = {"mlp.down_proj": h.clone()}
    with torch.no_grad():
        qz.apply_scale(layer, [("norm", ("mlp.gate_proj", "mlp.up_proj"), torch.from_numpy(d["scales_gate_up"].copy())),
                               ("mlp.up_proj", ("mlp.down_proj",), torch.from_numpy(d["scales_down"].copy()))], feats)
    for got, key in ((layer.norm.weight, "fold_norm_w"), (layer.norm.bias, "fold_norm_b"), (mlp.gate_proj.weight, "fold_gate"),
                     (mlp.up_proj.weight, "fold_up"), (mlp.down_proj.weight, "fold_down")):
        assert np.array_equal(got.detach().numpy(), d[key]), key
    assert torch.equal(feats["mlp.down_proj"], h / torch.from_numpy(d["scales_down"]).view(1, -1))
    qz.apply_clip(layer, [("mlp.down_proj", torch.from_numpy(d["clip_down"].copy()))])
    assert np.array_equal(mlp.down_proj.weight.detach().numpy(), d["clip_applied_down"])


def test_apply_scale_preserves_function():
    torch.manual_seed(3)
    H, I = 128, 256

    class Blk(nn.Module):
        def __init__(self):
            super().__init__()
            self.norm = nn.LayerNorm(H)
            self.mlp = _GoldMLP(H, I)

        def forward(self, x):
            return self.mlp(self.norm(x))

    blk = Blk().double()
    x = torch.randn(16, H, dtype=torch.float64)
    y0 = blk(x)
    s_in, s_mid = torch.rand(H, dtype=torch.float64) + 0.5, torch.rand(I, dtype=torch.float64) + 0.5
    feats = {"mlp.down_proj": torch.ones(4, I, dtype=torch.float64)}
    qz.apply_scale(blk, [("norm", ("mlp.gate_proj", "mlp.up_proj"), s_in), ("mlp.up_proj", ("mlp.down_proj",), s_mid)], feats)
    torch.testing.assert_close(blk(x), y0, rtol=1e-9, atol=1e-9)
    torch.testing.assert_close(feats["mlp.down_proj"], (1 / s_mid).expand(4, I))
    with pytest.raises(NotImplementedError):
        qz.apply_scale(blk, [("mlp", ("mlp.down_proj",), s_mid)])


def test_calib_dataset_forms():
    blocks = qz.get_calib_dataset([[1, 2, 3, 4, 5], [6, 7, 8], list(range(600)), [9, 10]], None, block_size=4)
    assert blocks.tolist() == [[1, 2, 3, 4], [5, 6, 7, 8]]            # > 512-token samples dropped, tail dropped

    class Tok:
        def encode(self, text):
            return [ord(c) for c in text]
    assert qz.get_calib_dataset(["abcd", " efgh "], Tok(), block_size=4).tolist() == [[97, 98, 99, 100], [101, 102, 103, 104]]
    with pytest.raises(EnvironmentError, match="no network"):
        qz.get_calib_dataset("pileval")
    with pytest.raises(ValueError):
        qz.get_calib_dataset([[1, 2]], None, block_size=4)


# ---------------------------------------------------------------------------------------------- config / sharding
def test_awq_config_roundtrip(tmp_path):
    c = AwqConfig.from_dict({"zero_point": True, "q_group_size": 64, "w_bit": 4, "version": "QUICK"})
    c.save_pretrained(str(tmp_path))
    on_disk = json.load(open(tmp_path / "quant_config.json"))
    assert on_disk == {"zero_point": True, "q_group_size": 64, "w_bit": 4, "version": "QUICK", "modules_to_not_convert": None}
    assert AwqConfig.from_pretrained(str(tmp_path)) == c
    assert c.to_transformers_dict() == {"quant_method": "awq", "zero_point": True, "group_size": 64, "bits": 4,
                                        "version": "quick", "modules_to_not_convert": None}
    assert AwqConfig.from_transformers_dict(c.to_transformers_dict()) == c
    assert AwqConfig.from_dict({}).version == "GEMM"             # the reference default (_config.py:15)
    os.remove(tmp_path / "quant_config.json")
    json.dump({"quantization_config": c.to_transformers_dict()}, open(tmp_path / "config.json", "w"))
    assert AwqConfig.from_pretrained(str(tmp_path)) == c           # config.json-only checkpoints


def test_shard_state_dict():
    sd = {f"t{i}": torch.zeros(256, dtype=torch.float16) for i in range(5)}          # 512 B each
    one, idx = shard_state_dict(sd, "10GB", "model.safetensors")
    assert list(one) == ["model.safetensors"] and idx is None
    many, idx = shard_state_dict(sd, 1024, "model.safetensors")
    assert list(many) == [f"model-0000{i}-of-00003.safetensors" for i in (1, 2, 3)]
    assert idx["metadata"]["total_size"] == 2560 and idx["weight_map"]["t4"] == "model-00003-of-00003.safetensors"
    assert [list(s) for s in many.values()] == [["t0", "t1"], ["t2", "t3"], ["t4"]]


# ---------------------------------------------------------------------------------------------- end to end on CPU
@pytest.mark.parametrize("family", ["llama", "mistral", "qwen2"])
def test_quantize_save_load_roundtrip(tmp_path, family):
    fp_path = _tiny_hf(tmp_path, family)
    model = AutoAWQForCausalLM.from_pretrained(fp_path, device_map="cpu", torch_dtype=torch.float32)
    assert not model.is_quantized
    torch.manual_seed(1)
    calib = torch.randint(0, 512, (4, 32))
    probe = torch.randint(0, 512, (2, 16))
    with torch.no_grad():
        logits_fp = model.model(probe).logits.float()
        import copy
        rtn = copy.deepcopy(model.model)                      # plain round-to-nearest of the same linears: the baseline AWQ must not lose to
        for layer in rtn.model.layers:
            for lin in (m for m in layer.modules() if isinstance(m, nn.Linear)):
                lin.weight.data = qz.pseudo_quantize_tensor(lin.weight.data, 4, 128)
        rel_rtn = (rtn(probe).logits.float() - logits_fp).norm() / logits_fp.norm()
    model.quantize(None, quant_config={"zero_point": True, "q_group_size": 128, "w_bit": 4, "version": "QUICK"}, calib_data=calib)
    assert model.is_quantized and model.quant_config.version == "QUICK"
    layer0 = model.model.model.layers[0]
    lin_names = ["self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj", "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj"]
    for n in lin_names:
        m = layer0.get_submodule(n)
        assert isinstance(m, WQLinear_QUICK) and m.qweight.dtype == torch.int32 and m.scales.dtype == torch.float16
        assert m.qweight.shape == (m.in_features // 4, m.out_features // 2)
        assert m.qzeros.shape == (m.in_features // 128, m.out_features // 4) and m.scales.shape == (m.in_features // 128, 2 * m.out_features)
    assert isinstance(model.model.lm_head, nn.Linear)                      # lm_head stays fp16 (reference base.py:396-405)
    scale_names = [(p, l) for p, l, _ in model.search_result["scale"]]
    assert ("model.layers.0.input_layernorm", ("model.layers.0.self_attn.q_proj", "model.layers.0.self_attn.k_proj", "model.layers.0.self_attn.v_proj")) in scale_names
    assert not any(p.endswith("v_proj") for p, _ in scale_names)              # GQA: v -> o pair skipped (llama.py:50-57)
    assert not any("q_proj" in n or "k_proj" in n for n, _ in model.search_result["clip"])
    with pytest.raises(Exception):                                          # no CPU fallback for the GEMM
        model.model(probe)
    # what the packed modules represent is a 4-bit model of the fp16 one: logits stay close
    with torch.no_grad():
        logits_q = _dequantized_copy(model.model)(probe).logits.float()
    rel = (logits_q - logits_fp).norm() / logits_fp.norm()
    assert rel < 0.35 and rel < 1.15 * rel_rtn, (rel, rel_rtn)

    out = str(tmp_path / "quick")
    model.save_quantized(out, shard_size="300KB")
    files = sorted(os.listdir(out))
    assert "quant_config.json" in files and "config.json" in files and "model.safetensors.index.json" in files
    assert json.load(open(os.path.join(out, "config.json")))["quantization_config"]["version"] == "quick"
    assert json.load(open(os.path.join(out, "quant_config.json")))["version"] == "QUICK"

    loaded = AutoAWQForCausalLM.from_quantized(out, device_map="cpu", fuse_layers=False, max_new_tokens=64)
    assert loaded.is_quantized and loaded.config.max_new_tokens == 64
    sd0, sd1 = model.model.state_dict(), loaded.model.state_dict()
    assert sd0.keys() == sd1.keys()
    for k in sd0:
        assert torch.equal(sd0[k].to(sd1[k].dtype), sd1[k]), k
    inv = loaded.model.model.rotary_emb.inv_freq
    torch.testing.assert_close(inv, 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64)))   # rebuilt, not to_empty() garbage

    # fuser structure (CPU): q‖k‖v and gate‖up are the QUICK-layout concatenations, o/down are taken over as they are
    ref_layer = loaded.model.model.layers[0]
    q_parts = [layout.unpack_quick(m.qweight, m.qzeros, m.scales) for m in (ref_layer.self_attn.q_proj, ref_layer.self_attn.k_proj, ref_layer.self_attn.v_proj)]
    o_ptr = ref_layer.self_attn.o_proj.qweight.data_ptr()
    fused = AutoAWQForCausalLM.from_quantized(out, device_map="cpu", fuse_layers=True, max_new_tokens=96, batch_size=3)
    runner = fused.model.model
    assert fused.model.qb200_fused and len(runner.blocks) == 2 and runner.batch == 3
    assert runner.cfg.max_seq_len == (64 if family == "mistral" else 96)        # capped at the sliding window
    blk = runner.blocks[0]
    assert blk.qkv_proj.out_features == 256 + 128 + 128 and blk.gate_up_proj.out_features == 1024
    fq, fz, fs = layout.unpack_quick(blk.qkv_proj.qweight, blk.qkv_proj.qzeros, blk.qkv_proj.scales)
    assert torch.equal(fq, torch.cat([p[0] for p in q_parts], 1)) and torch.equal(fz, torch.cat([p[1] for p in q_parts], 1))
    assert torch.equal(fs, torch.cat([p[2] for p in q_parts], 1))
    assert blk.cache_k.shape == (3, 2, runner.cfg.max_seq_len, 64)
    if family == "qwen2":                                                        # biased q / k / v: concatenated like the weights
        a = ref_layer.self_attn
        assert torch.equal(blk.qkv_proj.bias, torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias]))
    else:
        assert blk.qkv_proj.bias is None
    with pytest.raises(RuntimeError, match="un-fused"):
        fused.save_quantized(str(tmp_path / "nope"))
    with pytest.raises(ValueError, match="KV-cache batch"):
        fused.generate(torch.zeros(1, 4, dtype=torch.long), max_new_tokens=2)
    del o_ptr


def test_export_compatible_then_pack(tmp_path):
    fp_path = _tiny_hf(tmp_path, "llama", layers=1)
    model = AutoAWQForCausalLM.from_pretrained(fp_path, device_map="cpu", torch_dtype=torch.float32)
    calib = [[int(t) for t in row] for row in torch.randint(0, 512, (6, 40))]
    model.quantize(None, quant_config={"q_group_size": 64, "w_bit": 4, "version": "QUICK"}, calib_data=calib,
                   export_compatible=True, seqlen=48)
    assert isinstance(model.model.model.layers[0].mlp.down_proj, nn.Linear)     # scaled + clipped, still fp
    model.pack()
    m = model.model.model.layers[0].mlp.down_proj
    assert isinstance(m, WQLinear_QUICK) and m.group_size == 64 and m.scales.shape == (512 // 64, 2 * 256)


def test_gemm_checkpoint_converts_at_load(tmp_path):
    """A checkpoint in the AWQ "GEMM" layout (what public AWQ checkpoints ship; reference gemm.py:37-58) loads into
    QUICK modules that hold the same integers."""
    from oracle import quick_oracle as qo
    fp_path = _tiny_hf(tmp_path, "llama", layers=1, tie=True)
    model = AutoAWQForCausalLM.from_pretrained(fp_path, device_map="cpu", torch_dtype=torch.float32)
    model.quantize(None, quant_config={"q_group_size": 128, "w_bit": 4, "version": "QUICK"}, calib_data=torch.randint(0, 512, (2, 32)))
    out = str(tmp_path / "quick")
    model.save_quantized(out)
    assert "model.safetensors" in os.listdir(out)
    from safetensors.torch import load_file, save_file
    sd = load_file(os.path.join(out, "model.safetensors"))
    assert "lm_head.weight" not in sd                                           # tied to the embedding
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
    import shutil
    shutil.copytree(out, gdir)
    save_file(gemm_sd, os.path.join(gdir, "model.safetensors"), metadata={"format": "pt"})
    qc = json.load(open(os.path.join(gdir, "quant_config.json"))); qc["version"] = "GEMM"
    json.dump(qc, open(os.path.join(gdir, "quant_config.json"), "w"))
    loaded = AutoAWQForCausalLM.from_quantized(gdir, device_map="cpu", fuse_layers=False)
    assert loaded.quant_config.version == "QUICK"
    for k, v in model.model.state_dict().items():
        assert torch.equal(v.to(loaded.model.state_dict()[k].dtype), loaded.model.state_dict()[k]), k
    assert loaded.model.lm_head.weight.data_ptr() == loaded.model.model.embed_tokens.weight.data_ptr()


def test_auto_rejects_unknown_family_and_bare_ctor(tmp_path):
    import transformers
    transformers.GPT2Config(n_layer=1, n_head=2, n_embd=64).save_pretrained(str(tmp_path))
    with pytest.raises(TypeError, match="isn't supported yet"):
        AutoAWQForCausalLM.from_quantized(str(tmp_path))
    with pytest.raises(EnvironmentError):
        AutoAWQForCausalLM()
    with pytest.raises(FileNotFoundError):
        AutoAWQForCausalLM.from_quantized(str(tmp_path / "missing"))


def test_perplexity_windows_and_loss(tmp_path):
    """evaluate_perplexity == exp(mean next-token NLL over non-overlapping windows) (reference eval_utils.py:20-66),
    on the fp model (no GEMM kernel involved on CPU)."""
    from quick_b200.awq.evaluation import evaluate_perplexity
    model = AutoAWQForCausalLM.from_pretrained(_tiny_hf(tmp_path, layers=1), device_map="cpu", torch_dtype=torch.float32)
    ids = torch.randint(0, 512, (1, 100), generator=torch.Generator().manual_seed(2))
    ppl = evaluate_perplexity(model, None, ids, seqlen=32)
    nll = []
    with torch.no_grad():
        for i in range(3):                                       # 100 // 32 windows, the tail is dropped
            w = ids[:, 32 * i:32 * (i + 1)]
            lg = model.model(w).logits.float()
            nll.append(torch.nn.functional.cross_entropy(lg[0, :-1], w[0, 1:]) * 32)
    assert abs(ppl - float(torch.exp(torch.stack(nll).sum() / 96))) < 1e-3 * ppl
    assert 100 < ppl < 5000                                      # random-init model over a 512-token vocabulary
    assert evaluate_perplexity(model, None, ids, seqlen=32, max_windows=1) > 0
    with pytest.raises(ValueError):
        evaluate_perplexity(model, None, ids, seqlen=256)
    with pytest.raises(ValueError):
        evaluate_perplexity(model, None, "text without tokenizer")
