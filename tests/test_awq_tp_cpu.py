"""CPU, world_size 2, gloo: a checkpoint loaded through AutoAWQForCausalLM.from_quantized under torch.distributed is
tensor-parallel — every rank keeps q‖k‖v of ITS heads (attention and the KV cache are head-sharded), its N/2 output
columns of o_proj / down_proj and [gate | up] of its half of the MLP width (layout.slice_columns), and the runner
all-gathers the attention output, the MLP activation and the two column-parallel outputs (BASELINE config 5).  The kernel cannot run on CPU, so a dequantise + matmul stand-in
replaces WQLinear_QUICK.forward INSIDE THIS TEST ONLY (test scaffolding, like the injectable gemm_fn of
test_parallel_cpu.py); what is checked is the host logic: shard contents, gather order, identical logits and tokens on
both ranks, equal to the single-process model."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _install_cpu_stand_in():
    from quick_b200 import layout
    from quick_b200.awq.modules.linear.quick import WQLinear_QUICK

    def forward(self, x, residual=None):
        q, z, s = layout.unpack_quick(self.qweight, self.qzeros, self.scales)
        G = self.group_size
        w16 = ((q - z.repeat_interleave(G, 0)).half() * s.repeat_interleave(G, 0)).float()
        y = (x.float() @ w16).to(x.dtype)
        if self.bias is not None:
            y = y + self.bias
        return y if residual is None else residual + y
    WQLinear_QUICK.forward = forward


def _make_checkpoint(tmp):
    import transformers
    from quick_b200.awq import AutoAWQForCausalLM
    cfg = transformers.LlamaConfig(hidden_size=512, intermediate_size=640, num_hidden_layers=2, num_attention_heads=4,
                                   num_key_value_heads=2, vocab_size=512, max_position_embeddings=128)   # head_dim 128, GQA;
    # 640 = 5 tiles: not divisible by 128 x 2 ranks -> exercises the zero-weight padding of the MLP width (to 768)
    torch.manual_seed(0)
    transformers.LlamaForCausalLM(cfg).half().save_pretrained(os.path.join(tmp, "fp16"))
    m = AutoAWQForCausalLM.from_pretrained(os.path.join(tmp, "fp16"), device_map="cpu", torch_dtype=torch.float32)
    m.quantize(None, quant_config={"q_group_size": 128, "w_bit": 4, "version": "QUICK"}, calib_data=torch.randint(0, 512, (2, 32)))
    m.save_quantized(os.path.join(tmp, "quick"))
    return os.path.join(tmp, "quick")


def _worker(rank, world, port, ckpt, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), QB200_TP_MODE="nccl", LOCAL_RANK=str(rank))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quick_b200 import layout
    from quick_b200.awq import AutoAWQForCausalLM
    _install_cpu_stand_in()
    full = AutoAWQForCausalLM.from_quantized(ckpt, device_map="cpu", fuse_layers=False)
    layer = full.model.model.layers[0]
    model = AutoAWQForCausalLM.from_quantized(ckpt, device_map="cpu", fuse_layers=True, max_new_tokens=48, batch_size=2)
    blk = model.model.model.blocks[0]
    hd, nh_l, nkv_l = 128, 4 // world, 2 // world
    ok = blk.qkv_proj.out_features == (nh_l + 2 * nkv_l) * hd and blk.gate_up_proj.out_features == 2 * 768 // world
    ok = ok and blk.o_proj.out_features == 512 // world and blk.down_proj.in_features == 768 and blk.o_proj.tp_sharded
    ok = ok and tuple(blk.cache_k.shape[:2]) == (2, nkv_l)
    # this rank's q‖k‖v is [its query heads | its key heads | its value heads]
    qq, kk, vv = (layout.unpack_quick(m.qweight, m.qzeros, m.scales)[0] for m in (layer.self_attn.q_proj, layer.self_attn.k_proj, layer.self_attn.v_proj))
    want = torch.cat([qq[:, rank * nh_l * hd:(rank + 1) * nh_l * hd], kk[:, rank * nkv_l * hd:(rank + 1) * nkv_l * hd],
                      vv[:, rank * nkv_l * hd:(rank + 1) * nkv_l * hd]], 1)
    q_mine, _, _ = layout.unpack_quick(blk.qkv_proj.qweight, blk.qkv_proj.qzeros, blk.qkv_proj.scales)
    ok = ok and torch.equal(q_mine, want)
    # ... and [gate | up] of its slice of the (padded) MLP width; the padding channels are zero weights
    gg, uu = (layout.unpack_quick(m.qweight, m.qzeros, m.scales)[0] for m in (layer.mlp.gate_proj, layer.mlp.up_proj))
    gu_mine, _, s_mine = layout.unpack_quick(blk.gate_up_proj.qweight, blk.gate_up_proj.qzeros, blk.gate_up_proj.scales)
    I_l, lo = 768 // world, rank * (768 // world)
    for half, full in ((gu_mine[:, :I_l], gg), (gu_mine[:, I_l:], uu)):
        real = max(0, min(640, lo + I_l) - lo)
        ok = ok and torch.equal(half[:, :real], full[:, lo:lo + real]) and int(half[:, real:].abs().sum()) == 0
    ok = ok and float(s_mine[:, max(0, min(640, lo + I_l) - lo):I_l].abs().sum()) == 0.0
    ids = torch.randint(0, 512, (2, 12), generator=torch.Generator().manual_seed(5))
    logits = model(ids).logits
    seq = model.generate(ids, max_new_tokens=5)
    sampled = model.generate(ids, max_new_tokens=4, do_sample=True, top_k=8, generator=torch.Generator().manual_seed(100 + rank))
    ret[rank] = (bool(ok), logits.float(), seq, sampled)
    dist.destroy_process_group()


def test_from_quantized_is_tensor_parallel_under_torch_distributed(tmp_path):
    from quick_b200.awq import AutoAWQForCausalLM
    ckpt = _make_checkpoint(str(tmp_path))
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ckpt, ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0]
    assert torch.equal(ret[0][1], ret[1][1]) and torch.equal(ret[0][2], ret[1][2])
    assert torch.equal(ret[0][3], ret[1][3]), "sampled tokens must agree across ranks (rank 0's draw is broadcast)"

    saved = None
    try:        # the single-process model with the same stand-in
        from quick_b200.awq.modules.linear.quick import WQLinear_QUICK
        saved = WQLinear_QUICK.forward
        _install_cpu_stand_in()
        single = AutoAWQForCausalLM.from_quantized(ckpt, device_map="cpu", fuse_layers=True, max_new_tokens=48, batch_size=2)
        ids = torch.randint(0, 512, (2, 12), generator=torch.Generator().manual_seed(5))
        ref = single(ids).logits.float()
        ref_seq = single.generate(ids, max_new_tokens=5)
    finally:
        if saved is not None:
            WQLinear_QUICK.forward = saved
    rms = ref.pow(2).mean().sqrt()
    assert (ret[0][1] - ref).abs().max() <= 2e-2 * rms
    assert ret[0][2].shape == ref_seq.shape == (2, 17)
