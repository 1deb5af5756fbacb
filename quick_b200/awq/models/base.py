"""BaseAWQForCausalLM — the model-level plugin surface of the reference (quick/awq/models/base.py:68-458):
``from_pretrained`` → ``quantize`` → ``save_quantized`` and ``from_quantized`` → ``generate``, with the same argument
names and the same on-disk checkpoint (config.json with ``quantization_config``, ``quant_config.json``, safetensors shards
holding ``qweight`` / ``qzeros`` / ``scales`` per linear in the QUICK layout).

Written against transformers 5.x without accelerate (neither the reference's ``shard_checkpoint`` import, base.py:12,
nor ``load_checkpoint_and_dispatch``, base.py:309-316, exist in this environment):
  * the skeleton is built on the meta device, its decoder-layer linears are swapped for ``WQLinear_QUICK`` (init only,
    base.py:406-415), storage is allocated on ONE target device and the checkpoint tensors are copied in file by file;
  * ``device_map`` accepts None / "auto" / "balanced" / "cuda[:i]" / "cpu" / {"": dev} — one device.  Models that need
    several GPUs run tensor-parallel: under torchrun every rank loads the checkpoint, ``fuse_layers`` keeps its N/R
    output columns of every linear (layout.shard_columns) and the runner all-gathers (fused GEMM + all-gather over
    peer memory, or NCCL) — not layer-scattered;
  * AWQ "GEMM"-layout checkpoints (what public AWQ checkpoints ship) load too: each linear's tensors go through the
    bit-exact GEMM → QUICK converter (WQLinear_QUICK.from_awq_gemm) — the reference cannot do that without the fp16 model;
  * ``fuse_layers=True`` swaps the HF decoder for the library's fused runner (fused q‖k‖v and gate‖up GEMMs, fused glue
    kernels, static KV cache, CUDA-graph decode); ``generate`` then runs the runner's own loop.
Exllama / GEMV back-ends are other kernels' surfaces and are not part of this library.
"""
from __future__ import annotations

import gc
import json
import os
import re
from typing import Dict, List, Union

import torch
import torch.nn as nn

from ..modules.linear.quick import WQLinear_QUICK
from ..quantize.quantizer import AwqQuantizer
from ..utils.module import exclude_layers_to_not_quantize, get_named_linears, set_op_by_name
from ._config import AwqConfig

_SIZE_RE = re.compile(r"^\s*([0-9.]+)\s*([KMGT]?I?B)\s*$", re.I)
_UNITS = {"B": 1, "KB": 10 ** 3, "MB": 10 ** 6, "GB": 10 ** 9, "TB": 10 ** 12, "KIB": 2 ** 10, "MIB": 2 ** 20, "GIB": 2 ** 30,
          "TIB": 2 ** 40}


def _parse_size(size: Union[int, str]) -> int:
    if isinstance(size, int):
        return size
    m = _SIZE_RE.match(size)
    if not m:
        raise ValueError(f"cannot parse shard size {size!r}")
    return int(float(m.group(1)) * _UNITS[m.group(2).upper()])


def shard_state_dict(state_dict: Dict[str, torch.Tensor], max_shard_size: Union[int, str], weights_name: str):
    """Greedy split in key order (what the reference gets from transformers' ``shard_checkpoint``, base.py:179-181):
    returns ({file name: {key: tensor}}, index or None)."""
    limit = _parse_size(max_shard_size)
    shards, cur, cur_bytes = [], {}, 0
    for k, v in state_dict.items():
        nbytes = v.numel() * v.element_size()
        if cur and cur_bytes + nbytes > limit:
            shards.append(cur)
            cur, cur_bytes = {}, 0
        cur[k] = v
        cur_bytes += nbytes
    if cur or not shards:
        shards.append(cur)
    if len(shards) == 1:
        return {weights_name: shards[0]}, None
    stem, ext = os.path.splitext(weights_name)
    named = {f"{stem}-{i + 1:05d}-of-{len(shards):05d}{ext}": s for i, s in enumerate(shards)}
    total = sum(v.numel() * v.element_size() for v in state_dict.values())
    index = {"metadata": {"total_size": total}, "weight_map": {k: f for f, s in named.items() for k in s}}
    return named, index


def _resolve_device(device_map) -> torch.device:
    if isinstance(device_map, dict):
        vals = set(device_map.values())
        if len(vals) != 1:
            raise NotImplementedError("layer-scattered device maps need accelerate; use one device (or torchrun tensor "
                                      "parallelism for models larger than one GPU)")
        device_map = vals.pop()
    if device_map in (None, "auto", "balanced", "balanced_low_0", "sequential"):
        if not torch.cuda.is_available():
            return torch.device("cpu")
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and "LOCAL_RANK" in os.environ:
            return torch.device("cuda", int(os.environ["LOCAL_RANK"]))      # one process per GPU (torchrun)
        return torch.device("cuda", torch.cuda.current_device())
    if isinstance(device_map, int):
        return torch.device("cuda", device_map)
    return torch.device(device_map)


def _checkpoint_files(path: str, safetensors: bool) -> List[str]:
    if os.path.isfile(path):
        return [path]
    stem = "model.safetensors" if safetensors else "pytorch_model.bin"
    index = os.path.join(path, stem + ".index.json")
    if os.path.exists(index):
        with open(index) as f:
            return [os.path.join(path, n) for n in sorted(set(json.load(f)["weight_map"].values()))]
    single = os.path.join(path, stem)
    if os.path.exists(single):
        return [single]
    raise FileNotFoundError(f"no {stem} (or {stem}.index.json) under {path}")


def _read_checkpoint_file(file: str) -> Dict[str, torch.Tensor]:
    if file.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(file)
    return torch.load(file, map_location="cpu", weights_only=True)


class BaseAWQForCausalLM(nn.Module):
    layer_type: str = ""
    max_new_tokens_key: str = "max_position_embeddings"

    def __init__(self, model, model_type, is_quantized, config, quant_config, processor=None):
        super().__init__()
        self.model = model
        self.model_type: str = model_type
        self.is_quantized: bool = is_quantized
        self.search_result = None
        self.config = config
        self.quant_config: AwqConfig = quant_config
        self.processor = processor

    # ---- per-family hooks (reference llama.py:13-76)
    @staticmethod
    def get_model_layers(model):
        raise NotImplementedError

    @staticmethod
    def get_layers_for_scaling(module, input_feat, module_kwargs):
        raise NotImplementedError

    @staticmethod
    def get_act_for_scaling(module):
        return dict(is_scalable=False)

    @staticmethod
    def move_embed(model, device):
        raise NotImplementedError

    @staticmethod
    def fuse_layers(model, batch_size: int = 1):
        pass

    # ---- nn.Module plumbing
    def to(self, device):
        return self.model.to(device)

    def forward(self, *args, **kwargs):
        return self.model(*args, **kwargs)

    def generate(self, *args, **kwargs):
        with torch.inference_mode():
            return self.model.generate(*args, **kwargs)

    # ---- quantize / pack / save
    @torch.no_grad()
    def quantize(self, tokenizer=None, quant_config=None, calib_data="pileval", split="train", text_column="text",
                 duo_scaling=True, modules_to_not_convert=None, export_compatible=False, n_samples=128, seqlen=512):
        """reference base.py:92-122.  ``calib_data`` must be caller-supplied text / tokens (no dataset download)."""
        self.quant_config = AwqConfig.from_dict(quant_config or {})
        if modules_to_not_convert is not None:
            self.quant_config.modules_to_not_convert = modules_to_not_convert
        self.quantizer = AwqQuantizer(self, self.model, tokenizer, self.quant_config.w_bit, self.quant_config.q_group_size,
                                      self.quant_config.version, calib_data, split, text_column, duo_scaling,
                                      modules_to_not_convert=modules_to_not_convert, export_compatible=export_compatible,
                                      n_samples=n_samples, seqlen=seqlen)
        self.quantizer.quantize()
        self.search_result = self.quantizer.search_result
        self.is_quantized = True

    @torch.no_grad()
    def pack(self):
        """After ``quantize(export_compatible=True)``: turn the scaled / clipped fp16 linears into packed modules
        (reference base.py:124-138)."""
        self.quantizer.pack()

    def save_quantized(self, save_dir, safetensors=True, shard_size="10GB"):
        """reference base.py:144-193 — same files: config.json (+ quantization_config), generation_config.json,
        quant_config.json, model.safetensors[-0000i-of-0000n + index] or pytorch_model.bin."""
        if getattr(self.model, "qb200_fused", False):
            raise RuntimeError("save_quantized needs the un-fused module tree; load with fuse_layers=False to re-save")
        save_dir = save_dir.rstrip("/") or "/"
        os.makedirs(save_dir, exist_ok=True)
        self.model.config.quantization_config = self.quant_config.to_transformers_dict()
        self.model.config.save_pretrained(save_dir)
        gen_cfg = getattr(self.model, "generation_config", None)
        if gen_cfg is not None:
            gen_cfg.save_pretrained(save_dir)
        self.quant_config.save_pretrained(save_dir)
        if self.processor is not None:
            self.processor.save_pretrained(save_dir)

        state = self.model.state_dict()
        tied = getattr(self.model.config, "tie_word_embeddings", False)
        if tied and "lm_head.weight" in state and "model.embed_tokens.weight" in state:
            state.pop("lm_head.weight")       # safetensors refuses aliased storage; tie_weights() restores it at load
        name = "model.safetensors" if safetensors else "pytorch_model.bin"
        shards, index = shard_state_dict(state, shard_size, name)
        for fname, shard in shards.items():
            path = os.path.join(save_dir, fname)
            if safetensors:
                from safetensors.torch import save_file
                save_file({k: v.detach().cpu().clone().contiguous() for k, v in shard.items()}, path, metadata={"format": "pt"})
            else:
                torch.save({k: v.detach().cpu() for k, v in shard.items()}, path)
        if index is not None:
            with open(os.path.join(save_dir, name + ".index.json"), "w") as f:
                json.dump(index, f, indent=4)

    # ---- loading
    @classmethod
    def from_pretrained(cls, model_path, model_type, torch_dtype: torch.dtype = torch.float16, trust_remote_code=True,
                        safetensors=False, device_map=None, **model_init_kwargs):
        """The fp16 model to be quantized (reference base.py:195-238)."""
        import transformers
        _, config, quant_config = cls._load_config(cls, model_path, "", safetensors, trust_remote_code=trust_remote_code)
        model = transformers.AutoModelForCausalLM.from_pretrained(model_path, trust_remote_code=trust_remote_code,
                                                                  dtype=torch_dtype, **model_init_kwargs)
        if device_map is not None:          # transformers' own device_map needs accelerate: place the whole model
            model.to(_resolve_device(device_map))
        model.eval()
        return cls(model, model_type, is_quantized=False, config=config, quant_config=quant_config, processor=None)

    @classmethod
    def from_quantized(cls, model_path, model_type, model_filename="", max_new_tokens=None, torch_dtype=torch.float16,
                       trust_remote_code=True, safetensors=True, is_quantized=True, fuse_layers=False, use_exllama=False,
                       use_exllama_v2=False, version="QUICK", device_map="balanced", offload_folder=None, batch_size=1,
                       **config_kwargs):
        """reference base.py:240-339."""
        import transformers
        if use_exllama or use_exllama_v2:
            raise NotImplementedError("ExLlama kernels are not part of this library")
        config_kwargs.pop("use_quick", None)
        weights_path, config, quant_config = cls._load_config(cls, model_path, model_filename, safetensors, version,
                                                              trust_remote_code, max_new_tokens=max_new_tokens, **config_kwargs)
        if quant_config.version not in ("QUICK", "GEMM"):
            raise NotImplementedError(f"checkpoint version {quant_config.version!r}: QUICK (native) and GEMM (converted at "
                                      "load) are supported")
        device = _resolve_device(device_map)

        hf_quant = getattr(config, "quantization_config", None)
        if hf_quant is not None:             # keep transformers' own AWQ integration out of the skeleton build
            try:
                delattr(config, "quantization_config")
            except AttributeError:
                config.quantization_config = None
        with torch.device("meta"):
            model = transformers.AutoModelForCausalLM.from_config(config, dtype=torch_dtype, trust_remote_code=trust_remote_code)
        cls._load_quantized_modules(cls, model, quant_config, "QUICK", use_exllama=False, use_exllama_v2=False)
        model.to_empty(device=device)
        cls._reinit_nonpersistent(model, config, device)
        cls._load_checkpoint(model, weights_path, safetensors, quant_config, device, torch_dtype)
        model.tie_weights()
        model.eval()
        model.config.quantization_config = dict(quant_config.to_transformers_dict(), version="quick")
        loaded_config = AwqConfig.from_dict(dict(quant_config.to_dict(), version="QUICK"))
        if fuse_layers:
            cls.fuse_layers(model, batch_size=batch_size)
        return cls(model, model_type, is_quantized=is_quantized, config=config, quant_config=loaded_config, processor=None)

    def _load_config(self, model_path, model_filename, safetensors=True, version="QUICK", trust_remote_code=True,
                     max_new_tokens=4096, **config_kwargs):
        """reference base.py:341-387, local directories only."""
        import transformers
        if not os.path.isdir(model_path):
            raise FileNotFoundError(f"{model_path} is not a local directory (no hub access; download the checkpoint first)")
        weights_path = os.path.join(model_path, model_filename) if model_filename else model_path
        quant_config = AwqConfig.from_pretrained(model_path)
        config = transformers.AutoConfig.from_pretrained(model_path, trust_remote_code=trust_remote_code, **config_kwargs)
        if max_new_tokens is None and hasattr(self, "max_new_tokens_key"):
            config.max_new_tokens = getattr(config, self.max_new_tokens_key, 2048)
        else:
            config.max_new_tokens = 2048 if max_new_tokens is None else max_new_tokens
        return weights_path, config, quant_config

    def _load_quantized_modules(self, model, quant_config, version, use_exllama=False, use_exllama_v2=False):
        """Swap every decoder-layer nn.Linear for an empty WQLinear_QUICK (reference base.py:389-440; the A100
        split-K override there, :411-432, is a tuning of the Ampere kernel and has no meaning here)."""
        if not quant_config.zero_point:
            raise AssertionError("We only support zero_point quantization now.")
        if version != "QUICK":
            raise NotImplementedError(version)
        for layer in self.get_model_layers(model):
            named = exclude_layers_to_not_quantize(get_named_linears(layer), quant_config.modules_to_not_convert)
            for name, lin in named.items():
                set_op_by_name(layer, name, WQLinear_QUICK.from_linear(lin, quant_config.w_bit, quant_config.q_group_size, True))
        gc.collect()

    @staticmethod
    def _reinit_nonpersistent(model, config, device):
        """Buffers that are computed in __init__ and never saved (rotary inv_freq) are garbage after to_empty():
        rebuild the modules that own them."""
        for name, mod in list(model.named_modules()):
            nonpersistent = getattr(mod, "_non_persistent_buffers_set", set())
            if not nonpersistent or isinstance(mod, WQLinear_QUICK):
                continue
            try:
                fresh = type(mod)(config=config, device=device)
            except TypeError as e:
                raise NotImplementedError(f"cannot rebuild the non-persistent buffers {sorted(nonpersistent)} of "
                                          f"{name} ({type(mod).__name__})") from e
            set_op_by_name(model, name, fresh)

    @staticmethod
    def _load_checkpoint(model, weights_path, safetensors, quant_config, device, dtype):
        targets: Dict[str, torch.Tensor] = dict(model.state_dict(keep_vars=True))
        qmods = {n: m for n, m in model.named_modules() if isinstance(m, WQLinear_QUICK)}
        pending_gemm: Dict[str, Dict[str, torch.Tensor]] = {}
        seen = set()
        for file in _checkpoint_files(weights_path, safetensors):
            for key, t in _read_checkpoint_file(file).items():
                owner, _, leaf = key.rpartition(".")
                if quant_config.version == "GEMM" and owner in qmods and leaf in ("qweight", "qzeros", "scales"):
                    grp = pending_gemm.setdefault(owner, {})
                    grp[leaf] = t.to(device)
                    if len(grp) == 3:
                        conv = WQLinear_QUICK.from_awq_gemm(grp["qweight"], grp["qzeros"], grp["scales"])
                        m = qmods[owner]
                        for nm in ("qweight", "qzeros", "scales"):
                            src = getattr(conv, nm)
                            if src.shape != getattr(m, nm).shape:
                                raise ValueError(f"{owner}.{nm}: converted shape {tuple(src.shape)} != module "
                                                 f"{tuple(getattr(m, nm).shape)}")
                            getattr(m, nm).copy_(src)
                            seen.add(f"{owner}.{nm}")
                        del pending_gemm[owner]
                    continue
                dst = targets.get(key)
                if dst is None:
                    continue            # e.g. rotary inv_freq of old checkpoints
                if dst.shape != t.shape:
                    raise ValueError(f"{key}: checkpoint shape {tuple(t.shape)} != module {tuple(dst.shape)}")
                with torch.no_grad():
                    dst.copy_(t.to(device=device, dtype=dst.dtype))
                seen.add(key)
        if pending_gemm:
            raise KeyError(f"incomplete qweight/qzeros/scales triples for {sorted(pending_gemm)[:4]}")
        tied = getattr(model.config, "tie_word_embeddings", False)
        missing = [k for k in targets if k not in seen and not (tied and k == "lm_head.weight")]
        if missing:
            raise KeyError(f"{len(missing)} tensors missing from the checkpoint, e.g. {missing[:4]}")
