"""``quant_config.json`` — same file name, keys and defaults as the reference's AwqConfig
(quick/awq/models/_config.py:9-92) so checkpoints written by either side load in the other.  Local directories
only: there is no hub access in the environments this library targets, so a non-directory path raises instead of
being treated as a repo id."""
import json
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional


@dataclass
class AwqConfig:
    quant_method: str = field(default="awq")
    zero_point: bool = field(default=True)
    q_group_size: int = field(default=128)
    w_bit: int = field(default=4)
    version: str = field(default="GEMM")
    modules_to_not_convert: Optional[List] = None
    config_file_name = "quant_config.json"

    @classmethod
    def from_dict(cls, quant_config: Optional[Dict] = None):
        return cls(**quant_config) if quant_config else cls()

    @classmethod
    def from_transformers_dict(cls, d: Dict):
        """Inverse of to_transformers_dict: the ``quantization_config`` entry of config.json."""
        return cls(quant_method=d.get("quant_method", "awq"), zero_point=d.get("zero_point", True),
                   q_group_size=d.get("group_size", 128), w_bit=d.get("bits", 4),
                   version=str(d.get("version", "gemm")).upper(), modules_to_not_convert=d.get("modules_to_not_convert"))

    @classmethod
    def from_pretrained(cls, save_dir: str, **kwargs):
        if not os.path.isdir(save_dir):
            raise FileNotFoundError(f"{save_dir} is not a local directory (no hub access; download the checkpoint first)")
        path = os.path.join(save_dir, cls.config_file_name)
        if os.path.exists(path):
            with open(path, "r", encoding="utf-8") as f:
                return cls(**json.load(f))
        # newer AutoAWQ checkpoints carry the settings only in config.json -> quantization_config
        cfg_path = os.path.join(save_dir, "config.json")
        if os.path.exists(cfg_path):
            with open(cfg_path, "r", encoding="utf-8") as f:
                qc = json.load(f).get("quantization_config")
            if qc:
                return cls.from_transformers_dict(qc)
        return cls()

    def save_pretrained(self, save_dir: str, **kwargs):
        with open(os.path.join(save_dir, self.config_file_name), "w", encoding="utf-8") as f:
            json.dump(self.to_dict(), f, indent=4)

    def to_dict(self):
        return {"zero_point": self.zero_point, "q_group_size": self.q_group_size, "w_bit": self.w_bit,
                "version": self.version, "modules_to_not_convert": self.modules_to_not_convert}

    def to_transformers_dict(self):
        return {"quant_method": self.quant_method, "zero_point": self.zero_point, "group_size": self.q_group_size,
                "bits": self.w_bit, "version": self.version.lower(), "modules_to_not_convert": self.modules_to_not_convert}
