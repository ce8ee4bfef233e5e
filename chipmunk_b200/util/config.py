"""GLOBAL_CONFIG: same keys and YAML deep-merge behaviour as the reference
(src/chipmunk/util/config.py:4-107), so the examples' `--chipmunk-config` files load unchanged.

B200 difference: `offloading.global_disable_offloading` defaults to True.  The reference moves
each layer's caches to pinned host memory because an 80 GB H100 cannot hold them (731 MB per
layer x 60 layers at HunyuanVideo 720p); 180 GB of HBM3e can, so caches stay resident unless a
config file turns offloading back on -- and then `offloading.backing: peer` parks them in a neighbour GPU's HBM over
NVLink instead of in host memory (util/storage/).
"""
from __future__ import annotations

import copy
from typing import Any, Dict

_SKIP = {7, 11, 13, 14, 15, 17, 18, 19, 21, 22, 23, 25, 26, 27, 29, 31, 33, 34, 35, 37, 38, 39, 41, 42, 43}

BASE_CONFIG: Dict[str, Any] = {
    "num_model_invocations_per_inference_step": 1,
    "should_profile": False,
    "generation_index": 0,
    "steps": 50,
    "world_size": 1,
    "mlp": {
        "is_enabled": True,
        "is_fp8": False,
        "top_keys": "dd",            # must be overridden with a float by the model's config file
        "random_keys": 0.05,
        "full_step_every": 10,
        "block_mask_cache": 2,
        "first_n_dense_layers": 2,
        # kernel-coupled constants
        "counts_multiple_of": 256,
        "bm": 128,
        "mbm": 128,
    },
    "patchify": {"is_enabled": True, "chunk_size_1": 8, "chunk_size_2": 4},
    "attn": {
        "is_enabled": True,
        "top_keys": 0.05,
        "random_keys": 0.01,
        "local_voxels": 0,
        "local_1d_window": 0,
        "first_n_dense_layers": 2,
        "full_step_every": 10,
        "full_step_schedule": None,   # a set of step numbers overrides full_step_every
        "recompute_mask": True,
        "should_compress_indices": True,
        # kernel-coupled constants
        "counts_multiple_of": 128,
        "pad_qkv_before_kernel": True,
        "mbm": 192,
    },
    "offloading": {
        "global_disable_offloading": True,
        # B200 addition: where an offloaded cache lives.  "host" = pinned host memory (the reference's only option,
        # PCIe ~55 GB/s); "peer" = the HBM of another GPU of the NVSwitch domain (`peer_device`, NVLink ~770 GB/s per
        # direction): the residency manager for configurations whose caches exceed one GPU's 180 GB
        "backing": "host",
        "peer_device": None,
        "mlp.out_cache": False,
        "mlp.indices": False,
        "mlp.counts": False,
        "mlp.sparse_act_T": False,
        "mlp.blockmean_mid_cache": False,
        "attn.out_cache": True,
        "attn.indices": True,
        "attn.counts": False,
        "attn.lse_constants": False,
        "text_encoders": True,
    },
    "step_caching": {"is_enabled": True, "skip_step_schedule": set(_SKIP)},
}

GLOBAL_CONFIG: Dict[str, Any] = copy.deepcopy(BASE_CONFIG)


def update_global_config(config: Dict[str, Any]) -> None:
    """Shallow top-level update (reference config.py:80-85)."""
    GLOBAL_CONFIG.update(config)


def _merge(dst: Dict[str, Any], src: Dict[str, Any]) -> None:
    for key, val in src.items():
        if isinstance(val, dict) and isinstance(dst.get(key), dict):
            _merge(dst[key], val)
        else:
            dst[key] = val


def load_from_file(config_file: str) -> None:
    """Deep-merge a YAML file into GLOBAL_CONFIG (reference config.py:98-107)."""
    import yaml

    with open(config_file, "r") as f:
        loaded = yaml.safe_load(f)
    if loaded:
        _merge(GLOBAL_CONFIG, loaded)
        print(f"CHIPMUNK: using config file {config_file}")


def reset_to_defaults() -> None:
    GLOBAL_CONFIG.clear()
    GLOBAL_CONFIG.update(copy.deepcopy(BASE_CONFIG))
