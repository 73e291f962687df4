"""Checkpoint wire format of the reference.

train_shapenet.py:660-678 / train_modelnet.py:491-509 save, on the best validation epoch, either
    my_model.state_dict()                                      (DDP-wrapped: every key starts with "module.")
or, with dynamic bin boundaries,
    {"model_state_dict": my_model.state_dict(),
     "bin_boundaries":   [ds.bin_boundaries for ds in my_model.module.block.downsample_list]}
where each bin_boundaries is the [upper, lower] pair of (1,1,1,nb) tensors that utils/ops.py:174-236 maintains -- module state
that lives OUTSIDE the state_dict.  test_shapenet.py:173-187 loads it back, writes `upper[0,0,0,1:]` of every layer into the
config and switches the dynamic update off.  `load` does the same for our models (or for the reference's wiring with the
blocks patched in): parameters by name, boundaries onto the DownSampleToken modules, EMA frozen.
"""
from __future__ import annotations

from typing import Dict, List, Union

import torch
from torch import nn

from . import blocks

Tensor = torch.Tensor


def _strip_module(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    return {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}


def downsample_layers(model: nn.Module) -> List[nn.Module]:
    """the DownSampleToken modules in forward order (block.downsample_list of both models)."""
    core = model.module if hasattr(model, "module") else model
    return list(core.block.downsample_list)


def state(model: nn.Module, dynamic_boundaries: bool = True, ddp_prefix: bool = True) -> Union[Dict[str, Tensor], dict]:
    """What the reference's training script would torch.save for this model (train_shapenet.py:663-675)."""
    core = model.module if hasattr(model, "module") else model
    sd = {("module." + k if ddp_prefix else k): v for k, v in core.state_dict().items()}
    if not dynamic_boundaries:
        return sd
    return {"model_state_dict": sd, "bin_boundaries": [ds.bin_boundaries for ds in downsample_layers(core)]}


def save(model: nn.Module, path: str, dynamic_boundaries: bool = True) -> None:
    torch.save(state(model, dynamic_boundaries), path)


def load(model: nn.Module, ckpt, map_location=None, strict: bool = True) -> nn.Module:
    """ckpt: a path or an already loaded object in either format above.  Loads the parameters / buffers by name (with or
    without the DDP "module." prefix) and, when the checkpoint carries bin boundaries, installs them on the DownSample
    layers and freezes the EMA (test_shapenet.py:179-183: dynamic_boundaries = False, bin_boundaries = upper[0,0,0,1:])."""
    if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, "__fspath__"):
        ckpt = torch.load(ckpt, map_location=map_location, weights_only=False)
    core = model.module if hasattr(model, "module") else model
    bounds = None
    if isinstance(ckpt, dict) and "model_state_dict" in ckpt:
        sd, bounds = ckpt["model_state_dict"], ckpt.get("bin_boundaries")
    else:
        sd = ckpt
    core.load_state_dict(_strip_module(sd), strict=strict)
    if bounds is not None:
        layers = downsample_layers(core)
        if len(bounds) != len(layers):
            raise RuntimeError(f"checkpoint holds bin boundaries for {len(bounds)} DownSample layers, the model has {len(layers)}")
        dev = next(core.parameters()).device
        for ds, pair in zip(layers, bounds):
            upper = torch.as_tensor(pair[0], dtype=torch.float32).reshape(-1)
            nb = upper.numel()
            if isinstance(ds, blocks.DownSampleToken) and nb != ds.num_bins:
                raise RuntimeError(f"checkpoint boundaries have {nb} bins, the layer has {ds.num_bins}")
            cuts = upper[1:].tolist()              # the reference keeps exactly this list (test_shapenet.py:181-184)
            ds.bin_boundaries = [torch.tensor([float("inf")] + cuts, device=dev).reshape(1, 1, 1, nb),
                                 torch.tensor(cuts + [float("-inf")], device=dev).reshape(1, 1, 1, nb)]
            ds.dynamic_boundaries_enable = False
    return model
